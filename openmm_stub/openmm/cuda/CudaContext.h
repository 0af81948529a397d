#ifndef OPENMM_STUB_CUDACONTEXT_H_
#define OPENMM_STUB_CUDACONTEXT_H_
// Stand-in for the part of OpenMM's CudaContext a force plugin touches: the device buffers and their layout.
//   posq            float4 per slot (single / mixed) or double4 (double); posqCorrection float4 in mixed mode
//   force           signed 64-bit fixed point (scale 2^32), [x: padded][y: padded][z: padded]
//   energyBuffer    float (single) or double (mixed / double) accumulation buffer
//   atomIndex       slot -> original atom index (the CUDA platform re-orders atoms for locality)
#include "openmm/cuda/CudaArray.h"
#include "openmm/cuda/CudaForceInfo.h"
#include "openmm/System.h"
#include "openmm/Vec3.h"
#include <cmath>
#include <vector>

namespace OpenMM {

class CudaContext {
public:
    class ReorderListener {
    public:
        virtual ~ReorderListener() {}
        virtual void execute() = 0;
    };
    static const int TileSize = 32;

    CudaContext(const System& system, int deviceIndex, const std::string& precision, bool reverseOrder)
            : numAtoms(system.getNumParticles()), deviceIndex(deviceIndex), stream(nullptr),
              useDouble(precision == "double"), useMixed(precision == "mixed") {
        if (cudaSetDevice(deviceIndex) != cudaSuccess) throw OpenMMException("No CUDA device");
        paddedNumAtoms = ((numAtoms + TileSize - 1) / TileSize) * TileSize;
        posq.initialize(*this, paddedNumAtoms, useDouble ? 4 * sizeof(double) : 4 * sizeof(float), "posq");
        if (useMixed) posqCorrection.initialize(*this, paddedNumAtoms, 4 * sizeof(float), "posqCorrection");
        force.initialize(*this, 3 * (size_t) paddedNumAtoms, sizeof(long long), "force");
        energyBuffer.initialize(*this, 64, (useDouble || useMixed) ? sizeof(double) : sizeof(float), "energyBuffer");
        atomIndexDevice.initialize(*this, paddedNumAtoms, sizeof(int), "atomIndex");
        atomIndex.resize(paddedNumAtoms);
        for (int i = 0; i < paddedNumAtoms; i++) atomIndex[i] = i;
        if (reverseOrder)
            for (int i = 0; i < numAtoms; i++) atomIndex[i] = numAtoms - 1 - i;
        atomIndexDevice.upload(atomIndex);
        cudaStreamCreate(&stream);
    }
    ~CudaContext() {
        for (ReorderListener* l : reorderListeners) delete l;
        for (CudaForceInfo* f : forces) delete f;
        cudaStreamDestroy(stream);
    }
    void setAsCurrent() { cudaSetDevice(deviceIndex); }
    int getDeviceIndex() const { return deviceIndex; }
    int getContextIndex() const { return 0; }
    int getNumAtoms() const { return numAtoms; }
    int getPaddedNumAtoms() const { return paddedNumAtoms; }
    bool getUseDoublePrecision() const { return useDouble; }
    bool getUseMixedPrecision() const { return useMixed; }
    cudaStream_t getCurrentStream() { return stream; }
    CudaArray& getPosq() { return posq; }
    CudaArray& getPosqCorrection() { return posqCorrection; }
    CudaArray& getForce() { return force; }
    CudaArray& getEnergyBuffer() { return energyBuffer; }
    CudaArray& getAtomIndexArray() { return atomIndexDevice; }
    const std::vector<int>& getAtomIndex() const { return atomIndex; }
    void addForce(CudaForceInfo* info) { forces.push_back(info); }
    void addReorderListener(ReorderListener* listener) { reorderListeners.push_back(listener); }

    // ---- what OpenMM's own UpdateStateData / CalcForcesAndEnergy kernels do, for the stub Context ----
    void setPositions(const std::vector<Vec3>& positions) {
        std::vector<double> full(4 * (size_t) paddedNumAtoms, 0.0);
        for (int slot = 0; slot < numAtoms; slot++)
            for (int c = 0; c < 3; c++) full[4 * slot + c] = positions[atomIndex[slot]][c];
        if (useDouble) {
            posq.upload(full.data());
        } else {
            std::vector<float> hi(full.size()), lo(full.size());
            for (size_t i = 0; i < full.size(); i++) { hi[i] = (float) full[i]; lo[i] = (float) (full[i] - (double) hi[i]); }
            posq.upload(hi.data());
            if (useMixed) posqCorrection.upload(lo.data());
        }
    }
    void getPositions(std::vector<Vec3>& positions) {
        positions.assign(numAtoms, Vec3());
        if (useDouble) {
            std::vector<double> full(4 * (size_t) paddedNumAtoms);
            posq.download(full.data());
            for (int slot = 0; slot < numAtoms; slot++) positions[atomIndex[slot]] = Vec3(full[4 * slot], full[4 * slot + 1], full[4 * slot + 2]);
        } else {
            std::vector<float> hi(4 * (size_t) paddedNumAtoms), lo(4 * (size_t) paddedNumAtoms, 0.f);
            posq.download(hi.data());
            if (useMixed) posqCorrection.download(lo.data());
            for (int slot = 0; slot < numAtoms; slot++)
                positions[atomIndex[slot]] = Vec3((double) hi[4 * slot] + lo[4 * slot], (double) hi[4 * slot + 1] + lo[4 * slot + 1],
                                                  (double) hi[4 * slot + 2] + lo[4 * slot + 2]);
        }
    }
    void clearBuffers() {
        cudaMemsetAsync(reinterpret_cast<void*>(force.getDevicePointer()), 0, force.getSize() * sizeof(long long), stream);
        cudaMemsetAsync(reinterpret_cast<void*>(energyBuffer.getDevicePointer()), 0, energyBuffer.getSize() * energyBuffer.getElementSize(), stream);
    }
    void getForces(std::vector<Vec3>& out) {
        cudaStreamSynchronize(stream);
        std::vector<long long> raw;
        force.download(raw);
        out.assign(numAtoms, Vec3());
        const double scale = 1.0 / (double) 0x100000000LL;
        for (int slot = 0; slot < numAtoms; slot++)
            out[atomIndex[slot]] = Vec3(scale * raw[slot], scale * raw[slot + paddedNumAtoms], scale * raw[slot + 2 * paddedNumAtoms]);
    }
    double sumEnergy() {
        cudaStreamSynchronize(stream);
        double total = 0.0;
        if (useDouble || useMixed) {
            std::vector<double> e; energyBuffer.download(e);
            for (double v : e) total += v;
        } else {
            std::vector<float> e; energyBuffer.download(e);
            for (float v : e) total += v;
        }
        return total;
    }
    /** Change the slot order (what OpenMM does periodically) and tell every listener. */
    void reorderAtoms(const std::vector<int>& newAtomIndex) {
        std::vector<Vec3> positions;
        getPositions(positions);
        for (int i = 0; i < numAtoms; i++) atomIndex[i] = newAtomIndex[i];
        atomIndexDevice.upload(atomIndex);
        setPositions(positions);
        for (ReorderListener* l : reorderListeners) l->execute();
    }
private:
    int numAtoms, paddedNumAtoms, deviceIndex;
    cudaStream_t stream;
    bool useDouble, useMixed;
    CudaArray posq, posqCorrection, force, energyBuffer, atomIndexDevice;
    std::vector<int> atomIndex;
    std::vector<CudaForceInfo*> forces;
    std::vector<ReorderListener*> reorderListeners;
};

} // namespace OpenMM
#endif
