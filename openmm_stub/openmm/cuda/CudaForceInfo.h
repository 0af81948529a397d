#ifndef OPENMM_STUB_CUDAFORCEINFO_H_
#define OPENMM_STUB_CUDAFORCEINFO_H_
#include <vector>
namespace OpenMM {
// Tells the CUDA platform which particles a force couples, so that atom reordering keeps molecules intact.
class CudaForceInfo {
public:
    virtual ~CudaForceInfo() {}
    virtual bool areParticlesIdentical(int particle1, int particle2) { return true; }
    virtual int getNumParticleGroups() { return 0; }
    virtual void getParticlesInGroup(int index, std::vector<int>& particles) {}
    virtual bool areGroupsIdentical(int group1, int group2) { return true; }
};
} // namespace OpenMM
#endif
