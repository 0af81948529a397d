#ifndef OPENMM_STUB_REFERENCEPLATFORM_H_
#define OPENMM_STUB_REFERENCEPLATFORM_H_
#include "openmm/Platform.h"
#include "openmm/internal/ContextImpl.h"
#include "openmm/reference/RealVec.h"
#include "openmm/reference/SimTKOpenMMRealType.h"
#include <vector>

namespace OpenMM {

// CPU fp64 platform.  PlatformData keeps its arrays behind void* like OpenMM 6/7 did, which is
// what reference-platform plugins of that era reinterpret_cast from.
class ReferencePlatform : public Platform, public StubPlatformHooks {
public:
    class PlatformData {
    public:
        PlatformData(int numParticles) : numParticles(numParticles), stepCount(0), time(0.0) {
            positions = new std::vector<RealVec>(numParticles);
            velocities = new std::vector<RealVec>(numParticles);
            forces = new std::vector<RealVec>(numParticles);
            periodicBoxSize = new RealVec();
            periodicBoxVectors = new RealVec[3];
        }
        ~PlatformData() {
            delete (std::vector<RealVec>*) positions;
            delete (std::vector<RealVec>*) velocities;
            delete (std::vector<RealVec>*) forces;
            delete (RealVec*) periodicBoxSize;
            delete[] (RealVec*) periodicBoxVectors;
        }
        int numParticles, stepCount;
        double time;
        void* positions;
        void* velocities;
        void* forces;
        void* periodicBoxSize;
        void* periodicBoxVectors;
    };
    const std::string& getName() const { static const std::string name = "Reference"; return name; }
    void contextCreated(ContextImpl& context, const std::map<std::string, std::string>&) const {
        context.setPlatformData(new PlatformData(context.getSystem().getNumParticles()));
    }
    void contextDestroyed(ContextImpl& context) const { delete data(context); }
    void stubSetPositions(ContextImpl& context, const std::vector<Vec3>& positions) const {
        *(std::vector<RealVec>*) data(context)->positions = positions;
    }
    void stubGetPositions(ContextImpl& context, std::vector<Vec3>& positions) const {
        positions = *(std::vector<RealVec>*) data(context)->positions;
    }
    void stubGetForces(ContextImpl& context, std::vector<Vec3>& forces) const {
        forces = *(std::vector<RealVec>*) data(context)->forces;
    }
    void stubBeginComputation(ContextImpl& context) const {
        for (RealVec& f : *(std::vector<RealVec>*) data(context)->forces) f = RealVec();
    }
    double stubFinishComputation(ContextImpl&) const { return 0.0; }
private:
    static PlatformData* data(ContextImpl& context) { return reinterpret_cast<PlatformData*>(context.getPlatformData()); }
};

} // namespace OpenMM
#endif
