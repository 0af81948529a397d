#ifndef OPENMM_STUB_CONTEXT_H_
#define OPENMM_STUB_CONTEXT_H_
#include "openmm/internal/ContextImpl.h"
#include "openmm/State.h"
#include <map>
#include <string>

namespace OpenMM {

class Context {
public:
    Context(const System& system, Integrator& integrator, Platform& platform) {
        impl = new ContextImpl(*this, system, integrator, &platform, std::map<std::string, std::string>());
        integrator.owner = this;
    }
    Context(const System& system, Integrator& integrator, Platform& platform,
            const std::map<std::string, std::string>& properties) {
        impl = new ContextImpl(*this, system, integrator, &platform, properties);
        integrator.owner = this;
    }
    ~Context() { delete impl; }
    const System& getSystem() const { return impl->getSystem(); }
    Platform& getPlatform() { return impl->getPlatform(); }
    void setPositions(const std::vector<Vec3>& positions) {
        if ((int) positions.size() != impl->getSystem().getNumParticles())
            throw OpenMMException("Called setPositions() on a Context with the wrong number of positions");
        impl->setPositions(positions);
    }
    State getState(int types, bool enforcePeriodicBox = false, int groups = 0xFFFFFFFF) {
        State state;
        bool wantForces = (types & State::Forces) != 0, wantEnergy = (types & State::Energy) != 0;
        if (wantForces || wantEnergy) {
            double energy = impl->calcForcesAndEnergy(wantForces, wantEnergy, groups);
            if (wantEnergy) state.potentialEnergy = energy;
            if (wantForces) impl->getForces(state.forces);
        }
        if (types & State::Positions) impl->getPositions(state.positions);
        return state;
    }
    ContextImpl& getImpl() { return *impl; }
private:
    friend class Force;
    Context(const Context&);
    Context& operator=(const Context&);
    ContextImpl* impl;
};

} // namespace OpenMM
#endif
