#ifndef OPENMM_STUB_CONTEXTIMPL_H_
#define OPENMM_STUB_CONTEXTIMPL_H_
#include "openmm/Vec3.h"
#include "openmm/Platform.h"
#include "openmm/System.h"
#include "openmm/Integrator.h"
#include "openmm/internal/ForceImpl.h"
#include <map>
#include <string>
#include <vector>

namespace OpenMM {

class Context;

// The stub adds five "stub*" virtuals to the platforms it ships (see StubPlatformHooks) in place
// of OpenMM's UpdateStateData / CalcForcesAndEnergy kernels: they move positions in, forces out,
// and bracket one force evaluation.
class StubPlatformHooks {
public:
    virtual ~StubPlatformHooks() {}
    virtual void stubSetPositions(ContextImpl& context, const std::vector<Vec3>& positions) const = 0;
    virtual void stubGetPositions(ContextImpl& context, std::vector<Vec3>& positions) const = 0;
    virtual void stubGetForces(ContextImpl& context, std::vector<Vec3>& forces) const = 0;
    virtual void stubBeginComputation(ContextImpl& context) const = 0;        // zero force / energy accumulators
    virtual double stubFinishComputation(ContextImpl& context) const = 0;     // energy accumulated on the device, if any
};

class ContextImpl {
public:
    ContextImpl(Context& owner, const System& system, Integrator& integrator, Platform* platform,
                const std::map<std::string, std::string>& properties)
            : owner(owner), system(system), integrator(integrator), platform(platform), platformData(nullptr) {
        hooks = dynamic_cast<StubPlatformHooks*>(platform);
        if (hooks == nullptr)
            throw OpenMMException("openmm_stub: platform '" + platform->getName() + "' cannot host a Context");
        platform->contextCreated(*this, properties);
        for (int i = 0; i < system.getNumForces(); i++)
            forceImpls.push_back(system.getForce(i).createImpl());
        for (ForceImpl* impl : forceImpls) {
            if (!platform->supportsKernels(impl->getKernelNames()))
                throw OpenMMException("Specified a Platform for a Context which does not support all required kernels");
            impl->initialize(*this);
        }
    }
    ~ContextImpl() {
        for (ForceImpl* impl : forceImpls) delete impl;
        platform->contextDestroyed(*this);
    }
    Context& getOwner() { return owner; }
    const System& getSystem() const { return system; }
    Integrator& getIntegrator() { return integrator; }
    Platform& getPlatform() { return *platform; }
    void* getPlatformData() { return platformData; }
    const void* getPlatformData() const { return platformData; }
    void setPlatformData(void* data) { platformData = data; }
    std::vector<ForceImpl*>& getForceImpls() { return forceImpls; }

    void setPositions(const std::vector<Vec3>& positions) { hooks->stubSetPositions(*this, positions); }
    void getPositions(std::vector<Vec3>& positions) { hooks->stubGetPositions(*this, positions); }
    void getForces(std::vector<Vec3>& forces) { hooks->stubGetForces(*this, forces); }

    double calcForcesAndEnergy(bool includeForces, bool includeEnergy, int groups = 0xFFFFFFFF) {
        hooks->stubBeginComputation(*this);
        double energy = 0.0;
        for (ForceImpl* impl : forceImpls)
            energy += impl->calcForcesAndEnergy(*this, includeForces, includeEnergy, groups);
        energy += hooks->stubFinishComputation(*this);
        return energy;
    }
private:
    Context& owner;
    const System& system;
    Integrator& integrator;
    Platform* platform;
    StubPlatformHooks* hooks;
    void* platformData;
    std::vector<ForceImpl*> forceImpls;
};

} // namespace OpenMM
#endif
