#ifndef OPENMM_STUB_VERLETINTEGRATOR_H_
#define OPENMM_STUB_VERLETINTEGRATOR_H_
#include "openmm/Integrator.h"
#include "openmm/OpenMMException.h"

namespace OpenMM {

// Only what plugin tests need: a step size to construct a Context with.  No dynamics.
class VerletIntegrator : public Integrator {
public:
    explicit VerletIntegrator(double stepSize) { setStepSize(stepSize); }
    void step(int) { throw OpenMMException("openmm_stub: VerletIntegrator::step is not implemented"); }
};

} // namespace OpenMM
#endif
