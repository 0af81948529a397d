#ifndef ANN_OPENMM_KERNELS_H_
#define ANN_OPENMM_KERNELS_H_
// Abstract platform kernel of ANN_Force — the same contract as the reference's
// /root/reference/openmmapi/include/openmm/ANN_Kernels.h:19-54 (name "CalcANN_Force").
#include "openmm/ANN_Force.h"
#include "openmm/KernelImpl.h"
#include "openmm/Platform.h"
#include "openmm/System.h"
#include <string>

namespace OpenMM {

class CalcANN_ForceKernel : public KernelImpl {
public:
    static std::string Name() { return "CalcANN_Force"; }
    CalcANN_ForceKernel(std::string name, const Platform& platform) : KernelImpl(name, platform) {}
    /** Read every parameter of `force`, build whatever the platform needs. */
    virtual void initialize(const System& system, const ANN_Force& force) = 0;
    /** Add this force's contribution to the context's forces / energy; returns the energy if the platform
     *  computes it on the host, 0 if it accumulates it on the device. */
    virtual double execute(ContextImpl& context, bool includeForces, bool includeEnergy) = 0;
    /** Re-read the umbrella centre and force constant from `force`. */
    virtual void copyParametersToContext(ContextImpl& context, const ANN_Force& force) = 0;
};

} // namespace OpenMM
#endif
