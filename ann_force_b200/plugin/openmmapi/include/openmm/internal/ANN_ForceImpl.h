#ifndef OPENMM_ANN_FORCE_IMPL_H_
#define OPENMM_ANN_FORCE_IMPL_H_
// ForceImpl glue between ANN_Force and the platform kernel
// (role of /root/reference/openmmapi/include/openmm/internal/ANN_ForceImpl.h:43-65).
#include "openmm/ANN_Force.h"
#include "openmm/Kernel.h"
#include "openmm/internal/ForceImpl.h"
#include <map>
#include <string>
#include <vector>

namespace OpenMM {

class ANN_ForceImpl : public ForceImpl {
public:
    explicit ANN_ForceImpl(const ANN_Force& owner);
    ~ANN_ForceImpl();
    void initialize(ContextImpl& context);
    const ANN_Force& getOwner() const { return owner; }
    void updateContextState(ContextImpl& context) {}
    double calcForcesAndEnergy(ContextImpl& context, bool includeForces, bool includeEnergy, int groups);
    std::map<std::string, double> getDefaultParameters() { return std::map<std::string, double>(); }
    std::vector<std::string> getKernelNames();
    void updateParametersInContext(ContextImpl& context);
private:
    const ANN_Force& owner;
    Kernel kernel;
};

} // namespace OpenMM
#endif
