// ForceImpl glue (role of /root/reference/openmmapi/src/ANN_ForceImpl.cpp:21-41).
#include "openmm/internal/ANN_ForceImpl.h"
#include "openmm/ANN_Kernels.h"
#include "openmm/internal/ContextImpl.h"

using namespace OpenMM;

ANN_ForceImpl::ANN_ForceImpl(const ANN_Force& owner) : owner(owner) {}
ANN_ForceImpl::~ANN_ForceImpl() {}

void ANN_ForceImpl::initialize(ContextImpl& context) {
    kernel = context.getPlatform().createKernel(CalcANN_ForceKernel::Name(), context);
    kernel.getAs<CalcANN_ForceKernel>().initialize(context.getSystem(), owner);
}

double ANN_ForceImpl::calcForcesAndEnergy(ContextImpl& context, bool includeForces, bool includeEnergy, int groups) {
    const bool in_group = (groups & (1 << owner.getForceGroup())) != 0;
    return in_group ? kernel.getAs<CalcANN_ForceKernel>().execute(context, includeForces, includeEnergy) : 0.0;
}

std::vector<std::string> ANN_ForceImpl::getKernelNames() { return std::vector<std::string>(1, CalcANN_ForceKernel::Name()); }

void ANN_ForceImpl::updateParametersInContext(ContextImpl& context) {
    kernel.getAs<CalcANN_ForceKernel>().copyParametersToContext(context, owner);
}
