// Plugin entry points + kernel factory (same exported symbols and behaviour as
// /root/reference/platforms/cuda/src/CudaANN_KernelFactory.cpp:43-72).
#include "CudaANN_KernelFactory.h"
#include "CudaANN_Kernels.h"
#include "openmm/OpenMMException.h"
#include "openmm/cuda/CudaPlatform.h"
#include "openmm/internal/ContextImpl.h"
#include "openmm/internal/windowsExport.h"
#include <exception>

using namespace OpenMM;

extern "C" OPENMM_EXPORT void registerPlatforms() {}

extern "C" OPENMM_EXPORT void registerKernelFactories() {
    try {
        Platform& platform = Platform::getPlatformByName("CUDA");
        platform.registerKernelFactory(CalcANN_ForceKernel::Name(), new CudaANN_KernelFactory());
    } catch (const std::exception&) {
        // no CUDA platform in this OpenMM build: nothing to register (the reference swallows this too)
    }
}

extern "C" OPENMM_EXPORT void registerANN_CudaKernelFactories() {
    try {
        Platform::getPlatformByName("CUDA");
    } catch (...) {
        Platform::registerPlatform(new CudaPlatform());
    }
    registerKernelFactories();
}

KernelImpl* CudaANN_KernelFactory::createKernelImpl(std::string name, const Platform& platform, ContextImpl& context) const {
    CudaContext& cu = *static_cast<CudaPlatform::PlatformData*>(context.getPlatformData())->contexts[0];
    if (name == CalcANN_ForceKernel::Name()) return new CudaCalcANN_ForceKernel(name, platform, cu, context.getSystem());
    throw OpenMMException("Tried to create kernel with illegal kernel name '" + name + "'");
}
