#ifndef OPENMM_CUDA_ANN_KERNEL_FACTORY_H_
#define OPENMM_CUDA_ANN_KERNEL_FACTORY_H_
// Factory + plugin entry points; same symbols as /root/reference/platforms/cuda/include/CudaANN_KernelFactory.h.
#include "openmm/KernelFactory.h"

namespace OpenMM {

class CudaANN_KernelFactory : public KernelFactory {
public:
    KernelImpl* createKernelImpl(std::string name, const Platform& platform, ContextImpl& context) const;
};

} // namespace OpenMM

extern "C" OPENMM_EXPORT void registerPlatforms();
extern "C" OPENMM_EXPORT void registerKernelFactories();
extern "C" OPENMM_EXPORT void registerANN_CudaKernelFactories();
#endif
