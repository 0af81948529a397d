#ifndef CUDA_ANN_KERNELS_H_
#define CUDA_ANN_KERNELS_H_
// CUDA-platform kernel of ANN_Force, B200 edition.
//
// Same class name, constructor and three virtuals as the reference's
// /root/reference/platforms/cuda/include/CudaANN_Kernels.h:45-125, so CudaANN_KernelFactory and OpenMM see
// no difference.  Everything behind them is new: instead of uploading eleven CudaArrays and string-generating
// a CUDA snippet for CudaBondedUtilities (CudaANN_Kernels.cpp:81-365), initialize() hands the parameters to
// the C-ABI library (include/annforce_b200.h) and execute() enqueues ONE fused sm_100a kernel on the
// context's stream, directly on OpenMM's posq / force / energy buffers.
#include "OpenMM_ANN.h"
#include "annforce_b200.h"
#include "openmm/ANN_Kernels.h"
#include "openmm/cuda/CudaArray.h"
#include "openmm/cuda/CudaContext.h"

namespace OpenMM {

class CudaCalcANN_ForceKernel : public CalcANN_ForceKernel {
public:
    CudaCalcANN_ForceKernel(std::string name, const Platform& platform, CudaContext& cu, const System& system)
            : CalcANN_ForceKernel(name, platform), cu(cu), system(system), handle(nullptr) {}
    ~CudaCalcANN_ForceKernel();
    void initialize(const System& system, const ANN_Force& force);
    double execute(ContextImpl& context, bool includeForces, bool includeEnergy);
    void copyParametersToContext(ContextImpl& context, const ANN_Force& force);
    /** Called by the context's reorder listener: atom slots changed. */
    void atomsWereReordered();
private:
    CudaContext& cu;
    const System& system;
    annf_handle handle;
};

} // namespace OpenMM
#endif
