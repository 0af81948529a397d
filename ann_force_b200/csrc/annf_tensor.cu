// Layer-wise tensor-core path (tcgen05, 3xTF32) for wide layers — placeholder until the kernels land.
#include <string>
#include <vector>

#include "annf_common.cuh"
#include "annf_profiler.h"

namespace annf {

struct TensorPlan {};

TensorPlan* tensor_plan_create(const NetView&, const std::vector<std::vector<double>>&, std::string* why_not) {
    if (why_not) *why_not = "tensor-core path not built into this library yet";
    return nullptr;
}

void tensor_plan_destroy(TensorPlan* plan) { delete plan; }

cudaError_t tensor_plan_run(TensorPlan*, Profiler*, const NetView&, int, const float*, int, const float*, float*, double*,
                            cudaStream_t, long long*) {
    return cudaErrorNotSupported;
}

}  // namespace annf
