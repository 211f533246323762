// tcgen05 / TMEM / TMA dense contractions (TF32 operands, fp32 accumulate).  Placeholder until the
// tensor-core kernels land: returning false routes the call to the fp32 CUDA-core kernels.
#include "kernels.cuh"
namespace vd {
bool gemm_tn_tc(LaunchCtx&, int, int, int, const float*, int64_t, const int32_t*, const float*, int64_t, float*, int64_t,
                float, const float*, int) { return false; }
bool gemm_atb_tc(LaunchCtx&, int, int, int64_t, const float*, int64_t, const int32_t*, const float*, int64_t, float*,
                 int64_t) { return false; }
}  // namespace vd
