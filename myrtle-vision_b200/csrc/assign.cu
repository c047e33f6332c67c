// assign.cu — batched bipartite matching on the device: the detection loss's
// Hungarian step (reference src/myrtle_vision/models/matcher.py:83-86, SciPy on the host after a
// device->host copy) as one kernel launch, one warp per image, so that the detection train step has
// no host synchronisation between forward and backward and can be captured as a single CUDA graph.
#include "common.cuh"
#include "assign_core.cuh"
#include "../../include/mv_b200.h"

namespace mv {

extern int64_t g_launches;

__global__ void __launch_bounds__(32) assign_kernel(const float* __restrict__ cost, const int* __restrict__ sizes,
                                                    int Q, int Tmax, int* __restrict__ match, int* __restrict__ flag) {
    __shared__ mv_assign::Work work;
    const int b = blockIdx.x;
    int nt = sizes[b];
    nt = nt < 0 ? 0 : (nt > Tmax ? Tmax : nt);
    mv_assign::match_block(cost + int64_t(b) * Q * Tmax, Q, nt, Tmax, match + int64_t(b) * Tmax, Tmax, flag,
                           work, threadIdx.x, 32);
}

}  // namespace mv

using namespace mv;

extern "C" int mv_linear_sum_assignment(const float* cost, const int* sizes, int B, int Q, int Tmax, int* match,
                                        int* flag, void* stream) {
    MV_CHECK(B >= 0 && Q >= 0 && Tmax >= 0, "mv_linear_sum_assignment: negative extent");
    if (B == 0 || Tmax == 0) return 0;
    MV_CHECK(cost && sizes && match, "mv_linear_sum_assignment: null pointer");
    MV_CHECK(Q <= mv_assign::kMaxSide && Tmax <= mv_assign::kMaxSide,
             "mv_linear_sum_assignment: at most %d predictions / targets per image (got %d / %d)",
             mv_assign::kMaxSide, Q, Tmax);
    assign_kernel<<<B, 32, 0, static_cast<cudaStream_t>(stream)>>>(cost, sizes, Q, Tmax, match, flag);
    g_launches++;
    return check_cuda(cudaGetLastError(), "assign launch");
}
