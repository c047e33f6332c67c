// optim.cu — multi-tensor AdamW fused with the next step's weight fake-quantisation.
//
// the step on the far side of backward (reference classification/train.py:274-277,
// timm AdamW).  One launch updates every parameter of the model; for the Linear weights the same thread
// that wrote the new fp32 weight also emits the tensor-core operands of the NEXT forward / backward —
// q(W) [out,in] and q(W)^T [in,out] in the fp16 (or fp32 for the 32-bit formats) container — so the
// per-step weight_fake_quant pass (torch.nn.qat.Linear re-quantises W every forward) never re-reads
// the weights from HBM.  HBM-bound: 28 B per parameter (p, g, m, v read; p, m, v written) + 4 B per
// Linear weight for the two fp16 operand copies.
//
// Work is cut into chunks of 1024 elements — a run of 1024 consecutive elements of a plain tensor,
// or one 32x32 tile of a weight matrix (so the transposed operand is written coalesced through
// shared memory).  Hyper-parameters that change between steps live in device memory: a captured
// CUDA graph replays the same launch with new values.
#include <type_traits>

#include "common.cuh"
#include "quant_dev.cuh"
#include "../../include/mv_b200.h"

namespace mv {
extern int64_t g_launches;

template <typename T> __device__ __forceinline__ T opt_cvt(float v);
template <> __device__ __forceinline__ float opt_cvt<float>(float v) { return v; }
template <> __device__ __forceinline__ __half opt_cvt<__half>(float v) { return __float2half_rn(v); }

struct AdamHyper { float lr_mul, beta1, beta2, eps, bc1, bc2_rsqrt, inv_scale; bool skip; };

__device__ __forceinline__ float adamw_elem(float p, float g, float& m, float& v, float lr, float wd,
                                            const AdamHyper& h) {
    g *= h.inv_scale;
    m = h.beta1 * m + (1.f - h.beta1) * g;
    v = h.beta2 * v + (1.f - h.beta2) * g * g;
    const float denom = sqrtf(v) * h.bc2_rsqrt + h.eps;
    p *= 1.f - lr * wd;
    return p - (lr / h.bc1) * (m / denom);
}

// hyper (device): [0] beta1 [1] beta2 [2] eps [3] step (float, >= 1) [4] inv_scale (gradient un-scale)
// [5] found_inf (non-zero: skip the update, GradScaler semantics; operands are still emitted)
__global__ void __launch_bounds__(256)
adamw_kernel(const mv_adamw_tensor* __restrict__ tensors, int n_tensors, int total_chunks,
             const float* __restrict__ hyper) {
    __shared__ float tile[32][33];
    AdamHyper h;
    h.beta1 = hyper[0]; h.beta2 = hyper[1]; h.eps = hyper[2];
    const float step = hyper[3];
    h.bc1 = 1.f - powf(h.beta1, step);
    h.bc2_rsqrt = rsqrtf(1.f - powf(h.beta2, step));
    h.inv_scale = hyper[4];
    h.skip = hyper[5] != 0.f;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    for (int chunk = blockIdx.x; chunk < total_chunks; chunk += gridDim.x) {
        // owner tensor: last one whose chunk0 <= chunk
        int lo = 0, hi = n_tensors - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (tensors[mid].chunk0 <= chunk) lo = mid; else hi = mid - 1;
        }
        const mv_adamw_tensor t = tensors[lo];
        const int local = chunk - t.chunk0;
        if (t.wq == nullptr) {
            const int64_t base = int64_t(local) * 1024;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int64_t i = base + j * 256 + tid;
                if (i < t.n && !h.skip) {
                    float m = t.exp_avg[i], v = t.exp_avg_sq[i];
                    t.param[i] = adamw_elem(t.param[i], t.grad[i], m, v, t.lr, t.weight_decay, h);
                    t.exp_avg[i] = m; t.exp_avg_sq[i] = v;
                }
            }
        } else {
            const int tiles_x = (t.cols + 31) >> 5;
            const int r0 = (local / tiles_x) * 32, c0 = (local % tiles_x) * 32;
            __syncthreads();                                   // previous tile fully consumed
            for (int j = threadIdx.y; j < 32; j += blockDim.y) {
                const int r = r0 + j, c = c0 + threadIdx.x;
                float p = 0.f;
                if (r < t.rows && c < t.cols) {
                    const int64_t i = int64_t(r) * t.cols + c;
                    p = t.param[i];
                    if (!h.skip) {
                        float m = t.exp_avg[i], v = t.exp_avg_sq[i];
                        p = adamw_elem(p, t.grad[i], m, v, t.lr, t.weight_decay, h);
                        t.param[i] = p; t.exp_avg[i] = m; t.exp_avg_sq[i] = v;
                    }
                    if (t.q_exp > 0) p = float_quantize_elem<false>(p, 0u, t.q_exp, t.q_man);
                }
                tile[j][threadIdx.x] = p;
            }
            __syncthreads();
            // q(W) row-major and q(W)^T through the tile, both coalesced (inline: keeps `tile` in the shared space)
            auto emit = [&](auto* wq, auto* wq_t) {
                using OutT = typename std::remove_pointer<decltype(wq)>::type;
                for (int j = threadIdx.y; j < 32; j += blockDim.y) {
                    const int r = r0 + j, c = c0 + threadIdx.x;
                    if (r < t.rows && c < t.cols) wq[int64_t(r) * t.cols + c] = opt_cvt<OutT>(tile[j][threadIdx.x]);
                }
                if (wq_t == nullptr) return;
                for (int j = threadIdx.y; j < 32; j += blockDim.y) {
                    const int c = c0 + j, r = r0 + threadIdx.x;
                    if (r < t.rows && c < t.cols) wq_t[int64_t(c) * t.rows + r] = opt_cvt<OutT>(tile[threadIdx.x][j]);
                }
            };
            if (t.wq_dtype == MV_F16) emit(static_cast<__half*>(t.wq), static_cast<__half*>(t.wq_t));
            else emit(static_cast<float*>(t.wq), static_cast<float*>(t.wq_t));
        }
    }
}

}  // namespace mv

using namespace mv;

extern "C" int mv_adamw_step(const mv_adamw_tensor* tensors_dev, int n_tensors, int total_chunks,
                             const float* hyper_dev, void* stream) {
    MV_CHECK(tensors_dev && hyper_dev && n_tensors > 0 && total_chunks > 0, "mv_adamw_step: bad arguments");
    const int grid = total_chunks < kNumSMs * 8 ? total_chunks : kNumSMs * 8;
    adamw_kernel<<<grid, dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(tensors_dev, n_tensors,
                                                                                total_chunks, hyper_dev);
    g_launches++;
    return check_cuda(cudaGetLastError(), "adamw launch");
}
