// segloss.cu — bilinear upsample + per-pixel cross-entropy, forward and backward in one pass.
//
// the reference's segmentation head upsamples the [B, C, gh, gw] patch
// logits to [B, C, H, W] (nn.Upsample(size=image_size, mode='bilinear'), models/vit.py:355, 371) and
// train.py applies CrossEntropyLoss to it (segmentation/train.py:188, 261): 1.14 GB of logits at
// B=256 / 17 classes / 256^2, written once and re-read by softmax, NLL and their backwards.  Here the
// full-resolution logits never exist: each thread owns one output column of a patch-row band, forms
// the C interpolated logits of a pixel in registers, takes log-sum-exp and the loss, and folds
// (softmax - onehot) straight back onto the <= 3 coarse rows x 2 coarse columns it touches
// (register accumulators along y, one shared-memory reduction along x, one global red.add per
// coarse cell and CTA).  HBM traffic: the labels (8 B / pixel) + the coarse tensors.
#include "common.cuh"
#include "../../include/mv_b200.h"

namespace mv {
extern int64_t g_launches;

// y, dy: [B, gh*gw, C] (the decoder Linear's output layout, patch-major); labels: [B, H, W] int64.
// acc: [0] += sum of pixel losses, [1] += number of non-ignored pixels.  dy += d(sum of losses)/dy.
template <int CMAX>
__global__ void __launch_bounds__(256)
upsample_ce_kernel(const float* __restrict__ y, const int64_t* __restrict__ labels, float* __restrict__ dy,
                   float* __restrict__ acc, int C, int gh, int gw, int H, int W, int64_t ignore_index) {
    extern __shared__ float sm[];
    float* s_y = sm;                         // [3][gw][C] coarse logits of rows r0-1, r0, r0+1 (clamped)
    float* s_g = sm + 3 * gw * C;            // [3][gw][C] gradient accumulators
    __shared__ float s_red[2][8];
    const int bands = gh;                    // one CTA per (image, coarse row): output rows [band*sy, (band+1)*sy)
    const int b = blockIdx.x / bands, band = blockIdx.x % bands;
    const int rows_per_band = H / gh;
    const float sy_inv = float(gh) / float(H), sx_inv = float(gw) / float(W);
    const int cell = gw * C;
    for (int i = threadIdx.x; i < 3 * cell; i += blockDim.x) {
        const int r = min(max(band - 1 + i / cell, 0), gh - 1);
        s_y[i] = y[(int64_t(b) * gh + r) * cell + i % cell];
        s_g[i] = 0.f;
    }
    __syncthreads();
    float loss = 0.f, count = 0.f;
    for (int ox = threadIdx.x; ox < W; ox += blockDim.x) {
        const float fx = fmaxf(sx_inv * (float(ox) + 0.5f) - 0.5f, 0.f);
        const int ix0 = int(fx), ix1 = ix0 + (ix0 < gw - 1 ? 1 : 0);
        const float lx1 = fx - float(ix0), lx0 = 1.f - lx1;
        float g[3][CMAX];                    // d(sum loss) / d(x-interpolated coarse row r), per class
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < CMAX; c++) g[r][c] = 0.f;
        for (int j = 0; j < rows_per_band; j++) {
            const int oy = band * rows_per_band + j;
            const int64_t lab = labels[(int64_t(b) * H + oy) * W + ox];
            if (lab == ignore_index) continue;
            const float fy = fmaxf(sy_inv * (float(oy) + 0.5f) - 0.5f, 0.f);
            const int iy0 = int(fy), iy1 = iy0 + (iy0 < gh - 1 ? 1 : 0);
            const float ly1 = fy - float(iy0), ly0 = 1.f - ly1;
            const int r0 = iy0 - (band - 1), r1 = iy1 - (band - 1);       // rows of s_y, in 0..2
            const float* a0 = s_y + (r0 * gw + ix0) * C; const float* a1 = s_y + (r0 * gw + ix1) * C;
            const float* b0 = s_y + (r1 * gw + ix0) * C; const float* b1 = s_y + (r1 * gw + ix1) * C;
            float v[CMAX], mx = -INFINITY;
#pragma unroll
            for (int c = 0; c < CMAX; c++) {
                if (c < C) {
                    v[c] = ly0 * (lx0 * a0[c] + lx1 * a1[c]) + ly1 * (lx0 * b0[c] + lx1 * b1[c]);
                    mx = fmaxf(mx, v[c]);
                }
            }
            float se = 0.f, vl = 0.f;
#pragma unroll
            for (int c = 0; c < CMAX; c++) {
                if (c < C) {
                    if (c == int(lab)) vl = v[c];
                    v[c] = __expf(v[c] - mx);
                    se += v[c];
                }
            }
            loss += __logf(se) + mx - vl;
            count += 1.f;
            const float inv = 1.f / se;
#pragma unroll
            for (int c = 0; c < CMAX; c++) {
                if (c < C) {
                    const float d = v[c] * inv - (c == int(lab) ? 1.f : 0.f);
#pragma unroll
                    for (int r = 0; r < 3; r++) {                         // r0, r1 are thread-varying: select, no indexing
                        const float wy = (r == r0 ? ly0 : 0.f) + (r == r1 ? ly1 : 0.f);
                        g[r][c] = fmaf(wy, d, g[r][c]);
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < 3; r++) {
#pragma unroll
            for (int c = 0; c < CMAX; c++) {
                if (c < C && g[r][c] != 0.f) {
                    atomicAdd(s_g + (r * gw + ix0) * C + c, lx0 * g[r][c]);
                    if (lx1 != 0.f) atomicAdd(s_g + (r * gw + ix1) * C + c, lx1 * g[r][c]);
                }
            }
        }
    }
    // CTA totals of loss / count
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        loss += __shfl_xor_sync(0xffffffffu, loss, o);
        count += __shfl_xor_sync(0xffffffffu, count, o);
    }
    if (lane == 0) { s_red[0][warp] = loss; s_red[1][warp] = count; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float l = 0.f, n = 0.f;
        for (int w = 0; w < int(blockDim.x >> 5); w++) { l += s_red[0][w]; n += s_red[1][w]; }
        atomicAdd(acc, l); atomicAdd(acc + 1, n);
    }
    for (int i = threadIdx.x; i < 3 * cell; i += blockDim.x) {
        const int r = band - 1 + i / cell;
        if (r < 0 || r >= gh) continue;      // clamped copies carry no gradient of their own: weights never select them
        const float v = s_g[i];
        if (v != 0.f) atomicAdd(dy + (int64_t(b) * gh + r) * cell + i % cell, v);
    }
}

}  // namespace mv

using namespace mv;

extern "C" int mv_upsample_ce(const float* y, const int64_t* labels, float* dy, float* acc, int B, int C, int gh,
                              int gw, int H, int W, int64_t ignore_index, void* stream) {
    MV_CHECK(y && labels && dy && acc && B > 0 && C > 0 && gh > 0 && gw > 0, "mv_upsample_ce: bad arguments");
    MV_CHECK(H % gh == 0 && W % gw == 0, "mv_upsample_ce: output size must be a multiple of the patch grid");
    MV_CHECK(C <= 32, "mv_upsample_ce: at most 32 classes");
    const size_t smem = size_t(6) * gw * C * sizeof(float);
    MV_CHECK(smem <= 48 * 1024, "mv_upsample_ce: patch row too wide");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = B * gh;
    if (C <= 8) upsample_ce_kernel<8><<<grid, 256, smem, st>>>(y, labels, dy, acc, C, gh, gw, H, W, ignore_index);
    else if (C <= 16) upsample_ce_kernel<16><<<grid, 256, smem, st>>>(y, labels, dy, acc, C, gh, gw, H, W, ignore_index);
    else if (C <= 24) upsample_ce_kernel<24><<<grid, 256, smem, st>>>(y, labels, dy, acc, C, gh, gw, H, W, ignore_index);
    else upsample_ce_kernel<32><<<grid, 256, smem, st>>>(y, labels, dy, acc, C, gh, gw, H, W, ignore_index);
    g_launches++;
    return check_cuda(cudaGetLastError(), "upsample_ce launch");
}
