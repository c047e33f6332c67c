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
#include <type_traits>
#include "common.cuh"
#include "../../include/mv_b200.h"

namespace mv {
extern int64_t g_launches;

// y, dy: [B, gh*gw, C] (the decoder Linear's output layout, patch-major); labels: [B, H, W] int64.
// acc: [0] += sum of pixel losses, [1] += number of non-ignored pixels.  dy += d(sum of losses)/dy.
// kExact: C == CMAX, so no per-class guard survives in the unrolled class loops (they were a third of the instructions)
template <int CMAX, bool kExact = false>
__global__ void __launch_bounds__(256, CMAX <= 24 ? 2 : 1)
upsample_ce_kernel(const float* __restrict__ y, const int64_t* __restrict__ labels, float* __restrict__ dy,
                   float* __restrict__ acc, int C_rt, int gh, int gw, int H, int W, int64_t ignore_index) {
    const int C = kExact ? CMAX : C_rt;
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
    // (whole warps stay in the loop — the x fold below is a warp collective; lanes past the last column carry zeros)
    for (int ox0 = threadIdx.x & ~31; ox0 < W; ox0 += blockDim.x) {
        const int oxr = ox0 + (threadIdx.x & 31);
        const bool live = oxr < W;
        const int ox = live ? oxr : W - 1;
        const float fx = fmaxf(sx_inv * (float(ox) + 0.5f) - 0.5f, 0.f);
        const int ix0 = int(fx), ix1 = ix0 + (ix0 < gw - 1 ? 1 : 0);
        const float lx1 = fx - float(ix0), lx0 = 1.f - lx1;
        // Coarse rows band - 1, band, band + 1 (s_y rows 0, 1, 2) are used two at a time: output rows whose upper source
        // row is band - 1 first (rows 0 / 1), then those whose upper source row is band (rows 1 / 2; at the bottom edge the
        // lower neighbour is the clamped copy in row 2, whose gradient the write-back folds into row gh - 1).  Per phase the
        // two rows interpolated along x at this column (xa, xb) and their gradients (ga, gb) live in registers.
        float xa[CMAX], xb[CMAX], ga[CMAX], gb[CMAX];
        auto load_row = [&](float (&dst)[CMAX], int r) {
#pragma unroll
            for (int c = 0; c < CMAX; c++)
                dst[c] = (kExact || c < C) ? lx0 * s_y[(r * gw + ix0) * C + c] + lx1 * s_y[(r * gw + ix1) * C + c] : 0.f;
        };
        // Along x the 32 columns of a warp fold onto the few coarse cells they touch (ix0 of lane 0 .. ix1 of lane 31: four
        // at 16 pixels per cell).  Per cell: every lane's weight of that cell, a butterfly sum over the warp, and ONE shared
        // atomic per (cell, class), issued by lane c so that the compare-and-swap loops of a row run side by side.  (Shared
        // fp32 atomicAdd is a CAS loop: with one atomic per lane, 16 lanes of a warp spun on the same word for each of
        // 102 values — 3.85 ms of the 20.3 ms segmentation step.)
        const int cell_lo = __shfl_sync(0xffffffffu, ix0, 0), cell_hi = __shfl_sync(0xffffffffu, ix1, 31);
        // The CMAX per-class sums of a cell are reduced TOGETHER: at every butterfly step a lane keeps half of its values
        // (the even or the odd ones, by its lane bit) and receives the partner's copy of the same — 10 + 5 + 3 + 2 + 1
        // shuffles for 20 classes instead of 100 — and ends with the warp total of class bitrev5(lane).
        const int my_class = int(__brev(threadIdx.x & 31u) >> 27);
        auto fold_row = [&](const float (&gr)[CMAX], int r) {
            for (int cx = cell_lo; cx <= cell_hi; cx++) {
                const float wx = (ix0 == cx ? lx0 : 0.f) + (ix1 == cx ? lx1 : 0.f);
                float t[32];
#pragma unroll
                for (int c = 0; c < 32; c++) t[c] = (c < CMAX && (kExact || c < C)) ? wx * gr[c] : 0.f;
#pragma unroll
                for (int o = 16, n = 32; o > 0; o >>= 1, n >>= 1) {
                    const bool up = (threadIdx.x & o) != 0;
#pragma unroll
                    for (int i = 0; i < n / 2; i++) {
                        if (2 * i < CMAX || o < 16) {          // (values past CMAX are zero on every lane: nothing to exchange)
                            const float keep = up ? t[2 * i + 1] : t[2 * i], send = up ? t[2 * i] : t[2 * i + 1];
                            t[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                        } else {
                            t[i] = 0.f;
                        }
                    }
                }
                if (my_class < C && t[0] != 0.f) atomicAdd(s_g + (r * gw + cx) * C + my_class, t[0]);
            }
        };
        auto pixel = [&](float ly0, float ly1, int lab) {
            float v[CMAX], mx = -INFINITY;
#pragma unroll
            for (int c = 0; c < CMAX; c++) {
                if (kExact || c < C) {
                    v[c] = fmaf(ly1, xb[c], ly0 * xa[c]);
                    mx = fmaxf(mx, v[c]);
                }
            }
            float se = 0.f, vl = 0.f;
#pragma unroll
            for (int c = 0; c < CMAX; c++) {
                if (kExact || c < C) {
                    if (c == lab) vl = v[c];
                    v[c] = ex2_fast((v[c] - mx) * 1.4426950408889634f);
                    se += v[c];
                }
            }
            loss += __logf(se) + mx - vl;
            count += 1.f;
            const float inv = 1.f / se;
#pragma unroll
            for (int c = 0; c < CMAX; c++) {
                if (kExact || c < C) {
                    const float d = fmaf(v[c], inv, c == lab ? -1.f : 0.f);
                    ga[c] = fmaf(ly0, d, ga[c]);
                    gb[c] = fmaf(ly1, d, gb[c]);
                }
            }
        };
        load_row(xa, 0); load_row(xb, 1);
#pragma unroll
        for (int c = 0; c < CMAX; c++) { ga[c] = 0.f; gb[c] = 0.f; }
        int j = 0;
        // labels one row ahead (a row's label load is otherwise exposed in front of every pixel)
        const int64_t* lp = labels + (int64_t(b) * H + band * rows_per_band) * W + ox;
        int64_t lab_next = live ? lp[0] : ignore_index;
#pragma unroll 1
        for (int phase = 0; phase < 2; phase++) {
            // source rows of an output row are the same for every column: iy0 is band - 1 (phase 0) or band (phase 1)
            for (; j < rows_per_band; j++) {
                const int oy = band * rows_per_band + j;
                const float fy = fmaxf(sy_inv * (float(oy) + 0.5f) - 0.5f, 0.f);
                const int iy0 = int(fy);
                if (phase == 0 && iy0 >= band) break;
                const float ly1 = fy - float(iy0), ly0 = 1.f - ly1;
                const int64_t lab = lab_next;
                if (live && j + 1 < rows_per_band) lab_next = lp[int64_t(j + 1) * W];
                if (lab != ignore_index) pixel(ly0, ly1, int(lab));
            }
            fold_row(ga, phase);
            if (phase == 0) {
#pragma unroll
                for (int c = 0; c < CMAX; c++) { xa[c] = xb[c]; ga[c] = gb[c]; gb[c] = 0.f; }
                load_row(xb, 2);
            }
        }
        fold_row(gb, 2);
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
        const int r = min(max(band - 1 + i / cell, 0), gh - 1);      // a clamped copy's gradient belongs to the row it copies
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
    if (C == 17) upsample_ce_kernel<17, true><<<grid, 256, smem, st>>>(y, labels, dy, acc, C, gh, gw, H, W, ignore_index);   // config 4
    else if (C <= 8) upsample_ce_kernel<8><<<grid, 256, smem, st>>>(y, labels, dy, acc, C, gh, gw, H, W, ignore_index);
    else if (C <= 16) upsample_ce_kernel<16><<<grid, 256, smem, st>>>(y, labels, dy, acc, C, gh, gw, H, W, ignore_index);
    else if (C <= 20) upsample_ce_kernel<20><<<grid, 256, smem, st>>>(y, labels, dy, acc, C, gh, gw, H, W, ignore_index);
    else if (C <= 24) upsample_ce_kernel<24><<<grid, 256, smem, st>>>(y, labels, dy, acc, C, gh, gw, H, W, ignore_index);
    else upsample_ce_kernel<32><<<grid, 256, smem, st>>>(y, labels, dy, acc, C, gh, gw, H, W, ignore_index);
    g_launches++;
    return check_cuda(cudaGetLastError(), "upsample_ce launch");
}
