// assign_core.cuh — rectangular linear sum assignment for one cost block, shared by the CUDA kernel
// (assign.cu, one warp per image, lanes striped over the columns) and by a host build with one lane
// (tests/test_assign_host.py compiles this file with g++ to check the algorithm against SciPy).
//
// Replaces scipy.optimize.linear_sum_assignment at the reference's call site
// (src/myrtle_vision/models/matcher.py:83-86): minimum-cost matching of the rows and columns of a
// [nq, nt] block, min(nq, nt) pairs.  Shortest augmenting paths with dual variables in double
// precision (the algorithm SciPy documents: Crouse 2016, "On implementing 2D rectangular assignment
// algorithms"); the smaller side is augmented row by row and the scan over the larger side is the
// parallel part.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define MV_ASSIGN_FN __device__ __forceinline__
#else
#define MV_ASSIGN_FN static inline
#endif

namespace mv_assign {

constexpr int kMaxSide = 1024;           // capacity of the per-image working arrays
constexpr double kInf = 1e300;

struct Work {                            // shared memory on the device, a plain struct on the host
    double u[kMaxSide];                  // duals of the augmented (smaller) side
    double v[kMaxSide];                  // duals of the scanned (larger) side
    double dist[kMaxSide];               // shortest path cost to every scanned-side element
    int pred[kMaxSide];                  // augmented-side element the shortest path to j comes from
    int owner[kMaxSide];                 // scanned side j -> augmented side i it is matched to, -1
    int mate[kMaxSide];                  // augmented side i -> scanned side j it is matched to, -1
    unsigned char in_tree_i[kMaxSide];
    unsigned char in_tree_j[kMaxSide];
};

#ifdef __CUDACC__
MV_ASSIGN_FN void lane_sync() { __syncwarp(); }
// (value, unmatched-first, lowest index) minimum over the warp
MV_ASSIGN_FN void lane_argmin(double& val, int& idx, int& is_free) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, val, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        const int of = __shfl_xor_sync(0xffffffffu, is_free, o);
        const bool take = ov < val || (ov == val && (of > is_free || (of == is_free && oi < idx)));
        if (take) { val = ov; idx = oi; is_free = of; }
    }
}
#else
MV_ASSIGN_FN void lane_sync() {}
MV_ASSIGN_FN void lane_argmin(double&, int&, int&) {}
#endif

// cost(i, j) = c[i * si + j * sj] with i on the augmented side (na elements) and j on the scanned side
// (ns >= na elements).  `lane` / `lanes`: this thread's index and the number of cooperating threads.
// On return w.mate[i] is the scanned-side partner of every i.  Returns 0, or 1 if no finite matching exists.
MV_ASSIGN_FN int solve(const float* c, int64_t si, int64_t sj, int na, int ns, Work& w, int lane, int lanes) {
    for (int j = lane; j < ns; j += lanes) { w.v[j] = 0.0; w.owner[j] = -1; }
    for (int i = lane; i < na; i += lanes) { w.u[i] = 0.0; w.mate[i] = -1; }
    lane_sync();
    for (int root = 0; root < na; root++) {
        for (int j = lane; j < ns; j += lanes) { w.dist[j] = kInf; w.in_tree_j[j] = 0; }
        for (int i = lane; i < na; i += lanes) w.in_tree_i[i] = 0;
        lane_sync();
        double reach = 0.0;              // length of the shortest path to the tree's frontier
        int i = root, sink = -1;
        while (sink < 0) {
            if (lane == 0) w.in_tree_i[i] = 1;
            const double ui = w.u[i];
            double best = kInf;
            int best_j = 0x7fffffff, best_free = 0;
            for (int j = lane; j < ns; j += lanes) {
                if (w.in_tree_j[j]) continue;
                const double r = reach + double(c[i * si + j * sj]) - ui - w.v[j];
                double d = w.dist[j];
                if (r < d) { d = r; w.dist[j] = r; w.pred[j] = i; }
                const int fr = w.owner[j] < 0;
                if (d < best || (d == best && fr > best_free)) { best = d; best_j = j; best_free = fr; }
            }
            lane_argmin(best, best_j, best_free);
            if (!(best < kInf)) return 1;
            reach = best;
            if (lane == 0) w.in_tree_j[best_j] = 1;
            const int o = w.owner[best_j];
            if (o < 0) sink = best_j; else i = o;
            lane_sync();
        }
        // dual update over the alternating tree
        for (int k = lane; k < na; k += lanes) {
            if (k == root) w.u[k] += reach;
            else if (w.in_tree_i[k]) w.u[k] += reach - w.dist[w.mate[k]];
        }
        for (int j = lane; j < ns; j += lanes)
            if (w.in_tree_j[j]) w.v[j] -= reach - w.dist[j];
        lane_sync();
        // augment along the predecessor chain (serial, at most `root + 1` hops)
        if (lane == 0) {
            int j = sink;
            while (true) {
                const int p = w.pred[j];
                w.owner[j] = p;
                const int next = w.mate[p];
                w.mate[p] = j;
                j = next;
                if (p == root) break;
            }
        }
        lane_sync();
    }
    return 0;
}

// One image: cost block [nq, ld] of which the first nt columns are valid.  match[t] = matched
// prediction of target t or -1; flag: set non-zero when the block admits no finite matching.
MV_ASSIGN_FN void match_block(const float* cost, int nq, int nt, int ld, int* match, int match_len, int* flag,
                              Work& w, int lane, int lanes) {
    for (int t = lane; t < match_len; t += lanes) match[t] = -1;
    lane_sync();
    if (nt <= 0 || nq <= 0) return;
    int rc;
    if (nt <= nq) {                      // targets are augmented, predictions scanned
        rc = solve(cost, 1, ld, nt, nq, w, lane, lanes);
        if (rc == 0) for (int t = lane; t < nt; t += lanes) match[t] = w.mate[t];
    } else {                             // more targets than predictions: predictions are augmented
        rc = solve(cost, ld, 1, nq, nt, w, lane, lanes);
        if (rc == 0) for (int q = lane; q < nq; q += lanes) match[w.mate[q]] = q;
    }
    if (rc != 0 && lane == 0 && flag) *flag = 1;
    lane_sync();
}

}  // namespace mv_assign
