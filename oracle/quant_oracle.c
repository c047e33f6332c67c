/*
 * quant_oracle.c — CPU restatement of the QPyTorch 0.3.0 fake-quantisation
 * arithmetic that myrtle-vision calls (reference call sites:
 * src/myrtle_vision/utils/quantize.py:4-6 imports, :47-72 formats, :84 invocation).
 *
 * THIS IS TEST INFRASTRUCTURE (the oracle), not the product.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.  The product path is the CUDA library under
 * myrtle-vision_b200/csrc and never links or calls this file.
 *
 * PARITY UNPINNED: qtorch==0.3.0 (setup.py:10 of the reference) is a pip
 * dependency whose source is not under /root/reference and cannot be fetched
 * here.  The reference ships no golden vectors or tests for this boundary
 * (SURVEY.md §4, §8c).  The algorithm below restates QPyTorch's published
 * quant_cpu/quant_cuda kernels (float_kernel / fixed_point_kernel /
 * block_kernel / bit_helper) as recorded in SURVEY.md Appendix B; every point
 * where upstream releases are known to differ is a named runtime switch
 * (mvo_set_switch) so it can be flipped if the real source becomes available.
 * Self-consistency anchors that do not depend on QPyTorch source (IEEE fp16
 * round trip away from ties, idempotence, monotonicity) are checked in
 * tests/test_oracle_quant.py.
 *
 * Plain C, single-threaded scalar loops — like the original quant_cpu.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

/* ---------------------------------------------------------------- switches */
/* U1: QPyTorch 0.3.0 has a dedicated subnormal branch (add ±2^min_exp, round,
 *     subtract).  0 = older behaviour (clip_exponent flushes small values). */
static int sw_subnormal_branch = 1;
/* U2: top exponent code is not used for finite values (fp16 max = 65504).
 *     0 = older behaviour max_e = (1<<(e-1)) + 127 (fp16 max 131008). */
static int sw_reserve_top_exponent = 1;
/* U3: fixed-point nearest tie rule.  0 = CUDA kernel floor(x+0.5) (ties to
 *     +inf, what the reference runs on GPU); 1 = CPU file nearbyint(x)
 *     (ties to even). */
static int sw_fixed_nearest_even = 0;

int mvo_set_switch(const char* name, int value) {
    if (!strcmp(name, "subnormal_branch")) { sw_subnormal_branch = value; return 0; }
    if (!strcmp(name, "reserve_top_exponent")) { sw_reserve_top_exponent = value; return 0; }
    if (!strcmp(name, "fixed_nearest_even")) { sw_fixed_nearest_even = value; return 0; }
    return -1;
}

/* -------------------------------------------------------------- bit helpers */
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* bit_helper: round_bitwise_nearest — add half an output ulp, clear the tail.
 * Ties move away from zero on the magnitude; a mantissa carry bumps the exponent. */
static inline uint32_t round_bitwise_nearest(uint32_t t, int man_bits) {
    uint32_t mask = (1u << (23 - man_bits)) - 1u;
    uint32_t half = 1u << (23 - man_bits - 1);
    return (t + half) & ~mask;
}

/* bit_helper: round_bitwise_stochastic — add random tail bits, clear the tail. */
static inline uint32_t round_bitwise_stochastic(uint32_t t, uint32_t r, int man_bits) {
    uint32_t mask = (1u << (23 - man_bits)) - 1u;
    return (t + (r & mask)) & ~mask;
}

/* bit_helper: clip_exponent — saturate to ±max finite, (older releases) flush small. */
static inline uint32_t clip_exponent(int exp_bits, int man_bits, uint32_t old_num, uint32_t q) {
    if (q == 0) return 0;
    int e = (int)((q << 1) >> 24);
    int max_e = (1 << (exp_bits - 1)) + 127 - (sw_reserve_top_exponent ? 1 : 0);
    int min_e = -((1 << (exp_bits - 1)) - 2) + 127;
    uint32_t sign = old_num & 0x80000000u;
    if (e > max_e) {
        uint32_t max_man = (0x007FFFFFu >> (23 - man_bits)) << (23 - man_bits);
        return sign | ((uint32_t)max_e << 23) | max_man;
    }
    if (e < min_e) {
        /* only reachable when the subnormal branch is disabled */
        uint32_t min_num = (uint32_t)min_e << 23;
        uint32_t mid = (uint32_t)(min_e - 1) << 23;
        return ((q & 0x7FFFFFFFu) > mid) ? (sign | min_num) : 0u;
    }
    return q;
}

static inline float float_quantize_one(float a, uint32_t r, int stochastic, int man_bits, int exp_bits) {
    uint32_t target = f2u(a);
    int target_exp = (int)((target << 1) >> 24) - 127;
    int min_exp = -((1 << (exp_bits - 1)) - 2);
    if (sw_subnormal_branch && target_exp < min_exp) {
        /* shift into the lowest normal binade, round there, shift back */
        uint32_t shift_bits = ((uint32_t)(127 + min_exp) << 23) | (target & 0x80000000u);
        float shift = u2f(shift_bits);
        volatile float val = a + shift;          /* fp32 add, no excess precision */
        uint32_t vb = f2u(val);
        uint32_t qb = stochastic ? round_bitwise_stochastic(vb, r, man_bits)
                                 : round_bitwise_nearest(vb, man_bits);
        volatile float out = u2f(qb) - shift;
        return out;
    }
    uint32_t qb = stochastic ? round_bitwise_stochastic(target, r, man_bits)
                             : round_bitwise_nearest(target, man_bits);
    qb = clip_exponent(exp_bits, man_bits, target, qb);
    return u2f(qb);
}

/* quant_cpu: float_quantize_nearest(a, man_bits, exp_bits) */
void mvo_float_quantize_nearest(const float* a, float* o, int64_t n, int man_bits, int exp_bits) {
    for (int64_t i = 0; i < n; i++) o[i] = float_quantize_one(a[i], 0u, 0, man_bits, exp_bits);
}

/* quant_cuda: float_quantize_stochastic(a, man_bits, exp_bits) with the random
 * integer tensor made explicit: r[i] is the 32 random bits of element i. */
void mvo_float_quantize_stochastic(const float* a, const uint32_t* r, float* o, int64_t n,
                                   int man_bits, int exp_bits) {
    for (int64_t i = 0; i < n; i++) o[i] = float_quantize_one(a[i], r[i], 1, man_bits, exp_bits);
}

/* ------------------------------------------------------------- fixed point */
static inline void fixed_min_max(int wl, int fl, int symmetric, float* t_min, float* t_max) {
    int sigma = -fl;
    *t_min = -ldexpf(1.0f, wl - fl - 1);
    *t_max = -*t_min - ldexpf(1.0f, sigma);
    if (symmetric) *t_min = *t_min + ldexpf(1.0f, sigma);
}

/* sim_helper: round(a, r, sigma).  CUDA: floor(a*2^-sigma + r); CPU: nearbyint(.. + r - 0.5). */
static inline float fixed_round(float a, float r, int sigma, int nearest) {
    float s = ldexpf(a, -sigma);
    if (nearest && sw_fixed_nearest_even) s = nearbyintf(s);
    else { volatile float t = s + r; s = floorf(t); }
    return ldexpf(s, sigma);
}

static inline float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* quant_cpu: fixed_point_quantize_nearest(a, wl, fl, use_clamp, symmetric) */
void mvo_fixed_point_quantize_nearest(const float* a, float* o, int64_t n, int wl, int fl,
                                      int use_clamp, int symmetric) {
    float t_min, t_max; fixed_min_max(wl, fl, symmetric, &t_min, &t_max);
    for (int64_t i = 0; i < n; i++) {
        float v = fixed_round(a[i], 0.5f, -fl, 1);
        o[i] = use_clamp ? clampf(v, t_min, t_max) : v;
    }
}

/* quant_cpu: fixed_point_quantize_stochastic; r[i] in [0,1) is rand_like(a) made explicit */
void mvo_fixed_point_quantize_stochastic(const float* a, const float* r, float* o, int64_t n, int wl,
                                         int fl, int use_clamp, int symmetric) {
    float t_min, t_max; fixed_min_max(wl, fl, symmetric, &t_min, &t_max);
    for (int64_t i = 0; i < n; i++) {
        float v = fixed_round(a[i], r[i], -fl, 0);
        o[i] = use_clamp ? clampf(v, t_min, t_max) : v;
    }
}

/* quant_cpu: fixed_point_quantize_{nearest,stochastic}_mask — always clamps, mask[i]=1 where clamped */
void mvo_fixed_point_quantize_mask(const float* a, const float* r, float* o, uint8_t* m, int64_t n,
                                   int wl, int fl, int symmetric) {
    float t_min, t_max; fixed_min_max(wl, fl, symmetric, &t_min, &t_max);
    for (int64_t i = 0; i < n; i++) {
        float v = r ? fixed_round(a[i], r[i], -fl, 0) : fixed_round(a[i], 0.5f, -fl, 1);
        m[i] = (uint8_t)(v < t_min || v > t_max);
        o[i] = clampf(v, t_min, t_max);
    }
}

/* ------------------------------------------------------------ block float */
/* block_kernel: per element, given the max |a| of its block:
 *   max_exp = the power of two of the block max; base = 6*2^e; round (a+base)
 *   to `wl` mantissa bits (as QPyTorch passes man_bits = wl) and subtract base. */
static inline float block_quantize_one(float a, float max_entry, uint32_t r, int stochastic, int wl) {
    uint32_t max_exp = ((f2u(max_entry) << 1) >> 24) << 23;
    volatile float base = 6.0f * u2f(max_exp);
    volatile float t = a + base;
    uint32_t tb = f2u(t);
    uint32_t qb = stochastic ? round_bitwise_stochastic(tb, r, wl) : round_bitwise_nearest(tb, wl);
    volatile float out = u2f(qb) - base;
    return out;
}

/* quant_cpu: block_quantize_{nearest,stochastic}(a, wl, dim).
 * The tensor is viewed as [outer, dsize, inner] with `dim` the middle axis.
 * dim < 0 : one block = the whole tensor.  Otherwise one block per index of
 * `dim` (max over all other axes), as QPyTorch's transpose(0,dim).view(size,-1).max(1). */
void mvo_block_quantize(const float* a, const uint32_t* r, float* o, int64_t outer, int64_t dsize,
                        int64_t inner, int whole_tensor, int wl) {
    int64_t n = outer * dsize * inner;
    if (whole_tensor) {
        float mx = 0.0f;
        for (int64_t i = 0; i < n; i++) { float v = fabsf(a[i]); if (v > mx) mx = v; }
        for (int64_t i = 0; i < n; i++) o[i] = block_quantize_one(a[i], mx, r ? r[i] : 0u, r != 0, wl);
        return;
    }
    for (int64_t d = 0; d < dsize; d++) {
        float mx = 0.0f;
        for (int64_t u = 0; u < outer; u++)
            for (int64_t v = 0; v < inner; v++) {
                float x = fabsf(a[(u * dsize + d) * inner + v]);
                if (x > mx) mx = x;
            }
        for (int64_t u = 0; u < outer; u++)
            for (int64_t v = 0; v < inner; v++) {
                int64_t i = (u * dsize + d) * inner + v;
                o[i] = block_quantize_one(a[i], mx, r ? r[i] : 0u, r != 0, wl);
            }
    }
}

/* ------------------------------------------- counter RNG (Philox4x32-10) */
/* The product CUDA kernels define "same seed/offset" at the kernel interface
 * (SURVEY.md §7 hard parts): element i consumes word (i & 3) of
 * Philox4x32-10(key = seed, counter = {lo32(i>>2), hi32(i>>2), lo32(offset), hi32(offset)}).
 * This is the published Philox algorithm (Salmon et al., SC'11), restated. */
static inline void philox_round(uint32_t c[4], uint32_t k[2]) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

static inline void philox4x32_10(uint64_t seed, uint64_t ctr, uint64_t offset, uint32_t out[4]) {
    uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)offset, (uint32_t)(offset >> 32)};
    uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    for (int i = 0; i < 10; i++) {
        philox_round(c, k);
        if (i < 9) { k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u; }
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

/* the 32 random bits of elements [0, n) */
void mvo_philox_bits(uint32_t* r, int64_t n, uint64_t seed, uint64_t offset) {
    uint32_t w[4];
    for (int64_t i = 0; i < n; i++) {
        if ((i & 3) == 0 || i == 0) philox4x32_10(seed, (uint64_t)(i >> 2), offset, w);
        r[i] = w[i & 3];
    }
}

/* uniform [0,1) floats derived from the same bits: (bits >> 8) * 2^-24 */
/* the 16-bit stream (float_quantize, man_bits >= 7): element i takes half-word (i & 7) of philox(seed, i >> 3, offset) —
 * low half of word (i & 7) >> 1 for even i, high half for odd i */
void mvo_philox_bits16(uint32_t* r, int64_t n, uint64_t seed, uint64_t offset) {
    uint32_t w[4] = {0, 0, 0, 0};
    for (int64_t i = 0; i < n; i++) {
        if ((i & 7) == 0) philox4x32_10(seed, (uint64_t)(i >> 3), offset, w);
        const uint32_t word = w[(i & 7) >> 1];
        r[i] = (i & 1) ? (word >> 16) : (word & 0xFFFFu);
    }
}

void mvo_philox_uniform(float* r, int64_t n, uint64_t seed, uint64_t offset) {
    uint32_t w[4];
    for (int64_t i = 0; i < n; i++) {
        if ((i & 3) == 0 || i == 0) philox4x32_10(seed, (uint64_t)(i >> 2), offset, w);
        r[i] = (float)(w[i & 3] >> 8) * (1.0f / 16777216.0f);
    }
}
