// Counter-based Poisson sampler shared by the detector kernels.
//
// Reference: Detector.py:113-115 draws numpy.random.RandomState(wall-clock seed).poisson(image); the
// stream is irreproducible by design, so parity is statistical.  Here every draw is a pure
// function of (seed, sequence, pixel): Philox4x32-10 + inversion (lam < 10) / Hoermann's PTRS.
#pragma once
#include "common.cuh"

namespace paresis {

// ---------------------------------------------------------------------------------------------
// Poisson noise: Philox4x32-10 counter-based generator + inversion (lam < 10) / PTRS (lam >= 10)
// ---------------------------------------------------------------------------------------------
struct Philox {
    uint32_t c[4], k[2];
    __device__ __forceinline__ void round() {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k[0], n2 = hi0 ^ c[3] ^ k[1];
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    }
    __device__ __forceinline__ void generate(uint32_t out[4]) {
        Philox s = *this;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            s.round();
            s.k[0] += 0x9E3779B9u;
            s.k[1] += 0xBB67AE85u;
        }
        out[0] = s.c[0]; out[1] = s.c[1]; out[2] = s.c[2]; out[3] = s.c[3];
    }
};

__device__ __forceinline__ float u01f(uint32_t x) {
    return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f);   // 24-bit uniform in (0, 1)
}
__device__ __forceinline__ double u01d(uint32_t hi, uint32_t lo) {
    const uint64_t x = ((uint64_t)hi << 21) ^ (uint64_t)(lo >> 11);   // 53-bit uniform in (0, 1)
    return ((double)(x & ((1ull << 53) - 1)) + 0.5) * (1.0 / 9007199254740992.0);
}

// log(k!) for k < 16
static __constant__ float LOG_FACT[16] = {0.f, 0.f, 0.69314718f, 1.79175947f, 3.17805383f, 4.78749174f, 6.57925121f, 8.52516136f,
                                   10.60460290f, 12.80182748f, 15.10441257f, 17.50230785f, 19.98721450f, 22.55216385f,
                                   25.19122118f, 27.89927138f};

// log of the Poisson pmf at k for mean lam >= 10, in fp32 without cancellation:
//   k >= 16: Stirling,  log p = lam * g(x) - log(2 pi k)/2 - 1/(12k) + 1/(360k^3),  x = (k - lam)/lam,
//            g(x) = x - (1+x) log(1+x) = sum_{n>=2} (-1)^(n+1) x^n / (n (n-1))  (the series up to x^12 for |x| < 1/4,
//            i.e. for every candidate of a mean above ~600: the two O(lam) terms never meet, no branch, no log1p);
//            the O(log k) terms use MUFU.LG2 / MUFU.RCP (absolute error ~1e-6 on a log-likelihood)
//   k <  16: -lam + k log(lam) - log(k!) from the table (all terms are small there).
__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float log_poisson_pmf(float k, float lam) {
    if (k < 16.f) return -lam + k * logf(lam) - LOG_FACT[(int)k];
    const float x = (k - lam) * rcp_approx(lam);
    float g;
    if (fabsf(x) < 0.25f) {
        float h = 1.f / 132.f;                       // n = 12
        h = fmaf(h, x, -1.f / 110.f);
        h = fmaf(h, -x, -1.f / 90.f);
        h = fmaf(h, -x, -1.f / 72.f);
        h = fmaf(h, -x, -1.f / 56.f);
        h = fmaf(h, -x, -1.f / 42.f);
        h = fmaf(h, -x, -1.f / 30.f);
        h = fmaf(h, -x, -1.f / 20.f);
        h = fmaf(h, -x, -1.f / 12.f);
        h = fmaf(h, -x, -1.f / 6.f);
        h = fmaf(h, -x, -0.5f);
        g = h * x * x;
    } else {
        g = x - (1.f + x) * log1pf(x);
    }
    const float ik = rcp_approx(k);
    return lam * g - 0.5f * __logf(6.28318530718f * k) - ik * (1.f / 12.f - ik * ik * (1.f / 360.f));
}

// One Poisson variate with mean lam, a pure function of (seed, sequence, pixel).
//   lam < 10 : inversion by sequential search on one uniform;
//   lam >= 10: PTRS -- W. Hoermann, "The transformed rejection method for generating Poisson random
//              variables", Insur. Math. Econ. 12 (1993).  ~86 % of the draws end at the quick
//              acceptance test; the full test uses the cancellation-free fp32 log-pmf above, so a
//              warp never waits on fp64 transcendentals.
// Counter layout: one Philox block serves the pixel PAIR (p >> 1); pixel p uses words 2*(p&1) and
// 2*(p&1)+1 of it.  c[1] carries the PTRS trial number.  A draw is therefore a pure function of
// (seed, sequence, pixel), whatever kernel, tile shape or GPU count produced it.
__device__ __forceinline__ Philox poisson_stream(uint64_t seed, uint64_t seq, uint64_t pixel) {
    const uint64_t pair = pixel >> 1;
    Philox g;
    g.c[0] = (uint32_t)pair;
    g.c[1] = 0u;
    g.c[2] = (uint32_t)seq;
    g.c[3] = (uint32_t)(seq >> 32) ^ (uint32_t)(pair >> 32);
    g.k[0] = (uint32_t)seed;
    g.k[1] = (uint32_t)(seed >> 32);
    return g;
}

__device__ __forceinline__ float sqrt_fast(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rcp_fast(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// PTRS constants of one mean.  sqrt / reciprocal are the 1-2 ulp MUFU approximations: the same values
// feed the hat function and both acceptance tests, so the method stays self-consistent.
struct PtrsSetup {
    float b, a, vr;
    __device__ __forceinline__ explicit PtrsSetup(float lam) {
        b = fmaf(2.53f, sqrt_fast(lam), 0.931f);
        a = fmaf(0.02483f, b, -0.059f);
        vr = fmaf(-3.6224f, rcp_fast(b - 2.0f), 0.9277f);
    }
    // k = floor((2a/us + b) U + lam + 0.43).  lam is split into floor(lam) + fraction so that the
    // sum keeps unit resolution in fp32 for any mean below 2^24 (no fp64 on the hot path).
    __device__ __forceinline__ float candidate(float lam, float U, float us) const {
        const float t = fmaf(fmaf(2.0f * a, rcp_fast(us), b), U, 0.43f);
        const float li = floorf(lam);
        return li + floorf((lam - li) + t);
    }
};

// Decide one draw of a LARGE mean (lam >= 10) from its two random words when the PTRS candidate
// passes the quick acceptance test (~86 % of such draws).  Everything else -- small or non-positive
// means, rejected candidates -- returns false and is finished by poisson_slow().
__device__ __forceinline__ bool poisson_decide(float lam, uint32_t ra, uint32_t rb, float& result) {
    const PtrsSetup t(lam);
    const float U = u01f(ra) - 0.5f, V = u01f(rb);
    const float us = 0.5f - fabsf(U);
    result = t.candidate(lam, U, us);
    return lam >= 10.f && us >= 0.07f && V <= t.vr;
}

__device__ __forceinline__ bool poisson_quick(float lam, uint64_t seed, uint64_t seq, uint64_t pixel, float& result) {
    Philox g = poisson_stream(seed, seq, pixel);
    uint32_t r[4];
    g.generate(r);
    const bool hi = pixel & 1;
    return poisson_decide(lam, hi ? r[2] : r[0], hi ? r[3] : r[1], result);
}

// Two neighbouring pixels (p0 even, p0 + 1) from ONE Philox block.
__device__ __forceinline__ void poisson_quick2(float lam0, float lam1, uint64_t seed, uint64_t seq, uint64_t p0,
                                               float& x0, float& x1, bool& ok0, bool& ok1) {
    Philox g = poisson_stream(seed, seq, p0);
    uint32_t r[4];
    g.generate(r);
    ok0 = poisson_decide(lam0, r[0], r[1], x0);
    ok1 = poisson_decide(lam1, r[2], r[3], x1);
}

// The full PTRS test of one candidate (U, V from two random words): true = accepted, k in `result`.
__device__ __forceinline__ bool ptrs_trial(const PtrsSetup& t, float lam, float log_invalpha, uint32_t ra, uint32_t rb, float& result) {
    const float U = u01f(ra) - 0.5f, V = u01f(rb);
    const float us = 0.5f - fabsf(U);
    const float k = t.candidate(lam, U, us);
    result = k;
    if (us >= 0.07f && V <= t.vr) return true;
    if (k < 0.f || (us < 0.013f && V > us)) return false;
    // lg2-based logs (abs. error ~1e-6) on the hat side; the pmf side is the accurate one
    return __logf(V) + log_invalpha - __logf(fmaf(t.a, rcp_fast(us * us), t.b)) <= log_poisson_pmf(k, lam);
}

// Everything the quick test leaves open, in STEPS that a kernel can run on dense warps:
//   step 0 (poisson_first) : the two words of the pair block that the quick test saw -- lam <= 0 -> 0, lam < 10 ->
//             inversion by sequential search, else the full PTRS test of candidate 0;
//   step n >= 1 (poisson_block): Philox block (pixel, n | 2^31) = TWO more candidates (words 0,1 then 2,3).
// Each returns true and the variate once the draw is decided.  A draw stays a pure function of (seed, sequence,
// pixel): poisson_round() / poisson_draw() below walk the same steps.
__device__ __forceinline__ bool poisson_first(float lam, uint32_t ra, uint32_t rb, float& result) {
    if (!(lam > 0.f)) { result = 0.f; return true; }
    if (lam < 10.f) {
        const float u = (float)u01d(ra, rb);
        float p = expf(-lam), F = p;
        int x = 0;
        while (u > F && x < 200) {
            ++x;
            p *= lam / (float)x;
            F += p;
        }
        result = (float)x;
        return true;
    }
    const PtrsSetup t(lam);
    const float log_invalpha = __logf(1.1239f + 1.1328f * rcp_fast(t.b - 3.4f));
    return ptrs_trial(t, lam, log_invalpha, ra, rb, result);
}

__device__ __forceinline__ bool poisson_block(float lam, uint64_t seed, uint64_t seq, uint64_t pixel, uint32_t block, float& result) {
    const PtrsSetup t(lam);
    const float log_invalpha = __logf(1.1239f + 1.1328f * rcp_fast(t.b - 3.4f));
    Philox g;
    g.c[0] = (uint32_t)pixel;
    g.c[1] = block | 0x80000000u;
    g.c[2] = (uint32_t)seq;
    g.c[3] = (uint32_t)(seq >> 32) ^ (uint32_t)(pixel >> 32);
    g.k[0] = (uint32_t)seed;
    g.k[1] = (uint32_t)(seed >> 32);
    uint32_t r[4];
    g.generate(r);
    if (ptrs_trial(t, lam, log_invalpha, r[0], r[1], result)) return true;
    return ptrs_trial(t, lam, log_invalpha, r[2], r[3], result);
}

// The two random words of `pixel` in its pair block (what poisson_quick / poisson_quick2 consumed).
__device__ __forceinline__ void poisson_words(uint64_t seed, uint64_t seq, uint64_t pixel, uint32_t& ra, uint32_t& rb) {
    Philox g = poisson_stream(seed, seq, pixel);
    uint32_t r[4];
    g.generate(r);
    const bool hi = pixel & 1;
    ra = hi ? r[2] : r[0];
    rb = hi ? r[3] : r[1];
}

// Round 0 = step 0 and, if that rejects (~55 % of what reaches it), block 1 right away; round r >= 1 = block r + 1.
static __device__ __noinline__ bool poisson_round(float lam, uint64_t seed, uint64_t seq, uint64_t pixel, uint32_t round,
                                                  float& result) {
    if (round == 0) {
        uint32_t ra, rb;
        poisson_words(seed, seq, pixel, ra, rb);
        if (poisson_first(lam, ra, rb, result)) return true;
    }
    return poisson_block(lam, seed, seq, pixel, round + 1, result);
}

// Blocks first, first + 1, ... until the draw is decided (after step 0 rejected).
static __device__ __noinline__ float poisson_rest(float lam, uint64_t seed, uint64_t seq, uint64_t pixel, uint32_t first) {
    float x;
    for (uint32_t blk = first; blk < first + 32; ++blk)
        if (poisson_block(lam, seed, seq, pixel, blk, x)) return x;
    return floorf(lam + 0.5f);  // unreachable in practice (two candidates per block, acceptance > 0.88 each)
}

// The complete draw.
static __device__ float poisson_slow(float lam, uint64_t seed, uint64_t seq, uint64_t pixel) {
    float x;
    for (uint32_t round = 0; round < 32; ++round)
        if (poisson_round(lam, seed, seq, pixel, round, x)) return x;
    return floorf(lam + 0.5f);  // unreachable in practice (two candidates per round, acceptance > 0.88 each)
}

__device__ __forceinline__ float poisson_draw(float lam, uint64_t seed, uint64_t seq, uint64_t pixel) {
    float x;
    if (poisson_quick(lam, seed, seq, pixel, x)) return x;
    return poisson_slow(lam, seed, seq, pixel);
}

}  // namespace paresis
