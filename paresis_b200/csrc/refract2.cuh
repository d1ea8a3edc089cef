// Two source columns per thread: the production version of the fused refraction kernel.
//
// Same contract as refract_kernel (refraction.cu); what changes is the work decomposition.
// A thread owns the column pair (2t, 2t+1) of a 64-column strip and walks down the rows:
//   * maps are read as one 8-byte load per row, and half of the horizontal neighbours needed by
//     the gradient are already in the thread;
//   * the two rays of a thread almost always land side by side, so their deposits form three
//     cells per output row; the cell that is not part of the thread's 8-byte-aligned pair is handed
//     to the neighbouring lane with one shuffle, and the aligned pair leaves as ONE
//     REDG.E.ADD.F32x2 per row -- a warp writes 256 contiguous bytes per instruction;
//   * the lower-row pair is carried in registers and merged with the next source row's upper pair.
// Whenever a warp is not in that regular regime (image borders, torn rays) it flushes its carry
// and falls back to the per-ray code of splat.cuh, so results are the reference's for any input.
#pragma once
#include "splat.cuh"

namespace paresis {

__device__ __forceinline__ void red_add2(float* p, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}

struct PairSplatter {
    float* out;
    int ny, lane;
    int carry_key;       // key of the aligned pair waiting in registers (lower row of the last step)
    float c0, c1;
    int* flag;
    bool bad;
    Splatter<0> slow;    // per-ray fallback (stateless deposit)

    __device__ __forceinline__ void init(float* out_, int ny_, int* flag_) {
        out = out_; ny = ny_; flag = flag_;
        lane = threadIdx.x & 31;
        carry_key = NO_KEY; c0 = c1 = 0.f;
        bad = false;
        slow.init(out_, ny_, nullptr);
    }

    __device__ __forceinline__ void flush() {
        if (carry_key != NO_KEY) { red_add2(out + carry_key, c0, c1); carry_key = NO_KEY; }
    }

    // Regular regime: every lane of the warp holds two `simple` rays whose lower cells are adjacent
    // (q1.key == q0.key + 1).  ny is even, so the parity of a key is the parity of its column.
    __device__ __forceinline__ void put_pair(const FastRay& q0, const FastRay& q1) {
        bad |= !(fabsf((q0.w0 + q0.w3) + (q1.w0 + q1.w3)) <= 3.0e38f);
        const int k = q0.key;
        float t0 = q0.w0, t1 = q0.w1 + q1.w0, t2 = q1.w1;     // upper row: cells k, k+1, k+2
        float b0 = q0.w2, b1 = q0.w3 + q1.w2, b2 = q1.w3;     // lower row: cells k+ny, ...
        const bool odd = k & 1;
        const unsigned odd_mask = __ballot_sync(FULL_MASK, odd);
        bool given = false;   // my unaligned cell was taken over by a neighbour
        if (odd_mask == 0u) {
            // aligned pair = (k, k+1); cell k+2 is the next lane's first cell when it continues the run
            const int kl = __shfl_up_sync(FULL_MASK, k, 1);
            const float tl = __shfl_up_sync(FULL_MASK, t2, 1), bl = __shfl_up_sync(FULL_MASK, b2, 1);
            const bool absorb = lane > 0 && k == kl + 2;
            const unsigned m = __ballot_sync(FULL_MASK, absorb);
            given = ((m >> lane) >> 1) & 1u;
            if (absorb) { t0 += tl; b0 += bl; }
        } else if (odd_mask == FULL_MASK) {
            // aligned pair = (k+1, k+2); cell k is the previous lane's last cell
            const int kr = __shfl_down_sync(FULL_MASK, k, 1);
            const float tr = __shfl_down_sync(FULL_MASK, t0, 1), br = __shfl_down_sync(FULL_MASK, b0, 1);
            const bool absorb = lane < 31 && kr == k + 2;
            const unsigned m = __ballot_sync(FULL_MASK, absorb);
            given = lane > 0 && ((m >> (lane - 1)) & 1u);
            if (absorb) { t2 += tr; b2 += br; }
        }
        const int pk = k + (odd ? 1 : 0);
        float pt0 = odd ? t1 : t0, pt1 = odd ? t2 : t1;
        const float pb0 = odd ? b1 : b0, pb1 = odd ? b2 : b1;
        const float xt = odd ? t0 : t2, xb = odd ? b0 : b2;   // the unaligned cell of each row
        const int xk = odd ? k : k + 2;
        if (carry_key == pk) { pt0 += c0; pt1 += c1; }
        else flush();
        if (pt0 != 0.f || pt1 != 0.f) red_add2(out + pk, pt0, pt1);
        if (pb0 != 0.f || pb1 != 0.f) { carry_key = pk + ny; c0 = pb0; c1 = pb1; }
        else carry_key = NO_KEY;
        if (!given) {
            if (xt != 0.f) red_add(out + xk, xt);
            if (xb != 0.f) red_add(out + xk + ny, xb);
        }
    }

    // Anything else: per-ray deposits.  `live` lanes pass their (i, j) and raw ray data.
    __device__ __forceinline__ void put_slow(bool live, int i, int j, float v, float dx, float dy, const Frame& f) {
        const FastRay q = fast_ray(i, j, v, dx, dy, f.nx, f.ny);
        if (live && q.simple) slow.put_simple(q);
        else slow.put(live ? make_ray(i, j, v, dx, dy, f) : empty_ray());
        bad |= slow.bad;
    }

    __device__ __forceinline__ void finish() {
        flush();
        if (bad && flag) atomicOr(flag, FLAG_NONFINITE);
    }
};

constexpr int PAIR_BLOCK_COLS = 2 * BLOCK_THREADS;

template <int NM, bool DUAL, bool HAS_I, bool ATT>
__global__ void __launch_bounds__(BLOCK_THREADS)
refract2_kernel(const RefractArgs<float> a) {
    const Frame f = a.f;
    const int lane = threadIdx.x & 31;
    const int j = 2 * (blockIdx.x * BLOCK_THREADS + threadIdx.x);   // first column of the pair (even)
    const int i0 = blockIdx.y * a.rows;
    const int i1 = min(i0 + a.rows, f.nx);
    const bool live = j < f.ny;                                    // ny is even: both columns or none
    const int jc = live ? j : f.ny - 2;
    // "inner": all 64 columns of the warp exist and none is the first / last image column
    const bool inner_cols = __all_sync(FULL_MASK, live && j > 0 && j + 1 < f.ny - 1);

    float2 up[NM], mid[NM], dn[NM];
    const float2* row[NM];    // (row i+2, column pair jc)
    const float* halo[NM];    // (row i, column jc-1) for lane 0, (row i, column jc+2) for lane 31
    const bool edge_lane = lane == 0 || lane == 31;
    const int jh = min(max(lane == 0 ? jc - 1 : jc + 2, 0), f.ny - 1);
#pragma unroll
    for (int m = 0; m < NM; ++m) {
        const float* t = a.map[m] + jc;
        mid[m] = __ldg(reinterpret_cast<const float2*>(t + (size_t)i0 * f.ny));
        up[m] = i0 > 0 ? __ldg(reinterpret_cast<const float2*>(t + (size_t)(i0 - 1) * f.ny)) : mid[m];
        dn[m] = i0 + 1 < f.nx ? __ldg(reinterpret_cast<const float2*>(t + (size_t)(i0 + 1) * f.ny)) : mid[m];
        row[m] = reinterpret_cast<const float2*>(t + (size_t)(i0 + 2) * f.ny);
        halo[m] = a.map[m] + (size_t)i0 * f.ny + jh;
    }
    const float2* irow = HAS_I ? reinterpret_cast<const float2*>(a.I_in + (size_t)i0 * f.ny + jc) : nullptr;
    float2 vin = HAS_I ? __ldg(irow) : make_float2(a.I_uniform, a.I_uniform);
    const int half_pitch = f.ny >> 1;

    PairSplatter sp_obj, sp_ref;
    sp_obj.init(a.out_obj, f.ny, a.flag);
    if (DUAL) sp_ref.init(a.out_ref, f.ny, a.flag);

    for (int i = i0; i < i1; ++i) {
        float2 nxt[NM];
        const bool more = i + 2 < f.nx && i + 1 < i1;
#pragma unroll
        for (int m = 0; m < NM; ++m) {
            nxt[m] = more ? __ldg(row[m]) : dn[m];
            row[m] += half_pitch;
        }
        float2 vnext = vin;
        if (HAS_I && i + 1 < i1) { irow += half_pitch; vnext = __ldg(irow); }

        const bool inner = inner_cols && i > 0 && i < f.nx - 1;   // warp-uniform
        float dxo0 = 0.f, dyo0 = 0.f, dxo1 = 0.f, dyo1 = 0.f, dxr0 = 0.f, dyr0 = 0.f, dxr1 = 0.f, dyr1 = 0.f;
        float arg0 = 0.f, arg1 = 0.f;
#pragma unroll
        for (int m = 0; m < NM; ++m) {
            const float* t = a.map[m];
            float lf = __shfl_up_sync(FULL_MASK, mid[m].y, 1);     // column j-1
            float rt = __shfl_down_sync(FULL_MASK, mid[m].x, 1);   // column j+2
            if (edge_lane) {
                const float h = __ldg(halo[m]);
                if (lane == 0) lf = h; else rt = h;
            }
            halo[m] += f.ny;
            float gy0, gy1, gx0, gx1;
            if (inner) {
                gy0 = mid[m].y - lf;
                gy1 = rt - mid[m].x;
                gx0 = dn[m].x - up[m].x;
                gx1 = dn[m].y - up[m].y;
            } else {
                // np.gradient(edge_order=2) numerators times 2h (refractionFileNumba2.py:54)
                const float* r = t + (size_t)i * f.ny;
                gy0 = jc == 0 ? -3.f * mid[m].x + 4.f * mid[m].y - __ldg(r + 2) : mid[m].y - lf;
                gy1 = jc + 1 == f.ny - 1 ? 3.f * mid[m].y - 4.f * mid[m].x + __ldg(r + f.ny - 3) : rt - mid[m].x;
                if (i == 0) {
                    const float2 r2 = __ldg(reinterpret_cast<const float2*>(t + (size_t)2 * f.ny + jc));
                    gx0 = -3.f * mid[m].x + 4.f * dn[m].x - r2.x;
                    gx1 = -3.f * mid[m].y + 4.f * dn[m].y - r2.y;
                } else if (i == f.nx - 1) {
                    const float2 r3 = __ldg(reinterpret_cast<const float2*>(t + (size_t)(f.nx - 3) * f.ny + jc));
                    gx0 = 3.f * mid[m].x - 4.f * up[m].x + r3.x;
                    gx1 = 3.f * mid[m].y - 4.f * up[m].y + r3.y;
                } else {
                    gx0 = dn[m].x - up[m].x;
                    gx1 = dn[m].y - up[m].y;
                }
            }
            dxo0 = fmaf(a.g_obj[m], gx0, dxo0); dyo0 = fmaf(a.g_obj[m], gy0, dyo0);
            dxo1 = fmaf(a.g_obj[m], gx1, dxo1); dyo1 = fmaf(a.g_obj[m], gy1, dyo1);
            if (DUAL) {
                dxr0 = fmaf(a.g_ref[m], gx0, dxr0); dyr0 = fmaf(a.g_ref[m], gy0, dyr0);
                dxr1 = fmaf(a.g_ref[m], gx1, dxr1); dyr1 = fmaf(a.g_ref[m], gy1, dyr1);
            }
            if (ATT) { arg0 = fmaf(a.att[m], mid[m].x, arg0); arg1 = fmaf(a.att[m], mid[m].y, arg1); }
        }
        // Sample.py:347; ex2.approx keeps 2e-7 relative accuracy over the attenuation range
        float vo0 = ATT ? vin.x * __expf(-arg0) : vin.x, vo1 = ATT ? vin.y * __expf(-arg1) : vin.y;
        {
            // |D| < 1e-12 -> 0 and the |D| > N kill (refractionFileNumba2.py:59-64) change nothing for a
            // `simple` ray (its target is inside the image, so |D| < N); the slow path applies them.
            const FastRay q0 = fast_ray(i, j, vo0, dxo0, dyo0, f.nx, f.ny);
            const FastRay q1 = fast_ray(i, j + 1, vo1, dxo1, dyo1, f.nx, f.ny);
            if (inner_cols && __all_sync(FULL_MASK, q0.simple && q1.simple && q1.key == q0.key + 1)) {
                sp_obj.put_pair(q0, q1);
            } else {
                sp_obj.flush();
                clean(vo0, dxo0, dyo0, a.clamp_x, a.clamp_y);
                clean(vo1, dxo1, dyo1, a.clamp_x, a.clamp_y);
                sp_obj.put_slow(live, i, j, vo0, dxo0, dyo0, f);
                sp_obj.put_slow(live, i, j + 1, vo1, dxo1, dyo1, f);
            }
        }
        if (DUAL) {
            float vr0 = vin.x, vr1 = vin.y;
            const FastRay q0 = fast_ray(i, j, vr0, dxr0, dyr0, f.nx, f.ny);
            const FastRay q1 = fast_ray(i, j + 1, vr1, dxr1, dyr1, f.nx, f.ny);
            if (inner_cols && __all_sync(FULL_MASK, q0.simple && q1.simple && q1.key == q0.key + 1)) {
                sp_ref.put_pair(q0, q1);
            } else {
                sp_ref.flush();
                clean(vr0, dxr0, dyr0, a.clamp_x, a.clamp_y);
                clean(vr1, dxr1, dyr1, a.clamp_x, a.clamp_y);
                sp_ref.put_slow(live, i, j, vr0, dxr0, dyr0, f);
                sp_ref.put_slow(live, i, j + 1, vr1, dxr1, dyr1, f);
            }
        }
#pragma unroll
        for (int m = 0; m < NM; ++m) { up[m] = mid[m]; mid[m] = dn[m]; dn[m] = nxt[m]; }
        vin = vnext;
    }
    sp_obj.finish();
    if (DUAL) sp_ref.finish();
}

}  // namespace paresis
