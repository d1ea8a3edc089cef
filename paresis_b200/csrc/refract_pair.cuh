// Fused refraction hop, TWO source columns per thread (the production kernel of the per-energy loop;
// same contract as refract_kernel in refraction.cu).
//
// What bounds this hop on a membrane field is (a) instruction issue -- ~160 warp-instructions per 32
// pixels in the one-column kernel -- and (b) the rate at which L2 retires REDs, which is per 32-byte
// SECTOR touched, not per lane (~170 sector-ops/ns on B200: tools/redbench.cu; the displacement field
// is torn every few pixels, so a warp's REDs land in ~11 sectors per instruction).  A thread that owns
// the column pair (2t, 2t+1) attacks both:
//   * maps are read as 8-byte loads, half of the horizontal neighbours are already in the thread, the
//     loop / addressing / zero-fill overhead is shared by two pixels;
//   * the two rays of a thread almost always land side by side: their 2 x 3 cells leave as one
//     REDG.ADD.F32x2 (the 8-byte aligned pair) and one scalar RED per row -- 4 RED instructions per TWO
//     rays instead of 8, i.e. half the sector-ops;
//   * where object and reference beam coincide (outside the sample) the rays are formed once.
// A pair that tears apart, or touches the image border (reference edge rules, splat.cuh), falls back
// per thread to per-ray deposits, so the result is the reference's for any input.  Needs an even pitch.
#pragma once
#include "refract_common.cuh"

namespace paresis {

constexpr int PAIR_THREADS = 256;
constexpr int PAIR_COLS = 2 * PAIR_THREADS;

__device__ __forceinline__ void red_add2(float* p, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}

struct PairRay {
    int r, c;                 // lower cell
    float w0, w1, w2, w3;     // (r,c) (r,c+1) (r+1,c) (r+1,c+1)
    bool simple;              // all four cells strictly inside the image
};

__device__ __forceinline__ PairRay pair_ray(int i, int j, float v, float dx, float dy, int nx, int ny) {
    PairRay q;
    const float flx = floorf(dx), fly = floorf(dy);
    const float fx = dx - flx, fy = dy - fly;
    q.r = i + __float2int_rd(dx);    // saturating; wrap-around fails `simple`
    q.c = j + __float2int_rd(dy);
    q.simple = ((unsigned)q.r < (unsigned)(nx - 1)) & ((unsigned)q.c < (unsigned)(ny - 1));
    const float v1 = v * fx, v0 = v - v1;
    q.w1 = v0 * fy; q.w0 = v0 - q.w1;
    q.w3 = v1 * fy; q.w2 = v1 - q.w3;
    return q;
}

// Deposit the two rays of a thread into `out`.  Returns what was deposited inside the image.
// (v, dx, dy) are the raw values: the clean-up of refractionFileNumba2.py:59-64 only matters for rays
// that are not `simple` (see refract_kernel) and is applied on that path.
__device__ __forceinline__ float deposit_pair(float* __restrict__ out, const PairRay& q0, const PairRay& q1, int i, int j,
                                              float v0, float dx0, float dy0, float v1, float dx1, float dy1,
                                              const Frame& f, float cx, float cy, bool live, bool& bad) {
    if (!live) return 0.f;
    bad |= !(fabsf((q0.w0 + q0.w3) + (q1.w0 + q1.w3)) <= 3.0e38f);
    if (q0.simple & q1.simple) {
        float* p = out + (size_t)q0.r * f.ny + q0.c;
        if (q1.r == q0.r && q1.c == q0.c + 1) {
            // side by side: cells c .. c+2 of rows r and r+1; the 8-byte aligned pair goes as one RED
            const float t1 = q0.w1 + q1.w0, b1 = q0.w3 + q1.w2;
            float* pb = p + f.ny;
            if (q0.c & 1) {
                red_add(p, q0.w0); red_add2(p + 1, t1, q1.w1);
                red_add(pb, q0.w2); red_add2(pb + 1, b1, q1.w3);
            } else {
                red_add2(p, q0.w0, t1); red_add(p + 2, q1.w1);
                red_add2(pb, q0.w2, b1); red_add(pb + 2, q1.w3);
            }
        } else {
            float* s = out + (size_t)q1.r * f.ny + q1.c;
            red_add(p, q0.w0); red_add(p + 1, q0.w1); red_add(p + f.ny, q0.w2); red_add(p + f.ny + 1, q0.w3);
            red_add(s, q1.w0); red_add(s + 1, q1.w1); red_add(s + f.ny, q1.w2); red_add(s + f.ny + 1, q1.w3);
        }
        return v0 + v1;
    }
    // image border / far-flung rays: the reference's frame rules (splat.cuh), per cell
    float sum = 0.f;
    {
        clean(v0, dx0, dy0, cx, cy);
        clean(v1, dx1, dy1, cx, cy);
        const Ray a = make_ray(i, j, v0, dx0, dy0, f), b = make_ray(i, j + 1, v1, dx1, dy1, f);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (((a.ok >> k) & 1u) && a.w[k] != 0.f) { red_add(out + (size_t)(a.r + (k >> 1)) * f.ny + (a.c + (k & 1)), a.w[k]); sum += a.w[k]; }
            if (((b.ok >> k) & 1u) && b.w[k] != 0.f) { red_add(out + (size_t)(b.r + (k >> 1)) * f.ny + (b.c + (k & 1)), b.w[k]); sum += b.w[k]; }
        }
    }
    return sum;
}

template <int NM, bool DUAL, bool HAS_I, bool ATT>
__global__ void __launch_bounds__(PAIR_THREADS)
refract_pair_kernel(const RefractArgs<float> a) {
    const Frame f = a.f;
    const int lane = threadIdx.x & 31;
    const int j = 2 * (blockIdx.x * PAIR_THREADS + threadIdx.x);   // first column of the pair (even)
    const int i0 = blockIdx.y * a.rows;
    const int i1 = min(i0 + a.rows, f.nx);
    const bool live = j < f.ny;                                    // ny is even: both columns or none
    const int jc = live ? j : f.ny - 2;
    const int hp = f.ny >> 1;                                      // pitch in float2

    // rolling rows: up = i-1, mid = i, dn = i+1, n1 = i+2 (in flight); row i+3 is fetched at the top of a step
    float2 up[NM], mid[NM], dn[NM], n1[NM];
    int off = i0 * hp + (jc >> 1);                                 // float2 offset of (i, jc); nx*ny < 2^30
    const bool edge_lane = lane == 0 || lane == 31;
    const int jh = min(max(lane == 0 ? jc - 1 : jc + 2, 0), f.ny - 1);
    int offh = i0 * f.ny + jh;                                     // float offset of the column next to the strip
#pragma unroll
    for (int m = 0; m < NM; ++m) {
        const float2* t = reinterpret_cast<const float2*>(a.map[m]);
        mid[m] = __ldg(t + off);
        up[m] = i0 > 0 ? __ldg(t + off - hp) : mid[m];
        dn[m] = i0 + 1 < f.nx ? __ldg(t + off + hp) : mid[m];
        n1[m] = i0 + 2 < f.nx ? __ldg(t + off + 2 * hp) : dn[m];
    }
    // plain loads: with clear_input the same thread stores to this address after reading it
    const float2* iin = reinterpret_cast<const float2*>(a.I_in);
    float2 vin = HAS_I ? iin[off] : make_float2(a.I_uniform, a.I_uniform);
    float2 vn1 = (HAS_I && i0 + 1 < i1) ? iin[off + hp] : vin;

    bool bad = false;
    float ref_sum = 0.f;
    if (a.zero_scalar && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *a.zero_scalar = 0.0;

    for (int i = i0; i < i1; ++i) {
        float2 n2[NM];
#pragma unroll
        for (int m = 0; m < NM; ++m)
            n2[m] = (i + 3 < f.nx && i + 3 <= i1) ? __ldg(reinterpret_cast<const float2*>(a.map[m]) + off + 3 * hp) : n1[m];
        float2 vn2 = vn1;
        if (HAS_I && i + 2 < i1) vn2 = iin[off + 2 * hp];

        const bool inner_row = i > 0 && i < f.nx - 1;   // warp-uniform
        float dxo0 = 0.f, dyo0 = 0.f, dxo1 = 0.f, dyo1 = 0.f, dxr0 = 0.f, dyr0 = 0.f, dxr1 = 0.f, dyr1 = 0.f;
        float arg0 = 0.f, arg1 = 0.f;
#pragma unroll
        for (int m = 0; m < NM; ++m) {
            const float* t = a.map[m];
            float lf = __shfl_up_sync(FULL_MASK, mid[m].y, 1);     // column j-1
            float rt = __shfl_down_sync(FULL_MASK, mid[m].x, 1);   // column j+2
            if (edge_lane) {   // one predicated load serves both ends of the strip
                const float h = __ldg(t + offh);
                if (lane == 0) lf = h; else rt = h;
            }
            // np.gradient(edge_order=2) numerators times 2h (refractionFileNumba2.py:54)
            float gy0 = mid[m].y - lf, gy1 = rt - mid[m].x;
            if (jc == 0) gy0 = -3.f * mid[m].x + 4.f * mid[m].y - rt;                       // rt = column 2
            if (jc + 2 == f.ny) gy1 = 3.f * mid[m].y - 4.f * mid[m].x + lf;                 // lf = column ny-3
            float gx0, gx1;
            if (inner_row) {
                gx0 = dn[m].x - up[m].x;
                gx1 = dn[m].y - up[m].y;
            } else if (i == 0) {
                const float2 r2 = __ldg(reinterpret_cast<const float2*>(t + (size_t)2 * f.ny + jc));
                gx0 = -3.f * mid[m].x + 4.f * dn[m].x - r2.x;
                gx1 = -3.f * mid[m].y + 4.f * dn[m].y - r2.y;
            } else {
                const float2 r3 = __ldg(reinterpret_cast<const float2*>(t + (size_t)(f.nx - 3) * f.ny + jc));
                gx0 = 3.f * mid[m].x - 4.f * up[m].x + r3.x;
                gx1 = 3.f * mid[m].y - 4.f * up[m].y + r3.y;
            }
            dxo0 = fmaf(a.g_obj[m], gx0, dxo0); dyo0 = fmaf(a.g_obj[m], gy0, dyo0);
            dxo1 = fmaf(a.g_obj[m], gx1, dxo1); dyo1 = fmaf(a.g_obj[m], gy1, dyo1);
            if (DUAL) {
                dxr0 = fmaf(a.g_ref[m], gx0, dxr0); dyr0 = fmaf(a.g_ref[m], gy0, dyr0);
                dxr1 = fmaf(a.g_ref[m], gx1, dxr1); dyr1 = fmaf(a.g_ref[m], gy1, dyr1);
            }
            if (ATT) { arg0 = fmaf(a.att[m], mid[m].x, arg0); arg1 = fmaf(a.att[m], mid[m].y, arg1); }
        }
        const float vo0 = ATT ? vin.x * expf(-arg0) : vin.x, vo1 = ATT ? vin.y * expf(-arg1) : vin.y;   // Sample.py:347
        if (live) {
            const float2 z = make_float2(0.f, 0.f);
            if (a.zero[0]) reinterpret_cast<float2*>(a.zero[0])[off] = z;
            if (a.zero[1]) reinterpret_cast<float2*>(a.zero[1])[off] = z;
            if (a.zero[2]) reinterpret_cast<float2*>(a.zero[2])[off] = z;
            if (HAS_I && a.clear_input) reinterpret_cast<float2*>(const_cast<float*>(a.I_in))[off] = z;
        }
        const PairRay q0 = pair_ray(i, j, vo0, dxo0, dyo0, f.nx, f.ny), q1 = pair_ray(i, j + 1, vo1, dxo1, dyo1, f.nx, f.ny);
        deposit_pair(a.out_obj, q0, q1, i, j, vo0, dxo0, dyo0, vo1, dxo1, dyo1, f, a.clamp_x, a.clamp_y, live, bad);
        if (DUAL) {
            // outside the sample the two beams are the same rays: form them once, deposit them twice
            const bool same = dxo0 == dxr0 && dyo0 == dyr0 && dxo1 == dxr1 && dyo1 == dyr1 && vo0 == vin.x && vo1 == vin.y;
            if (__all_sync(FULL_MASK, same)) {
                ref_sum += deposit_pair(a.out_ref, q0, q1, i, j, vo0, dxo0, dyo0, vo1, dxo1, dyo1, f, a.clamp_x, a.clamp_y, live, bad);
            } else {
                const PairRay p0 = pair_ray(i, j, vin.x, dxr0, dyr0, f.nx, f.ny), p1 = pair_ray(i, j + 1, vin.y, dxr1, dyr1, f.nx, f.ny);
                ref_sum += deposit_pair(a.out_ref, p0, p1, i, j, vin.x, dxr0, dyr0, vin.y, dxr1, dyr1, f, a.clamp_x, a.clamp_y, live, bad);
            }
        }
#pragma unroll
        for (int m = 0; m < NM; ++m) { up[m] = mid[m]; mid[m] = dn[m]; dn[m] = n1[m]; n1[m] = n2[m]; }
        vin = vn1; vn1 = vn2;
        off += hp; offh += f.ny;
    }
    if (bad && a.flag) atomicOr(a.flag, FLAG_NONFINITE);
    if (DUAL && a.sum_ref) {   // one double atomic per warp
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) ref_sum += __shfl_xor_sync(FULL_MASK, ref_sum, d);
        if (lane == 0) atomicAdd(a.sum_ref, (double)ref_sum);
    }
}

template <int NM, bool DUAL, bool HAS_I, bool ATT>
static int launch_refract_pair(const RefractArgs<float>& a_in, int rows_override, cudaStream_t s) {
    RefractArgs<float> a = a_in;
    // rows per warp: ~1.5 waves of 148 SMs x 6 resident blocks; more rows = fewer prologue re-reads
    const int strips = div_up(a.f.ny, PAIR_COLS);
    int rows = (int)((long)a.f.nx * strips / (148L * 6 * 3 / 2));
    rows = rows < 8 ? 8 : (rows > 64 ? 64 : rows);
    if (rows_override > 0) rows = rows_override;
    a.rows = rows;
    dim3 grid(strips, div_up(a.f.nx, a.rows));
    refract_pair_kernel<NM, DUAL, HAS_I, ATT><<<grid, PAIR_THREADS, 0, s>>>(a);
    PARESIS_LAUNCH_CHECK("refract_pair_kernel");
    return PARESIS_OK;
}

}  // namespace paresis
