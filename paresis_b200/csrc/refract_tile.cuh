// Fused refraction hop with SHARED-MEMORY TILE accumulation in 32-bit FIXED POINT (the production
// kernel of the per-energy loop; same contract as refract_kernel in refraction.cu).
//
// Why a tile: on a membrane the displacement field is torn every few pixels (sphere caps), so the REDs
// of a warp land in ~11 different 32-byte sectors per instruction, and L2 retires atomics per SECTOR
// (~170 sector-ops/ns on B200, measured: tools/redbench.cu, profiles/).  The direct-to-L2 kernel is
// pinned at that rate.  Here a block of 8 warps owns TR source rows x 256 source columns and bins its
// rays into a shared-memory tile that covers the source tile plus a halo of H pixels; when the block
// is done the tile leaves the SM as dense 128-bit REDs (every sector full, all-zero quads skipped).
//
// Why fixed point: sm_100a has no native fp32 add on shared memory (atomicAdd compiles to an
// ATOMS.CAST.SPIN loop, ~5 instructions per try and ~2 tries per deposit on this field), but 32-bit
// integer ATOMS.ADD is native and fire-and-forget (tools/smembench.cu: 2.1 vs 4.8-13.6 cycles per
// warp-op).  A ray's intensity is converted once, V = round(v * S), and split EXACTLY between its four
// cells with 24-bit fractions, so a ray deposits precisely V units whatever the summation order: the
// tile contents are bit-reproducible.  S = 2^FIX_BITS / intensity_scale, and only rays with
// V < 2^32 / (rays per tile) take this path, so a cell cannot overflow whatever the field focuses.
// Quantisation: 2^-FIX_BITS of the beam intensity per deposit (~1e-6 relative L2 in the image).
// Brighter rays, rays that leave the halo, and rays that touch the image border (reference edge
// rules, splat.cuh) go straight to L2 in fp32 -- results are the reference's for any input.
//
// Other changes against refract_kernel: thickness rows are fetched two rows ahead; where the object
// and reference beams coincide (outside the sample) the ray is formed once and deposited twice.
#pragma once
#include "refract_common.cuh"

namespace paresis {

constexpr int TILE_COLS = 256;   // 8 warps x 32 lanes

__device__ __forceinline__ void red_add4(float* p, const float4& v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// r, c of the lower cell and the four weights; `simple` = all four cells strictly inside the image
struct TileRay {
    int r, c;
    float v, fx, fy;          // intensity and the fractional displacement (weights (1-f, f) per axis)
    bool simple;
};

constexpr int FIX_BITS = 19;  // with 16 x 256 rays per tile: rays up to 2 x intensity_scale stay in fixed point

__device__ __forceinline__ TileRay tile_ray(int i, int j, float v, float dx, float dy, int nx, int ny) {
    TileRay q;
    const float flx = floorf(dx), fly = floorf(dy);
    q.fx = dx - flx; q.fy = dy - fly;
    q.v = v;
    q.r = i + __float2int_rd(dx);    // saturating; wrap-around fails `simple`
    q.c = j + __float2int_rd(dy);
    q.simple = ((unsigned)q.r < (unsigned)(nx - 1)) & ((unsigned)q.c < (unsigned)(ny - 1));
    return q;
}

template <int SR, int SC>
struct TileSplatter {
    unsigned* tile;  // [SR][SC] shared, fixed point
    float* out;      // global image
    int rlo, clo, ny;
    float scale, vmax;   // fixed units per intensity unit; rays at or above vmax bypass the tile
    bool bad;

    __device__ __forceinline__ void init(unsigned* tile_, float* out_, int rlo_, int clo_, int ny_, float scale_, float vmax_) {
        tile = tile_; out = out_; rlo = rlo_; clo = clo_; ny = ny_; scale = scale_; vmax = vmax_; bad = false;
    }
    // all four cells inside the image (q.simple): tile if it reaches, else straight to L2
    __device__ __forceinline__ void put(const TileRay& q) {
        bad |= !(fabsf(fmaf(q.fx + q.fy, 0.f, q.v)) <= 3.0e38f);   // NaN / Inf in the intensity or the displacement
        const unsigned sr = (unsigned)(q.r - rlo), sc = (unsigned)(q.c - clo);
        if (sr < (unsigned)(SR - 1) && sc < (unsigned)(SC - 1) && q.v >= 0.f && q.v < vmax) {
            // exact split of V between the four cells: 24-bit fractions, V < 2^24
            const unsigned V = __float2uint_rn(q.v * scale);
            const unsigned fxi = __float2uint_rn(q.fx * 16777216.f), fyi = __float2uint_rn(q.fy * 16777216.f);
            const unsigned V1 = __umulhi(V << 8, fxi), V0 = V - V1;
            const unsigned W1 = __umulhi(V0 << 8, fyi), W3 = __umulhi(V1 << 8, fyi);
            unsigned* t = tile + sr * SC + sc;
            atomicAdd(t, V0 - W1);
            atomicAdd(t + 1, W1);
            atomicAdd(t + SC, V1 - W3);
            atomicAdd(t + SC + 1, W3);
        } else {
            const float v1 = q.v * q.fx, v0 = q.v - v1;
            const float w1 = v0 * q.fy, w0 = v0 - w1, w3 = v1 * q.fy, w2 = v1 - w3;
            float* p = out + (size_t)q.r * ny + q.c;
            if (w0 != 0.f) red_add(p, w0);
            if (w1 != 0.f) red_add(p + 1, w1);
            if (w2 != 0.f) red_add(p + ny, w2);
            if (w3 != 0.f) red_add(p + ny + 1, w3);
        }
    }
};

// Tile -> image: dense 128-bit REDs, all-zero quads skipped; image borders and odd pitches fall back to scalars.
template <int SR, int SC>
__device__ __forceinline__ void flush_tile(const unsigned* tile, float* out, int rlo, int clo, int nx, int ny, float inv_scale,
                                           bool vec_ok) {
    constexpr int Q = SC / 4;
    for (int idx = threadIdx.x; idx < SR * Q; idx += blockDim.x) {
        const int sr = idx / Q, q4 = idx - sr * Q;
        const int r = rlo + sr, c = clo + 4 * q4;
        if ((unsigned)r >= (unsigned)nx) continue;
        const uint4 u = *reinterpret_cast<const uint4*>(tile + sr * SC + 4 * q4);
        if ((u.x | u.y | u.z | u.w) == 0u) continue;
        const float4 v = make_float4((float)u.x * inv_scale, (float)u.y * inv_scale, (float)u.z * inv_scale, (float)u.w * inv_scale);
        float* p = out + (size_t)r * ny + c;
        if (vec_ok && c >= 0 && c + 3 < ny) {
            red_add4(p, v);
        } else {
            if (u.x && (unsigned)c < (unsigned)ny) red_add(p, v.x);
            if (u.y && (unsigned)(c + 1) < (unsigned)ny) red_add(p + 1, v.y);
            if (u.z && (unsigned)(c + 2) < (unsigned)ny) red_add(p + 2, v.z);
            if (u.w && (unsigned)(c + 3) < (unsigned)ny) red_add(p + 3, v.w);
        }
    }
}

template <int NM, bool DUAL, bool HAS_I, bool ATT, int TR, int H>
__global__ void __launch_bounds__(TILE_COLS)
refract_tile_kernel(const RefractArgs<float> a) {
    constexpr int SR = TR + 2 * H + 1, SC = TILE_COLS + 2 * H;
    static_assert(H % 4 == 0 && H >= 4, "halo must keep the tile 16-byte aligned");
    static_assert((long)TR * TILE_COLS << FIX_BITS <= (1L << 32), "a tile of the brightest rays must fit in 32 bits");
    extern __shared__ __align__(16) unsigned tile_smem[];
    unsigned* tobj = tile_smem;
    unsigned* tref = tile_smem + SR * SC;

    const Frame f = a.f;
    const int lane = threadIdx.x & 31;
    const int j = blockIdx.x * TILE_COLS + threadIdx.x;
    const int i0 = blockIdx.y * TR;
    const int i1 = min(i0 + TR, f.nx);
    const int rlo = i0 - H, clo = blockIdx.x * TILE_COLS - H;
    const bool live = j < f.ny;
    const int jc = live ? j : f.ny - 1;  // dead lanes read a valid address, contribute nothing
    const bool inner_cols = __all_sync(FULL_MASK, live && j > 0 && j < f.ny - 1);

    {   // zero the tile(s)
        uint4* z = reinterpret_cast<uint4*>(tile_smem);
        constexpr int NZ = SR * SC / 4 * (DUAL ? 2 : 1);
        for (int k = threadIdx.x; k < NZ; k += TILE_COLS) z[k] = make_uint4(0u, 0u, 0u, 0u);
    }

    // rolling rows: up = i-1, mid = i, dn = i+1, n1 = i+2 (in flight from the previous step); row i+3 is fetched now
    float up[NM], mid[NM], dn[NM], n1[NM];
    const bool edge_lane = lane == 0 || lane == 31;
    const int jh = min(max(lane == 0 ? jc - 1 : jc + 1, 0), f.ny - 1);
    int off = i0 * f.ny + jc;          // element offset of (i, jc); nx*ny < 2^30 (host check)
    int offh = i0 * f.ny + jh;         // (i, column next to the strip) for the edge lanes
#pragma unroll
    for (int m = 0; m < NM; ++m) {
        const float* t = a.map[m];
        mid[m] = __ldg(t + off);
        up[m] = i0 > 0 ? __ldg(t + off - f.ny) : mid[m];
        dn[m] = i0 + 1 < f.nx ? __ldg(t + off + f.ny) : mid[m];
        n1[m] = i0 + 2 < f.nx ? __ldg(t + off + 2 * f.ny) : dn[m];
    }
    // plain loads: with clear_input the same thread stores to this address after reading it
    float vin = HAS_I ? a.I_in[off] : a.I_uniform;
    float vn1 = (HAS_I && i0 + 1 < i1) ? a.I_in[off + f.ny] : vin;

    TileSplatter<SR, SC> sp_obj, sp_ref;
    // rays below vmax convert to V < 2^32 / (TR * 256): the whole tile cannot overflow one cell
    const float fix_scale = (float)(1u << FIX_BITS) / a.intensity_scale;
    const float fix_vmax = a.intensity_scale * (float)((1ull << 32) / ((unsigned long long)TR * TILE_COLS) - 2) / (float)(1u << FIX_BITS);
    sp_obj.init(tobj, a.out_obj, rlo, clo, f.ny, fix_scale, fix_vmax);
    if (DUAL) sp_ref.init(tref, a.out_ref, rlo, clo, f.ny, fix_scale, fix_vmax);
    Splatter<0> slow_obj, slow_ref;    // reference edge rules, straight to L2
    slow_obj.init(a.out_obj, f.ny, nullptr);
    if (DUAL) slow_ref.init(a.out_ref, f.ny, nullptr);
    float ref_sum = 0.f;
    if (a.zero_scalar && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *a.zero_scalar = 0.0;
    __syncthreads();

    for (int i = i0; i < i1; ++i) {
        float n2[NM];
#pragma unroll
        for (int m = 0; m < NM; ++m) n2[m] = (i + 3 < f.nx && i + 3 <= i1) ? __ldg(a.map[m] + off + 3 * f.ny) : n1[m];
        float vn2 = vn1;
        if (HAS_I && i + 2 < i1) vn2 = a.I_in[off + 2 * f.ny];

        const bool inner = inner_cols && i > 0 && i < f.nx - 1;   // warp-uniform
        float dxo = 0.f, dyo = 0.f, dxr = 0.f, dyr = 0.f, arg = 0.f;
#pragma unroll
        for (int m = 0; m < NM; ++m) {
            const float* t = a.map[m];
            float lf = __shfl_up_sync(FULL_MASK, mid[m], 1);
            float rt = __shfl_down_sync(FULL_MASK, mid[m], 1);
            if (edge_lane) {   // one predicated load serves both ends of the strip
                const float h = __ldg(t + offh);
                if (lane == 0) lf = h; else rt = h;
            }
            float gy, gx;
            if (inner) {
                gy = rt - lf;
                gx = dn[m] - up[m];
            } else {
                // np.gradient(edge_order=2) numerators times 2h (refractionFileNumba2.py:54)
                const float* r = t + (size_t)i * f.ny;
                if (jc == 0) gy = -3.f * mid[m] + 4.f * rt - __ldg(r + 2);
                else if (jc == f.ny - 1) gy = 3.f * mid[m] - 4.f * lf + __ldg(r + f.ny - 3);
                else gy = rt - lf;
                if (i == 0) gx = -3.f * mid[m] + 4.f * dn[m] - __ldg(t + (size_t)2 * f.ny + jc);
                else if (i == f.nx - 1) gx = 3.f * mid[m] - 4.f * up[m] + __ldg(t + (size_t)(f.nx - 3) * f.ny + jc);
                else gx = dn[m] - up[m];
            }
            dxo = fmaf(a.g_obj[m], gx, dxo);
            dyo = fmaf(a.g_obj[m], gy, dyo);
            if (DUAL) {
                dxr = fmaf(a.g_ref[m], gx, dxr);
                dyr = fmaf(a.g_ref[m], gy, dyr);
            }
            if (ATT) arg = fmaf(a.att[m], mid[m], arg);
        }
        float vo = ATT ? vin * expf(-arg) : vin;   // Sample.py:347
        float vr = vin;
        if (live) {
            if (a.zero[0]) a.zero[0][off] = 0.f;
            if (a.zero[1]) a.zero[1][off] = 0.f;
            if (a.zero[2]) a.zero[2][off] = 0.f;
            if (HAS_I && a.clear_input) const_cast<float*>(a.I_in)[off] = 0.f;
        }
        {
            // A `simple` ray lands strictly inside the image, so |D| < N: the kill rule of
            // refractionFileNumba2.py:61-64 cannot fire, and the |D| < 1e-12 -> 0 rule (:59-60) changes
            // nothing at fp32 resolution.  Everything else is cleaned and takes the reference's edge rules.
            const TileRay q = tile_ray(i, j, vo, dxo, dyo, f.nx, f.ny);
            const bool fast = inner_cols && __all_sync(FULL_MASK, q.simple);
            if (fast) sp_obj.put(q);
            else {
                clean(vo, dxo, dyo, a.clamp_x, a.clamp_y);
                slow_obj.put(live ? make_ray(i, j, vo, dxo, dyo, f) : empty_ray());
            }
            if (DUAL) {
                // outside the sample the two beams are the same ray: form it once, deposit it twice
                const bool same = fast && __all_sync(FULL_MASK, dxo == dxr && dyo == dyr && vo == vr);
                if (same) {
                    sp_ref.put(q);
                    ref_sum += vr;
                } else {
                    const TileRay qr = tile_ray(i, j, vr, dxr, dyr, f.nx, f.ny);
                    if (inner_cols && __all_sync(FULL_MASK, qr.simple)) {
                        sp_ref.put(qr);
                        ref_sum += vr;
                    } else {
                        clean(vr, dxr, dyr, a.clamp_x, a.clamp_y);
                        const Ray qs = live ? make_ray(i, j, vr, dxr, dyr, f) : empty_ray();
                        slow_ref.put(qs);
#pragma unroll
                        for (int k = 0; k < 4; ++k) ref_sum += ((qs.ok >> k) & 1u) ? qs.w[k] : 0.f;
                    }
                }
            }
        }
#pragma unroll
        for (int m = 0; m < NM; ++m) { up[m] = mid[m]; mid[m] = dn[m]; dn[m] = n1[m]; n1[m] = n2[m]; }
        vin = vn1; vn1 = vn2;
        off += f.ny; offh += f.ny;
    }
    const bool bad = sp_obj.bad | slow_obj.bad | (DUAL && (sp_ref.bad | slow_ref.bad));
    if (bad && a.flag) atomicOr(a.flag, FLAG_NONFINITE);
    if (DUAL && a.sum_ref) {   // one double atomic per warp
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) ref_sum += __shfl_xor_sync(FULL_MASK, ref_sum, d);
        if (lane == 0) atomicAdd(a.sum_ref, (double)ref_sum);
    }
    __syncthreads();
    const bool vec_ok = (f.ny & 3) == 0;
    const float inv_scale = a.intensity_scale / (float)(1u << FIX_BITS);
    flush_tile<SR, SC>(tobj, a.out_obj, rlo, clo, f.nx, f.ny, inv_scale, vec_ok && (reinterpret_cast<uintptr_t>(a.out_obj) & 15) == 0);
    if (DUAL) flush_tile<SR, SC>(tref, a.out_ref, rlo, clo, f.nx, f.ny, inv_scale, vec_ok && (reinterpret_cast<uintptr_t>(a.out_ref) & 15) == 0);
}

template <int NM, bool DUAL, bool HAS_I, bool ATT, int TR, int H>
static int launch_refract_tile(const RefractArgs<float>& a, cudaStream_t s) {
    constexpr int SR = TR + 2 * H + 1, SC = TILE_COLS + 2 * H;
    constexpr size_t smem = sizeof(unsigned) * SR * SC * (DUAL ? 2 : 1);
    static bool configured = false;
    if (!configured) {
        PARESIS_CUDA(cudaFuncSetAttribute(refract_tile_kernel<NM, DUAL, HAS_I, ATT, TR, H>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    dim3 grid(div_up(a.f.ny, TILE_COLS), div_up(a.f.nx, TR));
    refract_tile_kernel<NM, DUAL, HAS_I, ATT, TR, H><<<grid, TILE_COLS, smem, s>>>(a);
    PARESIS_LAUNCH_CHECK("refract_tile_kernel");
    return PARESIS_OK;
}

}  // namespace paresis
