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
constexpr int FIX_BITS = 19;  // with 16 x 256 rays per tile: rays up to 2 x intensity_scale stay in fixed point

__device__ __forceinline__ float ex2_fast(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// Block-uniform description of one output image: shared tile, global image, and the window of lower cells
// (r, c) whose four cells lie both inside the tile and strictly inside the image.
struct TileTarget {
    unsigned* tile;
    float* out;
    int r0, c0;            // first lower cell of the window (image coordinates)
    unsigned nr, nc;       // window extent: lower cells r0 .. r0+nr-1, c0 .. c0+nc-1
    int rlo, clo;          // image coordinates of tile cell (0, 0)
};

// One ray straight to L2 in fp32, cell by cell; cells outside the image are dropped, which is what zero-padding,
// scattering and cropping does (refractionFileNumba2.py:65-78; the |D| > N kill of :61-64 only removes rays that
// land outside anyway).  Returns what was deposited inside the image.  TWIN: the same ray goes to `out2` as well.
template <bool TWIN>
__device__ __forceinline__ float deposit_direct(float* out, float* out2, int i, int j, float v, float dx, float dy, int nx, int ny,
                                                bool& bad) {
    const float flx = floorf(dx), fly = floorf(dy);
    const float fx = dx - flx, fy = dy - fly;
    const int r = i + __float2int_rd(dx), c = j + __float2int_rd(dy);   // saturating
    const float v1 = v * fx, v0 = v - v1;
    const float w1 = v0 * fy, w0 = v0 - w1, w3 = v1 * fy, w2 = v1 - w3;
    bad |= !(fabsf(w0 + w3) <= 3.0e38f);
    const bool ra = (unsigned)r < (unsigned)nx, rb = (unsigned)(r + 1) < (unsigned)nx;
    const bool ca = (unsigned)c < (unsigned)ny, cb = (unsigned)(c + 1) < (unsigned)ny;
    const long long o = (long long)r * ny + c;
    float* p = out + o;
    float sum = 0.f;
    if (ra && ca && w0 != 0.f) { red_add(p, w0); sum += w0; }
    if (ra && cb && w1 != 0.f) { red_add(p + 1, w1); sum += w1; }
    if (rb && ca && w2 != 0.f) { red_add(p + ny, w2); sum += w2; }
    if (rb && cb && w3 != 0.f) { red_add(p + ny + 1, w3); sum += w3; }
    if (TWIN) {
        float* q = out2 + o;
        if (ra && ca && w0 != 0.f) red_add(q, w0);
        if (ra && cb && w1 != 0.f) red_add(q + 1, w1);
        if (rb && ca && w2 != 0.f) red_add(q + ny, w2);
        if (rb && cb && w3 != 0.f) red_add(q + ny + 1, w3);
    }
    return sum;
}

// One ray.  Fast path: fixed-point deposit into the tile (native ATOMS.ADD).  Everything else -- the ray leaves
// the tile, touches the image border, is too bright / negative / not finite -- goes to L2 in fp32, cell by cell;
// cells outside the image are dropped, which is what zero-padding, scattering and cropping does
// (refractionFileNumba2.py:65-78; the |D| > N kill of :61-64 only removes rays that land outside anyway).
// Returns what was deposited inside the image.
// TWIN: the same ray goes to a second image as well (`t2`: its tile sits `twin_off` words after the first).
template <int SC, bool TWIN>
__device__ __forceinline__ float deposit(const TileTarget& t, int i, int j, float v, float dx, float dy, int nx, int ny,
                                         float scale, unsigned vmax_bits, bool live, bool& bad, float* out2 = nullptr,
                                         int twin_off = 0) {
    const float flx = floorf(dx), fly = floorf(dy);
    const float fx = dx - flx, fy = dy - fly;
    const int r = i + __float2int_rd(dx), c = j + __float2int_rd(dy);   // saturating
    // v in [0, vmax) as one unsigned compare on the bit pattern (negative and NaN patterns are larger)
    const bool fast = live && (unsigned)(r - t.r0) < t.nr && (unsigned)(c - t.c0) < t.nc && __float_as_uint(v) < vmax_bits &&
                      fx + fy < 3.f;
    if (fast) {
        const float vs = v * scale;
        const float v1 = vs * fx, v0 = vs - v1;
        const float w1 = v0 * fy, w0 = v0 - w1, w3 = v1 * fy, w2 = v1 - w3;
        unsigned* p = t.tile + (r - t.rlo) * SC + (c - t.clo);
        const unsigned u0 = __float2uint_rn(w0), u1 = __float2uint_rn(w1), u2 = __float2uint_rn(w2), u3 = __float2uint_rn(w3);
        atomicAdd(p, u0);
        atomicAdd(p + 1, u1);
        atomicAdd(p + SC, u2);
        atomicAdd(p + SC + 1, u3);
        if (TWIN) {
            unsigned* q = p + twin_off;
            atomicAdd(q, u0);
            atomicAdd(q + 1, u1);
            atomicAdd(q + SC, u2);
            atomicAdd(q + SC + 1, u3);
        }
        return v;
    }
    return live ? deposit_direct<TWIN>(t.out, out2, i, j, v, dx, dy, nx, ny, bad) : 0.f;
}

// Tile -> image: dense 128-bit REDs, all-zero quads skipped; image borders and odd pitches fall back to scalars.
template <int SR, int SC>
__device__ __forceinline__ void flush_tile(const unsigned* tile, float* out, int rlo, int clo, int nx, int ny, float inv_scale,
                                           bool vec_ok, int nrows = SR) {
    constexpr int Q = SC / 4;
    for (int idx = threadIdx.x; idx < nrows * Q; idx += blockDim.x) {
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
    static_assert(((1ull << 32) / (TR * TILE_COLS)) >= (2ull << FIX_BITS), "rays up to 2 x intensity_scale must fit the fixed-point tile");
    extern __shared__ __align__(16) unsigned tile_smem[];

    const Frame f = a.f;
    const int lane = threadIdx.x & 31;
    const int j = blockIdx.x * TILE_COLS + threadIdx.x;
    const int i0 = blockIdx.y * a.rows;          // a.rows <= TR: picked on the host to fill whole waves
    const int i1 = min(i0 + a.rows, f.nx);
    const bool live = j < f.ny;
    const int jc = live ? j : f.ny - 1;  // dead lanes read a valid address, contribute nothing
    const bool inner_cols = __all_sync(FULL_MASK, live && j > 0 && j < f.ny - 1);

    TileTarget tobj, tref;
    tobj.tile = tile_smem; tobj.out = a.out_obj;
    tobj.rlo = i0 - H; tobj.clo = blockIdx.x * TILE_COLS - H;
    tobj.r0 = max(tobj.rlo, 0); tobj.c0 = max(tobj.clo, 0);
    tobj.nr = (unsigned)max(min(tobj.rlo + SR - 1, f.nx - 1) - tobj.r0, 0);
    tobj.nc = (unsigned)max(min(tobj.clo + SC - 1, f.ny - 1) - tobj.c0, 0);
    tref = tobj; tref.tile = tile_smem + SR * SC; tref.out = a.out_ref;

    {   // zero the tile(s)
        uint4* z = reinterpret_cast<uint4*>(tile_smem);
        constexpr int NZ = SR * SC / 4 * (DUAL ? 2 : 1);
        for (int k = threadIdx.x; k < NZ; k += TILE_COLS) z[k] = make_uint4(0u, 0u, 0u, 0u);
    }

    // Row ring: slot k of step s holds row i-1+k; rows i-1, i, i+1 are used and row i+2 is fetched now (one full
    // step ahead of its first use).  The loop is unrolled by the ring size so that the rotation is register renaming -- a rotation by
    // moves would wait for the load issued in the same step.  `hal` = the column next to the strip (edge lanes),
    // `inten` = the input intensity, same indexing.
    constexpr int RING = 4;
    float row[RING][NM], hal[RING][NM], inten[RING];
    const bool edge_lane = lane == 0 || lane == 31;
    const int jh = min(max(lane == 0 ? jc - 1 : jc + 1, 0), f.ny - 1);
    int off = i0 * f.ny + jc;          // element offset of (i, jc); nx*ny < 2^30 (host check)
    int offh = i0 * f.ny + jh;         // (i, column next to the strip) for the edge lanes
    const int last = (f.nx - 1) * f.ny;   // offsets are clamped to the image instead of predicating the loads
#pragma unroll
    for (int k = 0; k < RING - 1; ++k) {
        const int d = (k - 1) * f.ny;
#pragma unroll
        for (int m = 0; m < NM; ++m) {
            row[k][m] = __ldg(a.map[m] + min(max(off + d, jc), last + jc));
            hal[k][m] = edge_lane ? __ldg(a.map[m] + min(max(offh + d, jh), last + jh)) : 0.f;
        }
        // plain loads: with clear_input the same thread stores to this address after reading it
        inten[k] = HAS_I ? a.I_in[min(max(off + d, jc), last + jc)] : a.I_uniform;
    }

    // rays below vmax convert to less than 2^32 / (TR * 256) - 4 units: a whole tile cannot overflow one cell
    const float fix_scale = (float)(1u << FIX_BITS) / a.intensity_scale;
    const unsigned vmax_bits = __float_as_uint(a.intensity_scale * (float)((unsigned)((1ull << 32) / (TR * TILE_COLS)) - 8u) / (float)(1u << FIX_BITS));
    const float neg_log2e = -1.4426950408889634f;
    const bool zero_fill = a.zero[0] != nullptr;    // the host fills unused slots with a used pointer
    const bool clear_in = HAS_I && a.clear_input;
    bool bad = false;
    float ref_sum = 0.f;
    if (a.zero_scalar && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *a.zero_scalar = 0.0;
    __syncthreads();

    for (int ib = i0; ib < i1; ib += RING) {
#pragma unroll
        for (int s = 0; s < RING; ++s) {
            const int i = ib + s;
            if (i >= i1) break;                                   // warp-uniform
            const int kup = s % RING, kmid = (s + 1) % RING, kdn = (s + 2) % RING, knew = (s + 3) % RING;
            {
                const int d = 2 * f.ny;
#pragma unroll
                for (int m = 0; m < NM; ++m) {
                    row[knew][m] = __ldg(a.map[m] + min(off + d, last + jc));
                    hal[knew][m] = edge_lane ? __ldg(a.map[m] + min(offh + d, last + jh)) : 0.f;
                }
                inten[knew] = HAS_I ? a.I_in[min(off + d, last + jc)] : a.I_uniform;
            }
            const bool inner = inner_cols && i > 0 && i < f.nx - 1;   // warp-uniform
            float dxo = 0.f, dyo = 0.f, dxr = 0.f, dyr = 0.f, arg = 0.f;
#pragma unroll
            for (int m = 0; m < NM; ++m) {
                const float* t = a.map[m];
                const float mid = row[kmid][m], up = row[kup][m], dn = row[kdn][m];
                float lf = __shfl_up_sync(FULL_MASK, mid, 1);
                float rt = __shfl_down_sync(FULL_MASK, mid, 1);
                if (lane == 0) lf = hal[kmid][m];
                if (lane == 31) rt = hal[kmid][m];
                float gy, gx;
                if (inner) {
                    gy = rt - lf;
                    gx = dn - up;
                } else {
                    // np.gradient(edge_order=2) numerators times 2h (refractionFileNumba2.py:54)
                    const float* r = t + (size_t)i * f.ny;
                    if (jc == 0) gy = -3.f * mid + 4.f * rt - __ldg(r + 2);
                    else if (jc == f.ny - 1) gy = 3.f * mid - 4.f * lf + __ldg(r + f.ny - 3);
                    else gy = rt - lf;
                    if (i == 0) gx = -3.f * mid + 4.f * dn - __ldg(t + (size_t)2 * f.ny + jc);
                    else if (i == f.nx - 1) gx = 3.f * mid - 4.f * up + __ldg(t + (size_t)(f.nx - 3) * f.ny + jc);
                    else gx = dn - up;
                }
                dxo = fmaf(a.g_obj[m], gx, dxo);
                dyo = fmaf(a.g_obj[m], gy, dyo);
                if (DUAL) {
                    dxr = fmaf(a.g_ref[m], gx, dxr);
                    dyr = fmaf(a.g_ref[m], gy, dyr);
                }
                if (ATT) arg = fmaf(a.att[m], mid, arg);
            }
            const float vin = inten[kmid];
            // Sample.py:347; ex2.approx keeps ~2e-7 relative accuracy over the attenuation range
            const float vo = ATT ? vin * ex2_fast(arg * neg_log2e) : vin;
            if (live) {
                if (zero_fill) { a.zero[0][off] = 0.f; a.zero[1][off] = 0.f; a.zero[2][off] = 0.f; }
                if (clear_in) const_cast<float*>(a.I_in)[off] = 0.f;
            }
            if (DUAL) {
                // outside the sample the two beams are the same ray: form it once, deposit it twice
                const bool same = dxo == dxr && dyo == dyr && vo == vin;
                if (__all_sync(FULL_MASK, same)) {
                    ref_sum += deposit<SC, true>(tobj, i, j, vo, dxo, dyo, f.nx, f.ny, fix_scale, vmax_bits, live, bad, a.out_ref, SR * SC);
                } else {
                    deposit<SC, false>(tobj, i, j, vo, dxo, dyo, f.nx, f.ny, fix_scale, vmax_bits, live, bad);
                    ref_sum += deposit<SC, false>(tref, i, j, vin, dxr, dyr, f.nx, f.ny, fix_scale, vmax_bits, live, bad);
                }
            } else {
                deposit<SC, false>(tobj, i, j, vo, dxo, dyo, f.nx, f.ny, fix_scale, vmax_bits, live, bad);
            }
            off += f.ny; offh += f.ny;
        }
    }
    if (bad && a.flag) atomicOr(a.flag, FLAG_NONFINITE);
    if (DUAL && a.sum_ref) {   // one double atomic per warp
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) ref_sum += __shfl_xor_sync(FULL_MASK, ref_sum, d);
        if (lane == 0) atomicAdd(a.sum_ref, (double)ref_sum);
    }
    __syncthreads();
    const bool vec_ok = (f.ny & 3) == 0;
    const float inv_scale = a.intensity_scale / (float)(1u << FIX_BITS);
    flush_tile<SR, SC>(tobj.tile, a.out_obj, tobj.rlo, tobj.clo, f.nx, f.ny, inv_scale, vec_ok && (reinterpret_cast<uintptr_t>(a.out_obj) & 15) == 0);
    if (DUAL) flush_tile<SR, SC>(tref.tile, a.out_ref, tobj.rlo, tobj.clo, f.nx, f.ny, inv_scale, vec_ok && (reinterpret_cast<uintptr_t>(a.out_ref) & 15) == 0);
}

// Rows per block: all blocks cost the same, so a launch takes ceil(blocks / resident slots) rounds.  Pick the row count
// (<= TR, the size the tile was built for) whose last round is fullest, charging the fixed cost of a block (three
// extra thickness rows, halo flush) against short tiles.
inline int pick_tile_rows(int nx, int strips, int slots, int max_rows) {
    int best = max_rows;
    double best_score = -1.0;
    for (int rows = max_rows; rows >= (max_rows > 8 ? 6 : 4); --rows) {
        const long blocks = (long)strips * div_up(nx, rows);
        const long rounds = (blocks + slots - 1) / slots;
        const double score = (double)blocks / (double)(rounds * slots) * rows / (rows + 4.0);
        if (score > best_score + 1e-9) { best_score = score; best = rows; }
    }
    return best;
}

template <int NM, bool DUAL, bool HAS_I, bool ATT, int TR, int H>
static int launch_refract_tile(const RefractArgs<float>& a_in, cudaStream_t s) {
    constexpr int SR = TR + 2 * H + 1, SC = TILE_COLS + 2 * H;
    constexpr size_t smem = sizeof(unsigned) * SR * SC * (DUAL ? 2 : 1);
    static int slots = 0;
    if (!slots) {
        PARESIS_CUDA(cudaFuncSetAttribute(refract_tile_kernel<NM, DUAL, HAS_I, ATT, TR, H>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0, dev = 0, sms = 0;
        PARESIS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, refract_tile_kernel<NM, DUAL, HAS_I, ATT, TR, H>,
                                                                   TILE_COLS, smem));
        PARESIS_CUDA(cudaGetDevice(&dev));
        PARESIS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        slots = (per_sm > 0 ? per_sm : 1) * (sms > 0 ? sms : 148);
    }
    RefractArgs<float> a = a_in;
    const int strips = div_up(a.f.ny, TILE_COLS);
    a.rows = pick_tile_rows(a.f.nx, strips, slots, TR);
    dim3 grid(strips, div_up(a.f.nx, a.rows));
    refract_tile_kernel<NM, DUAL, HAS_I, ATT, TR, H><<<grid, TILE_COLS, smem, s>>>(a);
    PARESIS_LAUNCH_CHECK("refract_tile_kernel");
    return PARESIS_OK;
}

}  // namespace paresis
