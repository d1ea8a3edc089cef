// Fused refraction hop (thickness -> transmission -> gradient -> displacement -> scatter) as an OWNER-COMPUTES rolling-
// strip kernel: the production hop of the per-energy loop since round 2.
//
// Reference: Experiment.py:463-474 (the hop sequence), Sample.py:285-351 (setWaveRT), refractionFileNumba2.py:25-86
// (fastRefraction: np.gradient :54, displacement :55-56, clean-up :59-64, pad / scatter / crop :65-78) and :198-263
// (fastloopNumba).  fastRefraction scatters into a fresh zero array (:70, :77): the hop OVERWRITES its outputs; inside a
// detector bin the images of later energies are added to the accumulators (Experiment.py:482-483): ACC mode.
//
// What changes against refract_lean.cuh (source-owner tiles + dense REDs): no zero-fill of the accumulators by the
// previous kernel, no clearing of the input behind the pass, no read-modify-write of the outputs in L2, no per-tile
// set-up every 16 rows -- a block walks a whole row segment with a register ring of thickness rows (4 deep: a row is
// requested 2 rows before its first use), finished rows leave through a dedicated flushing warp (strip.cuh) while the
// eight depositing warps are already in the next rows, and several membrane positions share one launch (blockIdx.z).
// The machinery (ownership windows, circular fixed-point tile, per-warp ray lists + drain launch) is in strip.cuh.
#pragma once
#include <type_traits>

#include "tile_common.cuh"   // deposit_direct, ex2_fast
#include "strip.cuh"

namespace paresis {

constexpr int STRIP_MAX_BATCH = 8;

struct HopItem {
    const float* map[PARESIS_MAX_LAYERS];
    const float* I_in;
    float* out_obj;
    float* out_ref;
    double* sum_ref;      // += what the reference beam deposits inside the image (Experiment.py:485-486); may be null
};

struct HopArgs {
    HopItem item[STRIP_MAX_BATCH];
    float g_obj[PARESIS_MAX_LAYERS], g_ref[PARESIS_MAX_LAYERS], att[PARESIS_MAX_LAYERS];
    float I_uniform;
    float intensity_scale;
    Frame f;
    StripPlan p;
    uint4* far;
    unsigned* far_count;
    int* flag;
    bool vec;
};

template <int NM, bool DUAL, bool HAS_I, int H, bool ACC>
__global__ void __launch_bounds__(STRIP_BLOCK, DUAL ? 3 : 4)
refract_strip_kernel(const HopArgs a) {
    using S = Strip<H>;
    constexpr int U = STRIP_U, RING = 4, NT = DUAL ? 2 : 1;
    constexpr unsigned TILE_BYTES = S::TILE_WORDS * 4u;
    extern __shared__ __align__(16) unsigned strip_smem[];
    unsigned* const tiles = strip_smem;

    const HopItem& it = a.item[blockIdx.z];
    const Frame f = a.f;
    const StripPlan p = a.p;
    const int tid = threadIdx.x, lane = tid & 31;
    const int C0 = blockIdx.x * p.oc;
    const int oc = min(p.oc, f.ny - C0);
    const int R0 = blockIdx.y * p.seg_rows, R1 = min(R0 + p.seg_rows, f.nx);

    {
        uint4* z = reinterpret_cast<uint4*>(tiles);
        for (int k = tid; k < NT * S::TILE_WORDS / 4; k += STRIP_BLOCK) z[k] = make_uint4(0u, 0u, 0u, 0u);
    }
    // rays of [2^9, 2^(FIX+1)) units = [scale * 2^(9-FIX), 2 * scale) take the tiles
    const float fix_scale = (float)(1u << S::FIX) / a.intensity_scale, inv_scale = a.intensity_scale / (float)(1u << S::FIX);
    const unsigned vmin_bits = __float_as_uint(a.intensity_scale * (512.f / (float)(1u << S::FIX)));
    const unsigned vspan = __float_as_uint(2.f * a.intensity_scale) - vmin_bits;
    __syncthreads();

    if (tid >= STRIP_THREADS) {                                    // the flushing warp
        float* const outs[2] = {it.out_obj, it.out_ref};
        unsigned long long sum = strip_flusher<H, S::W, NT, ACC, DUAL>(tiles, outs, R0, R1, C0, oc, f.nx, f.ny, inv_scale, a.vec);
        if (DUAL && it.sum_ref) {
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(FULL_MASK, sum, d);
            if (lane == 0) atomicAdd(it.sum_ref, (double)sum * (double)inv_scale);
        }
        return;
    }

    const int j = C0 - H + tid;
    const bool live = j >= 0 && j < f.ny && tid < p.oc + 2 * H;
    const int jc = min(max(j, 0), f.ny - 1);
    const bool own_col = j >= C0 && j < C0 + oc;
    const unsigned wid = ((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * STRIP_WARPS + (tid >> 5);
    uint4* const slice = a.far + (size_t)wid * (p.far_cap * NT);
    unsigned n_far = 0u;
    // warp-uniform: no lane of this warp sits on the first / last image column or outside the image
    const bool inner_cols = __all_sync(FULL_MASK, j > 0 && j < f.ny - 1);
    const bool edge_lane = lane == 0 || lane == 31;
    const int dh = min(max(lane == 0 ? j - 1 : j + 1, 0), f.ny - 1) - jc;     // column next to the warp, for its edge lanes

    // Register ring of thickness rows: position k holds source row s_begin - 1 + k (mod 4); processing the row at
    // position k uses k (up), k+1 (this row), k+2 (down) and then refills k with the row 4 further on, two rows
    // before that one is first needed.  The ring is as deep as a chunk, so the positions are compile-time constants.
    const int s_begin = max(R0 - H, 0), s_end = min(R1 + H, f.nx);
    const int last_t = (f.nx - 1) * f.ny + jc;                     // nx * ny < 2^30 (host check)
    float row[RING][NM], hal[4][NM], inten[4];
    int pre_t = max(s_begin - 1, 0) * f.ny + jc;
#pragma unroll
    for (int k = 0; k < RING; ++k) {
#pragma unroll
        for (int m = 0; m < NM; ++m) strip_prefetch(row[k][m], it.map[m] + pre_t);
        if (k > 0 || s_begin > 0) pre_t = min(pre_t + f.ny, last_t);       // row -1 does not exist: position 0 repeats row 0
    }
    // the column next to the warp (edge lanes only), rows s_begin .. s_begin + 3, and the incoming intensity of the same rows
    int pre_i = s_begin * f.ny + jc;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int m = 0; m < NM; ++m) {
            hal[k][m] = 0.f;
            if (edge_lane) strip_prefetch(hal[k][m], it.map[m] + pre_i + dh);
        }
        inten[k] = a.I_uniform;
        if (HAS_I) strip_prefetch(inten[k], it.I_in + pre_i);
        pre_i = min(pre_i + f.ny, last_t);
    }

    const unsigned tile_s = (unsigned)__cvta_generic_to_shared(tiles);
    const ColWin cw = col_window<H>(j, C0, p.oc, f.ny, live, tile_s);
    const int in_lo = max(R0 + H - 1, H), in_hi = min(R1 - H, f.nx - 1 - H);
    const unsigned long long half = strip_half(f.nx);
    const float neg_log2e = -1.4426950408889634f;

    // a ray this block could not deposit although it owns the source pixel: tile ray of the neighbours, or one for the list
    auto maybe_push = [&](bool ok, bool own_row, unsigned flags, int i, unsigned bx, unsigned by, float v, float dx, float dy) {
        const bool cand = !ok && own_row && own_col && v != 0.f;
        if (__any_sync(FULL_MASK, cand)) {                          // warp-uniform; rare away from the strip's edges
            const bool push = cand && !(tile_class<H>(i, j, bx, by, f.nx, f.ny) && (__float_as_uint(v) - vmin_bits) < vspan);
            strip_push(slice, n_far, push, flags | (unsigned)(i * f.ny + j), v, dx, dy);
        }
    };

    // one source row; KC = its ring position (compile time), `interior` = full deposit window, rows strictly inside the image
    auto do_row = [&](auto KC, int i, bool interior) {
        constexpr int K = decltype(KC)::value;
        constexpr int kup = K & 3, kmid = (K + 1) & 3, kdn = (K + 2) & 3;
        const bool inner = inner_cols && (interior || (i > 0 && i < f.nx - 1));      // warp-uniform
        float dxo = 0.f, dyo = 0.f, dxr = 0.f, dyr = 0.f, arg = 0.f;
#pragma unroll
        for (int m = 0; m < NM; ++m) {
            const float mid = row[kmid][m], up = row[kup][m], dn = row[kdn][m];
            float lf = __shfl_up_sync(FULL_MASK, mid, 1), rt = __shfl_down_sync(FULL_MASK, mid, 1);
            if (lane == 0) lf = hal[K & 3][m];
            if (lane == 31) rt = hal[K & 3][m];
            float gx, gy;
            if (inner) {
                gy = rt - lf;
                gx = dn - up;
            } else {
                // np.gradient(edge_order=2) numerators times 2h (refractionFileNumba2.py:54)
                const float* t = it.map[m];
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
            arg = fmaf(a.att[m], mid, arg);
        }
        const float vin = inten[K & 3];
        // Sample.py:347; ex2.approx keeps ~2e-7 relative accuracy over the attenuation range
        const float vo = vin * ex2_fast(arg * neg_log2e);
        const RowWin rw = interior ? row_window_interior<H>(i) : row_window<H>(i, R0, R1, f.nx);
        const bool own_row = interior || (i >= R0 && i < R1);
        unsigned bx, by;
        if (!DUAL) {
            const bool ok = strip_deposit<S::W, false>(rw, cw, vo, dxo, dyo, fix_scale, vmin_bits, vspan, 0u, half, bx, by);
            maybe_push(ok, own_row, 0u, i, bx, by, vo, dxo, dyo);
        } else if (__all_sync(FULL_MASK, dxo == dxr && dyo == dyr && vo == vin)) {
            // outside the sample the two beams are the same ray: form it once, deposit it twice
            const bool ok = strip_deposit<S::W, true>(rw, cw, vo, dxo, dyo, fix_scale, vmin_bits, vspan, TILE_BYTES, half, bx, by);
            maybe_push(ok, own_row, FAR_TWIN, i, bx, by, vo, dxo, dyo);
        } else {
            const bool oko = strip_deposit<S::W, false>(rw, cw, vo, dxo, dyo, fix_scale, vmin_bits, vspan, 0u, half, bx, by);
            maybe_push(oko, own_row, 0u, i, bx, by, vo, dxo, dyo);
            const bool okr = strip_deposit<S::W, false>(rw, cw, vin, dxr, dyr, fix_scale, vmin_bits, vspan, TILE_BYTES, half, bx, by);
            maybe_push(okr, own_row, FAR_REF, i, bx, by, vin, dxr, dyr);
        }
        // refill: thickness row i + 3 into the slot of row i - 1; neighbour column and intensity of row i + 4 into those of row i
#pragma unroll
        for (int m = 0; m < NM; ++m) {
            strip_prefetch(row[kup][m], it.map[m] + pre_t);
            if (edge_lane) strip_prefetch(hal[K & 3][m], it.map[m] + pre_i + dh);
        }
        pre_t = min(pre_t + f.ny, last_t);
        if (HAS_I) strip_prefetch(inten[K & 3], it.I_in + pre_i);
        pre_i = min(pre_i + f.ny, last_t);
    };

    int k = 0;
    for (int s = s_begin; s < s_end; s += U, ++k) {
        if (k >= 2) named_sync(BAR_EMPTY, k & 1);                  // the slots this chunk deposits into have been flushed
        const bool interior = s >= in_lo && s + U - 1 <= in_hi;    // block-uniform
        if (interior || s + 0 < s_end) do_row(std::integral_constant<int, 0>{}, s + 0, interior);
        if (interior || s + 1 < s_end) do_row(std::integral_constant<int, 1>{}, s + 1, interior);
        if (interior || s + 2 < s_end) do_row(std::integral_constant<int, 2>{}, s + 2, interior);
        if (interior || s + 3 < s_end) do_row(std::integral_constant<int, 3>{}, s + 3, interior);
        named_arrive(BAR_FULL, k & 1);
    }
    if (lane == 0) a.far_count[wid] = n_far;
}

// The listed rays in fp32, cell by cell; cells outside the image are dropped, which is what zero-padding, scattering and
// cropping does (refractionFileNumba2.py:65-78).  One warp per list slice.
struct HopDrainArgs {
    float* out_obj[STRIP_MAX_BATCH];
    float* out_ref[STRIP_MAX_BATCH];
    double* sum_ref[STRIP_MAX_BATCH];
    const uint4* far;
    const unsigned* far_count;
    unsigned slice_cap, slices_per_item, n_slices;
    Frame f;
    int* flag;
};

__global__ void __launch_bounds__(128) refract_strip_drain_kernel(const HopDrainArgs a);

int dispatch_refract_strip(int n_layers, HopArgs& a, int n_items, bool dual, bool has_i, bool accumulate, int reach, void* work,
                           size_t work_bytes, cudaStream_t s);
size_t refract_strip_work_bytes(int nx, int ny, int n_layers, int n_items, bool dual, bool has_i, int reach);

}  // namespace paresis
