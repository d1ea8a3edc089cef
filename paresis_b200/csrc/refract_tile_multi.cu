// Object hop of SEVERAL energies of one detector bin in one pass (polychromatic spectra).
//
// Reference: Experiment.py:448-498 -- inside a detector bin every energy repeats setWaveRT + refraction on the
// SAME thickness maps and adds its images to the same accumulators (:482-483).  Only the scalars change:
// delta(E), beta(E), the incident intensity.  This kernel therefore forms the thickness gradients of a row once
// and loops over the energies of a group: each energy reads its own object-plane intensity I_bs[g], scales the
// gradients, and deposits its sample and reference rays into the SAME pair of shared-memory tiles
// (tile_common.cuh), which are zeroed and flushed once per group instead of once per energy.
//
// Fixed point: the unit is sum_g(intensity_g) / 2^19 and a ray of energy g takes the tile only below
// 2 x intensity_g, so a pixel contributes < 2^20 units over the whole group and a cell cannot overflow.
#include <string.h>

#include "tile_common.cuh"

namespace paresis {

constexpr int MULTI_MAX = PARESIS_MAX_GROUP;

struct MultiEnergy {
    float g_obj[PARESIS_MAX_LAYERS], g_ref[PARESIS_MAX_LAYERS], att[PARESIS_MAX_LAYERS];
    float* I_in;          // object-plane intensity of this energy; cleared behind the pass
    unsigned vmax_bits;   // rays at or above 2 x intensity_g bypass the tile
    double* sum_ref;      // += what the reference beam of this energy deposits inside the image (may be null)
};

struct MultiArgs {
    const float* map[PARESIS_MAX_LAYERS];
    MultiEnergy en[MULTI_MAX];
    int n_energies;
    float* out_obj;
    float* out_ref;
    Frame f;
    int rows;
    int* flag;
    float scale;          // fixed units per intensity unit
    float inv_scale;
};

template <int NM, int TR, int H>
__global__ void __launch_bounds__(TILE_COLS)
refract_tile_multi_kernel(const MultiArgs a) {
    constexpr int SR = TR + 2 * H + 1, SC = TILE_COLS + 2 * H;
    extern __shared__ __align__(16) unsigned tile_smem[];

    const Frame f = a.f;
    const int lane = threadIdx.x & 31;
    const int j = blockIdx.x * TILE_COLS + threadIdx.x;
    const int i0 = blockIdx.y * a.rows;
    const int i1 = min(i0 + a.rows, f.nx);
    const bool live = j < f.ny;
    const int jc = live ? j : f.ny - 1;
    const bool inner_cols = __all_sync(FULL_MASK, live && j > 0 && j < f.ny - 1);

    TileTarget tobj, tref;
    tobj.tile = tile_smem; tobj.out = a.out_obj;
    tobj.rlo = i0 - H; tobj.clo = blockIdx.x * TILE_COLS - H;
    tobj.r0 = max(tobj.rlo, 0); tobj.c0 = max(tobj.clo, 0);
    tobj.nr = (unsigned)max(min(tobj.rlo + SR - 1, f.nx - 1) - tobj.r0, 0);
    tobj.nc = (unsigned)max(min(tobj.clo + SC - 1, f.ny - 1) - tobj.c0, 0);
    tref = tobj; tref.tile = tile_smem + SR * SC; tref.out = a.out_ref;
    {
        uint4* z = reinterpret_cast<uint4*>(tile_smem);
        for (int k = threadIdx.x; k < SR * SC / 4 * 2; k += TILE_COLS) z[k] = make_uint4(0u, 0u, 0u, 0u);
    }

    constexpr int RING = 4;     // see refract_tile_kernel
    float row[RING][NM], hal[RING][NM];
    const bool edge_lane = lane == 0 || lane == 31;
    const int jh = min(max(lane == 0 ? jc - 1 : jc + 1, 0), f.ny - 1);
    int off = i0 * f.ny + jc;
    int offh = i0 * f.ny + jh;
    const int last = (f.nx - 1) * f.ny;
#pragma unroll
    for (int k = 0; k < RING - 1; ++k) {
        const int d = (k - 1) * f.ny;
#pragma unroll
        for (int m = 0; m < NM; ++m) {
            row[k][m] = __ldg(a.map[m] + min(max(off + d, jc), last + jc));
            hal[k][m] = edge_lane ? __ldg(a.map[m] + min(max(offh + d, jh), last + jh)) : 0.f;
        }
    }
    const float neg_log2e = -1.4426950408889634f;
    bool bad = false;
    float ref_sum[MULTI_MAX];
#pragma unroll
    for (int g = 0; g < MULTI_MAX; ++g) ref_sum[g] = 0.f;
    __syncthreads();

    for (int ib = i0; ib < i1; ib += RING) {
#pragma unroll
        for (int s = 0; s < RING; ++s) {
            const int i = ib + s;
            if (i >= i1) break;
            const int kup = s % RING, kmid = (s + 1) % RING, kdn = (s + 2) % RING, knew = (s + 3) % RING;
#pragma unroll
            for (int m = 0; m < NM; ++m) {
                row[knew][m] = __ldg(a.map[m] + min(off + 2 * f.ny, last + jc));
                hal[knew][m] = edge_lane ? __ldg(a.map[m] + min(offh + 2 * f.ny, last + jh)) : 0.f;
            }
            float v_next = a.en[0].I_in[off];      // plain loads: the same thread clears the address afterwards
            const bool inner = inner_cols && i > 0 && i < f.nx - 1;
            float gx[NM], gy[NM], mid[NM];
#pragma unroll
            for (int m = 0; m < NM; ++m) {
                const float* t = a.map[m];
                mid[m] = row[kmid][m];
                const float up = row[kup][m], dn = row[kdn][m];
                float lf = __shfl_up_sync(FULL_MASK, mid[m], 1);
                float rt = __shfl_down_sync(FULL_MASK, mid[m], 1);
                if (lane == 0) lf = hal[kmid][m];
                if (lane == 31) rt = hal[kmid][m];
                if (inner) {
                    gy[m] = rt - lf;
                    gx[m] = dn - up;
                } else {
                    // np.gradient(edge_order=2) numerators times 2h (refractionFileNumba2.py:54)
                    const float* r = t + (size_t)i * f.ny;
                    if (jc == 0) gy[m] = -3.f * mid[m] + 4.f * rt - __ldg(r + 2);
                    else if (jc == f.ny - 1) gy[m] = 3.f * mid[m] - 4.f * lf + __ldg(r + f.ny - 3);
                    else gy[m] = rt - lf;
                    if (i == 0) gx[m] = -3.f * mid[m] + 4.f * dn - __ldg(t + (size_t)2 * f.ny + jc);
                    else if (i == f.nx - 1) gx[m] = 3.f * mid[m] - 4.f * up + __ldg(t + (size_t)(f.nx - 3) * f.ny + jc);
                    else gx[m] = dn - up;
                }
            }
#pragma unroll 1
            for (int g = 0; g < a.n_energies; ++g) {
                const MultiEnergy& e = a.en[g];
                const float vin = v_next;
                if (g + 1 < a.n_energies) v_next = a.en[g + 1].I_in[off];
                float dxo = 0.f, dyo = 0.f, dxr = 0.f, dyr = 0.f, arg = 0.f;
#pragma unroll
                for (int m = 0; m < NM; ++m) {
                    dxo = fmaf(e.g_obj[m], gx[m], dxo); dyo = fmaf(e.g_obj[m], gy[m], dyo);
                    dxr = fmaf(e.g_ref[m], gx[m], dxr); dyr = fmaf(e.g_ref[m], gy[m], dyr);
                    arg = fmaf(e.att[m], mid[m], arg);
                }
                const float vo = vin * ex2_fast(arg * neg_log2e);                 // Sample.py:347
                if (live) e.I_in[off] = 0.f;
                float s_ref;
                const bool same = dxo == dxr && dyo == dyr && vo == vin;
                if (__all_sync(FULL_MASK, same)) {
                    s_ref = deposit<SC, true>(tobj, i, j, vo, dxo, dyo, f.nx, f.ny, a.scale, e.vmax_bits, live, bad, a.out_ref, SR * SC);
                } else {
                    deposit<SC, false>(tobj, i, j, vo, dxo, dyo, f.nx, f.ny, a.scale, e.vmax_bits, live, bad);
                    s_ref = deposit<SC, false>(tref, i, j, vin, dxr, dyr, f.nx, f.ny, a.scale, e.vmax_bits, live, bad);
                }
#pragma unroll
                for (int gg = 0; gg < MULTI_MAX; ++gg)
                    if (gg == g) ref_sum[gg] += s_ref;
            }
            off += f.ny; offh += f.ny;
        }
    }
    if (bad && a.flag) atomicOr(a.flag, FLAG_NONFINITE);
#pragma unroll
    for (int g = 0; g < MULTI_MAX; ++g) {
        if (g < a.n_energies && a.en[g].sum_ref) {
            float v = ref_sum[g];
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL_MASK, v, d);
            if (lane == 0) atomicAdd(a.en[g].sum_ref, (double)v);
        }
    }
    __syncthreads();
    const bool vec_ok = (f.ny & 3) == 0;
    flush_tile<SR, SC>(tobj.tile, a.out_obj, tobj.rlo, tobj.clo, f.nx, f.ny, a.inv_scale, vec_ok && (reinterpret_cast<uintptr_t>(a.out_obj) & 15) == 0);
    flush_tile<SR, SC>(tref.tile, a.out_ref, tobj.rlo, tobj.clo, f.nx, f.ny, a.inv_scale, vec_ok && (reinterpret_cast<uintptr_t>(a.out_ref) & 15) == 0);
}

template <int NM>
static int launch_multi(MultiArgs a, cudaStream_t s) {
    constexpr int TR = 16, H = 4;
    constexpr int SR = TR + 2 * H + 1, SC = TILE_COLS + 2 * H;
    constexpr size_t smem = sizeof(unsigned) * SR * SC * 2;
    static int slots_of[32] = {0};             // resident blocks x SMs, per device
    int dev = 0;
    PARESIS_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 32) dev = 0;
    if (!slots_of[dev]) {
        PARESIS_CUDA(cudaFuncSetAttribute(refract_tile_multi_kernel<NM, TR, H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0, sms = 0;
        PARESIS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, refract_tile_multi_kernel<NM, TR, H>, TILE_COLS, smem));
        PARESIS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        slots_of[dev] = (per_sm > 0 ? per_sm : 1) * (sms > 0 ? sms : 148);
    }
    const int slots = slots_of[dev];
    const int strips = div_up(a.f.ny, TILE_COLS);
    a.rows = pick_tile_rows(a.f.nx, strips, slots, TR);
    dim3 grid(strips, div_up(a.f.nx, a.rows));
    refract_tile_multi_kernel<NM, TR, H><<<grid, TILE_COLS, smem, s>>>(a);
    PARESIS_LAUNCH_CHECK("refract_tile_multi_kernel");
    return PARESIS_OK;
}

}  // namespace paresis

using namespace paresis;

extern "C" int paresis_refract_group(const paresis_group_energy* energies_host, int n_energies,
                                     float* out_obj, float* out_ref, int nx, int ny, int* flag, paresis_stream stream) {
    if (!energies_host || n_energies < 1 || n_energies > PARESIS_MAX_GROUP || !out_obj || !out_ref || nx < 3 || ny < 3 ||
        (long)nx * ny >= (1L << 30)) {
        set_last_error("paresis_refract_group: need 1..%d energies, two outputs and 3 <= n, nx*ny < 2^30", PARESIS_MAX_GROUP);
        return PARESIS_ERR_ARG;
    }
    MultiArgs a{};
    const int nl = energies_host[0].n_layers;
    if (nl < 1 || nl > PARESIS_MAX_LAYERS) { set_last_error("paresis_refract_group: 1..%d layers", PARESIS_MAX_LAYERS); return PARESIS_ERR_ARG; }
    double total = 0.0;
    for (int g = 0; g < n_energies; ++g) {
        const paresis_group_energy& e = energies_host[g];
        if (e.n_layers != nl || !e.intensity_in || !(e.intensity_scale > 0.f)) {
            set_last_error("paresis_refract_group: energy %d needs the same layer count, an intensity image and a positive scale", g);
            return PARESIS_ERR_ARG;
        }
        for (int m = 0; m < nl; ++m) {
            if (!e.layers[m].thickness || (g > 0 && e.layers[m].thickness != energies_host[0].layers[m].thickness)) {
                set_last_error("paresis_refract_group: the energies of a group must share their thickness maps");
                return PARESIS_ERR_ARG;
            }
            a.en[g].g_obj[m] = e.layers[m].grad_obj;
            a.en[g].g_ref[m] = e.layers[m].grad_ref;
            a.en[g].att[m] = e.layers[m].atten;
        }
        a.en[g].I_in = e.intensity_in;
        a.en[g].sum_ref = e.sum_ref;
        total += e.intensity_scale;
    }
    for (int m = 0; m < nl; ++m) a.map[m] = energies_host[0].layers[m].thickness;
    const float unit_scale = (float)((double)(1u << FIX_BITS) / total);      // fixed units per intensity unit
    for (int g = 0; g < n_energies; ++g) {
        const float vmax = 2.0f * energies_host[g].intensity_scale * (1.0f - 1.0f / 65536.0f);
        unsigned bits;
        memcpy(&bits, &vmax, sizeof bits);
        a.en[g].vmax_bits = bits;
    }
    a.n_energies = n_energies;
    a.out_obj = out_obj;
    a.out_ref = out_ref;
    a.f = Frame{nx, ny, 15};
    a.flag = flag;
    a.scale = unit_scale;
    a.inv_scale = (float)(total / (double)(1u << FIX_BITS));
    cudaStream_t s = (cudaStream_t)stream;
    switch (nl) {
        case 1: return launch_multi<1>(a, s);
        case 2: return launch_multi<2>(a, s);
        case 3: return launch_multi<3>(a, s);
        default: return launch_multi<4>(a, s);
    }
}
