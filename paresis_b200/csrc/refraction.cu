// Ray-tracing refraction model: the bilinear splat and its fused producers.
//
// Reference: refractionFileNumba2.py:25-86 (fastRefraction), :198-263 (fastloopNumba),
// Sample.py:285-351 (setWaveRT), Experiment.py:463-474 (the per-energy hop sequence).
//
// Launch shape (all kernels here): a block is 8 warps side by side, each warp owns a strip of
// 32 consecutive source columns and walks `rows` consecutive source rows, one row per step.
// Loads are unit-stride 128 B per warp and issued one row ahead; deposits leave as REDG.ADD.F32
// after the warp/register aggregation described in splat.cuh.  HBM-bound integer/fp32 work:
// no tensor cores (there is no contraction on this path).
#include <math.h>

#include "refract_common.cuh"

namespace paresis {

constexpr int BLOCK_WARPS = 8;
constexpr int BLOCK_THREADS = 32 * BLOCK_WARPS;

// ---------------------------------------------------------------------------------------------
// fastloopNumba: scatter given I, Dx, Dy
// ---------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(BLOCK_THREADS)
splat_kernel(const float* __restrict__ I, const float* __restrict__ Dx, const float* __restrict__ Dy,
             float* __restrict__ out, Frame f, int rows, int* flag) {
    const int j = blockIdx.x * BLOCK_THREADS + threadIdx.x;
    const int i0 = blockIdx.y * rows;
    const int i1 = min(i0 + rows, f.nx);
    const bool live = j < f.ny;
    const bool full_warp = __all_sync(FULL_MASK, live);
    Splatter<MODE> sp;
    sp.init(out, f.ny, flag);
    constexpr int U = 4;
    const size_t col = live ? j : 0;
    for (int ib = i0; ib < i1; ib += U) {
        float v[U], dx[U], dy[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = min(ib + u, i1 - 1);
            const size_t p = (size_t)i * f.ny + col;
            v[u] = __ldg(I + p);
            dx[u] = __ldg(Dx + p);
            dy[u] = __ldg(Dy + p);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = ib + u;
            if (i < i1) {  // warp-uniform
                const FastRay q = fast_ray(i, j, v[u], dx[u], dy[u], f.nx, f.ny);
                if (full_warp && __all_sync(FULL_MASK, q.simple)) sp.put_simple(q);
                else sp.put(live ? make_ray(i, j, v[u], dx[u], dy[u], f) : empty_ray());
            }
        }
    }
    sp.finish();
}

// ---------------------------------------------------------------------------------------------
// Fused: thickness (or phase) maps -> transmission -> gradient -> displacement -> scatter
// ---------------------------------------------------------------------------------------------
template <typename T, int NM, bool DUAL, bool HAS_I, bool ATT, bool WRITE_D, int MODE>
__global__ void __launch_bounds__(BLOCK_THREADS)
refract_kernel(const RefractArgs<T> a) {
    const Frame f = a.f;
    const int lane = threadIdx.x & 31;
    const int j = blockIdx.x * BLOCK_THREADS + threadIdx.x;
    const int i0 = blockIdx.y * a.rows;
    const int i1 = min(i0 + a.rows, f.nx);
    const bool live = j < f.ny;
    const int jc = live ? j : f.ny - 1;  // dead lanes read a valid address, contribute nothing
    // the strip is "inner" when no lane sits on the first / last image column and every lane is live
    const bool inner_cols = __all_sync(FULL_MASK, live && j > 0 && j < f.ny - 1);

    // rolling rows: up = row i-1, mid = row i, dn = row i+1 (difference first, scale after)
    T up[NM], mid[NM], dn[NM];
    const T* row[NM];    // points at (row i+2, column jc): the next row to fetch
    const T* halo[NM];   // points at (row i, column jc -+ 1) for the first / last lane of the strip
    const bool edge_lane = lane == 0 || lane == 31;
    const int jh = min(max(lane == 0 ? jc - 1 : jc + 1, 0), f.ny - 1);
#pragma unroll
    for (int m = 0; m < NM; ++m) {
        const T* t = a.map[m] + jc;
        mid[m] = __ldg(t + (size_t)i0 * f.ny);
        up[m] = i0 > 0 ? __ldg(t + (size_t)(i0 - 1) * f.ny) : mid[m];
        dn[m] = i0 + 1 < f.nx ? __ldg(t + (size_t)(i0 + 1) * f.ny) : mid[m];
        row[m] = t + (size_t)(i0 + 2) * f.ny;
        halo[m] = a.map[m] + (size_t)i0 * f.ny + jh;
    }
    const float* irow = HAS_I ? a.I_in + (size_t)i0 * f.ny + jc : nullptr;
    // plain loads: with clear_input the same thread stores to this address after reading it
    float vin = HAS_I ? *irow : a.I_uniform;

    Splatter<MODE> sp_obj, sp_ref;
    sp_obj.init(a.out_obj, f.ny, a.flag);
    if (DUAL) sp_ref.init(a.out_ref, f.ny, a.flag);
    float ref_sum = 0.f;
    if (a.zero_scalar && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *a.zero_scalar = 0.0;

    for (int i = i0; i < i1; ++i) {
        // fetch the row after next and the next intensity before touching this row
        T nxt[NM];
        const bool more = i + 2 < f.nx && i + 1 < i1;
#pragma unroll
        for (int m = 0; m < NM; ++m) {
            nxt[m] = more ? __ldg(row[m]) : dn[m];
            row[m] += f.ny;
        }
        float vnext = vin;
        if (HAS_I && i + 1 < i1) { irow += f.ny; vnext = *irow; }

        const bool inner = inner_cols && i > 0 && i < f.nx - 1;   // warp-uniform
        float dxo = 0.f, dyo = 0.f, dxr = 0.f, dyr = 0.f, arg = 0.f;
#pragma unroll
        for (int m = 0; m < NM; ++m) {
            const T* t = a.map[m];
            // neighbours along the row come from the adjacent lanes; the strip edges load them
            T lf = __shfl_up_sync(FULL_MASK, mid[m], 1);
            T rt = __shfl_down_sync(FULL_MASK, mid[m], 1);
            if (edge_lane) {   // one predicated load serves both ends of the strip
                const T h = __ldg(halo[m]);
                if (lane == 0) lf = h; else rt = h;
            }
            halo[m] += f.ny;
            T gy, gx;
            if (inner) {
                gy = rt - lf;
                gx = dn[m] - up[m];
            } else {
                // np.gradient(edge_order=2) numerators times 2h (refractionFileNumba2.py:54)
                if (jc == 0) gy = -(T)3 * mid[m] + (T)4 * rt - __ldg(t + (size_t)i * f.ny + 2);
                else if (jc == f.ny - 1) gy = (T)3 * mid[m] - (T)4 * lf + __ldg(t + (size_t)i * f.ny + f.ny - 3);
                else gy = rt - lf;
                if (i == 0) gx = -(T)3 * mid[m] + (T)4 * dn[m] - __ldg(t + (size_t)2 * f.ny + jc);
                else if (i == f.nx - 1) gx = (T)3 * mid[m] - (T)4 * up[m] + __ldg(t + (size_t)(f.nx - 3) * f.ny + jc);
                else gx = dn[m] - up[m];
            }
            const float gxf = (float)gx, gyf = (float)gy;
            dxo = fmaf(a.g_obj[m], gxf, dxo);
            dyo = fmaf(a.g_obj[m], gyf, dyo);
            if (DUAL) {
                dxr = fmaf(a.g_ref[m], gxf, dxr);
                dyr = fmaf(a.g_ref[m], gyf, dyr);
            }
            if (ATT) arg = fmaf(a.att[m], (float)mid[m], arg);
        }
        float vo = ATT ? vin * expf(-arg) : vin;   // Sample.py:347
        float vr = vin;
        if (live) {
            const size_t p = (size_t)i * f.ny + j;
            if (a.zero[0]) a.zero[0][p] = 0.f;
            if (a.zero[1]) a.zero[1][p] = 0.f;
            if (a.zero[2]) a.zero[2][p] = 0.f;
            if (HAS_I && a.clear_input) const_cast<float*>(a.I_in)[p] = 0.f;
        }
        if (WRITE_D) {
            clean(vo, dxo, dyo, a.clamp_x, a.clamp_y);
            if (live) {
                const size_t pp = (size_t)(i + f.margin) * (f.ny + 2 * f.margin) + (j + f.margin);
                a.dx_pad[pp] = dxo;
                a.dy_pad[pp] = dyo;
            }
        }
        {
            // A `simple` ray lands strictly inside the image, so |D| < N: the kill rule of
            // refractionFileNumba2.py:61-64 cannot fire, and the |D| < 1e-12 -> 0 rule (:59-60) changes
            // nothing at fp32 resolution.  Everything else is cleaned and takes the reference's edge rules.
            const FastRay q = fast_ray(i, j, vo, dxo, dyo, f.nx, f.ny);
            if (inner_cols && __all_sync(FULL_MASK, q.simple)) sp_obj.put_simple(q);
            else {
                if (!WRITE_D) clean(vo, dxo, dyo, a.clamp_x, a.clamp_y);
                sp_obj.put(live ? make_ray(i, j, vo, dxo, dyo, f) : empty_ray());
            }
        }
        if (DUAL) {
            const FastRay q = fast_ray(i, j, vr, dxr, dyr, f.nx, f.ny);
            if (inner_cols && __all_sync(FULL_MASK, q.simple)) {
                sp_ref.put_simple(q);
                ref_sum += vr;
            } else {
                clean(vr, dxr, dyr, a.clamp_x, a.clamp_y);
                const Ray qr = live ? make_ray(i, j, vr, dxr, dyr, f) : empty_ray();
                sp_ref.put(qr);
#pragma unroll
                for (int k = 0; k < 4; ++k) ref_sum += ((qr.ok >> k) & 1u) ? qr.w[k] : 0.f;
            }
        }
#pragma unroll
        for (int m = 0; m < NM; ++m) { up[m] = mid[m]; mid[m] = dn[m]; dn[m] = nxt[m]; }
        vin = vnext;
    }
    sp_obj.finish();
    if (DUAL) {
        sp_ref.finish();
        if (a.sum_ref) {   // one double atomic per warp
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) ref_sum += __shfl_xor_sync(FULL_MASK, ref_sum, d);
            if (lane == 0) atomicAdd(a.sum_ref, (double)ref_sum);
        }
    }
}

// Tuning knobs (paresis_set_tuning): deposit mode of the fused kernels, rows per warp.
static int g_fused_mode = 2;
static int g_rows_override = 0;
// which fused hop kernel runs: 0 = fixed-point shared-memory tiles (refract_lean.cu; production, needs
// paresis_refract_extras.intensity_scale), -1 = one column per thread straight to L2 (this file; what runs for everything
// else: displacement output, phase input, callers that give no intensity scale)
static int g_tile_config = 0;
int g_lean_rows_override = 0;      // paresis_set_tuning(3, rows)

// rows per warp: enough blocks for ~2 waves of 148 SMs x 8 resident blocks, few halo re-reads
static int pick_rows(int nx, int ny) {
    const int strips = div_up(ny, BLOCK_THREADS);
    const long target_blocks = 148L * 8 * 2;
    int rows = (int)((long)nx * strips / target_blocks);
    if (rows < 8) rows = 8;
    if (rows > 64) rows = 64;
    if (g_rows_override > 0) rows = g_rows_override;
    return rows;
}

static int check_frame(int nx, int ny, int margin) {
    if (nx < 3 || ny < 3 || margin < 0 || (long)nx * ny >= (1L << 30)) {
        set_last_error("bad frame %d x %d (margin %d): need 3 <= n and nx*ny < 2^30", nx, ny, margin);
        return PARESIS_ERR_ARG;
    }
    return PARESIS_OK;
}

template <typename T, int NM, bool DUAL, bool HAS_I, bool ATT>
static int launch_refract(const RefractArgs<T>& a, bool write_d, cudaStream_t s) {
    dim3 grid(div_up(a.f.ny, BLOCK_THREADS), div_up(a.f.nx, a.rows));
    if (write_d) refract_kernel<T, NM, DUAL, HAS_I, ATT, true, 2><<<grid, BLOCK_THREADS, 0, s>>>(a);
    else if (g_fused_mode == 0) refract_kernel<T, NM, DUAL, HAS_I, ATT, false, 0><<<grid, BLOCK_THREADS, 0, s>>>(a);
    else refract_kernel<T, NM, DUAL, HAS_I, ATT, false, 2><<<grid, BLOCK_THREADS, 0, s>>>(a);
    PARESIS_LAUNCH_CHECK("refract_kernel");
    return PARESIS_OK;
}

template <int NM>
static int dispatch_layers(const RefractArgs<float>& a, cudaStream_t s) {
    const bool dual = a.out_ref != nullptr, has_i = a.I_in != nullptr;
    const bool wd = a.dx_pad != nullptr;
    if (!wd && g_tile_config == 0 && a.intensity_scale > 0.f) return dispatch_refract_lean(NM, a, s);
    if (dual) return has_i ? launch_refract<float, NM, true, true, true>(a, false, s)
                           : launch_refract<float, NM, true, false, true>(a, false, s);
    return has_i ? launch_refract<float, NM, false, true, true>(a, wd, s)
                 : launch_refract<float, NM, false, false, true>(a, wd, s);
}

}  // namespace paresis

using namespace paresis;

extern "C" int paresis_set_tuning(int key, int value) {
    switch (key) {
        case 0: g_fused_mode = value == 0 ? 0 : 2; return PARESIS_OK;
        case 1: g_rows_override = value; return PARESIS_OK;
        case 2: g_tile_config = value < 0 ? -1 : 0; return PARESIS_OK;
        case 3: paresis::g_lean_rows_override = value > 0 ? value : 0; return PARESIS_OK;
        default: set_last_error("paresis_set_tuning: unknown key %d", key); return PARESIS_ERR_ARG;
    }
}

extern "C" int paresis_splat(const float* intensity, const float* dx, const float* dy, float* out,
                             int nx, int ny, int margin, int variant, int* flag, paresis_stream stream) {
    if (!intensity || !dx || !dy || !out) { set_last_error("paresis_splat: null pointer"); return PARESIS_ERR_ARG; }
    if (nx < 1 || ny < 1 || margin < 0 || (long)nx * ny >= (1L << 30)) {
        set_last_error("paresis_splat: bad frame %d x %d margin %d", nx, ny, margin);
        return PARESIS_ERR_ARG;
    }
    Frame f{nx, ny, margin};
    const int rows = pick_rows(nx, ny);
    dim3 grid(div_up(ny, BLOCK_THREADS), div_up(nx, rows));
    cudaStream_t s = (cudaStream_t)stream;
    switch (variant) {
        case 0: splat_kernel<0><<<grid, BLOCK_THREADS, 0, s>>>(intensity, dx, dy, out, f, rows, flag); break;
        case 1: splat_kernel<1><<<grid, BLOCK_THREADS, 0, s>>>(intensity, dx, dy, out, f, rows, flag); break;
        case 2: splat_kernel<2><<<grid, BLOCK_THREADS, 0, s>>>(intensity, dx, dy, out, f, rows, flag); break;
        case 3: return launch_splat_tile(intensity, dx, dy, out, f, flag, s);
        case 4: return launch_splat_strip(intensity, dx, dy, out, f, flag, false, s);    // out  = splat (owner computes, plain stores)
        case 5: return launch_splat_strip(intensity, dx, dy, out, f, flag, true, s);     // out += splat (owner adds its rows)
        default: set_last_error("paresis_splat: unknown variant %d", variant); return PARESIS_ERR_ARG;
    }
    PARESIS_LAUNCH_CHECK("splat_kernel");
    return PARESIS_OK;
}

extern "C" int paresis_refract_phi(const float* intensity, const double* phi, float* out,
                                   float* dx_pad, float* dy_pad, int nx, int ny, int margin,
                                   double distance_m, double energy_kev, double magnification, double pixel_um,
                                   double clamp_px, int* flag, paresis_stream stream) {
    if (!intensity || !phi || !out || ((dx_pad == nullptr) != (dy_pad == nullptr))) {
        set_last_error("paresis_refract_phi: null pointer");
        return PARESIS_ERR_ARG;
    }
    int rc = check_frame(nx, ny, margin);
    if (rc) return rc;
    // refractionFileNumba2.py:47-48, :55-56 -- D = dphi * z / k / (h * M), dphi = numerator / (2h)
    const double lambda = 6.626 * 1e-34 * 2.998e8 / (energy_kev * 1000 * 1.6e-19);
    const double k = 2 * M_PI / lambda;
    const double h = pixel_um * 1e-6;
    RefractArgs<double> a{};
    a.map[0] = phi;
    a.g_obj[0] = (float)(distance_m / k / (h * magnification) / (2.0 * h));
    a.I_in = intensity;
    a.out_obj = out;
    a.dx_pad = dx_pad;
    a.dy_pad = dy_pad;
    a.f = Frame{nx, ny, margin};
    a.clamp_x = clamp_px > 0 ? (float)clamp_px : (float)nx;
    a.clamp_y = clamp_px > 0 ? (float)clamp_px : (float)ny;
    a.rows = pick_rows(nx, ny);
    a.flag = flag;
    return launch_refract<double, 1, false, true, false>(a, dx_pad != nullptr, (cudaStream_t)stream);
}

extern "C" int paresis_refract_layers(const float* intensity_in, float intensity_uniform,
                                      const paresis_layer* layers_host, int n_layers,
                                      float* out_obj, float* out_ref, float* dx_pad, float* dy_pad,
                                      int nx, int ny, int margin, int* flag, paresis_stream stream) {
    return paresis_refract_layers_ex(intensity_in, intensity_uniform, layers_host, n_layers, out_obj, out_ref, dx_pad, dy_pad,
                                     nx, ny, margin, flag, nullptr, stream);
}

extern "C" int paresis_refract_layers_ex(const float* intensity_in, float intensity_uniform,
                                         const paresis_layer* layers_host, int n_layers,
                                         float* out_obj, float* out_ref, float* dx_pad, float* dy_pad,
                                         int nx, int ny, int margin, int* flag,
                                         const paresis_refract_extras* extras, paresis_stream stream) {
    if (!layers_host || !out_obj || n_layers < 1 || n_layers > PARESIS_MAX_LAYERS) {
        set_last_error("paresis_refract_layers: need 1..%d layers and an output", PARESIS_MAX_LAYERS);
        return PARESIS_ERR_ARG;
    }
    if (((dx_pad == nullptr) != (dy_pad == nullptr)) || (dx_pad && out_ref)) {
        set_last_error("paresis_refract_layers: dx_pad/dy_pad come together and only without out_ref");
        return PARESIS_ERR_ARG;
    }
    int rc = check_frame(nx, ny, margin);
    if (rc) return rc;
    RefractArgs<float> a{};
    for (int m = 0; m < n_layers; ++m) {
        if (!layers_host[m].thickness) { set_last_error("paresis_refract_layers: null map %d", m); return PARESIS_ERR_ARG; }
        a.map[m] = layers_host[m].thickness;
        a.g_obj[m] = layers_host[m].grad_obj;
        a.g_ref[m] = layers_host[m].grad_ref;
        a.att[m] = layers_host[m].atten;
    }
    a.I_in = intensity_in;
    a.I_uniform = intensity_uniform;
    a.out_obj = out_obj;
    a.out_ref = out_ref;
    a.dx_pad = dx_pad;
    a.dy_pad = dy_pad;
    a.f = Frame{nx, ny, margin};
    a.clamp_x = (float)nx;
    a.clamp_y = (float)ny;
    a.rows = pick_rows(nx, ny);
    a.flag = flag;
    cudaStream_t s = (cudaStream_t)stream;
    if (extras && extras->mode != 0) {
        // owner-computes rolling strips (refract_strip.cu): out = / out += without atomics on the image
        if (dx_pad) { set_last_error("paresis_refract_layers: the displacement maps come from mode 0 only"); return PARESIS_ERR_ARG; }
        paresis_hop_item item{};
        item.intensity_in = intensity_in;
        for (int m = 0; m < n_layers; ++m) item.thickness[m] = layers_host[m].thickness;
        item.out_obj = out_obj;
        item.out_ref = out_ref;
        item.sum_ref = out_ref ? extras->sum_ref : nullptr;
        return paresis_refract_hop_batch(&item, 1, layers_host, n_layers, intensity_uniform, extras->intensity_scale, extras->mode == 2,
                                         extras->reach > 0 ? extras->reach : 12, nx, ny, nullptr, 0, flag, stream);
    }
    if (extras) {
        // compact the zero-fill list; kernels that test only slot 0 get every slot filled (a repeated store is harmless)
        int nz = 0;
        for (int k = 0; k < 3; ++k) if (extras->zero_fill[k]) a.zero[nz++] = extras->zero_fill[k];
        for (int k = nz; k < 3 && nz > 0; ++k) a.zero[k] = a.zero[0];
        a.clear_input = extras->clear_input != 0 && intensity_in != nullptr;
        a.sum_ref = out_ref ? extras->sum_ref : nullptr;
        a.zero_scalar = extras->zero_scalar;
        a.intensity_scale = extras->intensity_scale > 0.f && extras->intensity_scale < 3.0e38f ? extras->intensity_scale : 0.f;
        a.full_tiles = extras->throughput != 0;
    }
    switch (n_layers) {
        case 1: return dispatch_layers<1>(a, s);
        case 2: return dispatch_layers<2>(a, s);
        case 3: return dispatch_layers<3>(a, s);
        default: return dispatch_layers<4>(a, s);
    }
}
