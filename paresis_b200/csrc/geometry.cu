// Projected-thickness maps: membrane of spherical grains, sphere and rotated-cylinder samples.
//
// Reference: Samples/getMembraneFromFile.py:127-168 (pure-Python triple loop, the dominant cost
// per membrane position in the reference), Samples/createSampGeom.py:41-53 and :87-106.
// Geometry is evaluated in fp64 (r^2 - d^2 cancels at the cap edge) and stored as fp32 metres.
#include <math.h>

#include "common.cuh"

namespace paresis {

constexpr int MAX_MEMBRANE_LAYERS = 16;

struct LayerOffsets {
    long long x[MAX_MEMBRANE_LAYERS];
    long long y[MAX_MEMBRANE_LAYERS];
};

// Two kernels, no atomics on the image.
//   bin:    one thread per (layer, sphere) candidate tests the reference's acceptance window (:151)
//           and the field of view; a survivor (~3 %) appends its id to the list of every 32 x 32 pixel
//           tile its bounding box (:155-159) touches.
//   gather: one block per tile; the tile's spheres are staged in shared memory and every thread sums
//           the caps over its own four pixels in registers (the cancelling r^2 - d^2 stays in fp64),
//           then writes them with plain coalesced stores.  The map is written exactly once: no
//           zero-fill pass, no read-modify-write.
// A tile whose list overflows RASTER_CAP falls back to scanning all candidates itself (correct,
// slow, and not reachable with physical membranes: ~6 spheres per tile at the bundled density).
constexpr int RASTER_THREADS = 256;
constexpr int RASTER_TILE = 32;
constexpr int RASTER_CAP = 256;

struct Cap {
    double xf, yf, r2;       // centre (canvas coordinates) and squared radius, pixels
    int rlo, rhi, clo, chi;  // bounding box in image coordinates, inclusive (:155-159)
};

// Sphere `s` of layer `layer`: acceptance test and bounding box.  Returns false when it adds nothing.
__device__ __forceinline__ bool make_cap(const double* __restrict__ spheres, int s, int layer, double pix, const LayerOffsets& off,
                                         int dim_x, int dim_y, int margin, Cap& c) {
    const double rad = spheres[3 * s + 2] / pix;
    c.xf = spheres[3 * s + 1] / pix - (double)off.x[layer];
    c.yf = spheres[3 * s + 0] / pix - (double)off.y[layer];
    c.r2 = rad * rad;
    const long long x = __double2ll_rn(c.xf), y = __double2ll_rn(c.yf);   // np.round: half to even (:149-150)
    const long long reach = (long long)floor(rad) + 1;
    const int margin2 = margin / 2;
    const bool accepted = margin2 < x && x < dim_x + margin + margin2 && margin2 < y && y < dim_y + margin + margin2;
    const bool visible = x + reach > margin && x - reach < dim_x + margin && y + reach > margin && y - reach < dim_y + margin;
    if (!(accepted && visible) || reach > 500) return false;   // (a 500-pixel grain would already exceed the canvas margin)
    // canvas rows x-reach .. x+reach-1 (:155-159), cropped by `margin` (:161)
    c.rlo = (int)(x - reach) - margin; c.rhi = (int)(x + reach) - 1 - margin;
    c.clo = (int)(y - reach) - margin; c.chi = (int)(y + reach) - 1 - margin;
    return true;
}

__global__ void __launch_bounds__(RASTER_THREADS)
raster_bin_kernel(const double* __restrict__ spheres, int n, double pix, LayerOffsets off, int n_layers, int dim_x, int dim_y,
                  int margin, int tiles_y, int* __restrict__ count, int* __restrict__ entries, int* __restrict__ n_caps,
                  Cap* __restrict__ caps) {
    const long long id = (long long)blockIdx.x * RASTER_THREADS + threadIdx.x;
    if (id >= (long long)n * n_layers) return;
    Cap c;
    if (!make_cap(spheres, (int)(id % n), (int)(id / n), pix, off, dim_x, dim_y, margin, c)) return;
    const int slot_c = atomicAdd(n_caps, 1);     // < n * n_layers = capacity of `caps`
    caps[slot_c] = c;
    const int tr0 = max(c.rlo, 0) / RASTER_TILE, tr1 = min(c.rhi, dim_x - 1) / RASTER_TILE;
    const int tc0 = max(c.clo, 0) / RASTER_TILE, tc1 = min(c.chi, dim_y - 1) / RASTER_TILE;
    for (int tr = tr0; tr <= tr1; ++tr)
        for (int tc = tc0; tc <= tc1; ++tc) {
            const int tile = tr * tiles_y + tc;
            const int slot = atomicAdd(count + tile, 1);
            if (slot < RASTER_CAP) entries[(size_t)tile * RASTER_CAP + slot] = slot_c;
        }
}

__global__ void __launch_bounds__(RASTER_THREADS)
raster_gather_kernel(const int* __restrict__ n_caps_all, const Cap* __restrict__ caps_all, int dim_x, int dim_y, int margin,
                     float two_scale, int tiles_y, const int* __restrict__ count, const int* __restrict__ entries,
                     float* __restrict__ out) {
    __shared__ Cap caps[RASTER_CAP];
    __shared__ int n_caps;
    const int tile = blockIdx.y * tiles_y + blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r0 = blockIdx.y * RASTER_TILE + 4 * warp;      // this thread: rows r0 .. r0+3, column c
    const int c = blockIdx.x * RASTER_TILE + lane;
    const int t_rlo = blockIdx.y * RASTER_TILE, t_rhi = t_rlo + RASTER_TILE - 1;
    const int t_clo = blockIdx.x * RASTER_TILE, t_chi = t_clo + RASTER_TILE - 1;
    const int listed = count[tile];
    const bool overflow = listed > RASTER_CAP;               // then: scan every accepted cap
    const int total = overflow ? *n_caps_all : listed;
    const double cd = (double)(c + margin), rd = (double)(r0 + margin);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int base = 0; base < total; base += RASTER_CAP) {
        if (threadIdx.x == 0) n_caps = 0;
        __syncthreads();
        const int e = base + threadIdx.x;
        if (e < total) {
            const Cap cp = caps_all[overflow ? e : entries[(size_t)tile * RASTER_CAP + e]];
            if (!overflow || (cp.rlo <= t_rhi && cp.rhi >= t_rlo && cp.clo <= t_chi && cp.chi >= t_clo))
                caps[overflow ? atomicAdd(&n_caps, 1) : threadIdx.x] = cp;
        }
        if (!overflow && threadIdx.x == 0) n_caps = min(total - base, RASTER_CAP);
        __syncthreads();
        const int m = n_caps;
        for (int k = 0; k < m; ++k) {
            const Cap& cp = caps[k];
            if (cp.rhi < r0 || cp.rlo > r0 + 3) continue;          // warp-uniform: rows of this warp
            if (c < cp.clo || c > cp.chi) continue;
            const double ey = cd - cp.yf;
            const double rem = cp.r2 - ey * ey;
            const double ex0 = rd - cp.xf;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int r = r0 + q;
                const double ex = ex0 + (double)q;
                const float diff = (float)(rem - ex * ex);
                if (diff > 0.f && r >= cp.rlo && r <= cp.rhi) {
                    float inv;   // sqrt(x) = x * rsqrt(x): one MUFU, ~2 ulp
                    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(diff));
                    acc[q] = fmaf(two_scale, diff * inv, acc[q]);
                }
            }
        }
        __syncthreads();
    }
    if (c < dim_y) {
        float* o = out + (size_t)r0 * dim_y + c;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (r0 + q < dim_x) o[(size_t)q * dim_y] = acc[q];
    }
}

// One membrane position from the once-rasterised sphere field: the grain map is translation
// invariant (getMembraneFromFile.py:139-161 only shifts the list by integer pixels per layer), so
// out[r][c] = sum over layers of field[ox_l + margin + r][oy_l + margin + c].
// Window starts are arbitrary, so loads are scalar; a thread owns four columns 256 apart, which keeps every
// load and store instruction of a warp on 32 consecutive floats.
// (the kernel is membrane_from_field_batch_kernel, further down: one launch for up to 8 positions)

__global__ void sphere_map_kernel(double rad, double scale_m, int dim_x, int dim_y, float* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= dim_y) return;
    const double a = dim_x / 2.0 - i, b = dim_y / 2.0 - j;   // createSampGeom.py:47
    const double d2 = a * a + b * b;
    out[(size_t)i * dim_y + j] = d2 < rad * rad ? (float)(2.0 * sqrt(rad * rad - d2) * scale_m) : 0.f;
}

struct Affine {
    double m[6];  // inverse map, OpenCV layout: src_x = m0*x + m1*y + m2, src_y = m3*x + m4*y + m5
};

__device__ __forceinline__ double cylinder_profile(long long col, long long row, int nxp, int nyp, double rad) {
    // createSampGeom.py:97-99: every row of the canvas holds the chord profile; outside the canvas
    // cv2.warpAffine's constant border reads 0.
    if (col < 0 || col >= nyp || row < 0 || row >= nxp) return 0.0;
    const double d = nyp / 2.0 - (double)col;
    return fabs(d) < rad ? 2.0 * sqrt(rad * rad - d * d) : 0.0;
}

// cv2.warpAffine (INTER_LINEAR, BORDER_CONSTANT) of the 2N x 2N canvas, evaluated only on the
// centre crop (:101): coordinates in 1/1024 fixed point, truncated to 1/32 pixel, bilinear taps.
__global__ void cylinder_map_kernel(Affine inv, double rad, double scale_m, int dim_x, int dim_y, float* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= dim_y) return;
    const int nxp = 2 * dim_x, nyp = 2 * dim_y;
    const int row = i + (nxp - dim_x) / 2, col = j + (nyp - dim_y) / 2;   // position on the canvas
    const long long adelta = __double2ll_rn(inv.m[0] * col * 1024.0);
    const long long bdelta = __double2ll_rn(inv.m[3] * col * 1024.0);
    const long long x0 = __double2ll_rn((inv.m[1] * row + inv.m[2]) * 1024.0) + 16;
    const long long y0 = __double2ll_rn((inv.m[4] * row + inv.m[5]) * 1024.0) + 16;
    const long long X = (x0 + adelta) >> 5, Y = (y0 + bdelta) >> 5;
    const long long sx = X >> 5, sy = Y >> 5;
    const double fx = (double)(X & 31) / 32.0, fy = (double)(Y & 31) / 32.0;
    const double v = cylinder_profile(sx, sy, nxp, nyp, rad) * ((1.0 - fx) * (1.0 - fy))
                   + cylinder_profile(sx + 1, sy, nxp, nyp, rad) * (fx * (1.0 - fy))
                   + cylinder_profile(sx, sy + 1, nxp, nyp, rad) * ((1.0 - fx) * fy)
                   + cylinder_profile(sx + 1, sy + 1, nxp, nyp, rad) * (fx * fy);
    out[(size_t)i * dim_y + j] = (float)(v * scale_m);
}

// The two demo phantoms of createSampGeom.py:110-260: two spheres (materials 0 and 1) inside a vertical
// tube (material 2 = tube - spheres); kind 0 = cylinder on the study grid, kind 1 = rounded parallelepiped
// built on a canvas with a margin of max(dim)/2, rotated by 15 degrees (imutils.rotate = cv2.warpAffine,
// same fixed-point taps as cylinder_map_kernel) and cropped.  Evaluated analytically: no canvas in memory.
struct Phantom {
    int kind, nxp, nyp;          // canvas
    int pos_x0, pos_x1, pos_y;   // sphere centres
    int ps2;                     // half patch size, ceil(r)
    double r_sphere, r_tube;     // pixels
};

__device__ __forceinline__ double phantom_sphere(const Phantom& ph, long long row, long long col, int pos_x) {
    const long long i = row - (pos_x - ph.ps2), j = col - (ph.pos_y - ph.ps2);      // patch coordinates (:149-150, :236-237)
    if (i < 0 || i >= 2 * ph.ps2 || j < 0 || j >= 2 * ph.ps2) return 0.0;
    const double di = (double)ph.ps2 - (double)i, dj = (double)ph.ps2 - (double)j;
    const double dist = di * di + dj * dj;
    return dist < ph.r_sphere * ph.r_sphere ? 2.0 * sqrt(ph.r_sphere * ph.r_sphere - dj * dj - di * di) : 0.0;
}

__device__ __forceinline__ double phantom_tube(const Phantom& ph, long long col) {
    const double r = ph.r_tube, c = ph.nyp / 2.0;
    if (ph.kind == 0) {                                                      // :141-143
        const double d = c - (double)col;
        return fabs(d) < r ? 2.0 * sqrt(r * r - d * d) : 0.0;
    }
    if (col >= ph.nxp) return 0.0;                                           // `for j in range(dimX)` (:222)
    double t = 0.0;                                                          // :223-228, later tests overwrite earlier ones
    const double j = (double)col;
    if (fabs(c - j) < r * 3 / 4) t = r * 2;
    if (r > (j - c) && (j - c) >= r * 3 / 4) { const double e = j - (c + r * 3 / 4); t = r / 2 * 3 + 2.0 * sqrt((r / 4) * (r / 4) - e * e); }
    if (-r < (j - c) && (j - c) <= -r * 3 / 4) { const double e = j - (c - r * 3 / 4); t = r / 2 * 3 + 2.0 * sqrt((r / 4) * (r / 4) - e * e); }
    return t;
}

__device__ __forceinline__ void phantom_at(const Phantom& ph, long long col, long long row, double v[3]) {
    v[0] = v[1] = v[2] = 0.0;
    if (col < 0 || col >= ph.nyp || row < 0 || row >= ph.nxp) return;      // BORDER_CONSTANT
    v[0] = phantom_sphere(ph, row, col, ph.pos_x0);
    v[1] = phantom_sphere(ph, row, col, ph.pos_x1);
    v[2] = phantom_tube(ph, col) - v[0] - v[1];
}

__global__ void phantom_kernel(Phantom ph, Affine inv, int rotate, int row0, int col0, double scale_m, int dim_x, int dim_y,
                               float* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= dim_y) return;
    const int row = i + row0, col = j + col0;
    double acc[3];
    if (!rotate) {
        phantom_at(ph, col, row, acc);
    } else {
        const long long adelta = __double2ll_rn(inv.m[0] * col * 1024.0);
        const long long bdelta = __double2ll_rn(inv.m[3] * col * 1024.0);
        const long long x0 = __double2ll_rn((inv.m[1] * row + inv.m[2]) * 1024.0) + 16;
        const long long y0 = __double2ll_rn((inv.m[4] * row + inv.m[5]) * 1024.0) + 16;
        const long long X = (x0 + adelta) >> 5, Y = (y0 + bdelta) >> 5;
        const long long sx = X >> 5, sy = Y >> 5;
        const double fx = (double)(X & 31) / 32.0, fy = (double)(Y & 31) / 32.0;
        double a[3], b[3], c[3], d[3];
        phantom_at(ph, sx, sy, a); phantom_at(ph, sx + 1, sy, b); phantom_at(ph, sx, sy + 1, c); phantom_at(ph, sx + 1, sy + 1, d);
        for (int m = 0; m < 3; ++m)
            acc[m] = a[m] * ((1.0 - fx) * (1.0 - fy)) + b[m] * (fx * (1.0 - fy)) + c[m] * ((1.0 - fx) * fy) + d[m] * (fx * fy);
    }
    const size_t n = (size_t)dim_x * dim_y, p = (size_t)i * dim_y + j;
    for (int m = 0; m < 3; ++m) out[m * n + p] = (float)(acc[m] * scale_m);
}

}  // namespace paresis

using namespace paresis;

extern "C" size_t paresis_raster_work_bytes(int n_spheres, int n_layers, int dim_x, int dim_y) {
    const size_t tiles = (size_t)div_up(dim_x, RASTER_TILE) * div_up(dim_y, RASTER_TILE);
    const size_t cands = (size_t)(n_spheres > 0 ? n_spheres : 0) * (size_t)(n_layers > 0 ? n_layers : 1);
    // [tile counts + cap counter | tile lists | accepted caps]
    return ((tiles + 1 + 3) / 4 * 4 + tiles * RASTER_CAP) * sizeof(int) + (cands + 1) * sizeof(Cap);
}

extern "C" int paresis_raster_spheres(const double* spheres, int n_spheres, double pix_um,
                                      const int64_t* offsets_host, int n_layers, int dim_x, int dim_y,
                                      int margin, float* thickness_out, void* work, size_t work_bytes,
                                      paresis_stream stream) {
    if (!spheres || !offsets_host || !thickness_out || n_spheres < 0 || n_layers < 1 || n_layers > MAX_MEMBRANE_LAYERS ||
        dim_x < 1 || dim_y < 1 || margin < 0 || !(pix_um > 0) || !work ||
        work_bytes < paresis_raster_work_bytes(n_spheres, n_layers, dim_x, dim_y) || (long long)n_spheres * n_layers >= (1LL << 31)) {
        set_last_error("paresis_raster_spheres: bad arguments (layers 1..%d, work buffer of paresis_raster_work_bytes required)",
                       MAX_MEMBRANE_LAYERS);
        return PARESIS_ERR_ARG;
    }
    cudaStream_t s = (cudaStream_t)stream;
    LayerOffsets off{};
    for (int l = 0; l < n_layers; ++l) {
        off.x[l] = offsets_host[2 * l];
        off.y[l] = offsets_host[2 * l + 1];
    }
    const int tiles_x = div_up(dim_x, RASTER_TILE), tiles_y = div_up(dim_y, RASTER_TILE);
    const size_t tiles = (size_t)tiles_x * tiles_y;
    int* count = (int*)work;
    int* n_caps = count + tiles;
    int* entries = count + (tiles + 1 + 3) / 4 * 4;
    Cap* caps = (Cap*)(entries + tiles * RASTER_CAP);
    PARESIS_CUDA(cudaMemsetAsync(count, 0, sizeof(int) * (tiles + 1), s));
    const long long candidates = (long long)n_spheres * n_layers;
    if (candidates > 0) {
        const int blocks = (int)((candidates + RASTER_THREADS - 1) / RASTER_THREADS);
        raster_bin_kernel<<<blocks, RASTER_THREADS, 0, s>>>(spheres, n_spheres, pix_um, off, n_layers, dim_x, dim_y, margin,
                                                             tiles_y, count, entries, n_caps, caps);
        PARESIS_LAUNCH_CHECK("raster_bin_kernel");
    }
    raster_gather_kernel<<<dim3(tiles_y, tiles_x), RASTER_THREADS, 0, s>>>(n_caps, caps, dim_x, dim_y, margin,
                                                                            2.0f * (float)(pix_um * 1e-6), tiles_y, count, entries,
                                                                            thickness_out);
    PARESIS_LAUNCH_CHECK("raster_gather_kernel");
    return PARESIS_OK;
}

extern "C" int paresis_raster_field(const double* spheres, int n_spheres, double pix_um, int field_x, int field_y,
                                    float* field, void* work, size_t work_bytes, paresis_stream stream) {
    const int64_t zero[2] = {0, 0};
    return paresis_raster_spheres(spheres, n_spheres, pix_um, zero, 1, field_x, field_y, 0, field, work, work_bytes, stream);
}

extern "C" int paresis_membrane_from_field(const float* field, int field_x, int field_y, const int64_t* offsets_host,
                                           int n_layers, int margin, int dim_x, int dim_y, float* thickness_out,
                                           paresis_stream stream) {
    if (!field || !offsets_host || !thickness_out || field_x < 1 || field_y < 1 || n_layers < 1 || n_layers > MAX_MEMBRANE_LAYERS ||
        margin < 0 || dim_x < 1 || dim_y < 1) {
        set_last_error("paresis_membrane_from_field: bad arguments (layers 1..%d)", MAX_MEMBRANE_LAYERS);
        return PARESIS_ERR_ARG;
    }
    const int64_t* offs[1] = {offsets_host};
    float* outs[1] = {thickness_out};
    return paresis_membrane_from_field_batch(field, field_x, field_y, offs, outs, 1, n_layers, margin, dim_x, dim_y, stream);
}

// The same for several positions in one launch (blockIdx.z): the windows of different positions overlap in the field,
// so a batch also re-uses what the previous position pulled into L2.
struct MembraneBatch {
    LayerOffsets off[PARESIS_MAX_HOP_BATCH];
    float* out[PARESIS_MAX_HOP_BATCH];
};

template <int ROWS>
__global__ void __launch_bounds__(256)
membrane_from_field_batch_kernel(const float* __restrict__ field, int field_x, int field_y, const MembraneBatch b, int n_layers, int margin,
                                 int dim_x, int dim_y) {
    // a thread owns four columns 256 apart (every load a full 128-byte line per warp) of ROWS consecutive rows: all the
    // loads of a layer are issued before the first add, so ROWS x 4 lines are in flight per warp and layer
    const int c0 = blockIdx.x * 1024 + threadIdx.x;
    const int r0 = blockIdx.y * ROWS;
    const LayerOffsets& off = b.off[blockIdx.z];
    float acc[ROWS][4];
#pragma unroll
    for (int rr = 0; rr < ROWS; ++rr)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[rr][k] = 0.f;
    for (int l = 0; l < n_layers; ++l) {
        const long long fc = off.y[l] + margin + c0;
        float v[ROWS][4];
#pragma unroll
        for (int rr = 0; rr < ROWS; ++rr) {
            const long long fr = off.x[l] + margin + r0 + rr;
            const bool row_ok = fr >= 0 && fr < field_x && r0 + rr < dim_x;
            const float* src = field + (size_t)(row_ok ? fr : 0) * field_y;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const long long f = fc + 256 * k;
                v[rr][k] = (row_ok && c0 + 256 * k < dim_y && f >= 0 && f < field_y) ? __ldg(src + f) : 0.f;
            }
        }
#pragma unroll
        for (int rr = 0; rr < ROWS; ++rr)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[rr][k] += v[rr][k];
    }
#pragma unroll
    for (int rr = 0; rr < ROWS; ++rr) {
        if (r0 + rr >= dim_x) break;
        float* o = b.out[blockIdx.z] + (size_t)(r0 + rr) * dim_y + c0;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (c0 + 256 * k < dim_y) o[256 * k] = acc[rr][k];
    }
}

extern "C" int paresis_membrane_from_field_batch(const float* field, int field_x, int field_y, const int64_t* const* offsets_host,
                                                 float* const* thickness_out_host, int n_items, int n_layers, int margin, int dim_x,
                                                 int dim_y, paresis_stream stream) {
    if (!field || !offsets_host || !thickness_out_host || n_items < 1 || n_items > PARESIS_MAX_HOP_BATCH || field_x < 1 || field_y < 1 ||
        n_layers < 1 || n_layers > MAX_MEMBRANE_LAYERS || margin < 0 || dim_x < 1 || dim_y < 1) {
        set_last_error("paresis_membrane_from_field_batch: bad arguments (items 1..%d, layers 1..%d)", PARESIS_MAX_HOP_BATCH, MAX_MEMBRANE_LAYERS);
        return PARESIS_ERR_ARG;
    }
    MembraneBatch b{};
    for (int z = 0; z < n_items; ++z) {
        if (!offsets_host[z] || !thickness_out_host[z]) { set_last_error("paresis_membrane_from_field_batch: item %d is incomplete", z); return PARESIS_ERR_ARG; }
        for (int l = 0; l < n_layers; ++l) {
            b.off[z].x[l] = offsets_host[z][2 * l];
            b.off[z].y[l] = offsets_host[z][2 * l + 1];
        }
        b.out[z] = thickness_out_host[z];
    }
    constexpr int ROWS = 4;
    membrane_from_field_batch_kernel<ROWS><<<dim3(div_up(dim_y, 1024), div_up(dim_x, ROWS), n_items), 256, 0, (cudaStream_t)stream>>>(
        field, field_x, field_y, b, n_layers, margin, dim_x, dim_y);
    PARESIS_LAUNCH_CHECK("membrane_from_field_batch_kernel");
    return PARESIS_OK;
}

extern "C" int paresis_sphere_map(double radius_um, int dim_x, int dim_y, double pix_um, float* out,
                                  paresis_stream stream) {
    if (!out || dim_x < 1 || dim_y < 1 || !(pix_um > 0)) { set_last_error("paresis_sphere_map: bad arguments"); return PARESIS_ERR_ARG; }
    sphere_map_kernel<<<dim3(div_up(dim_y, 128), dim_x), 128, 0, (cudaStream_t)stream>>>(radius_um / pix_um, pix_um * 1e-6, dim_x, dim_y, out);
    PARESIS_LAUNCH_CHECK("sphere_map_kernel");
    return PARESIS_OK;
}

// cv2.getRotationMatrix2D((w//2, h//2), angle, 1.0) followed by cv2.invertAffineTransform
static Affine inverse_rotation(int width, int height, double angle_deg) {
    const double ang = angle_deg * M_PI / 180.0;
    const double al = cos(ang), be = sin(ang);
    const double cx = width / 2, cy = height / 2;
    const double m00 = al, m01 = be, m02 = (1 - al) * cx - be * cy;
    const double m10 = -be, m11 = al, m12 = be * cx + (1 - al) * cy;
    double det = m00 * m11 - m01 * m10;
    det = det != 0 ? 1.0 / det : 0.0;
    Affine inv;
    inv.m[0] = m11 * det;
    inv.m[1] = -m01 * det;
    inv.m[3] = -m10 * det;
    inv.m[4] = m00 * det;
    inv.m[2] = -inv.m[0] * m02 - inv.m[1] * m12;
    inv.m[5] = -inv.m[3] * m02 - inv.m[4] * m12;
    return inv;
}

extern "C" int paresis_two_sphere_phantom(int kind, int dim_x, int dim_y, double pix_um, float* out3, paresis_stream stream) {
    if (!out3 || dim_x < 1 || dim_y < 1 || !(pix_um > 0) || (kind != 0 && kind != 1)) {
        set_last_error("paresis_two_sphere_phantom: bad arguments (kind 0 = spheres in cylinder, 1 = spheres in parallelepiped)");
        return PARESIS_ERR_ARG;
    }
    const double r0_um = 500.0;                          // createSampGeom.py:126, :196
    Phantom ph{};
    ph.kind = kind;
    ph.r_sphere = r0_um / pix_um;
    ph.r_tube = 2 * r0_um / pix_um;
    ph.ps2 = (int)ceil(ph.r_sphere);
    int row0 = 0, col0 = 0, rotate = 0;
    Affine inv{};
    if (kind == 0) {
        ph.nxp = dim_x; ph.nyp = dim_y;
        ph.pos_y = dim_y / 2;                                                 // :128
        ph.pos_x0 = (int)nearbyint(r0_um * 3 / pix_um);                        // :129-130 (np.round: half to even)
        ph.pos_x1 = (int)nearbyint(r0_um * 7 / pix_um);
    } else {
        const int margin = (dim_x > dim_y ? dim_x : dim_y) / 2;               // :201-203
        ph.nxp = dim_x + 2 * margin; ph.nyp = dim_y + 2 * margin;
        ph.pos_y = ph.nyp / 2;
        ph.pos_x0 = ph.nxp * 2 / 5;                                           // :205-206
        ph.pos_x1 = ph.nxp * 3 / 5;
        row0 = col0 = margin;
        rotate = 1;
        inv = inverse_rotation(ph.nyp, ph.nxp, 15.0);                         // :197, :243-245
    }
    if (2 * ph.r_sphere > ph.nxp || 2 * ph.r_sphere > ph.nyp || 2 * ph.r_tube > ph.nxp || 2 * ph.r_tube > ph.nyp) {
        set_last_error("The sample is too big for the detector field of view (increase dimX, dimY)");   // :139, :147
        return PARESIS_ERR_ARG;
    }
    // the reference assigns the sphere patches by slicing: they must fit the canvas
    if (ph.pos_x0 - ph.ps2 < 0 || ph.pos_x1 + ph.ps2 > ph.nxp || ph.pos_y - ph.ps2 < 0 || ph.pos_y + ph.ps2 > ph.nyp) {
        set_last_error("paresis_two_sphere_phantom: the sphere patches do not fit the canvas (could not broadcast upstream)");
        return PARESIS_ERR_ARG;
    }
    phantom_kernel<<<dim3(div_up(dim_y, 128), dim_x), 128, 0, (cudaStream_t)stream>>>(ph, inv, rotate, row0, col0, pix_um * 1e-6,
                                                                                      dim_x, dim_y, out3);
    PARESIS_LAUNCH_CHECK("phantom_kernel");
    return PARESIS_OK;
}

extern "C" int paresis_cylinder_map(double radius_um, double angle_deg, int dim_x, int dim_y, double pix_um,
                                    float* out, paresis_stream stream) {
    if (!out || dim_x < 1 || dim_y < 1 || !(pix_um > 0)) { set_last_error("paresis_cylinder_map: bad arguments"); return PARESIS_ERR_ARG; }
    const int nxp = 2 * dim_x, nyp = 2 * dim_y;
    const double rad = radius_um / pix_um;
    if (2 * rad > nxp || 2 * rad > nyp) {   // createSampGeom.py:94-95
        set_last_error("The sample is too big for the detector field of view (increase dimX, dimY)");
        return PARESIS_ERR_ARG;
    }
    // cv2.getRotationMatrix2D((w//2, h//2), angle, 1.0), then cv2.invertAffineTransform
    const double ang = angle_deg * M_PI / 180.0;
    const double al = cos(ang), be = sin(ang);
    const double cx = nyp / 2, cy = nxp / 2;
    const double m00 = al, m01 = be, m02 = (1 - al) * cx - be * cy;
    const double m10 = -be, m11 = al, m12 = be * cx + (1 - al) * cy;
    double det = m00 * m11 - m01 * m10;
    det = det != 0 ? 1.0 / det : 0.0;
    Affine inv;
    inv.m[0] = m11 * det;
    inv.m[1] = -m01 * det;
    inv.m[3] = -m10 * det;
    inv.m[4] = m00 * det;
    inv.m[2] = -inv.m[0] * m02 - inv.m[1] * m12;
    inv.m[5] = -inv.m[3] * m02 - inv.m[4] * m12;
    cylinder_map_kernel<<<dim3(div_up(dim_y, 128), dim_x), 128, 0, (cudaStream_t)stream>>>(inv, rad, pix_um * 1e-6, dim_x, dim_y, out);
    PARESIS_LAUNCH_CHECK("cylinder_map_kernel");
    return PARESIS_OK;
}
