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

// Two kernels.  cull: one thread per (layer, sphere) candidate tests the reference's acceptance
// window (:151) and the field of view; a survivor (~3 %) cuts its bounding box (:155-159) into
// chunks of RASTER_CHUNK cells and appends one work item per chunk to a list in `work`.
// fill: warps walk that list, so the cost of a rare 100-pixel grain is spread over many warps
// instead of stalling one block.  Each cell adds 2*sqrt(r^2 - d^2) with REDG.ADD.F32; the
// cancelling difference r^2 - d^2 stays in fp64.
constexpr int RASTER_THREADS = 256;
constexpr int RASTER_CHUNK = 512;

struct RasterItem {
    double xf, yf, rad;
    int x, y, chunk, pad_;
};

__device__ __forceinline__ void raster_chunk(const RasterItem& h, int lane, int nlanes, double scale_m, int dim_x, int dim_y,
                                             int margin, float* __restrict__ out) {
    const int reach = (int)floor(h.rad) + 1;
    const int w = 2 * reach;
    const double r2 = h.rad * h.rad;
    const double ex0 = (double)h.x - h.xf, ey0 = (double)h.y - h.yf;
    const unsigned magic = 0xffffffffu / (unsigned)w + 1u;   // idx / w == umulhi(idx, magic) for idx < w*w <= 2^20
    const int r0 = h.x - reach - margin, c0 = h.y - reach - margin;   // canvas -> cropped field of view (:161)
    const int end = min((h.chunk + 1) * RASTER_CHUNK, w * w);
    const float two_scale = 2.0f * (float)scale_m;
    for (int idx = h.chunk * RASTER_CHUNK + lane; idx < end; idx += nlanes) {
        const int q = (int)__umulhi((unsigned)idx, magic);
        const int p = idx - q * w;
        const int r = r0 + q, c = c0 + p;
        if ((unsigned)r >= (unsigned)dim_x || (unsigned)c >= (unsigned)dim_y) continue;
        const double ex = (double)(q - reach) + ex0, ey = (double)(p - reach) + ey0;
        const float diff = (float)(r2 - (ex * ex + ey * ey));
        if (diff > 0.f) {
            float inv;   // sqrt(x) = x * rsqrt(x): one MUFU, ~2 ulp, no IEEE fix-up code in the inner loop
            asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(diff));
            red_add(out + (size_t)r * dim_y + c, two_scale * (diff * inv));
        }
    }
}

__global__ void __launch_bounds__(RASTER_THREADS)
raster_cull_kernel(const double* __restrict__ spheres, int n, double pix, double scale_m, LayerOffsets off, int n_layers,
                   int dim_x, int dim_y, int margin, float* __restrict__ out, int* __restrict__ count,
                   RasterItem* __restrict__ items, int capacity) {
    const long long id = (long long)blockIdx.x * RASTER_THREADS + threadIdx.x;
    if (id >= (long long)n * n_layers) return;
    const int layer = (int)(id / n), s = (int)(id % n);
    const int margin2 = margin / 2;
    RasterItem h;
    h.rad = spheres[3 * s + 2] / pix;
    h.xf = spheres[3 * s + 1] / pix - (double)off.x[layer];
    h.yf = spheres[3 * s + 0] / pix - (double)off.y[layer];
    const long long x = __double2ll_rn(h.xf), y = __double2ll_rn(h.yf);   // np.round: half to even (:149-150)
    const long long reach = (long long)floor(h.rad) + 1;
    const bool accepted = margin2 < x && x < dim_x + margin + margin2 && margin2 < y && y < dim_y + margin + margin2;
    const bool visible = x + reach > margin && x - reach < dim_x + margin && y + reach > margin && y - reach < dim_y + margin;
    if (!(accepted && visible) || reach > 500) return;   // (a 500-pixel grain would already exceed the canvas margin)
    h.x = (int)x; h.y = (int)y; h.pad_ = 0;
    const int cells = (int)(4 * reach * reach);
    const int chunks = (cells + RASTER_CHUNK - 1) / RASTER_CHUNK;
    const int base = atomicAdd(count, chunks);
    for (int k = 0; k < chunks; ++k) {
        h.chunk = k;
        if (base + k < capacity) items[base + k] = h;
        else raster_chunk(h, 0, 1, scale_m, dim_x, dim_y, margin, out);   // list full: do it here, slowly but correctly
    }
}

__global__ void __launch_bounds__(RASTER_THREADS)
raster_fill_kernel(const int* __restrict__ count, const RasterItem* __restrict__ items, int capacity, double scale_m,
                   int dim_x, int dim_y, int margin, float* __restrict__ out) {
    const int n_items = min(*count, capacity);
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * RASTER_THREADS + threadIdx.x) >> 5, n_warps = (gridDim.x * RASTER_THREADS) >> 5;
    for (int e = warp; e < n_items; e += n_warps) raster_chunk(items[e], lane, 32, scale_m, dim_x, dim_y, margin, out);
}

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

}  // namespace paresis

using namespace paresis;

extern "C" size_t paresis_raster_work_bytes(int n_spheres, int n_layers, int dim_x, int dim_y) {
    (void)n_spheres;
    const size_t items = (size_t)dim_x * dim_y * (size_t)(n_layers > 0 ? n_layers : 1) / 128 + 65536;
    return 16 + items * sizeof(RasterItem);
}

extern "C" int paresis_raster_spheres(const double* spheres, int n_spheres, double pix_um,
                                      const int64_t* offsets_host, int n_layers, int dim_x, int dim_y,
                                      int margin, float* thickness_out, void* work, size_t work_bytes,
                                      paresis_stream stream) {
    if (!spheres || !offsets_host || !thickness_out || n_spheres < 0 || n_layers < 1 || n_layers > MAX_MEMBRANE_LAYERS ||
        dim_x < 1 || dim_y < 1 || margin < 0 || !(pix_um > 0) || !work || work_bytes < 16 + sizeof(RasterItem)) {
        set_last_error("paresis_raster_spheres: bad arguments (layers 1..%d, work buffer required)", MAX_MEMBRANE_LAYERS);
        return PARESIS_ERR_ARG;
    }
    cudaStream_t s = (cudaStream_t)stream;
    PARESIS_CUDA(cudaMemsetAsync(thickness_out, 0, sizeof(float) * (size_t)dim_x * dim_y, s));
    if (n_spheres == 0) return PARESIS_OK;
    LayerOffsets off{};
    for (int l = 0; l < n_layers; ++l) {
        off.x[l] = offsets_host[2 * l];
        off.y[l] = offsets_host[2 * l + 1];
    }
    int* count = (int*)work;
    RasterItem* items = (RasterItem*)((char*)work + 16);
    const int capacity = (int)((work_bytes - 16) / sizeof(RasterItem));
    PARESIS_CUDA(cudaMemsetAsync(count, 0, sizeof(int), s));
    const long long candidates = (long long)n_spheres * n_layers;
    const int blocks = (int)((candidates + RASTER_THREADS - 1) / RASTER_THREADS);
    raster_cull_kernel<<<blocks, RASTER_THREADS, 0, s>>>(spheres, n_spheres, pix_um, pix_um * 1e-6, off, n_layers, dim_x,
                                                          dim_y, margin, thickness_out, count, items, capacity);
    PARESIS_LAUNCH_CHECK("raster_cull_kernel");
    raster_fill_kernel<<<148 * 4, RASTER_THREADS, 0, s>>>(count, items, capacity, pix_um * 1e-6, dim_x, dim_y, margin,
                                                           thickness_out);
    PARESIS_LAUNCH_CHECK("raster_fill_kernel");
    return PARESIS_OK;
}

extern "C" int paresis_sphere_map(double radius_um, int dim_x, int dim_y, double pix_um, float* out,
                                  paresis_stream stream) {
    if (!out || dim_x < 1 || dim_y < 1 || !(pix_um > 0)) { set_last_error("paresis_sphere_map: bad arguments"); return PARESIS_ERR_ARG; }
    sphere_map_kernel<<<dim3(div_up(dim_y, 128), dim_x), 128, 0, (cudaStream_t)stream>>>(radius_um / pix_um, pix_um * 1e-6, dim_x, dim_y, out);
    PARESIS_LAUNCH_CHECK("sphere_map_kernel");
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
