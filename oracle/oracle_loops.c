/*
 * TEST INFRASTRUCTURE ONLY -- scalar fp64 CPU restatement of the PARESIS loops that are
 * too slow for numpy.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product never does.
 *
 * Each function names the reference lines (relative to /root/reference/CodePython) whose
 * behaviour it restates.  Arrays are C-contiguous fp64, axis 0 = "x" = row.
 *
 * Build: see oracle/Makefile  (gcc -O2 -shared -fPIC).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

/* ------------------------------------------------------------------------------------
 * Bilinear forward scatter of refractionFileNumba2.py:198-263 (fastloopNumba; the v1 copy
 * at refractionFileNumba.py:70-135 is the same body).
 *
 * The frame the reference loops over is the zero-padded one (fastRefraction pads by 15,
 * :65-67, and crops afterwards, :78).  Here the padding is virtual: `margin` says how far
 * the loop frame extends beyond the stored nx*ny arrays; deposits that land in the margin
 * are simply not stored (they would be cropped).  margin = 0 gives the raw kernel.
 *
 * Per ray: zero displacement short-circuits (:222-224).  |D| > 1 moves the base cell by
 * floor(D) and keeps the fractional part (:228-233); otherwise the signed sub-pixel shift
 * is kept and its sign picks the neighbour side.  The base must be inside the loop frame
 * (:235-236).  The three neighbour deposits happen only when BOTH the row neighbour and the
 * column neighbour exist inside the loop frame (the nested tests at :238-262), which is the
 * reference's edge quirk (last case).
 * ---------------------------------------------------------------------------------- */
void oracle_splat(int nx, int ny, int margin, const double *I, const double *Dx, const double *Dy,
                  double *out)
{
    const long fx_n = (long)nx + 2L * margin, fy_n = (long)ny + 2L * margin;
    for (int i = 0; i < nx; ++i) {
        for (int j = 0; j < ny; ++j) {
            const size_t p = (size_t)i * ny + j;
            const double v = I[p];
            double dx = Dx[p], dy = Dy[p];
            if (dx == 0.0 && dy == 0.0) {
                out[p] += v;
                continue;
            }
            long r = i, c = j;
            if (fabs(dx) > 1.0) { double f = floor(dx); r += (long)f; dx -= f; }
            if (fabs(dy) > 1.0) { double f = floor(dy); c += (long)f; dy -= f; }
            const long rp = r + margin, cp = c + margin; /* position in the loop frame */
            if (rp < 0 || rp >= fx_n || cp < 0 || cp >= fy_n) continue;
            const double ax = fabs(dx), ay = fabs(dy);
            const long sr = dx >= 0.0 ? 1 : -1, sc = dy >= 0.0 ? 1 : -1;
#define PUT(rr, cc, w)                                                              \
    do {                                                                            \
        long _r = (rr), _c = (cc);                                                  \
        if (_r >= 0 && _r < nx && _c >= 0 && _c < ny) out[(size_t)_r * ny + _c] += (w); \
    } while (0)
            PUT(r, c, v * (1.0 - ax) * (1.0 - ay));
            const int row_ok = sr > 0 ? rp < fx_n - 1 : rp > 0;
            const int col_ok = sc > 0 ? cp < fy_n - 1 : cp > 0;
            if (row_ok && col_ok) {
                PUT(r + sr, c, v * ax * (1.0 - ay));
                PUT(r + sr, c + sc, v * ax * ay);
                PUT(r, c + sc, v * (1.0 - ax) * ay);
            }
#undef PUT
        }
    }
}

/* ------------------------------------------------------------------------------------
 * Detector.py:185-198 (resize): s = int(Nx / sizeX); every output pixel is the SUM of an
 * s*s block.  The caller handles the identity short-cut (:188-189).
 * ---------------------------------------------------------------------------------- */
void oracle_bin_sum(int nx, int ny, const double *img, int sx, int sy, double *out)
{
    const int s = nx / sx;
    for (int a = 0; a < sx; ++a)
        for (int b = 0; b < sy; ++b) {
            double acc = 0.0;
            for (int u = 0; u < s; ++u)
                for (int w = 0; w < s; ++w) {
                    int r = a * s + u, c = b * s + w;
                    if (r < nx && c < ny) acc += img[(size_t)r * ny + c];
                }
            out[(size_t)a * sy + b] = acc;
        }
}

/* ------------------------------------------------------------------------------------
 * Samples/getMembraneFromFile.py:143-159: rasterise one layer of spherical caps.
 *
 * `spheres` holds n rows [c0, c1, radius] already rescaled and shifted to a top-left origin
 * (:95-124).  canvas is (dim_x + 2*margin) x (dim_y + 2*margin), in pixel units; the caller
 * crops the margin and converts to metres (:161, :168).  Centre rounding is numpy's
 * round-half-even (:149-150) == nearbyint in the default rounding mode.
 * ---------------------------------------------------------------------------------- */
void oracle_raster_layer(int n, const double *spheres, double pix, long off_x, long off_y,
                         int dim_x, int dim_y, int margin, double *canvas)
{
    const int margin2 = margin / 2;
    const long cx_n = (long)dim_x + 2L * margin, cy_n = (long)dim_y + 2L * margin;
    for (int s = 0; s < n; ++s) {
        const double rad = spheres[3 * s + 2] / pix;
        const long rint_ = (long)floor(rad) + 1;
        const double xf = spheres[3 * s + 1] / pix - (double)off_x;
        const double yf = spheres[3 * s + 0] / pix - (double)off_y;
        const long x = (long)nearbyint(xf), y = (long)nearbyint(yf);
        if (!(margin2 < x && x < dim_x + margin + margin2 && margin2 < y && y < dim_y + margin + margin2))
            continue;
        for (long ii = -rint_; ii < rint_; ++ii)
            for (long jj = -rint_; jj < rint_; ++jj) {
                const double ex = (double)(ii + x) - xf, ey = (double)(jj + y) - yf;
                const double dist = sqrt(ex * ex + ey * ey);
                if (dist < rad) {
                    long r = x + ii, c = y + jj;
                    if (r < 0) r += cx_n; /* numpy negative-index wrap; lands in the cropped margin */
                    if (c < 0) c += cy_n;
                    if (r >= 0 && r < cx_n && c >= 0 && c < cy_n)
                        canvas[(size_t)r * cy_n + c] += 2.0 * sqrt(rad * rad - dist * dist);
                }
            }
    }
}
