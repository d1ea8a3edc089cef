"""TEST INFRASTRUCTURE ONLY -- fp64 CPU oracle for the PARESIS image-formation hot path.

A plain numpy (+ three C loops, ``oracle_loops.c``) restatement of the reference algorithm,
written from its behaviour; every function cites the reference lines it follows (paths are
relative to ``/root/reference/CodePython``).  It is pinned against the unmodified reference
run in the build container: ``oracle/make_golden.py`` imports the reference through
``oracle/ref_harness.py`` and stores its outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this module against those vectors.  (The reference has
no tests or golden vectors of its own)

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product (``paresis_b200``) never
does: it calls the CUDA library through the C ABI and fails loudly without it.
"""
import ctypes
import os
import subprocess

import numpy as np
from scipy.signal import fftconvolve

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle_loops.so")
_lib = None


def build_loops(force=False):
    """Compile oracle_loops.c (gcc) into oracle/_build/; returns the .so path."""
    src = os.path.join(_HERE, "oracle_loops.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_LIB_PATH), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", _LIB_PATH, src, "-lm"])
    return _LIB_PATH


def _loops():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build_loops())
        dp = ctypes.POINTER(ctypes.c_double)
        _lib.oracle_splat.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, dp, dp, dp, dp]
        _lib.oracle_bin_sum.argtypes = [ctypes.c_int, ctypes.c_int, dp, ctypes.c_int, ctypes.c_int, dp]
        _lib.oracle_raster_layer.argtypes = [ctypes.c_int, dp, ctypes.c_double, ctypes.c_long, ctypes.c_long,
                                             ctypes.c_int, ctypes.c_int, ctypes.c_int, dp]
        for f in (_lib.oracle_splat, _lib.oracle_bin_sum, _lib.oracle_raster_layer):
            f.restype = None
    return _lib


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _c64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


# --------------------------------------------------------------------------- constants
PLANCK = 6.626e-34      # getk.py:16 -- the reference's own rounded constants, to the digit
LIGHT = 2.998e8         # getk.py:17
CHARGE = 1.6e-19        # getk.py:18
PAD_REFRACTION = 15     # refractionFileNumba2.py:50 (margin2); v1 uses 10, refractionFileNumba.py:36
PAD_FRESNEL = 15        # Experiment.py:236
PAD_DETECTOR = 15       # Detector.py:92


def wavenumber(energy_ev):
    """getk.py:12-20; same expression inline at Sample.py:265,300."""
    return 2 * np.pi * energy_ev * CHARGE / (PLANCK * LIGHT)


def wavenumber_from_lambda(energy_kev):
    """refractionFileNumba2.py:47-48 forms k through lambda; differs from getk by rounding only."""
    lam = 6.626 * 1e-34 * 2.998e8 / (energy_kev * 1000 * 1.6e-19)
    return 2 * np.pi / lam


# --------------------------------------------------------------------------- kernels
def gaussian_kernel(sigma):
    """Detector.py:201-220 / refractionFileNumba2.py:14-23: (2*round(3s)+1)^2 normalised Gaussian.
    ``round`` is Python's round-half-even."""
    half = round(sigma * 3)
    ax = np.arange(-half, half + 1, dtype=np.float64)
    g = np.exp(-(ax[None, :] ** 2 / 2.0 / sigma ** 2 + ax[:, None] ** 2 / 2.0 / sigma ** 2))
    return g / np.sum(g)


def gradient2(f, h):
    """np.gradient(f, h, edge_order=2) as used at refractionFileNumba2.py:54.
    Interior: central difference; borders: second-order one-sided."""
    f = _c64(f)
    out = []
    for axis in (0, 1):
        g = np.empty_like(f)
        a = np.moveaxis(f, axis, 0)
        b = np.moveaxis(g, axis, 0)
        b[1:-1] = (a[2:] - a[:-2]) / (2.0 * h)
        b[0] = -(3.0 * a[0] - 4.0 * a[1] + a[2]) / (2.0 * h)
        b[-1] = (3.0 * a[-1] - 4.0 * a[-2] + a[-3]) / (2.0 * h)
        out.append(g)
    return out


def displacement(intensity, phi, distance, energy_kev, magnification, pixel_um):
    """refractionFileNumba2.py:47-64: phase gradient -> displacement in pixels, with the
    reference's clean-up.  Returns (I', Dx, Dy); I' has rays with |D| > N zeroed (:61-62)."""
    k = wavenumber_from_lambda(energy_kev)
    nx, ny = intensity.shape
    h = pixel_um * 1e-6
    gx, gy = gradient2(phi, h)
    dx = gx * distance / k / (h * magnification)
    dy = gy * distance / k / (h * magnification)
    dx[np.abs(dx) < 1e-12] = 0
    dy[np.abs(dy) < 1e-12] = 0
    out_i = np.array(intensity, dtype=np.float64)
    out_i[np.abs(dx) > nx] = 0
    out_i[np.abs(dy) > ny] = 0
    dx[np.abs(dx) > nx] = 0
    dy[np.abs(dy) > ny] = 0
    return out_i, dx, dy


def splat(intensity, dx, dy, margin=0):
    """refractionFileNumba2.py:198-263 on a frame virtually zero-padded by ``margin``."""
    i = _c64(intensity)
    out = np.zeros_like(i)
    _loops().oracle_splat(i.shape[0], i.shape[1], int(margin), _dp(i), _dp(_c64(dx)), _dp(_c64(dy)), _dp(out))
    return out


def splat_python(intensity, dx, dy, margin=0):
    """Pure-Python twin of ``splat`` (small cases only; cross-checks the C loop)."""
    nx, ny = intensity.shape
    fx_n, fy_n = nx + 2 * margin, ny + 2 * margin
    out = np.zeros((nx, ny))

    def put(r, c, w):
        if 0 <= r < nx and 0 <= c < ny:
            out[r, c] += w

    for i in range(nx):
        for j in range(ny):
            v, ddx, ddy = float(intensity[i, j]), float(dx[i, j]), float(dy[i, j])
            if ddx == 0 and ddy == 0:
                out[i, j] += v
                continue
            r, c = i, j
            if abs(ddx) > 1:
                f = np.floor(ddx); r += int(f); ddx -= f
            if abs(ddy) > 1:
                f = np.floor(ddy); c += int(f); ddy -= f
            rp, cp = r + margin, c + margin
            if not (0 <= rp < fx_n and 0 <= cp < fy_n):
                continue
            ax, ay = abs(ddx), abs(ddy)
            sr, sc = (1 if ddx >= 0 else -1), (1 if ddy >= 0 else -1)
            put(r, c, v * (1 - ax) * (1 - ay))
            row_ok = rp < fx_n - 1 if sr > 0 else rp > 0
            col_ok = cp < fy_n - 1 if sc > 0 else cp > 0
            if row_ok and col_ok:
                put(r + sr, c, v * ax * (1 - ay))
                put(r + sr, c + sc, v * ax * ay)
                put(r, c + sc, v * (1 - ax) * ay)
    return out


class InsaneValues(Exception):
    pass


def fast_refraction(intensity, phi, distance, energy_kev, magnification, pixel_um, margin=PAD_REFRACTION):
    """refractionFileNumba2.py:25-86.  Returns (I2[N,N], Dx[N+2m,N+2m], Dy[...]) -- the
    displacement maps come back zero-padded, as in the reference (:65-66, :86)."""
    i2, dx, dy = displacement(intensity, phi, distance, energy_kev, magnification, pixel_um)
    out = splat(i2, dx, dy, margin)
    if np.isnan(out).any() or np.any(np.abs(out) > 1e50):
        raise InsaneValues("The calculated intensity refractive includes some nans or insane values")
    return out, np.pad(dx, margin), np.pad(dy, margin)


def dark_field_angle(thickness_m, delta, model):
    """Sample.py:322-343: mean scattering angle (rad) of the Lung / 'cylinder_beeds' micro-sphere model and the
    volume fraction the thickness is scaled by.  ``model`` = "Lung" (alveoli 47 um, fraction 0.5) or
    "cylinder_beeds" (15 um, 0.6)."""
    radius, fraction = (47.0, 0.5) if model == "Lung" else (15.0, 0.6)
    geom = np.asarray(thickness_m, dtype=np.float64) * 1e6
    n_vol = fraction * 3 / 4 / np.pi / (radius ** 3)
    n_sphere = n_vol ** (1 / 3) * geom
    return 2 * delta * n_sphere ** (1 / 2) * np.sqrt(np.log(2 / delta) + 1), fraction


def set_wave_rt_df(intensity, phi, thickness, deltas, betas, energy_kev, models):
    """Sample.py:285-351 WITH the dark-field branch: ``models[m]`` is None, "Lung" or "cylinder_beeds".
    Returns (I, phi, newDf); newDf is the int 0 when no material scatters (as upstream)."""
    k = 2 * np.pi * energy_kev * 1000 * 1.6e-19 / (6.626e-34 * 2.998e8)
    out_i, out_phi, new_df = intensity, phi, 0
    for t, d, b, model in zip(thickness, deltas, betas, models):
        geometry = t
        if model is not None:
            new_df, fraction = dark_field_angle(t, d, model)
            geometry = t * fraction
        out_i = np.exp(-2 * k * b * geometry) * out_i
        out_phi = out_phi - k * d * geometry
    return out_i, out_phi, new_df


def fast_refraction_df(intensity, phi, distance, energy_kev, magnification, pixel_um, dark_field):
    """refractionFileNumba2.py:88-196.  Returns (I3[N,N], Dx, Dy) with the displacement maps zero-padded by
    margin2 = ceil(6 max(DF)) (:116, :126-127, :196)."""
    nx, ny = np.asarray(intensity).shape
    df = np.array(dark_field, dtype=np.float64) * distance / (pixel_um * 1e-6 * magnification)      # :114
    margin2 = int(np.ceil(df.max() * 6))                                                            # :115-117
    i2, dx, dy = displacement(intensity, phi, distance, energy_kev, magnification, pixel_um)       # :120-129
    df[df > nx / 4] = 0                                                                             # :130
    plain = np.where(df != 0, 0.0, i2)                                                              # :143-146
    scat = np.where(df == 0, 0.0, i2)
    out = splat(plain, dx, dy, margin2)                                                             # :154
    moved = splat(scat, dx, dy, margin2)                                                            # :155
    for i in range(nx):                                                                             # :171-186
        for j in range(ny):
            v = moved[i, j]
            if v == 0:
                continue
            if df[i, j] != 0:
                patch = gaussian_kernel(df[i, j] / 2)
                h = patch.shape[0] // 2
                r0, r1, c0, c1 = max(i - h, 0), min(i + h + 1, nx), max(j - h, 0), min(j + h + 1, ny)   # cropped at :189
                out[r0:r1, c0:c1] += v * patch[r0 - i + h:r1 - i + h, c0 - j + h:c1 - j + h]
            else:
                out[i, j] += v
    if np.isnan(out).any() or np.any(np.abs(out) > 1e50):
        raise InsaneValues("The calculated intensity refractive includes some nans or insane values")
    return out, np.pad(dx, margin2), np.pad(dy, margin2)


def bin_sum(image, size_x, size_y):
    """Detector.py:185-198 (resize)."""
    img = _c64(image)
    if img.shape == (size_x, size_y):
        return img
    out = np.empty((size_x, size_y))
    _loops().oracle_bin_sum(img.shape[0], img.shape[1], _dp(img), int(size_x), int(size_y), _dp(out))
    return out


def detection(image, source_fwhm_px, oversampling, det_dims, psf_sigma, poisson=None):
    """Detector.py:79-119.  ``poisson`` is a callable lam->counts (None = noise-free)."""
    m = PAD_DETECTOR
    img = np.pad(_c64(image), m * oversampling, mode="reflect")
    if source_fwhm_px != 0:
        img = fftconvolve(img, gaussian_kernel(source_fwhm_px / 2.355), mode="same")
    img = bin_sum(img, det_dims[0] + 2 * m, det_dims[1] + 2 * m)
    if psf_sigma != 0:
        img = fftconvolve(img, gaussian_kernel(psf_sigma), mode="same")
    if poisson is not None:
        img = poisson(img)
    return img[m:det_dims[0] + m, m:det_dims[1] + m]


def set_wave_rt(intensity, phi, thickness, deltas, betas, energy_kev):
    """Sample.py:285-351 without the Lung / cylinder_beeds dark-field branch:
    I *= exp(-2 k beta_m t_m); phi -= k delta_m t_m, material by material."""
    k = 2 * np.pi * energy_kev * 1000 * 1.6e-19 / (6.626e-34 * 2.998e8)
    out_i, out_phi = intensity, phi
    for t, d, b in zip(thickness, deltas, betas):
        out_i = np.exp(-2 * k * b * t) * out_i
        out_phi = out_phi - k * d * t
    return out_i, out_phi


def set_wave(wave, thickness, deltas, betas, energy_kev):
    """Sample.py:248-282: complex transmission exp((-i k delta - k beta) t)."""
    k = 2 * np.pi * energy_kev * 1000 * 1.6e-19 / (6.626e-34 * 2.998e8)
    out = wave
    for t, d, b in zip(thickness, deltas, betas):
        out = np.exp((-1j * k * d - k * b) * t) * out
    return out


def wave_propagation(wave, distance, energy_kev, magnification, study_dims, pixel_um, margin=PAD_FRESNEL):
    """Experiment.py:219-252.  The frequency step uses the UNPADDED study dimensions (:246-247)
    while the transform runs on the reflect-padded array."""
    if distance == 0:
        return wave
    w = np.pad(wave, margin, mode="reflect")
    k = wavenumber(energy_kev * 1000)
    nx, ny = w.shape
    u = (np.arange(nx) - nx // 2) * 2 * np.pi / (study_dims[0] * pixel_um * 1e-6)
    v = (np.arange(ny) - ny // 2) * 2 * np.pi / (study_dims[1] * pixel_um * 1e-6)
    uv2 = u[:, None] ** 2 + v[None, :] ** 2
    kernel = np.exp(-1j * distance * uv2 / (2 * k * magnification))
    spec = np.fft.fftshift(np.fft.fft2(w))
    out = np.exp(1j * k * distance / magnification) * np.fft.ifft2(np.fft.ifftshift(kernel * spec))
    return out[margin:nx - margin, margin:ny - margin]


# --------------------------------------------------------------------------- geometry
def membrane_sphere_table(rows, mean_radius, dim_x, dim_y, pix):
    """Samples/getMembraneFromFile.py:84-124: rescale the sphere list to the wanted mean
    radius, move the origin to the top-left corner, tile along x then y until the list covers
    the field of view.  Returns (table[n,3], extent_x, extent_y) in micrometres."""
    corr = mean_radius / 12.8
    ext_x = int(np.floor(8102)) * corr + mean_radius
    ext_y = int(np.floor(9740)) * corr + mean_radius
    tab = np.asarray(rows, dtype=np.float64) * corr
    tab[:, 1] += ext_x / 2
    tab[:, 0] += ext_y / 2
    base, step = tab.copy(), ext_x
    while ext_x / pix - dim_x < 0:
        shifted = base.copy()
        shifted[:, 1] += ext_x
        tab = np.concatenate((tab, shifted), axis=0)
        ext_x += step
    base, step = tab.copy(), ext_y
    while ext_y / pix - dim_y < 0:
        shifted = base.copy()
        shifted[:, 0] += ext_y
        tab = np.concatenate((tab, shifted), axis=0)
        ext_y += step
    return tab, ext_x, ext_y


def membrane_margin(mean_radius, pix):
    """Samples/getMembraneFromFile.py:81-82."""
    margin = int(np.ceil(10 * mean_radius / pix))
    return margin, int(np.floor(margin / 2))


def draw_membrane_offsets(n_layers, mean_radius, ext_x, ext_y, dim_x, dim_y, pix, randint=None):
    """Samples/getMembraneFromFile.py:135-140: two ``np.random.randint`` draws per layer, x first.
    Uses the global numpy stream unless ``randint`` is given, so ``np.random.seed`` reproduces
    the reference draw for draw."""
    randint = randint or np.random.randint
    _, margin2 = membrane_margin(mean_radius, pix)
    offs = []
    for _ in range(n_layers):
        ox = randint(margin2, ext_x / pix - dim_x - margin2)
        oy = randint(margin2, ext_y / pix - dim_y - margin2)
        offs.append((int(ox), int(oy)))
    return offs


def membrane_segmented(rows, mean_radius, n_layers, dim_x, dim_y, pix, support_um, offsets=None):
    """Samples/getMembraneFromFile.py:60-171.  Returns [grains, support] thickness in metres."""
    tab, ext_x, ext_y = membrane_sphere_table(rows, mean_radius, dim_x, dim_y, pix)
    margin, _ = membrane_margin(mean_radius, pix)
    if offsets is None:
        offsets = draw_membrane_offsets(n_layers, mean_radius, ext_x, ext_y, dim_x, dim_y, pix)
    canvas = np.zeros((dim_x + 2 * margin, dim_y + 2 * margin))
    tab = _c64(tab)
    for ox, oy in offsets:
        _loops().oracle_raster_layer(tab.shape[0], _dp(tab), float(pix), int(ox), int(oy),
                                     int(dim_x), int(dim_y), int(margin), _dp(canvas))
    grains = canvas[margin:-margin, margin:-margin]
    return np.array([grains * pix * 1e-6, np.ones(grains.shape) * support_um * 1e-6])


def sample_sphere(radius_um, dim_x, dim_y, pix):
    """Samples/createSampGeom.py:41-53: centred sphere, thickness in metres."""
    r = radius_um / pix
    i = np.arange(dim_x, dtype=np.float64)[:, None]
    j = np.arange(dim_y, dtype=np.float64)[None, :]
    d2 = (dim_x / 2 - i) ** 2 + (dim_y / 2 - j) ** 2
    t = np.where(d2 < r ** 2, 2 * np.sqrt(np.maximum(r ** 2 - d2, 0.0)), 0.0)
    return t[None] * pix * 1e-6


def rotation_matrix(center, angle_deg, scale=1.0):
    """OpenCV ``getRotationMatrix2D`` (called by imutils.rotate, createSampGeom.py:100)."""
    a = np.deg2rad(angle_deg)
    al, be = scale * np.cos(a), scale * np.sin(a)
    return np.array([[al, be, (1 - al) * center[0] - be * center[1]],
                     [-be, al, be * center[0] + (1 - al) * center[1]]])


def warp_affine_linear(src, mat, out_w, out_h):
    """OpenCV ``warpAffine(src, M, (w, h))`` for a float64 image with its defaults
    (INTER_LINEAR, BORDER_CONSTANT 0) -- the third-party step behind ``imutils.rotate``
    at createSampGeom.py:100 (OpenCV 4.x, imgwarp.cpp WarpAffineInvoker + remapBilinear).

    Published algorithm restated: the matrix is inverted; source coordinates are formed in
    1/1024 fixed point (round-half-even of M*x*1024 per column plus a per-row constant plus the
    rounding offset 16), truncated to 1/32 pixel; the four taps are blended with the exact
    products of the 1/32 weights; taps outside the source read 0.
    """
    m = np.array(mat, dtype=np.float64)
    det = m[0, 0] * m[1, 1] - m[0, 1] * m[1, 0]
    det = 1.0 / det if det != 0 else 0.0
    a11, a22 = m[1, 1] * det, m[0, 0] * det
    inv = np.empty((2, 3))
    inv[0, 0], inv[0, 1], inv[1, 0], inv[1, 1] = a11, -m[0, 1] * det, -m[1, 0] * det, a22
    inv[0, 2] = -inv[0, 0] * m[0, 2] - inv[0, 1] * m[1, 2]
    inv[1, 2] = -inv[1, 0] * m[0, 2] - inv[1, 1] * m[1, 2]
    xs = np.arange(out_w, dtype=np.float64)
    ys = np.arange(out_h, dtype=np.float64)
    adelta = np.rint(inv[0, 0] * xs * 1024).astype(np.int64)
    bdelta = np.rint(inv[1, 0] * xs * 1024).astype(np.int64)
    x0 = np.rint((inv[0, 1] * ys + inv[0, 2]) * 1024).astype(np.int64) + 16
    y0 = np.rint((inv[1, 1] * ys + inv[1, 2]) * 1024).astype(np.int64) + 16
    X = (x0[:, None] + adelta[None, :]) >> 5
    Y = (y0[:, None] + bdelta[None, :]) >> 5
    sx, sy = X >> 5, Y >> 5
    fx, fy = (X & 31) / 32.0, (Y & 31) / 32.0
    h, w = src.shape

    def tap(yy, xx):
        ok = (yy >= 0) & (yy < h) & (xx >= 0) & (xx < w)
        return np.where(ok, src[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)], 0.0)

    return (tap(sy, sx) * ((1 - fx) * (1 - fy)) + tap(sy, sx + 1) * (fx * (1 - fy))
            + tap(sy + 1, sx) * ((1 - fx) * fy) + tap(sy + 1, sx + 1) * (fx * fy))


def sample_cylinder(radius_um, orientation_deg, dim_x, dim_y, pix):
    """Samples/createSampGeom.py:87-106: column profile on a 2N x 2N canvas, rotated about the
    canvas centre by ``imutils.rotate`` (cv2.warpAffine), centre crop, metres."""
    nxp, nyp = 2 * dim_x, 2 * dim_y
    r = radius_um / pix
    if 2 * r > nxp or 2 * r > nyp:
        raise Exception("The sample is too big for the detector field of view (increase dimX, dimY)")
    j = np.arange(nyp, dtype=np.float64)
    prof = np.where(np.abs(nyp / 2 - j) < r, 2 * np.sqrt(np.maximum(r ** 2 - (nyp / 2 - j) ** 2, 0.0)), 0.0)
    canvas = np.broadcast_to(prof[None, :], (nxp, nyp)).copy()
    rot = warp_affine_linear(canvas, rotation_matrix((nyp // 2, nxp // 2), orientation_deg), nyp, nxp)
    dx0, dy0 = int((nxp - dim_x) / 2), int((nyp - dim_y) / 2)
    return rot[None, dx0:dx0 + dim_x, dy0:dy0 + dim_y] * pix * 1e-6


def sample_two_spheres(kind, dim_x0, dim_y0, pix):
    """Samples/createSampGeom.py:110-172 (kind 0, spheres in a cylinder) and :174-260 (kind 1, spheres in a rounded
    parallelepiped on a canvas with margin max(dim)//2, rotated by 15 degrees, cropped).  [3, dimX, dimY] metres."""
    r0 = 500.0
    margin = 0 if kind == 0 else max(dim_x0, dim_y0) // 2
    dim_x, dim_y = dim_x0 + 2 * margin, dim_y0 + 2 * margin
    rad = r0 / pix
    ps2 = int(np.ceil(rad))
    ps = 2 * ps2
    if 2 * rad > dim_x or 2 * rad > dim_y:
        raise Exception("The sample is too big for the detector field of view (increase dimX, dimY)")
    i = np.arange(ps, dtype=np.float64)
    dist = (ps / 2 - i[:, None]) ** 2 + (ps / 2 - i[None, :]) ** 2
    patch = np.where(dist < rad ** 2, 2 * np.sqrt(np.maximum(rad ** 2 - (ps / 2 - i[None, :]) ** 2 - (ps / 2 - i[:, None]) ** 2, 0.0)), 0.0)
    rt = 2 * r0 / pix
    if 2 * rt > dim_x or 2 * rt > dim_y:
        raise Exception("The sample is too big for the detector field of view (increase dimX, dimY)")
    tube = np.zeros((dim_x, dim_y))
    if kind == 0:
        pos_y = dim_y // 2
        pos_a, pos_b = int(np.round(r0 * 3 / pix)), int(np.round(r0 * 7 / pix))
        j = np.arange(dim_y, dtype=np.float64)
        tube[:, :] = np.where(np.abs(dim_y / 2 - j) < rt, 2 * np.sqrt(np.maximum(rt ** 2 - (dim_y / 2 - j) ** 2, 0.0)), 0.0)[None, :]
    else:
        pos_y = dim_y // 2
        pos_a, pos_b = dim_x * 2 // 5, dim_x * 3 // 5
        for j in range(dim_x):                                               # :222 (sic: columns, bounded by dimX)
            if abs(dim_y / 2 - j) < rt * 3 / 4:
                tube[:, j] = rt * 2
            if rt > (j - dim_y / 2) >= rt * 3 / 4:
                tube[:, j] = rt / 2 * 3 + 2 * np.sqrt((rt / 4) ** 2 - (j - (dim_y / 2 + rt * 3 / 4)) ** 2)
            if -rt < j - dim_y / 2 <= -rt * 3 / 4:
                tube[:, j] = rt / 2 * 3 + 2 * np.sqrt((rt / 4) ** 2 - (j - (dim_y / 2 - rt * 3 / 4)) ** 2)
    out = np.zeros((3, dim_x, dim_y))
    out[0, pos_a - ps2:pos_a + ps2, pos_y - ps2:pos_y + ps2] = patch
    out[1, pos_b - ps2:pos_b + ps2, pos_y - ps2:pos_y + ps2] = patch
    out[2] = tube - out[0] - out[1]
    if kind == 1:
        mat = rotation_matrix((dim_y // 2, dim_x // 2), 15)
        out = np.stack([warp_affine_linear(out[m], mat, dim_y, dim_x) for m in range(3)])
        out = out[:, margin:dim_x - margin, margin:dim_y - margin]
    return out * pix * 1e-6


# --------------------------------------------------------------------------- orchestration
class Setup:
    """The scalars the per-energy loop needs (what Experiment.__init__ derives from the XML,
    Experiment.py:81-100, 204-216)."""

    def __init__(self, d_source_membrane, d_membrane_object, d_object_detector, det_dims, det_pixel_um,
                 oversampling, mean_shot_count, spectrum, source_size_um, psf_sigma, energy_sampling=1,
                 bin_thresholds=()):
        self.d1, self.d2, self.d3 = d_source_membrane, d_membrane_object, d_object_detector
        self.det_dims = (int(det_dims[0]), int(det_dims[1]))
        self.det_pixel_um = det_pixel_um
        self.os = int(oversampling)
        self.mean_shot_count = mean_shot_count
        self.spectrum = list(spectrum)
        self.source_size_um = source_size_um
        self.psf_sigma = psf_sigma
        self.energy_sampling = energy_sampling
        self.magnification = (self.d1 + self.d3 + self.d2) / (self.d1 + self.d2)
        self.study_dims = (self.det_dims[0] * self.os, self.det_dims[1] * self.os)
        self.study_pixel_um = det_pixel_um / self.os / self.magnification
        self.membrane_pixel_um = self.study_pixel_um * self.d1 / (self.d1 + self.d2)
        # Experiment.py:429 -- the last spectrum energy closes the last detector bin
        self.thresholds = list(bin_thresholds) + [self.spectrum[-1][0]]

    def effective_source_fwhm(self):
        """Experiment.py:503."""
        return self.source_size_um * self.d3 / (self.d1 + self.d2) / self.det_pixel_um * self.os


def compute_rt(setup, membrane_t, membrane_db, sample_t, sample_db, point_num, poisson=None):
    """Experiment.py:407-526 for a vacuum set-up without plate or scintillator.

    ``membrane_db`` / ``sample_db``: {energy: (deltas, betas)} per material.
    Returns (Sample, Reference, Propag, White)[nbins, dx, dy] plus the pre-detection
    accumulators of the last bin (for stage-level checks)."""
    s = setup
    nb = len(s.thresholds)
    shape = (nb,) + s.det_dims
    sample_img, ref_img, propag_img, white_img = (np.zeros(shape) for _ in range(4))
    n = s.study_dims
    i0 = np.ones(n) * (s.mean_shot_count / s.os ** 2)
    zeros = np.zeros(n)
    acc_s, acc_r, acc_p, acc_w = (np.zeros(n) for _ in range(4))
    ibin = 0
    pre = {}
    for energy, flux in s.spectrum:
        inc = i0 * flux
        md, mb = membrane_db[energy]
        sd, sb = sample_db[energy]
        i_m, phi_m = set_wave_rt(inc, zeros, membrane_t, md, mb, energy)
        i_bs, _, _ = fast_refraction(np.abs(i_m), phi_m, s.d2, energy, s.magnification, s.study_pixel_um)
        i_s, phi_ms = set_wave_rt(i_bs, phi_m, sample_t, sd, sb, energy)
        img_s, _, _ = fast_refraction(np.abs(i_s), phi_ms, s.d3, energy, s.magnification, s.study_pixel_um)
        img_r, _, _ = fast_refraction(np.abs(i_bs), phi_m, s.d3, energy, s.magnification, s.study_pixel_um)
        acc_s += img_s
        acc_r += img_r
        if point_num == 0:
            i_p, phi_p = set_wave_rt(inc, zeros, sample_t, sd, sb, energy)
            img_p, _, _ = fast_refraction(np.abs(i_p), phi_p, s.d3, energy, s.magnification, s.study_pixel_um)
            acc_w += inc
            acc_p += img_p
        if energy > s.thresholds[ibin] - s.energy_sampling / 2:
            fwhm = s.effective_source_fwhm()
            pre = dict(sample=acc_s.copy(), reference=acc_r.copy(), propag=acc_p.copy(), white=acc_w.copy())
            sample_img[ibin] = detection(acc_s, fwhm, s.os, s.det_dims, s.psf_sigma, poisson)
            ref_img[ibin] = detection(acc_r, fwhm, s.os, s.det_dims, s.psf_sigma, poisson)
            if point_num == 0:
                propag_img[ibin] = detection(acc_p, fwhm, s.os, s.det_dims, s.psf_sigma, poisson)
            white_img[ibin] = detection(acc_w, fwhm, s.os, s.det_dims, s.psf_sigma, poisson)
            acc_s, acc_r, acc_p, acc_w = (np.zeros(n) for _ in range(4))
            ibin += 1
    return sample_img, ref_img, propag_img, white_img, pre


def compute_fresnel(setup, membrane_t, membrane_db, sample_t, sample_db, point_num, poisson=None):
    """Experiment.py:279-405 for a vacuum set-up without plate or scintillator."""
    s = setup
    nb = len(s.thresholds)
    shape = (nb,) + s.det_dims
    sample_img, ref_img, propag_img, white_img = (np.zeros(shape) for _ in range(4))
    n = s.study_dims
    i0 = np.ones(n) * (s.mean_shot_count / s.os ** 2)
    acc_s, acc_r, acc_p, acc_w = (np.zeros(n) for _ in range(4))
    mag_mem_obj = (s.d1 + s.d2) / s.d1
    ibin = 0
    pre = {}

    def prop(w, z, e, m):
        return wave_propagation(w, z, e, m, s.study_dims, s.study_pixel_um)

    for energy, flux in s.spectrum:
        wave0 = np.sqrt(i0 * flux)
        md, mb = membrane_db[energy]
        sd, sb = sample_db[energy]
        after_mem = set_wave(wave0, membrane_t, md, mb, energy)
        before_sample = prop(after_mem, s.d2, energy, mag_mem_obj)
        after_sample = set_wave(before_sample, sample_t, sd, sb, energy)
        acc_s += np.abs(prop(after_sample, s.d3, energy, s.magnification)) ** 2
        acc_r += np.abs(prop(after_mem, s.d3 + s.d2, energy, s.magnification)) ** 2
        if point_num == 0:
            acc_p += np.abs(prop(set_wave(wave0, sample_t, sd, sb, energy), s.d3, energy, s.magnification)) ** 2
            acc_w += wave0 ** 2
        if energy > s.thresholds[ibin] - s.energy_sampling / 2:
            fwhm = s.effective_source_fwhm()
            pre = dict(sample=acc_s.copy(), reference=acc_r.copy(), propag=acc_p.copy(), white=acc_w.copy())
            sample_img[ibin] = detection(acc_s, fwhm, s.os, s.det_dims, s.psf_sigma, poisson)
            ref_img[ibin] = detection(acc_r, fwhm, s.os, s.det_dims, s.psf_sigma, poisson)
            if point_num == 0:
                propag_img[ibin] = detection(acc_p, fwhm, s.os, s.det_dims, s.psf_sigma, poisson)
            white_img[ibin] = detection(acc_w, fwhm, s.os, s.det_dims, s.psf_sigma, poisson)
            acc_s, acc_r, acc_p, acc_w = (np.zeros(n) for _ in range(4))
            ibin += 1
    return sample_img, ref_img, propag_img, white_img, pre
