"""Host-array entry points: numpy in -> CUDA kernels -> numpy out.

These back the module-level functions and methods of the drop-in ``CodePython`` shim whose
reference signatures take and return numpy arrays.  Each call copies
its inputs to the GPU, runs the kernels through the C ABI and copies the result back; the
per-position hot loop does not go through here (see ``engine.py``), these are the
compatibility surface for user code that calls the pieces one by one.
"""
import time

import numpy as np
import torch

from . import _cabi as abi
from . import hostmath as hm


def device():
    if not torch.cuda.is_available():
        raise RuntimeError("paresis_b200 needs a CUDA device (B200); there is no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


def to_dev(a, dtype=torch.float32):
    from . import transfer
    return transfer.upload(a, dtype, device())


def to_host(t, dtype=np.float64):
    from . import transfer
    torch_dtype = {np.float64: torch.float64, np.float32: torch.float32, np.complex128: torch.complex128}.get(dtype)
    return transfer.fetch(t, torch_dtype) if torch_dtype is not None else t.cpu().numpy().astype(dtype)


class InsaneValues(Exception):
    pass


def splat(nx, ny, intensity, out, dy, dx):
    """fastloopNumba(Nx, Ny, I, I2, Dy, Dx, DxFloor, DyFloor) -- refractionFileNumba2.py:198-263.
    ``out`` is updated in place and returned, like the Numba kernel does."""
    intensity = np.asarray(intensity)
    if intensity.shape != (nx, ny):
        raise ValueError("intensity shape %s does not match (Nx, Ny) = (%d, %d)" % (intensity.shape, nx, ny))
    d_out = to_dev(out)
    abi.splat(to_dev(intensity), to_dev(dx), to_dev(dy), d_out, margin=0)
    out[...] = to_host(d_out, out.dtype)
    return out


def fast_refraction(intensity, phi, distance, energy_kev, magnification, pixel_um, margin, clamp_px=0.0):
    """fastRefraction -- refractionFileNumba2.py:25-86 (margin 15) / refractionFileNumba.py:11-68
    (margin 10, clamp 1e3).  Returns (I2[N,N], Dx[N+2m,N+2m], Dy) as float64 host arrays."""
    intensity = np.asarray(intensity, dtype=np.float64)
    nx, ny = intensity.shape
    dev = device()
    out = torch.zeros((nx, ny), device=dev, dtype=torch.float32)
    dxp = torch.zeros((nx + 2 * margin, ny + 2 * margin), device=dev, dtype=torch.float32)
    dyp = torch.zeros_like(dxp)
    flag = torch.zeros(1, device=dev, dtype=torch.int32)
    abi.refract_phi(to_dev(intensity), to_dev(phi, torch.float64), out, distance, energy_kev, magnification, pixel_um,
                    margin, dxp, dyp, flag, clamp_px)
    if int(flag.item()) & abi.FLAG_NONFINITE:
        raise InsaneValues("The calculated intensity refractive includes some nans or insane values")
    return to_host(out), to_host(dxp), to_host(dyp)


def bin_sum(image, size_x, size_y):
    """resize -- Detector.py:185-198."""
    image = np.asarray(image)
    if image.shape == (size_x, size_y):
        return image
    out = torch.empty((int(size_x), int(size_y)), device=device(), dtype=torch.float32)
    abi.bin_sum(to_dev(image), int(size_x), int(size_y), out)
    return to_host(out)


def detection(image, source_fwhm_px, oversampling, det_dims, psf_sigma, poisson=True, seed=None, sequence=0):
    """Detector.detection -- Detector.py:79-119.  Returns float64 counts (noise-free expectation
    when ``poisson`` is False)."""
    image = np.asarray(image)
    dx, dy = int(det_dims[0]), int(det_dims[1])
    dev = device()
    src = psf = None
    if source_fwhm_px != 0 and hm.gaussian_half_width(source_fwhm_px / 2.355) > 0:
        src = to_dev(hm.gaussian_1d(source_fwhm_px / 2.355))
    if psf_sigma != 0 and hm.gaussian_half_width(psf_sigma) > 0:
        psf = to_dev(hm.gaussian_1d(psf_sigma))
    work = torch.empty(abi.detect_work_floats(image.shape[0], image.shape[1], int(oversampling), dx, dy),
                       device=dev, dtype=torch.float32)
    out = torch.empty((dx, dy), device=dev, dtype=torch.float32)
    if seed is None:
        seed = int(np.floor(time.time() * 100 % (2 ** 32 - 1)))   # Detector.py:113
    abi.detect_counts(to_dev(image), int(oversampling), dx, dy, src, psf, work, out, poisson, seed, sequence)
    return to_host(out)


_plans = {}


def wave_propagation(wave, distance, energy_kev, magnification, study_dims, pixel_um, margin=15):
    """Experiment.wavePropagation -- Experiment.py:219-252."""
    if distance == 0:
        return wave
    wave = np.asarray(wave)
    nx, ny = wave.shape
    key = (nx, ny, margin, torch.cuda.current_device())
    if key not in _plans:
        _plans[key] = abi.FresnelPlan(nx, ny, margin)
    hx, hy, phase = hm.fresnel_vectors(nx, ny, margin, study_dims, pixel_um, distance, energy_kev, magnification)
    out = torch.empty((nx, ny), device=device(), dtype=torch.complex64)
    _plans[key].propagate(to_dev(wave, torch.complex64), to_dev(hx, torch.complex64), to_dev(hy, torch.complex64),
                          phase, out, None)
    return to_host(out, np.complex128)
