"""Ray-tracing refraction model, B200 version -- drop-in for refractionFileNumba2.py.

Same module-level names and call signatures as the reference (refractionFileNumba2.py:14-333);
the arithmetic runs in hand-written sm_100a kernels behind the C ABI (``paresis_b200._cabi``).
"""
import numpy as np

import _paresis_path  # noqa: F401
from paresis_b200 import host_api, hostmath

MARGIN = 15  # refractionFileNumba2.py:50


def gaussian_shape(sigma):
    """refractionFileNumba2.py:14-23."""
    return hostmath.gaussian_2d(sigma)


def fastRefraction(intensityRefracted, phi, propagationDistance, Energy, magnification, studyPixelSize):
    """Intensity after free-space propagation from the refraction angles
    (refractionFileNumba2.py:25-86): returns (intensityRefracted2, Dx, Dy) with Dx, Dy
    zero-padded by 15 pixels, as the reference does.

    Raises:
        Exception: "The calculated intensity refractive includes some nans or insane values".
    """
    try:
        return host_api.fast_refraction(intensityRefracted, phi, propagationDistance, Energy, magnification,
                                        studyPixelSize, MARGIN)
    except host_api.InsaneValues as exc:
        raise Exception(str(exc))


def fastRefractionDF(intensityRefracted, phi, propagationDistance, Energy, magnification, studyPixelSize, darkField):
    """Dark-field variant (refractionFileNumba2.py:88-196)."""
    from paresis_b200 import darkfield
    return darkfield.fast_refraction_df(intensityRefracted, phi, propagationDistance, Energy, magnification,
                                        studyPixelSize, darkField)


def fastloopNumba(Nx, Ny, intensityRefracted, intensityRefracted2, Dy, Dx, DxFloor=None, DyFloor=None):
    """The bilinear scatter itself (refractionFileNumba2.py:198-263).  DxFloor / DyFloor are
    accepted and ignored, as in the reference kernel."""
    return host_api.splat(Nx, Ny, intensityRefracted, intensityRefracted2, Dy, Dx)


def fastloopNumbaDF(Nx, Ny, intensityRefracted, intensityRefracted2, Dy, Dx, DxFloor, DyFloor, DF):
    """refractionFileNumba2.py:266-333 (never called by the reference): rays with DF != 0 are skipped."""
    masked = np.where(np.asarray(DF) != 0, 0.0, np.asarray(intensityRefracted, dtype=np.float64))
    return host_api.splat(Nx, Ny, masked, intensityRefracted2, Dy, Dx)
