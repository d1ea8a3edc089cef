"""First-generation refraction module -- drop-in for refractionFileNumba.py.

Differs from refractionFileNumba2 only by the 10-pixel working margin (refractionFileNumba.py:36)
and the fixed 1e3-pixel ray clamp (:46-49).
"""
import _paresis_path  # noqa: F401
from paresis_b200 import host_api
from refractionFileNumba2 import fastloopNumba  # noqa: F401  (identical kernel, refractionFileNumba.py:70-135)

MARGIN = 10


def fastRefraction(intensityRefracted, phi, propagationDistance, Energy, magnification, studyPixelSize):
    """refractionFileNumba.py:11-68."""
    try:
        return host_api.fast_refraction(intensityRefracted, phi, propagationDistance, Energy, magnification,
                                        studyPixelSize, MARGIN, clamp_px=1e3)
    except host_api.InsaneValues as exc:
        raise Exception(str(exc))
