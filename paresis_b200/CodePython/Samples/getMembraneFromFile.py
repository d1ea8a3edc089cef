"""Membrane thickness maps -- drop-in for Samples/getMembraneFromFile.py.

The sphere-cap rasterisation, a pure-Python triple loop upstream
(getMembraneFromFile.py:143-159, ~30 s per position at 2048^2), is one CUDA kernel here.
"""
import glob

import numpy as np

import _paresis_path  # noqa: F401
from InputOutput.pagailleIO import openImage
from paresis_b200 import geometry


def getMembraneFromFile(myMembraneFile, studyDimensions, numPoint, supportThickness):
    """Load a pre-rendered thickness map (getMembraneFromFile.py:18-57).

    Raises:
        ValueError: The membrane you are trying to load does not have the correct dimensions.
    """
    paths = sorted(glob.glob(myMembraneFile + '/*.tif') + glob.glob(myMembraneFile + '/*.tiff') +
                   glob.glob(myMembraneFile + '/*.edf'))
    thickness = np.asarray(openImage(paths[numPoint]), dtype=float)
    if studyDimensions[0] != thickness.shape[0] or studyDimensions[1] != thickness.shape[1]:
        raise ValueError("The membrane you are trying to load does not have the correct dimensions")
    geom = geometry.from_host(np.array([thickness, np.ones(thickness.shape) * supportThickness * 1e-6]))
    return geom, {"Membrane geometry folder": (paths, ''), "Support thickness": (supportThickness, 'um')}


def getMembraneSegmentedFromFile(sample, dimX, dimY, pixSize, pointNum, supportThickness):
    """Membrane thickness from the segmented sphere list (getMembraneFromFile.py:60-171).

    Returns:
        geometry: [grains, support] thickness maps in metres (device-resident, ndarray-like).
        parameters_dic (dict): values written to the run report.
    """
    return geometry.membrane_segmented(sample, dimX, dimY, pixSize, pointNum, supportThickness, prefetch=True)
