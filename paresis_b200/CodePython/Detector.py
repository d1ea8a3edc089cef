"""Detector model, B200 version -- drop-in for the reference's Detector.py.

Same class, attribute and function names (Detector.py:19-220).  ``detection`` runs the
blur / bin / blur / Poisson chain in CUDA kernels (csrc/detector.cu); the scintillator
efficiency scalars are host arithmetic.
"""
import os

import numpy as np

import _paresis_path  # noqa: F401
from getk import getk
from paresis_b200 import host_api, hostmath
from paresis_b200.hostio import tables, xmlparams


class Detector:
    def __init__(self, exp_dict):
        self.xmlDetectorFileName = "xmlFiles/Detectors.xml"
        self.myName = ""
        self.det_param = {
            "myDimensions": (0, 0),
            "myPixelSize": 0.,             # um
            "myPSF": 0.,                   # pixels
            "myBinsThersholds": [],        # keV (spelling as in the reference)
            "myScintillatorMaterial": None,
            "myScintillatorThickness": 0.,  # um
            "photonCounting": True,
            "myDimensions_unit": "pixels",
            "myPixelSize_unit": "um",
            "myPSF_unit": "pixels",
            "myBinsThersholds_unit": "keV",
            "myScintillatorThickness_unit": "um",
        }
        self.mySpectralEfficiency = []
        self.beta = []
        # extensions (ignored by the reference): deterministic / noise-free detection
        self.poissonNoise = bool(exp_dict.get("poissonNoise", True)) if isinstance(exp_dict, dict) else True
        # test hook for drivers that build their own exp_dict (PARESIS's unmodified main.py): noise-free images
        if os.environ.get("PARESIS_B200_POISSON", "") == "0":
            self.poissonNoise = False
        self.seed = exp_dict.get("seed") if isinstance(exp_dict, dict) else None
        self._draws = 0

    def defineCorrectValuesDetector(self):
        """Detector.py:44-76.  Raises ValueError("detector not found in xml file")."""
        entry = xmlparams.find_entry(self.xmlDetectorFileName, "detector", self.myName)
        if entry is None:
            raise ValueError("detector not found in xml file")
        self.det_param["myDimensions"] = self.getMyDimensions(entry.element)
        self.det_param["myPixelSize"] = entry.get("myPixelSize", float)
        self.det_param["myPSF"] = entry.get("myPSF", float)
        if entry.has("myEnergyLimit"):
            self.myEnergyLimit = entry.get("myEnergyLimit", float)
        if entry.has("photonCounting"):
            self.det_param["photonCounting"] = bool(entry.get("photonCounting"))   # bool("False") is True, as upstream
        if entry.has("myBinsThersholds"):
            self.det_param["myBinsThersholds"] = [float(v) for v in entry.get("myBinsThersholds").split(",")]
        if entry.has("myScintillatorMaterial"):
            self.det_param["myScintillatorMaterial"] = entry.get("myScintillatorMaterial")
            self.det_param["myScintillatorThickness"] = entry.get("myScintillatorThickness", float)

    def detection(self, incidentWave, effectiveSourceSize, exp_param):
        """Source blur, binning to detector pixels, PSF blur, shot noise (Detector.py:79-119).

        Args:
            incidentWave (2d numpy array): intensity arriving at the detector (oversampled grid).
            effectiveSourceSize (float): projected source FWHM in oversampled pixels.
            exp_param (dict): needs 'overSampling'.
        Returns:
            2d numpy array [dimX, dimY] of counts.
        """
        self._draws += 1
        return host_api.detection(incidentWave, effectiveSourceSize, exp_param["overSampling"],
                                  self.det_param["myDimensions"], self.det_param["myPSF"],
                                  poisson=self.poissonNoise, seed=self.seed, sequence=(1 << 40) + self._draws)

    def getText(self, node):
        return xmlparams.text_of(node)

    def getMyDimensions(self, node):
        dimX = int(self.getText(node.getElementsByTagName("dimX")[0]))
        dimY = int(self.getText(node.getElementsByTagName("dimY")[0]))
        return np.array([dimX, dimY])

    def getBeta(self, sourceSpectrum):
        """Scintillator beta per spectrum energy (Detector.py:131-160)."""
        material = self.det_param["myScintillatorMaterial"]
        found = tables.interpolate(material, [e for e, _ in sourceSpectrum])
        if found is None:
            raise ValueError("The scintillator material has not been found in delta beta tables")
        self.beta.extend((e, b) for e, (_, b) in zip([e for e, _ in sourceSpectrum], found))

    def getSpectralEfficiency(self):
        """1 - exp(-2 k t beta) per energy (Detector.py:163-182)."""
        for energy, beta in self.beta:
            k = getk(energy * 1000)
            eff = 1 - np.exp(-2 * k * self.det_param["myScintillatorThickness"] * 1e-6 * beta)
            self.mySpectralEfficiency.append((energy, eff))
        if len(self.mySpectralEfficiency) == 1:
            print(f'Scintillator attenuation at {self.mySpectralEfficiency[0][0]} keV: {self.mySpectralEfficiency[0][1]}')


def resize(imageToResize, sizeX, sizeY):
    """Sum-binning to (sizeX, sizeY) (Detector.py:185-198); identity when the shape already matches."""
    return host_api.bin_sum(imageToResize, sizeX, sizeY)


def create_gaussian_shape(sigma):
    """Normalised (2*round(3*sigma)+1)^2 Gaussian (Detector.py:201-220)."""
    return hostmath.gaussian_2d(sigma)
