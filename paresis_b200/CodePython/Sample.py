"""Samples and membranes -- drop-in for the reference's Sample.py (Sample.py:22-353).

Thickness maps live in HBM (``paresis_b200.geometry.DeviceGeometry``); ``myGeometry`` still
indexes and converts like the reference's ``[n_mat, N, N]`` ndarray.  ``setWave`` / ``setWaveRT``
keep their numpy-in / numpy-out signatures and run as CUDA kernels.
"""
import numpy as np
import pandas as pd

import _paresis_path  # noqa: F401
from Samples.getMembraneFromFile import getMembraneFromFile, getMembraneSegmentedFromFile
from Samples.createSampGeom import (CreateSampleSpheresInParallelepiped, CreateSampleSpheresInCylinder,
                                    CreateSampleCylindre, CreateSampleSphere, CreateYourSampleGeometry,
                                    loadSampleGeometryFromImages)
from paresis_b200 import geometry, host_api
from paresis_b200.hostio import tables, xmlparams


class Sample:
    def __init__(self):
        self.xmlSampleFileName = "xmlFiles/Samples.xml"
        self.myName = ""
        self.myType = ""
        self.myMaterials = []
        self.myGeometry = []
        self.geom_parameters = None

    def defineCorrectValuesSample(self):
        """Sample.py:33-77.  Raises ValueError("Sample not found in the xml file")."""
        entry = xmlparams.find_entry(self.xmlSampleFileName, "sample", self.myName)
        if entry is None:
            print(self.myName)
            raise ValueError("Sample not found in the xml file")
        self.myType = entry.get("myType")
        self.myMaterials = list(entry.get("myMaterials").split(","))
        self.myGeometryFunction = fn = entry.get("myGeometryFunction")
        if fn == "getMembraneFromFile":
            self.myPMMAThickness = entry.get("myPMMAThickness", float)
            if entry.has("myMembraneFile"):     # upstream never reads it; harmless to accept
                self.myMembraneFile = entry.get("myMembraneFile")
        if fn == "getMembraneSegmentedFromFile":
            self.myMeanSphereRadius = entry.get("myMeanSphereRadius", float)
            self.myNbOfLayers = entry.get("myNbOfLayers", int)
            self.myPMMAThickness = entry.get("myPMMAThickness", float)
        if fn == "get_my_thickness" and self.myName != "air_volume":
            self.myThickness = entry.get("myThickness", float)
        if fn == "getSampleFromFile":
            self.mySampleFile = entry.get("mySampleFile")
        if fn == "loadSampleGeometryFromImages":
            self.myGeometryFolder = entry.get("myGeometryFolder")

    def getText(self, node):
        return xmlparams.text_of(node)

    def getDeltaBeta(self, sourceSpectrum):
        """delta and beta of every material at every spectrum energy (Sample.py:83-152).

        Materials listed in Samples/DeltaBeta/Materials.csv go through xraylib when it is
        installed (as upstream); otherwise, and for all other materials, the tabulated values are
        interpolated linearly like Sample.py:121-143."""
        energies = [e for e, _ in sourceSpectrum]
        formulas = None
        try:
            import xraylib as xrl
            formulas = pd.read_csv('Samples/DeltaBeta/Materials.csv').set_index('Material')
        except Exception:
            xrl = None
        for material in self.myMaterials:
            if xrl is not None and formulas is not None and material in formulas.index.tolist():
                print(f'{material} in Materials.csv')
                ns = [xrl.Refractive_Index(formulas['Formula'][material], e, formulas['Density'][material]) for e in energies]
                self.delta.append([(e, 1 - n.real) for e, n in zip(energies, ns)])
                self.beta.append([(e, n.imag) for e, n in zip(energies, ns)])
                continue
            found = tables.interpolate(material, energies)
            if found is not None:
                print(f'{material} in delta/beta tables')
                self.delta.append([(e, d) for e, (d, _) in zip(energies, found)])
                self.beta.append([(e, b) for e, (_, b) in zip(energies, found)])
        if np.shape(self.delta)[0] != len(self.myMaterials):
            raise ValueError("One or more materials have not been found in delta beta tables")


class AnalyticalSample(Sample):
    def __init__(self):
        Sample.__init__(self)
        self.delta = []
        self.beta = []

    def getMyGeometry(self, studyDimensions, studyPixelSize, oversamp, pointNum=0, number_of_positions=0):
        """Thickness map of each material, geometry[material, x, y] in metres (Sample.py:163-245).

        Raises:
            ValueError: Could not define sample geometry.
        """
        dx, dy = int(studyDimensions[0]), int(studyDimensions[1])
        fn = self.myGeometryFunction
        if self.myType == "sample_of_interest":
            makers = {"CreateSampleCylindre": CreateSampleCylindre, "CreateYourSampleGeometry": CreateYourSampleGeometry,
                      "CreateSampleSpheresInCylinder": CreateSampleSpheresInCylinder,
                      "CreateSampleSpheresInParallelepiped": CreateSampleSpheresInParallelepiped,
                      "CreateSampleSphere": CreateSampleSphere}
            if fn == "getSampleFromFile":
                self.myGeometry = np.load(self.mySampleFile)
                return
            if fn in makers:
                self.myGeometry, self.geom_parameters = makers[fn](self.myName, dx, dy, studyPixelSize)
                return
            if fn == "loadSampleGeometryFromImages":
                g, self.geom_parameters = loadSampleGeometryFromImages(self.myGeometryFolder, dx, dy, studyPixelSize)
                self.myGeometry = np.array(g)
                return
            if fn == "generateContrastPhantom":
                raise NotImplementedError("generateContrastPhantom needs scikit-image's radon transform and is outside "
                                          "the accelerated path (SURVEY.md section 2, #10)")
        if self.myType == "membrane":
            if fn == "getMembraneFromFile":
                self.myGeometry, self.geom_parameters = getMembraneFromFile(self.myMembraneFile, studyDimensions, pointNum,
                                                                            self.myPMMAThickness)
                return
            if fn == "getMembraneSegmentedFromFile":
                self.myGeometry, self.geom_parameters = getMembraneSegmentedFromFile(self, dx, dy, studyPixelSize, pointNum,
                                                                                     self.myPMMAThickness)
                return
        if fn == "get_my_thickness":
            self.myGeometry = geometry.DeviceGeometry([self.myThickness * 1e-6], (dx, dy))
            return
        raise ValueError("Could not define sample geometry")

    # ------------------------------------------------------------------ internals
    def _device_geometry(self):
        g = self.myGeometry
        if isinstance(g, geometry.DeviceGeometry):
            return g
        if np.ndim(g) != 3:
            raise Exception("Sample Geometry has the wrong nb of dim [material, x, y]")
        g = geometry.from_host(g)
        self.myGeometry = g
        return g

    def _coefficients(self, energy):
        """delta, beta per material by exact-energy lookup; a missing energy leaves 0 (Sample.py:266-277)."""
        delta = np.zeros(len(self.myMaterials))
        beta = np.zeros(len(self.myMaterials))
        for imat in range(len(self.myMaterials)):
            for e, v in self.delta[imat]:
                if e == energy:
                    delta[imat] = v
            for e, v in self.beta[imat]:
                if e == energy:
                    beta[imat] = v
        return delta, beta

    def _has_dark_field(self):
        return self.myType == "sample_of_interest" and ("Lung" in self.myMaterials or self.myName == 'cylinder_beeds')

    # ------------------------------------------------------------------ public arithmetic
    def setWave(self, incidentWave, energy):
        """Complex transmission through the object (Sample.py:248-282)."""
        geom = self._device_geometry()
        delta, beta = self._coefficients(energy)
        maps = geom.device_entries(materialise=True)
        return _set_wave_device(incidentWave, maps, delta, beta, energy, geom.map_shape)

    def setWaveRT(self, incidentIntensity, energy, incidentphi=0, incidentDf=0):
        """Attenuated intensity and accumulated phase behind the object (Sample.py:285-351).

        Returns (intensity, phi, darkField); darkField is the int 0 except for the Lung /
        'cylinder_beeds' dark-field model (:322-343)."""
        geom = self._device_geometry()
        delta, beta = self._coefficients(energy)
        if self._has_dark_field():
            from paresis_b200 import darkfield
            return darkfield.set_wave_rt(self, geom, incidentIntensity, energy, incidentphi, delta, beta)
        maps = geom.device_entries(materialise=True)
        i_out, phi_out = _set_wave_rt_device(incidentIntensity, incidentphi, maps, delta, beta, energy, geom.map_shape)
        return i_out, phi_out, 0


def _set_wave_rt_device(intensity, phi, maps, delta, beta, energy, shape):
    import torch
    from paresis_b200 import _cabi as abi, hostmath
    k = hostmath.wavenumber(energy * 1000)
    dev = host_api.device()
    i_in = host_api.to_dev(np.broadcast_to(np.asarray(intensity, dtype=np.float64), shape))
    phi_in = None
    if not (np.isscalar(phi) and phi == 0):
        phi_in = host_api.to_dev(np.broadcast_to(np.asarray(phi, dtype=np.float64), shape), torch.float64)
    i_out = torch.empty(shape, device=dev, dtype=torch.float32)
    phi_out = torch.empty(shape, device=dev, dtype=torch.float64)
    abi.transmit_rt(i_in, phi_in, maps, [2 * k * b for b in beta], [k * d for d in delta], i_out, phi_out)
    return host_api.to_host(i_out), host_api.to_host(phi_out)


def _set_wave_device(wave, maps, delta, beta, energy, shape):
    import torch
    from paresis_b200 import _cabi as abi, hostmath
    k = hostmath.wavenumber(energy * 1000)
    w_in = host_api.to_dev(np.broadcast_to(np.asarray(wave, dtype=np.complex128), shape), torch.complex64)
    w_out = torch.empty(shape, device=host_api.device(), dtype=torch.complex64)
    abi.transmit_wave(w_in, 0.0, maps, [k * b for b in beta], [k * d for d in delta], w_out)
    return host_api.to_host(w_out, np.complex128)
