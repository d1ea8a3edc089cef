"""X-ray source description -- drop-in for the reference's Source.py (Source.py:18-293).

Runs once per experiment and only yields ``mySpectrum = [(E_keV, weight), ...]``, so it stays
on the host.  Monochromatic sources and .xls spectra (read with the in-repo BIFF8 reader instead
of xlrd) are handled here; tube spectra need ``spekpy`` exactly like upstream.  One extension:
an inline ``<spectrumTable>E:w,E:w,...</spectrumTable>`` node for synthetic spectra.
"""
import numpy as np

import _paresis_path  # noqa: F401
from paresis_b200.hostio import biff8, xmlparams


class Source:
    def __init__(self):
        self.xmlSourcesFileName = "xmlFiles/Sources.xml"
        self.myName = ""
        self.mySpectrum = []
        self.source_dict = {
            "mySize": 0.,
            "myEnergySampling": 1,
            "myType": None,
            "mySize_unit": "um",
            "myEnergySampling_unit": "keV",
            "myVoltage_unit": "kVp",
            "Energy_unit": "keV",
            "filterThickness_unit": "mm",
        }
        self.spectrumFromXls = False
        self.spectrumTable = None

    def defineCorrectValuesSource(self):
        """Source.py:38-77.  Raises ValueError("Source not found in the xml file")."""
        entry = xmlparams.find_entry(self.xmlSourcesFileName, "source", self.myName)
        if entry is None:
            raise ValueError("Source not found in the xml file")
        self.currentSource = entry.element
        sd = self.source_dict
        sd["mySize"] = entry.get("mySize", float)
        sd["myType"] = entry.get("myType")
        if sd["myType"] == "Polychromatic":
            sd["filterMaterial"] = None
            sd["myEnergySampling"] = entry.get("myEnergySampling", float)
            if entry.has("sourceVoltage"):
                sd["myVoltage"] = entry.get("sourceVoltage", float)
            if entry.has("spectrumFromXls"):
                self.spectrumFromXls = bool(entry.get("spectrumFromXls"))
                for key in ("pathXlsSpectrum", "energyUnit", "energyColumnKey", "fluenceColumnKey"):
                    sd[key] = entry.get(key)
            if entry.has("filterMaterial"):
                sd["filterMaterial"] = entry.get("filterMaterial")
                sd["filterThickness"] = entry.get("filterThickness", float)
            if entry.has("myTargetMaterial"):
                sd["myTargetMaterial"] = entry.get("myTargetMaterial")
            if entry.has("spectrumTable"):
                pairs = [p.split(":") for p in entry.get("spectrumTable").split(",")]
                self.spectrumTable = [(float(e), float(w)) for e, w in pairs]
        if sd["myType"] == "Monochromatic":
            sd["myEnergySampling"] = 1
            sd["Energy"] = entry.get("myEnergy", float)

    def setMySpectrum(self, flu_fluEn=True):
        """Source.py:79-245."""
        sd = self.source_dict
        if sd["myType"] == "Monochromatic":
            self.mySpectrum.append((sd["Energy"], 1))
            return
        if sd["myType"] != "Polychromatic":
            raise ValueError("type of source not recognized")
        if self.spectrumTable is not None:
            total = sum(w for _, w in self.spectrumTable)
            self.mySpectrum.extend((e, w / total) for e, w in self.spectrumTable)
        elif self.spectrumFromXls:
            self._spectrum_from_xls()
        else:
            self._spectrum_from_spekpy(flu_fluEn)

    def _spectrum_from_spekpy(self, flu_fluEn):
        """Source.py:97-122: tungsten-anode tube spectrum, weights above 1e-4 kept."""
        try:
            import spekpy as sp
        except ImportError as exc:
            raise ImportError("tube spectra need the 'spekpy' package (as in PARESIS); "
                              "use an .xls spectrum or <spectrumTable> instead") from exc
        sd = self.source_dict
        sd.setdefault("myTargetMaterial", "W")
        s = sp.Spek(kvp=sd["myVoltage"], th=12, targ=sd["myTargetMaterial"], dk=sd["myEnergySampling"])
        if sd["filterMaterial"] is not None:
            s.filter(sd["filterMaterial"], sd["filterThickness"])
        energies, fluence = s.get_spectrum(flu=flu_fluEn)
        fluence = np.where(np.isnan(fluence), 0.0, fluence)
        total = float(np.sum(fluence))
        for e, f in zip(energies, fluence):
            if f / total > 0.0001:
                self.mySpectrum.append((e, f / total))

    def _spectrum_from_xls(self):
        """Source.py:131-231: read (energy, fluence) columns, re-bin to myEnergySampling keV,
        keep bins carrying more than 1e-3 of the flux."""
        sd = self.source_dict
        scale = {"eV": 0.001, "MeV": 1000}.get(sd["energyUnit"], 1)
        spectrum = []
        for sh in biff8.open_workbook(sd["pathXlsSpectrum"]).sheets():
            col_e = col_f = start = None
            for row in range(sh.nrows):
                for col in range(sh.ncols):
                    v = sh.cell(row, col).value
                    if v == sd["energyColumnKey"]:
                        col_e, start = col, row
                    if v == sd["fluenceColumnKey"]:
                        col_f = col
                if col_e is not None and col_f is not None:
                    break
            if col_e is None:
                raise Exception(f'Energy column key {sd["energyColumnKey"]} not found in the xls file')
            if col_f is None:
                raise Exception(f'Energy column key {sd["fluenceColumnKey"]} not found in the xls file')
            for row in range(start + 1, sh.nrows):
                spectrum.append([sh.cell(row, col_e).value * scale, sh.cell(row, col_f).value])
        step = spectrum[1][0] - spectrum[0][0]
        n_e = len(spectrum)
        n_bins = int((spectrum[-1][0] - spectrum[0][0]) // sd["myEnergySampling"])
        energies, weights = [], []
        n, tot = 0, 0
        for _ in range(n_bins - 1):
            width = w_bin = e_bin = 0
            while width < sd["myEnergySampling"]:
                w_bin += spectrum[n][1]
                e_bin += spectrum[n][1] * spectrum[n][0]
                n += 1
                width += step
            if w_bin != 0:
                energies.append(e_bin / w_bin)
                weights.append(w_bin)
            tot += w_bin
        w_bin = e_bin = 0
        while n < n_e:          # the tail bin is not added to the normalisation, as upstream
            w_bin += spectrum[n][1]
            e_bin += spectrum[n][1] * spectrum[n][0]
            n += 1
        if w_bin != 0:
            energies.append(e_bin / w_bin)
            weights.append(w_bin)
        for e, w in zip(energies, weights):
            if w / tot > 0.001:
                self.mySpectrum.append((e, w / tot))

    def getText(self, node):
        return xmlparams.text_of(node)
