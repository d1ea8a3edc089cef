"""TEST INFRASTRUCTURE ONLY -- loader for the unmodified PARESIS reference (/root/reference).

Only ``oracle/make_golden.py`` and the in-container oracle-pinning tests use this module.
It exists solely where ``/root/reference`` is mounted (the build container); nothing on the
GPU box may import it.  It never copies reference sources: it puts
``/root/reference/CodePython`` on ``sys.path``, runs from a scratch working directory whose
relative paths (``xmlFiles/``, ``Samples/DeltaBeta/``) are symlinks into the read-only tree,
and repairs only the non-numerical packaging gaps listed in SURVEY.md App. B-1:

* ``np.int`` / ``np.float`` aliases (removed from numpy >= 1.24; used at
  ``refractionFileNumba2.py:72-73``, ``getMembraneFromFile.py:43``),
* no-op ``matplotlib`` (the reference calls blocking ``plt.show()``),
* capturing ``fabio`` (TIFF/EDF writer), a 4-line ``imutils.rotate``, empty ``skimage`` /
  ``spekpy``,
* ``xlrd`` served by the repo's own BIFF8 reader and ``xraylib.Refractive_Index`` served from
  the same table, so both sides see one frozen (E, delta, beta) source,
* a seeded synthetic ``Samples/Membranes/CuSn.txt`` (the real list is a missing large blob).
"""
import json
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = "/root/reference/CodePython"
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _REPO not in sys.path:
    sys.path.insert(0, _REPO)

SAVED_IMAGES = {}  # filename -> ndarray captured by the fabio stub


def reference_available():
    return os.path.isdir(REFERENCE_ROOT)


def synthetic_sphere_rows(seed=0, count=60000):
    """Rows [c0 (y), c1 (x), radius] in the file's own units."""
    rng = np.random.default_rng(seed)
    rows = np.empty((count, 3))
    rows[:, 0] = rng.uniform(-4870.0, 4870.0, count)
    rows[:, 1] = rng.uniform(-4051.0, 4051.0, count)
    rows[:, 2] = rng.gamma(4.0, 3.2, count)
    return rows


class _Anything:
    """Attribute sink: every attribute is a callable returning another sink."""

    def __getattr__(self, name):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _install_stubs():
    if not hasattr(np, "int"):
        np.int = int
    if not hasattr(np, "float"):
        np.float = float

    plt = _module("matplotlib.pyplot")
    plt.__getattr__ = lambda name: (lambda *a, **k: _Anything())
    mpl = _module("matplotlib", pyplot=plt)
    mpl.__getattr__ = lambda name: _Anything()

    class _Img:
        def __init__(self, data=None, header=None):
            self.data = data
            self.header = header

        def write(self, filename):
            SAVED_IMAGES[str(filename)] = np.array(self.data)

    edf = _module("fabio.edfimage", EdfImage=_Img, EdfFile=_Img)
    tif = _module("fabio.tifimage", TifImage=_Img)

    def _open(filename):
        return _Img(data=SAVED_IMAGES[str(filename)])

    _module("fabio", edfimage=edf, tifimage=tif, open=_open)

    import cv2

    def rotate(image, angle, center=None, scale=1.0):
        h, w = image.shape[:2]
        if center is None:
            center = (w // 2, h // 2)
        return cv2.warpAffine(image, cv2.getRotationMatrix2D(center, angle, scale), (w, h))

    _module("imutils", rotate=rotate)
    tr = _module("skimage.transform")
    tr.__getattr__ = lambda name: _Anything()
    _module("skimage", transform=tr)
    sp = _module("spekpy")
    sp.__getattr__ = lambda name: _Anything()

    from paresis_b200.hostio import biff8

    _module("xlrd", open_workbook=biff8.open_workbook)

    table_cache = {}

    def _table(material):
        if not table_cache:
            wb = biff8.open_workbook(os.path.join(REFERENCE_ROOT, "Samples/DeltaBeta/TablesDeltaBeta.xls"))
            sh = wb.sheets()[0]
            for col in range(sh.ncols):
                name = sh.cell(0, col).value
                if isinstance(name, str) and name.strip():
                    rows = []
                    r = 3
                    while r < sh.nrows and col + 2 < sh.ncols:
                        trio = [sh.cell(r, col + k).value for k in range(3)]
                        if not all(isinstance(v, float) for v in trio):
                            break
                        rows.append(trio)
                        r += 1
                    if rows:
                        table_cache[name.strip()] = np.array(rows, dtype=float)
        return table_cache[material]

    formula_to_material = {"H0.080538C0.599848O0.319614": "PMMA", "C0.977O0.023": "CarbonFiber", "Cs1I1": "CsI", "C": "Carbon"}

    def Refractive_Index(formula, energy_kev, density):
        t = _table(formula_to_material[formula])
        d = np.interp(energy_kev * 1e3, t[:, 0], t[:, 1])
        b = np.interp(energy_kev * 1e3, t[:, 0], t[:, 2])
        return complex(1.0 - d, b)

    _module("xraylib", Refractive_Index=Refractive_Index)


def make_scratch(path, sphere_seed=0, sphere_count=60000):
    """CodePython-shaped working directory: symlinks for data, synthetic sphere list."""
    os.makedirs(os.path.join(path, "Samples", "Membranes"), exist_ok=True)
    for rel in ("xmlFiles", "Samples/DeltaBeta", "Sources"):
        dst = os.path.join(path, rel)
        if not os.path.lexists(dst):
            os.symlink(os.path.join(REFERENCE_ROOT, rel), dst)
    with open(os.path.join(path, "Samples", "Membranes", "CuSn.txt"), "w") as fh:
        json.dump(synthetic_sphere_rows(sphere_seed, sphere_count).tolist(), fh)
    os.makedirs(os.path.join(path, "..", "Results", "Fil_Nylon_ID17"), exist_ok=True)
    return path


_loaded = {}


def load_reference(scratch="/tmp/paresis_ref_scratch", sphere_seed=0, sphere_count=60000):
    """Import the reference modules (once) and return them in a namespace dict."""
    if _loaded:
        return _loaded
    if not reference_available():
        raise RuntimeError("reference tree is not mounted at " + REFERENCE_ROOT)
    _install_stubs()
    make_scratch(scratch, sphere_seed, sphere_count)
    os.chdir(scratch)
    sys.path.insert(0, REFERENCE_ROOT)
    import importlib

    for name in ("getk", "refractionFileNumba2", "refractionFileNumba", "Detector", "Sample", "Source", "Experiment"):
        _loaded[name] = importlib.import_module(name)
    _loaded["membrane"] = importlib.import_module("Samples.getMembraneFromFile")
    _loaded["geom"] = importlib.import_module("Samples.createSampGeom")
    _loaded["scratch"] = scratch
    return _loaded


class identity_poisson:
    """Context manager: RandomState.poisson -> identity, for noise-free end-to-end goldens
    (the reference seeds from the wall clock, Detector.py:113)."""

    def __enter__(self):
        class _RS:
            def __init__(self, seed=None):
                pass

            def poisson(self, lam):
                return lam

        self._orig = np.random.RandomState
        np.random.RandomState = _RS

    def __exit__(self, *exc):
        np.random.RandomState = self._orig
