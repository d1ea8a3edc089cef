"""delta / beta tables without xlrd or xraylib.

Resolution order for a material name:
  1. ``Samples/DeltaBeta/TablesDeltaBeta.xls`` in the working directory (the user's own
     PARESIS data file), read with the in-repo BIFF8 reader;
  2. ``Samples/DeltaBeta/delta_beta_tables.npz`` (working directory, then the copy shipped
     next to the shim), an export of the same sheet made by tools/export_delta_beta.py.
Interpolation follows Sample.py:121-143 step by step, including its stateful row cursor.
"""
import os

import numpy as np

from . import biff8

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIPPED = os.path.join(os.path.dirname(_HERE), "CodePython", "Samples", "DeltaBeta", "delta_beta_tables.npz")
_cache = {}


def _load_all():
    key = os.getcwd()
    if key in _cache:
        return _cache[key]
    out = {}
    xls = os.path.join("Samples", "DeltaBeta", "TablesDeltaBeta.xls")
    if os.path.exists(xls):
        for sheet in biff8.open_workbook(xls).sheets():
            for col in range(sheet.ncols):
                name = sheet.cell(0, col).value
                if not isinstance(name, str) or not name or col + 2 >= sheet.ncols or name in out:
                    continue
                rows, r = [], 3
                while r < sheet.nrows:
                    trio = [sheet.cell(r, col + k).value for k in range(3)]
                    if not all(isinstance(v, float) for v in trio):
                        break
                    rows.append(trio)
                    r += 1
                if len(rows) > 1:
                    out[name] = np.array(rows, dtype=np.float64)
    else:
        for path in (os.path.join("Samples", "DeltaBeta", "delta_beta_tables.npz"), _SHIPPED):
            if os.path.exists(path):
                with np.load(path) as z:
                    out = {k: z[k] for k in z.files}
                break
    _cache[key] = out
    return out


def has_material(name):
    return name in _load_all()


def interpolate(material, energies_kev):
    """[(delta, beta)] for each energy, or None when the material is unknown.
    Below the table's first energy the reference returns (0, 1) (Sample.py:124-128)."""
    table = _load_all().get(material)
    if table is None:
        return None
    out = []
    row = 0
    last = len(table) - 1
    for energy in energies_kev:
        e_ev = energy * 1000
        cur = table[row, 0]
        if e_ev < cur:
            print("No delta beta values under", cur, "eV")
            out.append((0, 1))
            continue
        nxt = table[row + 1, 0]
        while nxt < energy * 1e3:
            row += 1
            if row + 1 > last:
                raise IndexError("energy %g keV beyond the delta/beta table of %s" % (energy, material))
            cur, nxt = table[row, 0], table[row + 1, 0]
        step = nxt - cur
        wa, wb = abs(nxt - energy * 1e3) / step, abs(cur - energy * 1e3) / step
        out.append((wa * table[row, 1] + wb * table[row + 1, 1], wa * table[row, 2] + wb * table[row + 1, 2]))
    return out
