"""One-time export of the (E, delta, beta) tables PARESIS keeps in Samples/DeltaBeta/TablesDeltaBeta.xls
into a compact .npz the shim can read without xlrd.  Run where the PARESIS data file is available:

    python tools/export_delta_beta.py /root/reference/CodePython/Samples/DeltaBeta/TablesDeltaBeta.xls

Layout in the sheet (Sample.py:113-146): row 0 = material name, rows >= 3 = (E_eV, delta, beta) in
(col, col+1, col+2).  The .npz stores, per material, an [n, 3] float64 array under its name.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paresis_b200.hostio import biff8  # noqa: E402


def export(xls_path, out_path):
    sheet = biff8.open_workbook(xls_path).sheets()[0]
    tables = {}
    for col in range(sheet.ncols):
        name = sheet.cell(0, col).value
        if not isinstance(name, str) or not name.strip() or col + 2 >= sheet.ncols:
            continue
        rows, r = [], 3
        while r < sheet.nrows:
            trio = [sheet.cell(r, col + k).value for k in range(3)]
            if not all(isinstance(v, float) for v in trio):
                break
            rows.append(trio)
            r += 1
        if len(rows) > 1 and rows[0][0] >= 100 and all(b[0] > a[0] for a, b in zip(rows, rows[1:])):
            tables[name] = np.array(rows, dtype=np.float64)   # keep the sheet's exact header text as key
    np.savez_compressed(out_path, **tables)
    return tables


if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                       "paresis_b200", "CodePython", "Samples", "DeltaBeta", "delta_beta_tables.npz")
    t = export(sys.argv[1], out)
    print("wrote %s: %d materials, %d KiB" % (out, len(t), os.path.getsize(out) // 1024))
    print(sorted(t))
