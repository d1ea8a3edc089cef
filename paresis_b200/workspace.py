"""Working directories shaped like PARESIS's ``CodePython`` folder.

PARESIS resolves ``xmlFiles/...`` and ``Samples/...`` relative to the current directory
(Experiment.py:32, Sample.py:24,93, getMembraneFromFile.py:80), so it is run with
cwd = CodePython.  ``make_workspace`` builds such a directory anywhere: parameter files and
delta/beta tables are linked from the shim, and a sphere list is written (synthetic unless a
path to a real segmented membrane is given -- upstream's CuSn.txt is not redistributed).
"""
import json
import os
import sys

import numpy as np

SHIM_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "CodePython")


def synthetic_sphere_rows(seed=0, count=60000):
    """Rows [c0 (y), c1 (x), radius] in the sphere file's own units: uniform centres over its
    9740 x 8102 extent, Gamma(4, 3.2) radii (mean 12.8, the value upstream rescales from;
    SURVEY.md section 8d, config 1)."""
    rng = np.random.default_rng(seed)
    rows = np.empty((count, 3))
    rows[:, 0] = rng.uniform(-4870.0, 4870.0, count)
    rows[:, 1] = rng.uniform(-4051.0, 4051.0, count)
    rows[:, 2] = rng.gamma(4.0, 3.2, count)
    return rows


def make_workspace(path, sphere_rows=None, sphere_seed=0, sphere_count=60000, sphere_file=None):
    path = os.path.abspath(path)
    os.makedirs(os.path.join(path, "Samples", "Membranes"), exist_ok=True)
    for rel in ("xmlFiles", os.path.join("Samples", "DeltaBeta"), "Sources"):
        dst = os.path.join(path, rel)
        if not os.path.lexists(dst):
            os.symlink(os.path.join(SHIM_DIR, rel), dst)
    target = os.path.join(path, "Samples", "Membranes", "CuSn.txt")
    if sphere_file is not None:
        if os.path.lexists(target):
            os.remove(target)
        os.symlink(os.path.abspath(sphere_file), target)
    else:
        rows = synthetic_sphere_rows(sphere_seed, sphere_count) if sphere_rows is None else np.asarray(sphere_rows)
        with open(target, "w") as fh:
            json.dump(rows.tolist(), fh)
    return path


def enter(path):
    """chdir into a workspace and put the shim modules first on sys.path (flat imports, main.py:12-15)."""
    os.chdir(path)
    if SHIM_DIR not in sys.path:
        sys.path.insert(0, SHIM_DIR)
    return path
