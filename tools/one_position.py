"""A few membrane positions of the benchmark workload (for ncu): python tools/one_position.py [n_positions]"""
import contextlib
import io
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from paresis_b200 import geometry, workspace  # noqa: E402

n_pos = int(sys.argv[1]) if len(sys.argv) > 1 else 3
name = sys.argv[2] if len(sys.argv) > 2 else "B200_2048_mono"
ws = workspace.make_workspace(tempfile.mkdtemp())
workspace.enter(ws)
import Experiment as shim  # noqa: E402
import torch  # noqa: E402

d = dict(experimentName=name, filepath="x/", overSampling=2, nbExpPoints=n_pos, simulation_type="RayT", expID="p", seed=1)
with contextlib.redirect_stdout(io.StringIO()):
    exp = shim.Experiment(d)
mem = exp.myMembrane
n = int(exp.exp_dict["studyDimensions"][0])
eng = exp._get_engine()
grains = torch.empty((n, n), device="cuda")
thresholds = list(exp._open_bins(0))
np.random.seed(0)
for point in range(1, n_pos + 1):     # positions >= 1: the steady-state image-set
    geom, _ = geometry.membrane_segmented(mem, n, n, mem.membranePixelSize, point, mem.myPMMAThickness, out=grains)
    mem.myGeometry = geom
    eng.compute_rt(exp._scene(thresholds), point, want_mean=False)
torch.cuda.synchronize()
print("done", n_pos, "positions of", name)
