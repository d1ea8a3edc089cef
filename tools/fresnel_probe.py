"""One Fresnel membrane position (Experiment.py:279-405) at n^2 through the drop-in API, timed with CUDA events; run it
under `ncu --metrics gpu__time_duration.sum` for the launch list:  python tools/fresnel_probe.py [n] [points]"""
import contextlib
import io
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from paresis_b200 import workspace  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
points = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ws = workspace.make_workspace(tempfile.mkdtemp(prefix="paresis_fresnel_"))
workspace.enter(ws)
import Experiment as shim  # noqa: E402

with contextlib.redirect_stdout(io.StringIO()):
    exp = shim.Experiment(dict(experimentName="B200_%d_mono" % n, filepath=os.path.join(ws, "out", ""), overSampling=2,
                               nbExpPoints=points, simulation_type="Fresnel", expID="fp", seed=1, poissonNoise=True))
mem = exp.myMembrane
np.random.seed(0)
for point in range(points):   # point 0 closes the last detector bin in place (Experiment.py:300)
    mem.myGeometry = []
    with contextlib.redirect_stdout(io.StringIO()):
        mem.getMyGeometry(exp.exp_dict['studyDimensions'], mem.membranePixelSize, 2, point, points)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    with contextlib.redirect_stdout(io.StringIO()):
        res = exp.computeSampleAndReferenceImages_Fresnel(point)
    e1.record()
    torch.cuda.synchronize()
    print("point %d: %.3f ms GPU, %.3f ms wall, mean counts %.1f" % (point, e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3, res[1].mean()))

# where the wall time of a position goes: scene building, the device work, the copy back
thr = exp.myDetector.det_param["myBinsThersholds"]
for rep in range(3):
    mem.myGeometry = []
    with contextlib.redirect_stdout(io.StringIO()):
        mem.getMyGeometry(exp.exp_dict['studyDimensions'], mem.membranePixelSize, 2, 1, points)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    scene = exp._scene(thr)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    res = exp._get_engine().compute_fresnel(scene, 1)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    out = exp._finish(res)
    torch.cuda.synchronize(); t3 = time.perf_counter()
    print("scene %.2f ms, compute_fresnel %.2f ms, finish %.2f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))
