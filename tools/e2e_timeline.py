"""Where the end-to-end time of one position goes (API path of bench.py): python tools/e2e_timeline.py"""
import contextlib
import io
import os
import sys
import tempfile
import time


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from paresis_b200 import transfer, workspace  # noqa: E402

ws = workspace.make_workspace(tempfile.mkdtemp())
workspace.enter(ws)
import Experiment as shim  # noqa: E402
import torch  # noqa: E402

d = dict(experimentName="B200_2048_mono", filepath="x/", overSampling=2, nbExpPoints=20, simulation_type="RayT", expID="p", seed=1)
with contextlib.redirect_stdout(io.StringIO()):
    exp = shim.Experiment(d)
mem = exp.myMembrane
T = {"geometry": 0.0, "compute": 0.0, "thick": 0.0, "engine.compute_rt": 0.0, "_finish": 0.0, "_scene": 0.0}


def timed(obj, name, key):
    fn = getattr(obj, name)

    def wrap(*a, **k):
        t0 = time.perf_counter()
        try:
            return fn(*a, **k)
        finally:
            T[key] += time.perf_counter() - t0
    setattr(obj, name, wrap)


timed(exp._get_engine(), "compute_rt", "engine.compute_rt")
timed(exp, "_finish", "_finish")
timed(exp, "_scene", "_scene")


def job():
    exp.myDetector.det_param["myBinsThersholds"] = []
    with contextlib.redirect_stdout(io.StringIO()):
        for point in range(20):
            t0 = time.perf_counter()
            mem.myGeometry = []
            mem.getMyGeometry(exp.exp_dict['studyDimensions'], mem.membranePixelSize, 2, point, 20)
            t1 = time.perf_counter()
            res = exp.computeSampleAndReferenceImages_RT(point)
            t2 = time.perf_counter()
            thick = mem.myGeometry[0]
            t3 = time.perf_counter()
            T["geometry"] += t1 - t0; T["compute"] += t2 - t1; T["thick"] += t3 - t2
    torch.cuda.synchronize()


for _ in range(3):
    job()
for k in T:
    T[k] = 0.0
t0 = time.perf_counter()
n = 5
for _ in range(n):
    job()
dt = time.perf_counter() - t0
print("per position: total %.3f ms" % (dt / n / 20 * 1e3), {k: round(v / n / 20 * 1e3, 3) for k, v in T.items()},
      "pinned allocs", transfer.pinned_allocs)
# raw PCIe: one 16.7 MB and one 33.5 MB pinned copy
x = torch.empty((2048, 2048), device="cuda"); h = torch.empty((2048, 2048), pin_memory=True)
for nbytes, (src, dst) in {"16.7MB": (x, h)}.items():
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("D2H", nbytes, "%.1f GB/s" % (20 * src.numel() * 4 / dt / 1e9))
