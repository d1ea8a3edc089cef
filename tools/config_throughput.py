"""Device-resident throughput of the other BASELINE.json configurations on one GPU (information, not the bench line):
    python tools/config_throughput.py
config 3: 4096^2, 64 energies; config 5: 8192^2, 128 energies; config 4: Fresnel model at 4096^2 / 8192^2 (mono)."""
import contextlib
import io
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from paresis_b200 import geometry, workspace  # noqa: E402

ws = workspace.make_workspace(tempfile.mkdtemp())
workspace.enter(ws)
import Experiment as shim  # noqa: E402
import torch  # noqa: E402


def make(name, model):
    d = dict(experimentName=name, filepath="x/", overSampling=2, nbExpPoints=4, simulation_type=model, expID="p", seed=1)
    with contextlib.redirect_stdout(io.StringIO()):
        return shim.Experiment(d)


def ray_tracing(name, positions):
    exp = make(name, "RayT")
    n = int(exp.exp_dict["studyDimensions"][0])
    mem = exp.myMembrane
    plan = geometry.MembranePlan(mem, n, n, mem.membranePixelSize)
    thresholds = list(exp._open_bins(0))
    scene = exp._scene(thresholds, per_position_membrane=True)
    eng = exp._get_engine()
    np.random.seed(0)
    points = list(range(1, positions + 1))
    buffers = None
    times = []
    for rep in range(3):
        offsets = [plan.draw_offsets() for _ in points]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = eng.compute_rt_positions(scene, plan, offsets, points, n_slots=2, buffers=buffers)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
        buffers = res["buffers"]
    eng.check_flag()
    dt = min(times[1:])
    units = positions * len(scene.spectrum)
    return {"config": name, "grid": n, "energies": len(scene.spectrum), "positions": positions, "s_per_position": dt / positions,
            "positions_per_s": positions / dt, "energy_position_units_per_s": units / dt}


def fresnel(name):
    exp = make(name, "Fresnel")
    n = int(exp.exp_dict["studyDimensions"][0])
    mem = exp.myMembrane
    np.random.seed(0)
    times = []
    with contextlib.redirect_stdout(io.StringIO()):
        for point in range(3):
            mem.myGeometry = []
            mem.getMyGeometry(exp.exp_dict['studyDimensions'], mem.membranePixelSize, 2, point, 4)
            thresholds = exp._open_bins(point)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            exp._get_engine().compute_fresnel(exp._scene(thresholds), point)
            torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
    return {"config": name + " (Fresnel)", "grid": n, "s_per_position_point0": times[0], "s_per_position": min(times[1:]),
            "positions_per_s": 1.0 / min(times[1:])}


if __name__ == "__main__":
    for rec in (ray_tracing("B200_4096_poly64", 4), ray_tracing("B200_8192_poly128", 2), fresnel("B200_4096_mono"),
                fresnel("B200_8192_mono")):
        print(json.dumps(rec), flush=True)
