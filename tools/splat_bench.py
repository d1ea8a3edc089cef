"""The stand-alone splat figure of bench.py alone (no pipeline run): python tools/splat_bench.py"""
import contextlib
import io
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from paresis_b200 import _cabi as abi, geometry, workspace  # noqa: E402

ws = workspace.make_workspace(tempfile.mkdtemp(prefix="paresis_splat_"))
workspace.enter(ws)
import Experiment as shim  # noqa: E402

with contextlib.redirect_stdout(io.StringIO()):
    exp = shim.Experiment(dict(experimentName=bench.EXPERIMENT, filepath=os.path.join(ws, "out", ""), overSampling=2, nbExpPoints=1,
                               simulation_type="RayT", expID="splat", seed=1))
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda", dtype=torch.float32)
peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
peak = json.load(open(peaks))["hbm_gbs"] if os.path.exists(peaks) else 6650.0
print(json.dumps(bench.splat_roofline(abi, geometry, exp, torch, flush, peak)))
