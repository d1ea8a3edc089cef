"""Membrane hop + object hop of one energy, round-1 tile kernels against the owner-computes strip kernels (single position
and positions batched per launch), CUDA events, L2 flushed:  python tools/hop_bench.py [n ...]"""
import contextlib
import io
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from paresis_b200 import _cabi as abi, geometry, hostmath as hm, workspace  # noqa: E402

ws = workspace.make_workspace(tempfile.mkdtemp(prefix="paresis_hop_"))
workspace.enter(ws)
import Experiment as shim  # noqa: E402

peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
peak = json.load(open(peaks))["hbm_gbs"] if os.path.exists(peaks) else 6650.0
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda", dtype=torch.float32)


ONLY = os.environ.get("HOP_ONLY", "")


def timed(fn, reps=10):
    if ONLY:                       # one launch for ncu
        fn(); torch.cuda.synchronize()
        return 0.001
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)) * 1e3


for n in [int(a) for a in sys.argv[1:]] or [2048, 8192]:
    with contextlib.redirect_stdout(io.StringIO()):
        exp = shim.Experiment(dict(experimentName="B200_%d_mono" % n, filepath=os.path.join(ws, "out", ""), overSampling=2,
                                   nbExpPoints=1, simulation_type="RayT", expID="hop", seed=1))
    mem = exp.myMembrane
    batch = 4 if n <= 4096 else 2
    maps = []
    for b in range(batch):
        np.random.seed(b)
        geom, _ = geometry.membrane_segmented(mem, n, n, mem.membranePixelSize, 1, mem.myPMMAThickness)
        maps.append(geom.entries[0])
    t_s = exp.mySampleofInterest._device_geometry().device_entries(materialise=True)[0]
    E, pix, M = 52.0, float(exp.exp_dict["studyPixelSize"]), float(exp.exp_dict["magnification"])
    k = hm.wavenumber(E * 1000)
    g2 = hm.refraction_gradient_scale(1.6, M, pix); g3 = hm.refraction_gradient_scale(3.6, M, pix)
    dm, bm, ds, bs = 5.97e-7, 5.37e-9, 9.85e-8, 3.16e-12
    i0 = 7500.0
    f32 = dict(device="cuda", dtype=torch.float32)
    ibs = [torch.zeros((n, n), **f32) for _ in range(batch)]
    acc_s = [torch.zeros((n, n), **f32) for _ in range(batch)]
    acc_r = [torch.zeros((n, n), **f32) for _ in range(batch)]
    sums = torch.zeros(batch, device="cuda", dtype=torch.float64)
    hop1 = lambda b: [(maps[b], dm * g2, 0.0, 2 * k * bm)]
    hop2 = lambda b: [(maps[b], dm * g3, dm * g3, 0.0), (t_s, ds * g3, 0.0, 2 * k * bs)]
    res = {}
    # round-1 kernels (out +=: the pipeline zero-fills through the previous hop; here the buffers just keep growing)
    if not ONLY:
      res["tile_hop1"] = timed(lambda: abi.refract_layers(None, i0, hop1(0), ibs[0], intensity_scale=i0))
      res["tile_hop2"] = timed(lambda: abi.refract_layers(ibs[0], 0.0, hop2(0), acc_s[0], acc_r[0], sum_ref=sums[0:1], intensity_scale=i0))
    res["strip_hop1"] = timed(lambda: abi.refract_layers(None, i0, hop1(0), ibs[0], intensity_scale=i0, mode=1, reach=8))
    res["strip_hop2"] = timed(lambda: abi.refract_layers(ibs[0], 0.0, hop2(0), acc_s[0], acc_r[0], sum_ref=sums[0:1], intensity_scale=i0, mode=1))
    res["strip_hop2_acc"] = timed(lambda: abi.refract_layers(ibs[0], 0.0, hop2(0), acc_s[0], acc_r[0], sum_ref=sums[0:1], intensity_scale=i0, mode=2))
    c1 = [(dm * g2, 0.0, 2 * k * bm)]
    c2 = [(dm * g3, dm * g3, 0.0), (ds * g3, 0.0, 2 * k * bs)]
    it1 = [dict(thickness=[maps[b]], out_obj=ibs[b]) for b in range(batch)]
    it2 = [dict(intensity_in=ibs[b], thickness=[maps[b], t_s], out_obj=acc_s[b], out_ref=acc_r[b], sum_ref=sums[b:b + 1]) for b in range(batch)]
    w1 = torch.empty(abi.lib.paresis_refract_hop_work_bytes(n, n, 1, batch, 0, 0, 8), device="cuda", dtype=torch.uint8)
    w2 = torch.empty(abi.lib.paresis_refract_hop_work_bytes(n, n, 2, batch, 1, 1, 12), device="cuda", dtype=torch.uint8)
    res["strip_hop1_batch%d_per_pos" % batch] = timed(lambda: abi.refract_hop_batch(it1, c1, i0, i0, reach=8, work=w1)) / batch
    res["strip_hop2_batch%d_per_pos" % batch] = timed(lambda: abi.refract_hop_batch(it2, c2, 0.0, i0, reach=12, work=w2)) / batch
    for key, us in res.items():
        alg = (8.0 if "hop1" in key else 20.0) * n * n
        res[key] = {"us": round(us, 1), "frac": round(alg / (us * 1e-6) / 1e9 / peak, 3)}
    print(json.dumps({"n": n, "work_mb": [w1.numel() >> 20, w2.numel() >> 20], **res}))
    del maps, ibs, acc_s, acc_r, w1, w2
    geometry.drop_device_tables()
    torch.cuda.empty_cache()
