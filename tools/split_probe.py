"""Two-beam object hop in ONE launch against two single-beam launches (6 instead of 4 resident blocks each):
python tools/split_probe.py [sizes...]   (timing experiment; CUDA events, L2 flushed between iterations)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from paresis_b200 import _cabi as abi  # noqa: E402
from paresis_b200 import hostmath as hm  # noqa: E402
from microbench import timeit, synthetic_spheres  # noqa: E402

flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
for n in [int(v) for v in sys.argv[1:]] or [2048, 4096, 8192]:
    rows = synthetic_spheres()
    pix = 6.0 / 2 / 1.0254237288135593 * 140 / 141.6
    corr = 50.0 / 12.8
    tab = rows * corr
    ext_x, ext_y = 8102 * corr + 50.0, 9740 * corr + 50.0
    tab[:, 1] += ext_x / 2; tab[:, 0] += ext_y / 2
    reps_x = int(np.ceil(n * pix / ext_x)); reps_y = int(np.ceil(n * pix / ext_y))
    tab = np.concatenate([tab + np.array([jy * ext_y, ix * ext_x, 0.0]) for ix in range(reps_x) for jy in range(reps_y)])
    margin = int(np.ceil(10 * 50.0 / pix))
    offs = [(margin // 2 + 10 + 37 * l, margin // 2 + 20 + 53 * l) for l in range(3)]
    t_mem = torch.empty((n, n), device="cuda")
    abi.raster_spheres(torch.as_tensor(tab, device="cuda", dtype=torch.float64), pix, offs, n, n, margin, t_mem)
    k = hm.wavenumber(52e3)
    s2 = hm.refraction_gradient_scale(1.6, 1.0254, 2.9256)
    s3 = hm.refraction_gradient_scale(3.6, 1.0254, 2.9256)
    for sample in ("fibre", "sphere"):
        t_smp = torch.empty((n, n), device="cuda")
        if sample == "sphere":
            abi.sphere_map(1000.0 * n / 400, n, n, 2.9256, t_smp)
        else:
            abi.cylinder_map(700.0 * n / 2048, 30.0, n, n, 2.9256, t_smp)
        ibs = torch.zeros((n, n), device="cuda")
        abi.refract_layers(None, 7500.0, [(t_mem, 5.97e-7 * s2, 0.0, 2 * k * 5.37e-9)], ibs, intensity_scale=7500.0)
        o1 = torch.zeros((n, n), device="cuda"); o2 = torch.zeros((n, n), device="cuda")
        both = [(t_mem, 5.97e-7 * s3, 5.97e-7 * s3, 0.0), (t_smp, 9.85e-8 * s3, 0.0, 2 * k * 3.16e-12)]
        obj = [(t_mem, 5.97e-7 * s3, 0.0, 0.0), (t_smp, 9.85e-8 * s3, 0.0, 2 * k * 3.16e-12)]
        ref = [(t_mem, 5.97e-7 * s3, 0.0, 0.0)]
        m1, _ = timeit(lambda: abi.refract_layers(ibs, 0.0, both, o1, o2, intensity_scale=7500.0), flush=flush)

        def split():
            abi.refract_layers(ibs, 0.0, obj, o1, intensity_scale=7500.0)
            abi.refract_layers(ibs, 0.0, ref, o2, intensity_scale=7500.0)
        m2, _ = timeit(split, flush=flush)
        print(n, sample, "one launch %.1f us, two single-beam launches %.1f us" % (m1 * 1e6, m2 * 1e6), flush=True)
        del t_smp, ibs, o1, o2
    del t_mem
    torch.cuda.empty_cache()
