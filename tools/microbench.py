"""Kernel micro-benchmarks on one B200 (CUDA events, L2 flushed between iterations).

    python tools/microbench.py [--sizes 2048 4096 8192] [--out gpurun_out/microbench.json]

Reports per-kernel time and ALGORITHMIC GB/s (SURVEY.md section 8d figures), next to a plain
device copy measured the same way, so fractions are against this box's own copy bandwidth.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paresis_b200 import _cabi as abi  # noqa: E402
from paresis_b200 import hostmath as hm  # noqa: E402


def timeit(fn, iters=10, warmup=3, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    times = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        times.append(s.elapsed_time(e) * 1e-3)
    times.sort()
    return times[len(times) // 2], times[0]


def synthetic_spheres(seed=0, count=60000):
    rng = np.random.default_rng(seed)
    rows = np.empty((count, 3))
    rows[:, 0] = rng.uniform(-4870.0, 4870.0, count)
    rows[:, 1] = rng.uniform(-4051.0, 4051.0, count)
    rows[:, 2] = rng.gamma(4.0, 3.2, count)
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", type=int, nargs="+", default=[2048, 4096, 8192])
    ap.add_argument("--out", default="gpurun_out/microbench.json")
    ap.add_argument("--splat-only", action="store_true", help="copy / memset and the stand-alone splat variants only")
    args = ap.parse_args()
    dev = "cuda"
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)  # 256 MiB > 126 MB L2
    results = []

    def rec(name, n, med, best, alg_bytes, **extra):
        r = dict(kernel=name, n=n, ms=med * 1e3, ms_best=best * 1e3, alg_GBps=alg_bytes / med / 1e9,
                 alg_GBps_best=alg_bytes / best / 1e9, **extra)
        results.append(r)
        print(json.dumps(r), flush=True)

    for n in args.sizes:
        px = n * n
        a = torch.rand((n, n), device=dev)
        b = torch.empty_like(a)
        med, best = timeit(lambda: b.copy_(a), flush=flush)
        rec("copy_f32", n, med, best, 8 * px)
        med, best = timeit(lambda: b.zero_(), flush=flush)
        rec("memset", n, med, best, 4 * px)
        out = torch.zeros((n, n), device=dev)
        med, best = timeit(lambda: abi.fill(out, 0.0), flush=flush)
        rec("fill_kernel", n, med, best, 4 * px)

        x = torch.linspace(0, n / 50.0, n, device=dev)
        fields = {
            "zero": (torch.zeros((n, n), device=dev), torch.zeros((n, n), device=dev)),
            "smooth2": ((2.0 * torch.sin(x)[:, None] * torch.cos(0.5 * x)[None, :]).contiguous(),
                        (1.5 * torch.cos(0.7 * x)[:, None] * torch.sin(x)[None, :]).contiguous()),
            "smooth10": ((10.0 * torch.sin(x)[:, None] * torch.cos(0.5 * x)[None, :]).contiguous(),
                         (8.0 * torch.cos(0.7 * x)[:, None] * torch.sin(x)[None, :]).contiguous()),
            "random3": (3.0 * torch.randn((n, n), device=dev), 3.0 * torch.randn((n, n), device=dev)),
        }
        # membrane-like displacement: gradient of a rastered sphere membrane (the real workload)
        rows = synthetic_spheres()
        pix = 6.0 / 2 / 1.0254237288135593 * 140 / 141.6
        corr = 50.0 / 12.8
        tab = rows * corr
        ext_x, ext_y = 8102 * corr + 50.0, 9740 * corr + 50.0
        tab[:, 1] += ext_x / 2; tab[:, 0] += ext_y / 2
        reps_x = int(np.ceil(n * pix / ext_x)); reps_y = int(np.ceil(n * pix / ext_y))
        tabs = [tab + np.array([jy * ext_y, ix * ext_x, 0.0]) for ix in range(reps_x) for jy in range(reps_y)]
        tab = np.concatenate(tabs)
        margin = int(np.ceil(10 * 50.0 / pix))
        offs = [(margin // 2 + 10 + 37 * l, margin // 2 + 20 + 53 * l) for l in range(3)]
        sph = torch.as_tensor(tab, device=dev, dtype=torch.float64)
        t_mem = torch.empty((n, n), device=dev)
        med, best = timeit(lambda: abi.raster_spheres(sph, pix, offs, n, n, margin, t_mem), flush=flush)
        rec("raster_spheres", n, med, best, 4 * px, spheres=int(tab.shape[0]), layers=3)
        # displacement maps of that membrane (object hop) through the phase kernel
        k = hm.wavenumber(52e3)
        phi = (-(k * 5.97e-7) * t_mem.double()).contiguous()
        dxp = torch.zeros((n + 30, n + 30), device=dev); dyp = torch.zeros_like(dxp)
        tmp = torch.zeros((n, n), device=dev)
        abi.refract_phi(a, phi, tmp, 3.6, 52.0, 1.0254, 2.9256, 15, dxp, dyp)
        fields["membrane"] = (dxp[15:-15, 15:-15].contiguous(), dyp[15:-15, 15:-15].contiguous())
        del dxp, dyp, phi
        for fname, (Dx, Dy) in fields.items():
            frac_moved = float(((Dx != 0) | (Dy != 0)).float().mean().item())
            p99 = float(torch.quantile(torch.maximum(Dx.abs(), Dy.abs()).flatten()[:: max(1, px // 1000000)], 0.99).item())
            for variant in (0, 1, 2, 3):
                med, best = timeit(lambda: abi.splat(a, Dx, Dy, out, margin=15, variant=variant), flush=flush)
                rec("splat_v%d" % variant, n, med, best, 16 * px, field=fname, moved=frac_moved, p99=p99)
        del fields
        if args.splat_only:
            del a, b, out, t_mem, sph
            torch.cuda.empty_cache()
            continue
        # fused kernels on the real membrane map
        s2 = hm.refraction_gradient_scale(1.6, 1.0254, 2.9256)
        s3 = hm.refraction_gradient_scale(3.6, 1.0254, 2.9256)
        t_smp = torch.empty((n, n), device=dev)
        abi.sphere_map(1000.0 * n / 400, n, n, 2.9256, t_smp)
        ibs = torch.zeros((n, n), device=dev)
        o1 = torch.zeros((n, n), device=dev); o2 = torch.zeros((n, n), device=dev)
        layers = [(t_mem, 5.97e-7 * s3, 5.97e-7 * s3, 0.0), (t_smp, 9.85e-8 * s3, 0.0, 2 * k * 3.16e-12)]
        # hop kernels: the fp32 direct-to-L2 kernel (no intensity scale) and the fixed-point tile kernel (production)
        for label, scale in (("direct", 0.0), ("tile", 7500.0)):
            med, best = timeit(lambda: abi.refract_layers(None, 7500.0, [(t_mem, 5.97e-7 * s2, 0.0, 2 * k * 5.37e-9)], ibs,
                                                          intensity_scale=scale), flush=flush)
            rec("refract_membrane_hop", n, med, best, 8 * px, variant=label)
            med, best = timeit(lambda: abi.refract_layers(ibs, 0.0, layers, o1, o2, intensity_scale=scale), flush=flush)
            rec("refract_sample_ref_hop", n, med, best, 20 * px, variant=label)
        det = n // 2
        work = torch.empty(abi.detect_work_floats(n, n, 2, det, det), device=dev)
        expect = torch.empty((det, det), device=dev)
        src = torch.as_tensor(hm.gaussian_1d(0.424 / 2.355), device=dev, dtype=torch.float32)
        psf = torch.as_tensor(hm.gaussian_1d(1.2), device=dev, dtype=torch.float32)
        med, best = timeit(lambda: abi.detect(o1, 2, det, det, src, psf, work, expect), flush=flush)
        rec("detect_4pass", n, med, best, 4 * px + 4 * det * det)
        med, best = timeit(lambda: abi.detect_counts(o1, 2, det, det, src, psf, work, expect, False), flush=flush)
        rec("detect_fused_nonoise", n, med, best, 4 * px + 4 * det * det)
        med, best = timeit(lambda: abi.detect_counts(o1, 2, det, det, src, psf, work, expect, True, 1, 2), flush=flush)
        rec("detect_fused_poisson", n, med, best, 4 * px + 4 * det * det)
        e2 = torch.empty((det, det), device=dev)
        med, best = timeit(lambda: abi.detect_counts_multi([o1, o2], 2, det, det, src, psf, work, [expect, e2], True, 1, [2, 3]), flush=flush)
        rec("detect_two_images_poisson", n, med, best, 2 * (4 * px + 4 * det * det))
        counts = torch.empty((det, det), device=dev)
        med, best = timeit(lambda: abi.poisson(expect, counts, 1, 2), flush=flush)
        rec("poisson", n, med, best, 8 * det * det)
        del a, b, out, t_mem, t_smp, ibs, o1, o2, work, sph
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as fh:
        json.dump(results, fh, indent=1)


if __name__ == "__main__":
    main()
