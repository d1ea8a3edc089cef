"""One launch of each stand-alone splat variant at n^2 on a torn displacement field (for ncu):
python tools/splat_probe.py [n] [variants...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paresis_b200 import _cabi as abi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
variants = [int(v) for v in sys.argv[2:]] or [2, 3]
torch.manual_seed(0)
I = torch.rand((n, n), device="cuda") + 0.5
if os.environ.get("SPLAT_FIELD", "caps") == "random":
    Dx = 3.0 * torch.randn((n, n), device="cuda")
    Dy = 3.0 * torch.randn((n, n), device="cuda")
else:
    # membrane-like: gradient of a lattice of spherical caps (period 37 px), mean |D| ~ 1.5 px, torn at the cap edges
    x = torch.arange(n, device="cuda", dtype=torch.float32) / 37.0
    fx, fy = (x % 1.0 - 0.5)[:, None], ((x * 1.07) % 1.0 - 0.5)[None, :]
    t = torch.sqrt(torch.clamp(0.2 - fx * fx - fy * fy, min=0.0))
    Dx, Dy = torch.gradient(t)
    s = 1.5 / float(Dx.abs().mean())
    Dx = (s * Dx).contiguous(); Dy = (s * Dy).contiguous()
    del t
print("mean |D|", float(Dx.abs().mean()), float(Dy.abs().mean()), "max", float(Dx.abs().max()))
out = torch.zeros((n, n), device="cuda")
for v in variants:
    out.zero_()
    abi.splat(I, Dx, Dy, out, margin=15, variant=v)
    torch.cuda.synchronize()
    print("variant", v, "sum ratio", float(out.double().sum() / I.double().sum()))
