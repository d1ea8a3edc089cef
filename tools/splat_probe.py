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
Dx = 3.0 * torch.randn((n, n), device="cuda")
Dy = 3.0 * torch.randn((n, n), device="cuda")
out = torch.zeros((n, n), device="cuda")
for v in variants:
    out.zero_()
    abi.splat(I, Dx, Dy, out, margin=15, variant=v)
    torch.cuda.synchronize()
    print("variant", v, "sum ratio", float(out.double().sum() / I.double().sum()))
