#!/usr/bin/env python
"""Micro-benchmark of single kernels at depth 0 of the S5 grid (development aid; bench.py is the
contract).  Prints per-launch CUDA-event times and algorithmic GB/s."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import somar_b200 as sb

ap = argparse.ArgumentParser()
ap.add_argument("--nx", type=int, nargs=3, default=[1024, 1024, 256])
ap.add_argument("--iters", type=int, default=8)
ap.add_argument("--relax", type=int, default=6)
args = ap.parse_args()

ctx = sb.Context(0, 0, 1)
nx = np.array(args.nx)
dXi = np.array([16.0, 16.0, 1.0]) / np.array([1024, 1024, 256])
lo = np.array([0, 0, -nx[2]])
hi = lo + nx - 1
blo, bhi = sb.make_base_grids(lo, hi, (128, 128, 0), (1, 1, 0), 16)
op = sb.PoissonOp(ctx, lo, hi, dXi, blo, bhi, relax_method=args.relax)
n = int(np.prod(nx))
rng = np.random.default_rng(0)
res, cor, tmp = op.field(), op.field(), op.field()
res.upload(rng.standard_normal(tuple(nx)))
op.preCond(cor, res, 0)
op.relax(cor, res, 2)
ctx.profile(True)
op.relax(cor, res, args.iters)
op.residual(tmp, cor, res)
ctx.profile(False)
key = "vertline@0" if args.relax == 6 else "gsrb@0"
ms, cnt = ctx.profile_get(key)
print(f"{key}: {ms / cnt:.4f} ms per launch ({cnt} launches)  algorithmic {20.0 * n / (ms / cnt * 1e-3) / 1e9:.0f} GB/s")
ms, cnt = ctx.profile_get("residual@0")
print(f"residual@0: {ms / cnt:.4f} ms per launch  algorithmic {40.0 * n / (ms / cnt * 1e-3) / 1e9:.0f} GB/s")
ctx.timer_start()
op.relax(cor, res, args.iters)
t = ctx.timer_stop()
print(f"relax x{args.iters}: {t / args.iters:.4f} ms per iteration (2 colour passes + ghost fills)")
