#!/usr/bin/env python
"""Micro-benchmark of the line relaxation at depth 0 of the S5 grid (development aid; bench.py is the contract).
Prints per-launch CUDA-event times and GB/s on the 12 B per grid cell a colour pass has to move, for
  fused  : vertline_fused_k (SB_LINE_TMA=0)
  tma    : vertline_tma_k (default)
  mapped : vertline_tma_k<GENERAL> on a horizontally stretched map (ampl 0.05, 0.03, -0.1)"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import somar_b200 as sb

ap = argparse.ArgumentParser()
ap.add_argument("--nx", type=int, nargs=3, default=[1024, 1024, 256])
ap.add_argument("--iters", type=int, default=8)
ap.add_argument("--variants", default="fused,tma,mapped")
args = ap.parse_args()

ctx = sb.Context(0, 0, 1)
nx = np.array(args.nx)
L = np.array([16.0, 16.0, 1.0]) * nx / np.array([1024, 1024, 256])
dXi = L / nx
lo = np.array([0, 0, -nx[2]])
hi = lo + nx - 1
blo, bhi = sb.make_base_grids(lo, hi, (128, 128, 0), (1, 1, 0), 16)
n = int(np.prod(nx))
rng = np.random.default_rng(0)
r0 = rng.standard_normal(tuple(nx))
for variant in args.variants.split(","):
    os.environ["SB_LINE_TMA"] = "0" if variant == "fused" else "1"
    kw = {}
    if variant == "mapped":
        xmin = lo * dXi
        kw = dict(map_kind=sb.MAP_STRETCHED, map_xmin=xmin, map_xmax=xmin + L, map_ampl=(0.05, 0.03, -0.1))
    op = sb.PoissonOp(ctx, lo, hi, dXi, blo, bhi, relax_method=sb.RELAX_VERTLINE, **kw)
    res, cor = op.field(), op.field()
    res.upload(r0)
    op.preCond(cor, res, 0)
    op.relax(cor, res, 2)
    ctx.profile(True)
    op.relax(cor, res, args.iters)
    ctx.profile(False)
    ms, cnt = ctx.profile_get("vertline@0")
    per = ms / max(cnt, 1)
    print(f"{variant}: {per:.4f} ms per colour pass ({cnt} launches)  {12.0 * n / (per * 1e-3) / 1e9:.0f} GB/s on 12 B/cell", flush=True)
    res.free(); cor.free(); op.free()
