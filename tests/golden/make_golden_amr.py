#!/usr/bin/env python
"""Generates tests/golden/amr/*.npz: two-level composite solves by the reference's own AMRHybridSolver
(oracle/_ref/d3/somar_ref, mode amr).  Inputs for next round's CUDA path of SURVEY rows a15 / f2; run in the
build container:   python tests/golden/make_golden_amr.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from _oracle import run_ref  # noqa: E402
from amr_cases import AMR_CASES, composite_rhs, fine_shape, ref_kwargs_amr  # noqa: E402


def main():
    os.makedirs(os.path.join(HERE, "amr"), exist_ok=True)
    for name, c in AMR_CASES.items():
        r0, r1 = composite_rhs(c, 21)
        r = run_ref("amr", inp=[r0, r1], **ref_kwargs_amr(c))
        nf = fine_shape(c)
        np.savez_compressed(os.path.join(HERE, "amr", f"{name}.npz"), name=np.array(name), rhs0=r0, rhs1=r1,
                            phi0=r["phi0"].reshape(c["nx"], order="F"), phi1=r["phi1"].reshape(nf, order="F"),
                            status=int(r.kv["status"]), res_init_norm0=r.kv["res_initNorm0"], res_final_norm0=r.kv["res_finalNorm0"],
                            res_final_norm1=r.kv["res_finalNorm1"])
        print(name, "status", int(r.kv["status"]), "residual", r.kv["res_initNorm0"], "->", r.kv["res_finalNorm0"])


if __name__ == "__main__":
    main()
