"""Test-side driver of the oracle: runs oracle/_ref/d{2,3}/somar_ref (the reference's own C++
solver stack built by oracle/build_ref.sh) on inputs written to a temp dir and parses what it
dumps.  TEST INFRASTRUCTURE: nothing under somar_b200/ imports this."""
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DECK = os.path.join(ROOT, "oracle", "decks", "base3d.inputs")


def ref_binary(dim=3, shim=False):
    """shim: the same driver built with SOMAR's `using PoissonOp / LevelProjSolver` lines swapped for the B200 drop-ins of
    integration/ (somar_ref_b200: reference C++ API -> C ABI -> CUDA)."""
    return os.path.join(ROOT, "oracle", "_ref", f"d{dim}", "somar_ref_b200" if shim else "somar_ref")


def have_ref(dim=3, shim=False):
    return os.path.exists(ref_binary(dim, shim))


class RefResult:
    def __init__(self, prefix):
        self.kv, self.arrays, self.boxes = {}, {}, []
        data = np.fromfile(prefix + ".bin", dtype=np.float64)
        with open(prefix + ".txt") as f:
            for line in f:
                t = line.split()
                if not t:
                    continue
                if t[0] == "array":
                    off, n = int(t[2]), int(t[3])
                    self.arrays[t[1]] = data[off:off + n].copy()
                elif t[0] == "box":
                    v = [int(x) for x in t[2:]]
                    d = len(v) // 2
                    self.boxes.append((v[:d], v[d:]))
                elif len(t) >= 3 and t[1] == "=":
                    try:
                        self.kv[t[0]] = float(t[2])
                    except ValueError:
                        self.kv[t[0]] = t[2]

    def __getitem__(self, k):
        return self.arrays[k]


def run_ref(mode, nx, L, inp=None, max_box=(0, 0, 0), block_factor=None, offset=None, periodic=(0, 0, 0), relax=6,
            mapname="cartesian", ampl=(0, 0, 0), extra=None, timeout=600, dim=3, split_dirs=None, shim=False):
    """Run the reference driver.  nx, L, offset: 3-vectors (2-D: dim=2 and 2-vectors)."""
    nx = list(nx)
    D = len(nx)
    offset = list(offset) if offset is not None else [0] * (D - 1) + [-nx[-1]]
    if block_factor is None:
        block_factor = 1
        for d in range(D - 1):
            n = max_box[d] if max_box[d] else nx[d]
            block_factor = max(block_factor, 1)
        block_factor = min([mb for mb in max_box[:D - 1] if mb] + [min(nx[:D - 1])])
        while block_factor > 1 and any(n % block_factor for n in nx[:D - 1]):
            block_factor //= 2
        block_factor = max(1, min(block_factor, 16))
    split = list(split_dirs) if split_dirs is not None else [1] * (D - 1) + [0]
    vec = lambda v: " ".join(str(x) for x in v)
    with tempfile.TemporaryDirectory() as td:
        args = [ref_binary(dim, shim), DECK,
                f"base.nx={vec(nx)}", f"base.L={vec(L)}", f"base.nxOffset={vec(offset)}",
                f"base.isPeriodic={vec(list(periodic)[:D])}", f"base.splitDirs={vec(split)}",
                f"base.maxBaseGridSize={vec(list(max_box)[:D])}", f"base.blockFactor={block_factor}",
                f"proj.relaxMethod={relax}", f"drv.mode={mode}", f"drv.map={mapname}", f"drv.ampl={vec(list(ampl)[:D])}",
                f"drv.out={os.path.join(td, 'out')}"]
        if dim == 2:
            args += ["rhs.velBCTypeLo=2 2", "rhs.velBCTypeHi=2 2", "rhs.tempBCTypeLo=1 1", "rhs.tempBCTypeHi=1 1",
                     "rhs.salinityBCTypeLo=1 1", "rhs.salinityBCTypeHi=1 1"]
        if inp is not None:
            p = os.path.join(td, "in.bin")
            np.concatenate([np.asarray(a, dtype=np.float64).ravel(order="F") for a in inp]).tofile(p)
            args.append(f"drv.in={p}")
        for k, v in (extra or {}).items():
            args.append(f"{k}={v}")
        r = subprocess.run(args, capture_output=True, text=True, timeout=timeout)
        if r.returncode != 0:
            raise RuntimeError(f"somar_ref failed ({r.returncode}):\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")
        return RefResult(os.path.join(td, "out"))
