"""Worker for the multi-GPU parity test (launched by torchrun, one rank per GPU): solves one case
with the box list split over the ranks, gathers the pressure on rank 0 and compares it with the
oracle and, implicitly, with the single-GPU tests (same tolerances)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import somar_b200 as sb
from cases import CASES, geometry, rand_field, ref_kwargs, rel_err


def main():
    name, optset, out_path = sys.argv[1], sys.argv[2], sys.argv[3]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = sb.Context(local, rank, world)
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(sb.Context.unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    ctx.init_comm(idt.cpu().numpy().tobytes())

    import test_parity2d_gpu as t2
    from test_leptic_gpu import DJL_OPTS, LEPTIC2D, LEPTIC3D
    two_d = name in t2.CASES2D or name in LEPTIC2D
    if two_d:
        # 2-D build of the reference: directions (x, z) in slots 0 and 2, ranks split x only
        c = t2.CASES2D[name] if name in t2.CASES2D else LEPTIC2D[name][0]
        nx, L, dXi, lo, hi = t2.geometry(c)
        blo, bhi = sb.make_base_grids(lo, hi, (c["max_box"][0], 0, 0), (1, 0, 0), c["bf"])
        ranks = sb.assign_boxes_to_ranks(blo, bhi, world)
        xmin = lo * dXi
        kind = sb.MAP_CARTESIAN if c["map"] == "cartesian" else sb.MAP_STRETCHED
        op = sb.PoissonOp(ctx, lo, hi, dXi, blo, bhi, box_rank=ranks, periodic=(c["periodic"][0], 0, c["periodic"][1]), dim=2,
                          map_kind=kind, map_xmin=xmin, map_xmax=xmin + L, map_ampl=(c["ampl"][0], 0.0, c["ampl"][1]),
                          relax_method=c["relax"])
        rhs0 = t2.up(t2.rand_field(c, 4, zero_mean=True))
    else:
        c = CASES[name] if name in CASES else LEPTIC3D[name][0]
        nx, L, dXi, lo, hi = geometry(c)
        blo, bhi = sb.make_base_grids(lo, hi, c["max_box"], (1, 1, 0), c["bf"])
        ranks = sb.assign_boxes_to_ranks(blo, bhi, world)
        xmin = lo * dXi
        kind = sb.MAP_CARTESIAN if c["map"] == "cartesian" else sb.MAP_STRETCHED
        op = sb.PoissonOp(ctx, lo, hi, dXi, blo, bhi, box_rank=ranks, periodic=c["periodic"], map_kind=kind, map_xmin=xmin,
                          map_xmax=xmin + L, map_ampl=c["ampl"], relax_method=c["relax"])
        rhs0 = rand_field(c, 4, zero_mean=True)       # every rank builds the global field, uploads its part
    over = {} if optset == "defaults" else dict(numCycles=1, numSmoothDown=2, numSmoothUp=2, numSmoothBottom=2, prolongOrder=1,
                                                maxIters=20, relTol=1e-10)
    if name.startswith("c2_"):
        over = dict(DJL_OPTS)
    solver = sb.LevelHybridSolver(op, sb.default_options(**over))
    phi, rhs = op.field(), op.field()
    rhs.upload(rhs0)                                  # upload clips to this rank's tile (+ghosts)
    st = solver.solve(phi, rhs)
    mine = ranks == rank
    tlo, thi = blo[mine].min(axis=0), bhi[mine].max(axis=0)
    part = phi.download(tlo, thi)
    # gather tiles on rank 0 through a shared directory (keeps the test independent of tensor plumbing)
    np.save(f"{out_path}.phi{rank}.npy", part)
    np.save(f"{out_path}.box{rank}.npy", np.array([tlo, thi]))
    dist.barrier()
    if rank == 0:
        full = np.zeros(tuple(nx), order="F")
        for r in range(world):
            b = np.load(f"{out_path}.box{r}.npy")
            p = np.load(f"{out_path}.phi{r}.npy")
            sl = tuple(slice(int(b[0][d] - lo[d]), int(b[1][d] - lo[d]) + 1) for d in range(3))
            full[sl] = p
        np.save(f"{out_path}.phi.npy", full)
        with open(out_path, "w") as f:
            json.dump({"status": st.status, "norms": st.norms, "max_depth": st.max_depth, "solve_mode": st.solve_mode, "world": world,
                       "launches": ctx.launch_count(), "halo": op.halo_mode()}, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
