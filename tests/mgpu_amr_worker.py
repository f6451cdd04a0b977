"""Worker for the multi-GPU AMR parity test (launched by torchrun, one rank per GPU): every AMR level's boxes are split
over the ranks (horizontal tiles of the level's patch, as each level is load-balanced separately in the reference), the
composite solve runs with inter-level copies crossing ranks, and rank 0 assembles the pressure of every level."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import somar_b200 as sb
from amr_cases import AMR_CASES, C3_DECK, composite_rhs_levels, level_specs, make_amr_ops, ndim, num_levels


def main():
    name, out_path = sys.argv[1], sys.argv[2]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = sb.Context(local, rank, world)
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(sb.Context.unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    ctx.init_comm(idt.cpu().numpy().tobytes())

    c = C3_DECK if name == "c3_deck" else AMR_CASES[name]
    D, nl = ndim(c), num_levels(c)
    specs = level_specs(c)
    ranks = [sb.assign_boxes_to_ranks(s["box_lo"], s["box_hi"], world) for s in specs]
    ops = make_amr_ops(ctx, c, ranks=ranks)
    rhs0, _ = composite_rhs_levels(c, 3)
    up = (lambda a: a.reshape((a.shape[0], 1, a.shape[1]), order="F")) if D == 2 else (lambda a: a)
    solver = sb.AMRHybridSolver(ops, 0, nl - 1, sb.default_options())
    phi = [ops[l].field() for l in range(nl)]
    rhs = [ops[l].field() for l in range(nl)]
    for l in range(nl):
        rhs[l].upload(up(rhs0[l]))                     # the whole level's array; the upload clips to this rank's tile
    st = solver.solve(phi, rhs)
    for l in range(nl):
        mine = ranks[l] == rank
        tlo, thi = specs[l]["box_lo"][mine].min(axis=0), specs[l]["box_hi"][mine].max(axis=0)
        np.save(f"{out_path}.phi{l}_{rank}.npy", phi[l].download(tlo, thi))
        np.save(f"{out_path}.box{l}_{rank}.npy", np.array([tlo, thi]))
    dist.barrier()
    if rank == 0:
        for l in range(nl):
            lo, hi = specs[l]["reg_lo"], specs[l]["reg_hi"]
            full = np.zeros(tuple(int(v) for v in (hi - lo + 1)), order="F")
            for r in range(world):
                b = np.load(f"{out_path}.box{l}_{r}.npy")
                p = np.load(f"{out_path}.phi{l}_{r}.npy")
                full[tuple(slice(int(b[0][d] - lo[d]), int(b[1][d] - lo[d]) + 1) for d in range(3))] = p
            np.save(f"{out_path}.phi{l}.npy", full if D == 3 else full.reshape((full.shape[0], full.shape[2]), order="F"))
        with open(out_path, "w") as f:
            json.dump({"status": st.status, "norms": st.norms, "world": world, "launches": ctx.launch_count()}, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
