#!/usr/bin/env python
"""bench.py -- pressure-solve throughput of the B200 path: DOF/s per V-cycle on the synthetic
stratified 3-D Poisson problem S5 (1024 x 1024 x 256, fp64, 64 boxes of 128 x 128 x 256,
vertical-line relaxation), BASELINE.json's headline metric and configuration.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched with torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W

One "step" = one outer-iteration body of MGSolver::vCycle (reference MGSolverI.H:342-347):
preCond(cor, res) then vCycle_residualEq(cor, res, depth 0) with the reference's default
smoothing (16 down + 16 up line relaxations per depth, 2 at the bottom, BiCGStab bottom solve).
`value` times K steps with the residual resident in HBM; `e2e` is the same step called with host
buffers (pinned): residual H2D, V-cycle, correction D2H.  The fields (2.2 GB each) are far larger
than the 126 MB L2, so no explicit flush is needed between steps.

The reference arm times the reference's own CPU code (oracle/_ref: SOMAR's unmodified C++ solver
stack) on one S5 box (128 x 128 x 256, same dXi) per host core, all cores at once -- the
reference parallelises by MPI rank over boxes and there is no MPI here, so independent copies
stand in for ranks (an upper bound: no halo exchange is paid).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NX = (1024, 1024, 256)
LDOM = (16.0, 16.0, 1.0)
BOX = (128, 128, 0)
BLOCK_FACTOR = 16
LINE_KERNEL = "vertline_tma_k<8,4,4,false>"
METRIC = "pressure-solve DOF/s per V-cycle"
UNIT = "DOF/s"
# Bytes per cell of one full line-relaxation iteration (both colours).
#  * SURVEY.md 8(d) counts the reference's operand set: phi r+w, rhs, J, Dinv = 40 B.
#  * The kernel that runs (sb_line.cu) has no J / Dinv operand at all (one shared tridiagonal matrix,
#    DESIGN.md section 4): what one colour pass must move is phi of the other colour (read), rhs and
#    phi of its own colour (read, write) = 12 B per cell of the grid, 24 B per iteration.  The
#    roofline is reported against THAT figure (it is also what ncu measures as DRAM traffic); the
#    survey-convention number is given next to it.
RELAX_BYTES_PER_CELL_ITER_SURVEY = 40.0
RELAX_BYTES_PER_CELL_ITER = 24.0
# dram__bytes_read.sum + dram__bytes_write.sum of one depth-0 launch on one GPU come from the tracked ncu --set full
# summary of the round (profiles/traffic.json, written by tools/summarize_profile.py from the capture's raw CSV)
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "traffic.json")
# Decomposition-independent results of the N = 1 run (norms are per reference box, the right-hand side is built per
# reference box): every run at any N must reproduce them to 1e-10 of the initial residual norm.
CHECK_FILE = os.path.join(ROOT, "tests", "golden", "bench_s5_checks.json")


def ncu_traffic(kernel_prefix):
    try:
        with open(TRAFFIC_FILE) as f:
            d = json.load(f)
        for k, v in d.get("kernels", {}).items():
            if k.startswith(kernel_prefix):
                return float(v["dram_bytes_per_launch"]), d.get("source")
    except Exception:
        pass
    return None, None


def verify_checks(checks, world):
    """Compare this run's decomposition-independent numbers with the committed N = 1 values."""
    if not os.path.exists(CHECK_FILE):
        return {"ok": None, "note": "tests/golden/bench_s5_checks.json not committed yet"}
    with open(CHECK_FILE) as f:
        want = json.load(f)
    scale = want["res_norm_init"]
    bad = {}
    for k in ("res_norm_init", "cor_norm_after_step", "res_norm_after_step", "solve_final_res_norm", "dot_cor_res"):
        tol = 1e-10 * (scale if "res_norm" in k else abs(want[k]))
        if abs(checks[k] - want[k]) > tol:
            bad[k] = [checks[k], want[k]]
    if checks["solve_iters"] != want["solve_iters"] or checks["solve_status"] != want["solve_status"]:
        bad["solve_iters/status"] = [[checks["solve_iters"], checks["solve_status"]], [want["solve_iters"], want["solve_status"]]]
    return {"ok": not bad, "against": "tests/golden/bench_s5_checks.json (N = 1)", "tolerance": "1e-10 of the initial residual norm", "mismatch": bad}


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy bandwidth)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.index)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                t = [x.strip() for x in line.split(",")]
                if len(t) < 9:
                    continue
                try:
                    sm.append(float(t[1]))
                    mx.append(float(t[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), t[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def synthetic_residual(lo, hi, dXi, seed):
    """RHS-B of SURVEY.md 8(d) on one tile: divergence of a sheared, noisy, stratified face velocity
    (u = tanh shear + noise, v = noise, w = noise under a Gaussian pycnocline envelope).  The field is
    built reference box by reference box (BOX cells in x and y) with a generator seeded by the box's
    global position and no flow through box faces, so it is the same global field for every
    decomposition into tiles, and its integral vanishes (a solvable all-Neumann problem).
    Cartesian map: the advecting velocity equals the Cartesian one."""
    import numpy as np
    n = [int(h - l + 1) for l, h in zip(lo, hi)]
    bx = [BOX[0] or NX[0], BOX[1] or NX[1]]
    z_c = (np.arange(lo[2], hi[2] + 1) + 0.5) * dXi[2]
    z_f = np.arange(lo[2], hi[2] + 2) * dXi[2]
    res = np.zeros(n, order="F")
    for j0 in range(int(lo[1]), int(hi[1]) + 1, bx[1]):
        for i0 in range(int(lo[0]), int(hi[0]) + 1, bx[0]):
            m = [min(bx[0], int(hi[0]) + 1 - i0), min(bx[1], int(hi[1]) + 1 - j0), n[2]]
            rng = np.random.default_rng([seed, i0 // bx[0], j0 // bx[1]])
            blk = np.zeros(m)
            u = 0.05 * (2.0 * rng.random((m[0] + 1, m[1], m[2]), dtype=np.float64) - 1.0)
            u += np.tanh((z_c + 0.2) / 0.1)[None, None, :]
            u[0] = 0.0
            u[-1] = 0.0
            blk += (u[1:] - u[:-1]) * (1.0 / dXi[0])
            v = 0.05 * (2.0 * rng.random((m[0], m[1] + 1, m[2]), dtype=np.float64) - 1.0)
            v[:, 0] = 0.0
            v[:, -1] = 0.0
            blk += (v[:, 1:] - v[:, :-1]) * (1.0 / dXi[1])
            w = 0.05 * (2.0 * rng.random((m[0], m[1], m[2] + 1), dtype=np.float64) - 1.0)
            w *= np.exp(-((z_f + 0.2) / 0.1) ** 2)[None, None, :]
            w[:, :, 0] = 0.0
            w[:, :, -1] = 0.0
            blk += (w[:, :, 1:] - w[:, :, :-1]) * (1.0 / dXi[2])
            res[i0 - int(lo[0]):i0 - int(lo[0]) + m[0], j0 - int(lo[1]):j0 - int(lo[1]) + m[1], :] = blk
    return res


def run_ours(args):
    import numpy as np
    import torch

    import somar_b200 as sb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch N > 1 with torch.distributed.run")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = sb.Context(local, rank, world)
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(sb.Context.unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        ctx.init_comm(bytes(idt.cpu().numpy().tobytes()))

    nx = np.array(NX)
    dXi = np.array(LDOM) / nx
    lo = np.array([0, 0, -NX[2]])
    hi = lo + nx - 1
    blo, bhi = sb.make_base_grids(lo, hi, BOX, (1, 1, 0), BLOCK_FACTOR)
    ranks = sb.assign_boxes_to_ranks(blo, bhi, world)
    op = sb.PoissonOp(ctx, lo, hi, dXi, blo, bhi, box_rank=ranks, relax_method=sb.RELAX_VERTLINE)
    opt = sb.default_options()  # reference defaults: 16/16/2 smooths, FMG outer loop, BiCGStab bottom
    solver = sb.MGSolver(op, opt)
    sched = solver.schedule

    mine = ranks == rank
    tlo, thi = blo[mine].min(axis=0), bhi[mine].max(axis=0)
    ncell_tile = int(np.prod(thi - tlo + 1))
    ncell = int(np.prod(nx))
    res_h = torch.empty(ncell_tile, dtype=torch.float64, pin_memory=True)
    cor_h = torch.empty(ncell_tile, dtype=torch.float64, pin_memory=True)
    res_np = res_h.numpy().reshape(tuple(int(v) for v in thi - tlo + 1), order="F")
    res_np[...] = synthetic_residual(tlo, thi, dXi, 20250829)

    res, cor = op.field(), op.field()
    res.upload_ptr(res_h.data_ptr(), tlo, thi)

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def step():
        solver.precond_vcycle(cor, res)   # op.preCond(cor, res, 0); vCycle_residualEq(cor, res, 0)

    def step_e2e():
        res.upload_ptr(res_h.data_ptr(), tlo, thi)
        solver.precond_vcycle(cor, res)
        cor.download_ptr(cor_h.data_ptr(), tlo, thi)

    def timed(fn, steps):
        barrier()
        l0 = ctx.launch_count()
        t0 = time.perf_counter()
        ctx.timer_start()
        for _ in range(steps):
            fn()
        ms = ctx.timer_stop()
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        launches = ctx.launch_count() - l0
        if dist is not None:
            t = torch.tensor([ms, wall], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0]), float(t[1])
        return ms, wall, launches

    def pipelined_e2e(steps):
        """The same step for a stream of independent inputs: two device buffers per field, the upload
        of step n+1 and the download of step n-1 overlap the V-cycle of step n (asynchronous copies on
        the context's copy streams, ordered with sb_context_stream_wait).  Every step still moves its
        input H2D and its result D2H inside the timed region."""
        C, H, D = sb.STREAM_COMPUTE, sb.STREAM_H2D, sb.STREAM_D2H
        rs, cs = [res, op.field()], [cor, op.field()]
        outs = [cor_h, torch.empty(ncell_tile, dtype=torch.float64, pin_memory=True)]
        barrier()
        for s_ in (H, D):
            ctx.stream_sync(s_)
        t0 = time.perf_counter()
        ctx.stream_wait(H, C)
        rs[0].upload_ptr_async(res_h.data_ptr(), tlo, thi)
        for n in range(steps):
            cur, nxt = n % 2, (n + 1) % 2
            ctx.stream_wait(C, H)                    # input n is on the device
            ctx.stream_wait(H, C)                    # V-cycle n-1 has finished reading the other input buffer
            if n + 1 < steps:
                rs[nxt].upload_ptr_async(res_h.data_ptr(), tlo, thi)
            solver.precond_vcycle(cs[cur], rs[cur])
            ctx.stream_wait(C, D)                    # download n-1 done before V-cycle n+1 overwrites its buffer
            ctx.stream_wait(D, C)
            cs[cur].download_ptr_async(outs[cur].data_ptr(), tlo, thi)
        for s_ in (C, H, D):
            ctx.stream_sync(s_)
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        if dist is not None:
            t = torch.tensor([wall], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            wall = float(t[0])
        for f in (rs[1], cs[1]):
            f.free()
        return wall / steps

    failed = False
    if args.profile_mode:
        step()
        ms, wall_ms, launches = timed(step, args.steps)
        if rank == 0:
            print(json.dumps({"profile_mode": True, "ms_per_step_under_profiler": ms / args.steps, "gpu_launches": launches}))
        return
    # clocks are sampled every 50 ms; the sampler starts with the warm-up steps (same load) so that nvidia-smi's own
    # start-up does not eat the short timed region
    sampler = ClockSampler(local)
    nwarm = max(args.warmup, 3)
    for w in range(nwarm):
        if w == 0 and rank == 0:
            sampler.start()
        step()
    ms, wall_ms, launches = timed(step, args.steps)
    clocks = sampler.stop() if rank == 0 else {}

    # per-kernel accounting of the dominant kernel (line relaxation, depth 0) over one more step
    ctx.profile(True)
    step()
    ctx.profile(False)
    k_ms, k_n = ctx.profile_get("vertline@0")
    r_ms, r_n = ctx.profile_get("residual@0")
    tot_line = sum(ctx.profile_get(f"vertline@{d}")[0] for d in range(len(sched)))
    # where a step spends its time, phase by phase, on the production launch path (rank 0's stream; at N > 1 the waits for
    # neighbouring tiles and for the agglomerated depths are inside these numbers)
    ctx.profile(2)
    step()
    ctx.profile(0)
    phases = {}
    for d in range(len(sched)):
        for key in ("relax_down", "residual_restrict", "prolong", "relax_up", "bottom", "tail", "agglomerated",
                    "agg.relax_down", "agg.residual_restrict", "agg.prolong", "agg.relax_up", "agg.bottom", "agg.tail"):
            p_ms, p_n = ctx.profile_get(f"{key}@{d}")
            if p_n:
                phases[f"{key}@{d}"] = round(p_ms, 4)

    # the whole level solve the reference reports as "Solve time" (AMRNSLevelProject.cpp:315-324): MGSolver::solve with
    # the deck defaults (FMG outer iterations to relTol 1e-6) on the synthetic residual
    phi_s = op.field()
    st_solve = solver.solve(phi_s, res)
    solve_info = {"ms": float(st_solve.device_ms), "iters": int(st_solve.num_iters), "status": sb.STATUS_NAMES[st_solve.status],
                  "rel_res": float(st_solve.final_res_norm / st_solve.init_res_norm) if st_solve.init_res_norm > 0 else None,
                  "what": "MGSolver::solve (FMG, reference defaults) on the same grid and right-hand side, device time on rank 0"}
    phi_s.free()

    # Decomposition-independent checks: the same numbers at every N.
    step()
    tmp = op.field()
    op.residual(tmp, cor, res)
    checks = {"res_norm_init": op.norm(res, 2), "cor_norm_after_step": op.norm(cor, 2), "res_norm_after_step": op.norm(tmp, 2),
              "dot_cor_res": op.dotProduct(cor, res), "solve_final_res_norm": float(st_solve.final_res_norm),
              "solve_iters": int(st_solve.num_iters), "solve_status": int(st_solve.status)}
    tmp.free()

    # The same grid and step with a horizontally stretched map (StretchedMap, maps/StretchedMap.cpp:9-38, amplitudes
    # 0.05, 0.03, -0.1): every column has its own tridiagonal matrix, the line relaxation runs vertline_tma_k<GENERAL>
    # (per-column factorisation recomputed in the kernel).  N = 1 only; a second line of evidence, not the headline.
    mapped = None
    if world == 1 and not args.no_mapped:
        xmin = lo * dXi
        op_m = sb.PoissonOp(ctx, lo, hi, dXi, blo, bhi, box_rank=ranks, relax_method=sb.RELAX_VERTLINE, map_kind=sb.MAP_STRETCHED,
                            map_xmin=xmin, map_xmax=xmin + np.array(LDOM), map_ampl=(0.05, 0.03, -0.1))
        solver_m = sb.MGSolver(op_m, opt)
        res_m, cor_m = op_m.field(), op_m.field()
        res_m.upload_ptr(res_h.data_ptr(), tlo, thi)

        def step_m():
            solver_m.precond_vcycle(cor_m, res_m)
        for _ in range(2):
            step_m()
        m_ms, _, _ = timed(step_m, 3)
        ctx.profile(True)
        step_m()
        ctx.profile(False)
        mk_ms, mk_n = ctx.profile_get("vertline@0")
        m_launch = mk_ms / max(mk_n, 1)
        tmp_m = op_m.field()
        op_m.residual(tmp_m, cor_m, res_m)
        mapped = {"map": "StretchedMap ampl (0.05, 0.03, -0.1)", "ms_per_step": m_ms / 3, "value": ncell / (m_ms / 3 * 1e-3), "unit": UNIT,
                  "line_kernel": "vertline_tma_k<8,4,S,true> (per-column factorisation, no J / Dinv operands)",
                  "line_launch_ms": m_launch, "line_launches_timed": mk_n,
                  "line_achieved_gbs": (RELAX_BYTES_PER_CELL_ITER / 2.0) * ncell_tile / (m_launch * 1e-3) / 1e9 if m_launch > 0 else None,
                  "res_norm_init": op_m.norm(res_m, 2), "res_norm_after_step": op_m.norm(tmp_m, 2)}
        for f in (tmp_m, res_m, cor_m):
            f.free()
        solver_m.free()
        op_m.free()

    # end to end with host buffers
    step_e2e()
    e_ms, e_wall, _ = timed(step_e2e, max(1, min(args.steps, 3)))
    e_steps = max(1, min(args.steps, 3))
    p_steps = max(4, args.steps)
    pipelined_e2e(2)
    p_ms = pipelined_e2e(p_steps)

    if rank == 0:
        peak, peak_src = read_peaks()
        traffic, traffic_src = ncu_traffic("vertline_tma_k")
        ms_per_step = ms / args.steps
        value = ncell / (ms_per_step * 1e-3)
        cells_per_launch = ncell_tile  # one colour pass sweeps the whole tile (half the columns are solved)
        launch_ms = k_ms / max(k_n, 1)
        achieved = (RELAX_BYTES_PER_CELL_ITER / 2.0) * cells_per_launch / (launch_ms * 1e-3) / 1e9
        survey = (RELAX_BYTES_PER_CELL_ITER_SURVEY / 2.0) * cells_per_launch / (launch_ms * 1e-3) / 1e9
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "S5 synthetic stratified 3D Poisson 1024x1024x256 fp64, 64 boxes 128x128x256, "
                                   "VERTLINE relaxation, one preCond + vCycle_residualEq(depth 0) per step, reference default "
                                   "smoothing 16+16 (2 bottom) + BiCGStab bottom",
                       "mg_schedule": [list(r) for r in sched], "decomposition": f"{world} horizontal tile(s)",
                       "l2": "fields (2.2 GB each) exceed the 126 MB L2; no flush needed"},
            "e2e": {"value": ncell / (e_wall / e_steps * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 8 * ncell_tile * world,
                    "d2h_bytes_per_step": 8 * ncell_tile * world, "ms_per_step": e_wall / e_steps,
                    "what": "residual from pinned host -> device, preCond + V-cycle, correction -> pinned host, one step at a "
                            "time (copies not overlapped: the latency a single projection sees)",
                    "pipelined": {"value": ncell / (p_ms * 1e-3), "unit": UNIT, "ms_per_step": p_ms, "steps": p_steps,
                                  "what": "same bytes per step, independent inputs double-buffered: upload of step n+1 and "
                                          "download of step n-1 overlap the V-cycle of step n (sb_field_*_async)"}},
            "gpu_launches": launches,
            "phases_ms": phases,
            "mapped": (dict(mapped, line_frac=mapped["line_achieved_gbs"] / peak) if mapped and mapped["line_achieved_gbs"] else mapped),
            "halo": op.halo_mode(),
            "solve": solve_info,
            "wall_ms_per_step": wall_ms / args.steps,
            "roofline": {"bound": "hbm", "kernel": LINE_KERNEL + " (one colour pass of vertical line relaxation, depth 0)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic if world == 1 else None, "traffic_source": traffic_src,
                         "bytes_counted": "12 B per grid cell per colour pass: phi(other colour) read, rhs + phi(own colour) read/write; "
                                          "the kernel has no J/Dinv operands",
                         "survey_convention": {"bytes_per_cell_per_pass": RELAX_BYTES_PER_CELL_ITER_SURVEY / 2.0, "achieved": survey,
                                               "frac": survey / peak,
                                               "note": "SURVEY.md 8(d) counts the reference's operand set (phi, rhs, J, Dinv)"},
                         "peak_source": peak_src, "launch_ms": launch_ms, "launches_timed": k_n,
                         "algorithmic_bytes_per_launch": (RELAX_BYTES_PER_CELL_ITER / 2.0) * cells_per_launch,
                         "share_of_step": tot_line / (ms_per_step if ms_per_step > 0 else 1.0),
                         "residual_launch_ms": r_ms / max(r_n, 1)},
            "clocks": clocks,
            "checks": dict(checks, verify=verify_checks(checks, world)),
        }
        if args.write_checks and world == 1:
            with open(CHECK_FILE, "w") as f:
                json.dump(checks, f, indent=1)
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(cores=None, reps=1)
        print(json.dumps(out))
        if out["checks"]["verify"]["ok"] is False and not args.no_verify:
            sys.stderr.write("bench.py: decomposition-independent checks differ from the committed N = 1 values: %s\n"
                             % json.dumps(out["checks"]["verify"]["mismatch"]))
            failed = True
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if failed:
        raise SystemExit(3)


def ref_binary():
    return os.path.join(ROOT, "oracle", "_ref", "d3", "somar_ref")


def _run_ref_copies(cores, reps, td):
    """`cores` concurrent copies of the reference's V-cycle on one S5 box; returns per-copy time lists."""
    import numpy as np
    n = (BOX[0], BOX[1], NX[2])
    L = (LDOM[0] * n[0] / NX[0], LDOM[1] * n[1] / NX[1], LDOM[2])
    procs = []
    for c in range(cores):
        out = os.path.join(td, f"c{c}")
        cmd = [ref_binary(), os.path.join(ROOT, "oracle", "decks", "base3d.inputs"),
               f"base.nx={n[0]} {n[1]} {n[2]}", f"base.L={L[0]} {L[1]} {L[2]}", f"base.nxOffset=0 0 {-n[2]}",
               f"base.maxBaseGridSize={BOX[0]} {BOX[1]} 0", f"base.blockFactor={BLOCK_FACTOR}", "proj.relaxMethod=6",
               "drv.mode=vcycle", "drv.useMGSolver=1", f"drv.reps={reps}", f"drv.out={out}"]
        procs.append((out, subprocess.Popen(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)))
    times = []
    for out, p in procs:
        if p.wait() != 0:
            raise RuntimeError("oracle/_ref/d3/somar_ref failed")
        data = np.fromfile(out + ".bin")
        for line in open(out + ".txt"):
            t = line.split()
            if t[:2] == ["array", "vcycleTimes"]:
                times.append(data[int(t[2]):int(t[2]) + int(t[3])])
    return n, np.array(times)


def toolchain_probe():
    """the genuine MPI + Fortran build of the reference needs mpirun, gfortran and LAPACK."""
    import shutil
    have = {t: shutil.which(t) is not None for t in ("mpirun", "gfortran", "scons")}
    if all(have.values()):
        return "mpirun, gfortran and scons are on this host, but the oracle build (restated Fortran leaves, serial) is what is timed"
    return "no " + "/".join(t for t, h in have.items() if not h) + " on this host, so the MPI + Fortran build cannot be made"


def cpu_baseline(cores=None, reps=1):
    if not os.path.exists(ref_binary()):
        return {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "oracle/_ref/d3/somar_ref not built"}
    cores = cores or min(os.cpu_count() or 1, 32)
    with tempfile.TemporaryDirectory() as td:
        n, t = _run_ref_copies(cores, reps, td)
    ncell = n[0] * n[1] * n[2]
    per_copy = t.mean(axis=1)  # seconds per V-cycle of each copy
    value = float(sum(ncell / s for s in per_copy))
    return {"value": value, "unit": UNIT, "cores": cores, "kind": "reference",
            "sample": f"{cores} concurrent copies (1 per core; {toolchain_probe()}) of the reference's vCycle_residualEq on one S5 box "
                      f"{n[0]}x{n[1]}x{n[2]} (same dXi, same 16+16 smoothing), {reps} V-cycle(s) each; "
                      f"{float(per_copy.mean()):.2f} s per V-cycle per core"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if not os.path.exists(ref_binary()):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/d3/somar_ref not built (needs /root/reference at build time)"}))
        return
    cores = min(os.cpu_count() or 1, 32)
    # bounded: each step is one V-cycle on one box per core (~6 s): 3 warm-up + 5 timed whatever K and W ask for
    warm, steps = 3, 5
    t0 = time.perf_counter()
    with tempfile.TemporaryDirectory() as td:
        n, t = _run_ref_copies(cores, warm + steps, td)
    ncell = n[0] * n[1] * n[2]
    import numpy as np
    t = t[:, warm:]
    per_copy = np.median(t, axis=1)
    value = float(sum(ncell / s for s in per_copy))
    sample = (f"{cores} concurrent copies (1 per core; the reference parallelises by MPI rank over boxes; {toolchain_probe()}) of the "
              f"reference's vCycle_residualEq on one S5 box {n[0]}x{n[1]}x{n[2]} (same dXi and smoothing, no halo exchange paid); "
              f"{warm} warm-up + {steps} timed V-cycles each, median per copy")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": float(per_copy.mean() * 1e3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "S5 synthetic stratified 3D Poisson, per-box sample (see cpu_baseline.sample), VERTLINE, "
                               "reference default smoothing 16+16 (2 bottom) + BiCGStab bottom"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-mapped", action="store_true", help="skip the second measurement on the horizontally stretched map (N = 1)")
    ap.add_argument("--no-verify", action="store_true", help="report the decomposition-independent checks without failing on a mismatch")
    ap.add_argument("--write-checks", action="store_true", help="N = 1 only: (re)write tests/golden/bench_s5_checks.json from this run")
    ap.add_argument("--profile-mode", action="store_true",
                    help="for runs under ncu: one warm-up step, K timed steps, no e2e / cpu legs (numbers printed are not bench values)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
