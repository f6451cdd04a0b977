#!/bin/bash
# N-GPU check of the peer-memory halo exchange: the multi-rank parity tests with every depth distributed, then the bench
# line with and without SB_HALO_FORK.  usage: gpu_halo.sh N [pytest -k expression]
N=${1:-2}
K=${2:-test_two_rank_solve}
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
(SB_TEST_WORLD=$N timeout 1200 python -m pytest tests/test_multi_gpu.py -q -x --timeout 300 -k "$K" 2>&1 | tail -25) > gpurun_out/pytest_halo_n$N.log
cut -c1-400 gpurun_out/pytest_halo_n$N.log
run_bench() {  # tag, env...
    tag=$1; shift
    env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_halo_n${N}_$tag.json 2> gpurun_out/bench_halo_n${N}_$tag.err
    tail -2 gpurun_out/bench_halo_n${N}_$tag.err | cut -c1-300
    python - gpurun_out/bench_halo_n${N}_$tag.json <<'PY'
import json, sys
for line in open(sys.argv[1]):
    if line.startswith('{'):
        d = json.loads(line)
        print(sys.argv[1], d['n_gpus'], 'ms', round(d['ms_per_step'], 3), 'halo', d.get('halo'), 'frac', round(d['roofline']['frac'], 3), 'verify', d['checks']['verify']['ok'])
        print(' phases', json.dumps(d.get('phases_ms')))
PY
}
# variants: fused (default: stores and arrival signal inside vertline_tma_k), post (separate post kernel between the
# edge and interior parts of a pass), fork (post kernel on the second stream), nccl (grouped ncclSend / ncclRecv)
for v in ${VARIANTS:-fused post nccl}; do
    case $v in
        fused) run_bench fused SB_X=1 ;;
        post)  run_bench post SB_HALO_FUSED=0 ;;
        fork)  run_bench fork SB_HALO_FUSED=0 SB_HALO_FORK=1 ;;
        nccl)  run_bench nccl SB_PEER_HALO=0 ;;
        agg*)  run_bench $v SB_AGG_CELLS=${v#agg} ;;   # agglomeration threshold in cells, e.g. agg1048576
    esac
done
