#!/bin/bash
# N-GPU validation: multi-rank parity tests, then the bench under torchrun.  usage: gpu_call_mgpu.sh N
N=${1:-2}
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
(SB_TEST_WORLD=$N timeout 1500 python -m pytest tests/test_multi_gpu.py -q ${PYTEST_K:+-k "$PYTEST_K"} 2>&1 | tail -25) > gpurun_out/pytest_mgpu_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
[ -n "$SKIP_NOAGG" ] || SB_AGG_CELLS=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n${N}_noagg.json 2> gpurun_out/bench_n${N}_noagg.err
cat gpurun_out/pytest_mgpu_n$N.log | cut -c1-300; tail -3 gpurun_out/bench_n$N.err | cut -c1-300; cat gpurun_out/bench_n$N.json | cut -c1-400; [ -n "$SKIP_NOAGG" ] || cut -c1-300 gpurun_out/bench_n${N}_noagg.json
