#!/bin/bash
# bench only at N GPUs.  usage: gpu_bench_n.sh N
N=${1:-8}
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -3 gpurun_out/bench_n$N.err | cut -c1-300; grep '^{' gpurun_out/bench_n$N.json | cut -c1-300
