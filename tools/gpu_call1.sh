#!/bin/bash
# one-GPU validation pass: parity tests, bench line, ncu launch list of one step
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15) > gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --profile-mode --steps 1 > gpurun_out/prof_mode.log 2>&1
cat gpurun_out/pytest_gpu.log | cut -c1-300; tail -2 gpurun_out/bench_n1.err | cut -c1-300; cut -c1-2500 gpurun_out/bench_n1.json; tail -2 gpurun_out/prof_mode.log
