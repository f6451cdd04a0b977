#!/bin/bash
# one-GPU validation pass: parity tests, smoke, bench line, optional ncu launch list of one step
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15) > gpurun_out/pytest_gpu.log
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4) > gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
[ -n "$SKIP_NCU" ] || timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --profile-mode --steps 1 > gpurun_out/prof_mode.log 2>&1
cat gpurun_out/pytest_gpu.log | cut -c1-300; cat gpurun_out/smoke.log | cut -c1-300; tail -2 gpurun_out/bench_n1.err | cut -c1-300; cut -c1-1500 gpurun_out/bench_n1.json
