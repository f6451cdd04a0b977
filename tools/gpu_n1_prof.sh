#!/bin/bash
# one-GPU pass with profiles: tests, smoke, bench line, ncu launch list of one step, ncu --set full of the line kernel
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
[ -n "$SKIP_TESTS" ] || (timeout 2400 python -m pytest tests -m gpu -q ${PYTEST_K:+-k "$PYTEST_K"} --tb=short 2>&1 | tail -40) > gpurun_out/pytest_gpu.log
[ -n "$SKIP_TESTS" ] || cat gpurun_out/pytest_gpu.log | cut -c1-400
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4) > gpurun_out/smoke.log
timeout 900 python bench.py $BENCH_ARGS > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --profile-mode --steps 1 > gpurun_out/prof_mode.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${NCU_KERNEL:-vertline_tma_k} -s 8 -c 2 -o gpurun_out/prof_line -f python bench.py --profile-mode --steps 1 > gpurun_out/ncu_line.log 2>&1
cat gpurun_out/smoke.log | cut -c1-300; tail -2 gpurun_out/bench_n1.err | cut -c1-300; cut -c1-1200 gpurun_out/bench_n1.json; tail -3 gpurun_out/ncu_line.log
