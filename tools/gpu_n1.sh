#!/bin/bash
# one-GPU pass: parity tests, smoke, bench line (BENCH_ARGS e.g. --write-checks), optional ncu launch list
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
(timeout 2400 python -m pytest tests -m gpu -q ${PYTEST_K:+-k "$PYTEST_K"} --tb=short 2>&1 | tail -40) > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log | cut -c1-400
[ -n "$SKIP_SMOKE" ] || (timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4) > gpurun_out/smoke.log
[ -n "$SKIP_BENCH" ] || timeout 900 python bench.py $BENCH_ARGS > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
[ -n "$SKIP_BENCH" ] || cp tests/golden/bench_s5_checks.json gpurun_out/ 2>/dev/null
[ -z "$WITH_NCU" ] || timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --profile-mode --steps 1 > gpurun_out/prof_mode.log 2>&1
cat gpurun_out/smoke.log | cut -c1-300
[ -n "$SKIP_BENCH" ] || { tail -2 gpurun_out/bench_n1.err | cut -c1-300; cut -c1-3000 gpurun_out/bench_n1.json; }
true
