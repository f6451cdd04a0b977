#!/bin/bash
# one-GPU validation pass: parity tests, line-kernel launch-shape sweep, bench line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/pytest_gpu.log
for v in 0 1 3 4 5 7; do
  echo "== SB_LINE_VARIANT=$v" >> gpurun_out/variants.log
  SB_LINE_VARIANT=$v timeout 300 python tools/bench_relax.py --iters 8 >> gpurun_out/variants.log 2>&1
done
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
cat gpurun_out/pytest_gpu.log gpurun_out/variants.log gpurun_out/bench_n1.json
