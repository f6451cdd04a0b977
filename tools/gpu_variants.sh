#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_properties_gpu.py tests/test_parity_gpu.py -m gpu -q -x 2>&1 | tail -12) > gpurun_out/pytest_line.log
rm -f gpurun_out/variants.log
for v in ${VARIANTS:-0 1 2 3 4}; do
  echo "== SB_LINE_VARIANT=$v" >> gpurun_out/variants.log
  SB_LINE_VARIANT=$v timeout 300 python tools/bench_relax.py --iters 8 >> gpurun_out/variants.log 2>&1
done
cut -c1-300 gpurun_out/pytest_line.log; cat gpurun_out/variants.log
