#!/bin/bash
# GPU pass for the AMR path: the AMR parity tests (full tracebacks, short), then the whole GPU suite
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_amr_gpu.py -m gpu -q -x --tb=short 2>&1 | tail -60) > gpurun_out/pytest_amr.log
cat gpurun_out/pytest_amr.log | cut -c1-400
if [ -z "$ONLY_AMR" ]; then
  (timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_amr_gpu.py 2>&1 | tail -15) > gpurun_out/pytest_gpu.log
  cat gpurun_out/pytest_gpu.log | cut -c1-300
fi
