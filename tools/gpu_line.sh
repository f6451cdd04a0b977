#!/bin/bash
# line-kernel development pass: the kernel property tests under a hard timeout (a protocol error in the TMA pipeline
# traps, but a hang must not take the box with it), then per-pass timings
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_properties_gpu.py -m gpu -q -x --tb=short -k "tma or mapped or kernels_agree" 2>&1 | tail -30) > gpurun_out/pytest_line.log
cat gpurun_out/pytest_line.log | cut -c1-400
(timeout 600 python tools/bench_relax.py 2>&1 | tail -30) > gpurun_out/bench_relax.log
cat gpurun_out/bench_relax.log | cut -c1-300
