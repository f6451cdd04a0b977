#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
(timeout 120 python tools/dbg_tma.py 128 8 32 2>&1 | tail -5) > gpurun_out/dbg_tma.log
cat gpurun_out/dbg_tma.log | cut -c1-600
grep -q "^ok" gpurun_out/dbg_tma.log || exit 0
bash tools/gpu_line.sh
