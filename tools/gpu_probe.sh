#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
P=tools/probe/tma_probe
{
timeout 20 $P 32 1 4 4 1 0
timeout 20 $P 32 1 4 3 1 0
timeout 20 $P 34 3 4 4 1 0
timeout 20 $P 34 3 4 3 1 0
timeout 20 $P 36 3 4 2 1 0
timeout 20 $P 36 3 4 3 1 0
timeout 20 $P 34 1 4 3 1 0
timeout 20 $P 34 3 1 3 1 0
timeout 20 $P 32 3 4 4 1 0
timeout 20 $P 34 3 4 3 3 0
timeout 20 $P 34 3 4 35 3 4
timeout 20 $P 34 3 4 67 0 28
} > gpurun_out/probe.log 2>&1
cat gpurun_out/probe.log
