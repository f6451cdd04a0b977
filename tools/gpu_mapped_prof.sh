#!/bin/bash
# ncu --set full of the mapped-grid line kernel (vertline_tma_k<GENERAL>) at depth 0 of S5 with a horizontally stretched map
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python tools/bench_relax.py --variants ${VARIANTS:-mapped,tma} > gpurun_out/bench_relax.log 2>&1
cat gpurun_out/bench_relax.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:vertline_tma_k -s 6 -c 1 -o gpurun_out/prof_mapped -f python tools/bench_relax.py --variants mapped > gpurun_out/ncu_mapped.log 2>&1
tail -3 gpurun_out/ncu_mapped.log
