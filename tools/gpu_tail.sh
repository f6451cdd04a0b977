#!/bin/bash
# tail-kernel threshold sweep: bench phases for several SB_TINY_CELLS
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for c in ${CELLS:-8192 1024 128 0}; do
    SB_TINY_CELLS=$c timeout 600 python bench.py --no-cpu-baseline --no-mapped --steps 5 --warmup 3 > gpurun_out/bench_tail_$c.json 2> gpurun_out/bench_tail_$c.err
    python - gpurun_out/bench_tail_$c.json $c <<'PY'
import json, sys
for line in open(sys.argv[1]):
    if line.startswith('{'):
        d = json.loads(line)
        ph = d['phases_ms']
        print('cells', sys.argv[2], 'ms', round(d['ms_per_step'], 3), 'verify', d['checks']['verify']['ok'], {k: v for k, v in ph.items() if int(k.split('@')[1]) >= 4})
PY
done
