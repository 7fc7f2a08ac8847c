#!/bin/bash
# Timing-attribution sweeps for the tensor-core scan: tools/mma_debug_sweep.sh "<debug values>" [extra bench args]
# (BBQ_MMA_DEBUG bits: see MmaParams::debug in csrc/bbq_mma.cuh; results are WRONG with any bit set)
vals="$1"; shift
for d in $vals; do
  BBQ_MMA_DEBUG=$d timeout -s KILL 200 python bench.py --workload c3 --datagen device --no-cpu --steps 5 "$@" 2>/dev/null | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('debug=$d', 'ms/step', round(d['ms_per_step'],4), 'scan launch ms', round(r['avg_scan_launch_ms'],4))"
done
