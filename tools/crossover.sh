#!/bin/bash
# popcount vs tensor-core scan for mid-size query batches (C3 corpus): where should mma_plan switch engines?
for nq in ${NQS:-8 16 32 48}; do for eng in popc mma; do
  echo -n "nq=$nq engine=$eng: "
  BBQ_SCAN=$eng timeout -s KILL 120 python bench.py --workload c3 --datagen device --nq $nq --no-cpu --steps 5 2>/dev/null | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('QPS', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'scan ms', round(r['scan_ms_per_step'],4), r['scan_engine'])"
done; done
