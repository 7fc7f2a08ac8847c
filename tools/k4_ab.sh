#!/bin/bash
# A/B of the query quantiser forms on the single-query workloads (C2, C3-q1): BBQ_QQUANT=warp|cta (default: by batch size)
for w in ${WL:-c2 c3q1}; do for q in ${FORMS:-warp cta}; do
  echo -n "$w qquant=$q: "
  BBQ_QQUANT=$q timeout -s KILL 120 python bench.py --workload $w --no-cpu --no-secondary --steps ${STEPS:-200} --warmup 20 2>/dev/null | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('QPS', round(d['value']), 'us/query', round(d['ms_per_step']*1e3,1), 'e2e QPS', round(d['e2e']['value']), 'quantize us', round(r.get('quantize_ms_per_step',0)*1e3,1), 'scan us', round(r.get('scan_ms_per_step',0)*1e3,1), 'sample us', round(r.get('sample_ms_per_step',0)*1e3,1), 'select us', round(r.get('select_ms_per_step',0)*1e3,1))"
done; done
