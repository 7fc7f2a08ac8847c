#!/bin/bash
# A/B of the captured launch sequence (BBQ_GRAPH=0 off / default on) on the single-query workloads, end to end
for w in ${WL:-c2 c3q1}; do for g in ${GRAPHS:-0 16}; do
  echo -n "$w BBQ_GRAPH=$g: "
  BBQ_GRAPH=$g timeout -s KILL 120 python bench.py --workload $w --no-cpu --no-secondary --steps ${STEPS:-200} --warmup 20 2>/dev/null | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('device-entry QPS', round(d['value']), 'us/query', round(d['ms_per_step']*1e3,1), '| e2e (bbq_search, host buffers) QPS', round(d['e2e']['value']), 'us/query', round(1e6/d['e2e']['value'],1), '| quantize us', round(r.get('quantize_ms_per_step',0)*1e3,1), 'scan us', round(r.get('scan_ms_per_step',0)*1e3,1))"
done; done
