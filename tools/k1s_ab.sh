#!/bin/bash
# A/B of the streaming single-query scan (C3 corpus, one query): BBQ_CSA x BBQ_K1S_CTAS
for ctas in ${CTAS:-4 100000}; do for c in ${CSAS:-0 1 2}; do
  echo -n "ctas=$ctas csa=$c: "
  BBQ_K1S_CTAS=$ctas BBQ_CSA=$c python bench.py --workload c3 --nq ${NQ:-1} --no-cpu --steps 50 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['avg_scan_launch_ms'], d['roofline']['frac'], d['ms_per_step'])"
done; done
