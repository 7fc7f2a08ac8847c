#!/bin/bash
# same-box A/B of two builds of the library: tools/ab_libs.sh old.so new.so  (C3 and C4 scan launch times)
for rep in ${REPS:-1 2}; do for lib in "$@"; do
  cp "$lib" better-binary-quantization_b200/libbbq_b200.so
  for w in ${WL:-c3 c4}; do
    echo -n "$(basename $lib) $w: "
    timeout -s KILL 200 python bench.py --workload $w --datagen device --no-cpu --steps 5 2>/dev/null | tail -1 | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('QPS', round(d['value']), 'scan launch ms', round(r['avg_scan_launch_ms'],4))"
  done
done; done
