#!/bin/bash
# K2 configuration sweep (C3): expansion groups x register rebalancing x debug modes, same box, same data.
for cfg in ${CFGS:-"2 0" "1 0"}; do set -- $cfg
  for d in ${DEBUGS:-0 7 130}; do
    echo -n "groups=$1 rb=$2 "
    BBQ_EXP_GROUPS=$1 BBQ_MMA_RB=$2 BBQ_MMA_DEBUG=$d timeout -s KILL 120 python bench.py --workload c3 --datagen device --no-cpu --steps 5 2>/dev/null | tail -1 | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('debug=$d', 'ms/step', round(d['ms_per_step'],4), 'scan launch ms', round(r['avg_scan_launch_ms'],4))"
  done
done
