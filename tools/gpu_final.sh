#!/usr/bin/env bash
# Round-end style validation on one GPU: full GPU test suite, smoke, the default bench line, then the ncu evidence
# (launch list + one full capture of the dominant kernel) of that same bench command.  Outputs under gpurun_out/.
cd "$(dirname "$0")/.."
tag=${1:-r02z}
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest_gpu.log
tail -4 gpurun_out/${tag}_pytest_gpu.log
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"
timeout -s KILL 400 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench_c4.json 2> gpurun_out/${tag}_bench_c4.err
echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench_c4.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("C4 value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 2), "scan launch ms", round(r["avg_scan_launch_ms"], 2),
          "frac", round(r["frac"], 3), "int8 probe frac", r.get("frac_of_int8_probe"), "parity", (d.get("parity") or {}).get("ok"), "launches", d["gpu_launches"])
    for s in d.get("secondary", []):
        print("  ", s.get("workload", "")[:40], "value", s.get("value"), "e2e", (s.get("e2e") or {}).get("value"), "scan ms", (s.get("roofline") or {}).get("avg_scan_launch_ms"),
              "frac", (s.get("roofline") or {}).get("frac"), "warm", (s.get("warm_l2") or {}).get("value"), s.get("parity"), s.get("error"))
except Exception as e:
    print("bench parse failed:", e)
PY
[ "${NCU:-1}" = 1 ] || exit 0
# ncu: the same command, plain first (must exit 0), then the launch list of the search kernels and one full capture
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-secondary"
timeout -s KILL 300 $CMD > gpurun_out/${tag}_ncu_plain.log 2>&1 && \
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_scan|k_select|k_tau|k_osq_query|k_query|k_pack|k_validate|k_index_bounds' -c 200 \
    --csv --log-file gpurun_out/${tag}_c4_launches.csv $CMD > gpurun_out/${tag}_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:k_scan_mma -s 3 -c 1 \
    -o gpurun_out/${tag}_c4_k_scan_mma $CMD > gpurun_out/${tag}_ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out/${tag}_*
