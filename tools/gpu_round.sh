#!/usr/bin/env bash
# One gpurun call: JS-runtime probe, the GPU test suite, the default bench line.  Outputs under gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=${1:-r02}
bash tools/probe_js_runtime.sh > gpurun_out/${tag}_js_probe.txt 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest_gpu.log
tail -5 gpurun_out/${tag}_pytest_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench_c4.json 2> gpurun_out/${tag}_bench_c4.err
echo "bench rc=$?"
tail -c 600 gpurun_out/${tag}_bench_c4.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench_c4.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("C4 value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 2), "scan launch ms", round(r["avg_scan_launch_ms"], 2),
          "frac", round(r["frac"], 3), "int8 probe frac", r.get("frac_of_int8_probe"), "parity", d.get("parity"))
    for s in d.get("secondary", []):
        print("  ", s.get("workload", "")[:40], "value", s.get("value"), "e2e", (s.get("e2e") or {}).get("value"), "scan ms", (s.get("roofline") or {}).get("avg_scan_launch_ms"),
              "frac", (s.get("roofline") or {}).get("frac"), "warm", (s.get("warm_l2") or {}).get("value"), s.get("parity"), s.get("error"))
except Exception as e:
    print("bench parse failed:", e)
PY
