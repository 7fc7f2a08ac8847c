cd /root/repo
out=gpurun_out/r02j_c5.txt
echo -n "extension + wide-query tests: " | tee $out
timeout -s KILL 400 python -m pytest tests/test_gpu_round2.py -x -q -k "two_bit or wide_queries" 2>&1 | tail -4 | tee -a $out
echo -n "tensor-core regression subset: " | tee -a $out
timeout -s KILL 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -x -q -k "mma or 4096 or other_query_bits" 2>&1 | tail -2 | tee -a $out
timeout -s KILL 300 python bench.py --workload c5 --steps 3 --warmup 3 --cpu-max-queries 4 --parity-queries 8 > gpurun_out/r02j_bench_c5.json 2> gpurun_out/r02j_bench_c5.err
echo "bench c5 rc=$?" | tee -a $out
tail -c 300 gpurun_out/r02j_bench_c5.err | tee -a $out
python -c "
import json; d=json.loads(open('gpurun_out/r02j_bench_c5.json').read().strip().splitlines()[-1]); r=d['roofline']; print('C5 value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'],'scan ms',r['avg_scan_launch_ms'],'frac',r['frac'],'passes',r.get('passes_over_shard'),'ntile',r.get('queries_resident_per_pass'),'parity',d.get('parity'),'cpu',d.get('cpu_baseline'))" | tee -a $out
SAN_TIMEOUT=400 bash tools/sanitize.sh memcheck 2>&1 | tail -12 | tee -a $out
