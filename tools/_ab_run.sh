cd /root/repo
out=gpurun_out/r02p_ch4.txt
echo -n "tensor-core + extension tests: " | tee $out
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -x -q -k "mma or 4096 or two_bit or wide or large_k" 2>&1 | tail -1 | tee -a $out
echo "== 1M x 1024 EUCLIDEAN (c3), few queries: tensor-core scan with 4-chunk hand-offs (debug=0) vs pairs (debug=512), popcount tile scan" | tee -a $out
for nq in 8 48 128; do
  for d in 0 512; do echo -n "nq=$nq mma debug=$d: "; BBQ_SCAN=mma BBQ_MMA_DEBUG=$d timeout -s KILL 90 python bench.py --workload c3 --datagen device --nq $nq --no-cpu --no-secondary --steps 10 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('QPS', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'scan ms', round(r['scan_ms_per_step'],4), 'index GB/s', round(d['index_GBps']), 'frac hbm', round(d['index_frac_of_hbm_peak'],3))"; done
done 2>&1 | tee -a $out
echo -n "nq=8 popc: " | tee -a $out; BBQ_SCAN=popc timeout -s KILL 90 python bench.py --workload c3 --datagen device --nq 8 --no-cpu --no-secondary --steps 10 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('QPS', round(d['value']), 'scan ms', round(r['scan_ms_per_step'],4))" | tee -a $out
for d in 0 512; do echo -n "c5 debug=$d: "; BBQ_MMA_DEBUG=$d timeout -s KILL 200 python bench.py --workload c5 --no-cpu --steps 3 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('QPS', round(d['value']), 'ms/step', round(d['ms_per_step'],2), 'frac', round(r['frac'],3))"; done 2>&1 | tee -a $out
