cd /root/repo
N=${1:-8}
out=gpurun_out/r02q_n${N}.txt
BBQ_BENCH_WATCHDOG=200 timeout --kill-after=10 -s TERM 260 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02q_bench_c4_${N}gpu.json 2> gpurun_out/r02q_bench_c4_${N}gpu.err
echo "bench c4 x$N rc=$?" | tee $out
python -c "
import json; d=json.loads(open('gpurun_out/r02q_bench_c4_${N}gpu.json').read().strip().splitlines()[-1]); r=d['roofline']; print('c4 x$N value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'],'scan launch',r['avg_scan_launch_ms'],'sample',r['sample_ms_per_step'],'quant',r['quantize_ms_per_step'],'select',r['select_ms_per_step'],d['config']['sharding'][:100], d['clocks'])" 2>&1 | tail -1 | tee -a $out
grep -v "^frame\|^$\|Exception raised\|sendBytes\|should dump\|OMP_NUM\|\*\*\*\*\|recvValue\|recvBytes\|\[bench r" gpurun_out/r02q_bench_c4_${N}gpu.err | tail -5 | cut -c1-200 | tee -a $out
