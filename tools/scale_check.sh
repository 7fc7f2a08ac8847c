cd /root/repo
N=${1:-8}; WL=${2:-c4}
out=gpurun_out/r02s_${WL}_n${N}.txt
BBQ_BENCH_WATCHDOG=200 timeout --kill-after=10 -s TERM 260 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus $N --workload $WL --steps 10 --warmup 3 > gpurun_out/r02s_bench_${WL}_${N}gpu.json 2> gpurun_out/r02s_bench_${WL}_${N}gpu.err
echo "bench $WL x$N rc=$?" | tee $out
python -c "
import json; d=json.loads(open('gpurun_out/r02s_bench_${WL}_${N}gpu.json').read().strip().splitlines()[-1]); r=d['roofline']; print('$WL x$N value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'],'scan launch',r['avg_scan_launch_ms'],'sample',r['sample_ms_per_step'],'quant',r['quantize_ms_per_step'],'select',r['select_ms_per_step'], d['clocks'])" 2>&1 | tail -1 | tee -a $out
grep -v "^frame\|^$\|Exception raised\|sendBytes\|should dump\|OMP_NUM\|\*\*\*\*\|recvValue\|recvBytes\|\[bench r" gpurun_out/r02s_bench_${WL}_${N}gpu.err | tail -4 | cut -c1-200 | tee -a $out
