cd /root/repo
out=gpurun_out/r02m_n2.txt
: > $out
for wl in c3 c4; do
  BBQ_BENCH_WATCHDOG=150 timeout --kill-after=10 -s TERM 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --workload $wl --steps 5 --warmup 3 > gpurun_out/r02m_bench_${wl}_2gpu.json 2> gpurun_out/r02m_bench_${wl}_2gpu.err
  echo "bench $wl x2 rc=$?" | tee -a $out
  python -c "
import json; d=json.loads(open('gpurun_out/r02m_bench_${wl}_2gpu.json').read().strip().splitlines()[-1]); print('$wl x2 value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'],'scan launch',d['roofline']['avg_scan_launch_ms'],d['config']['sharding'][:90])" 2>&1 | tail -1 | tee -a $out
done
timeout -s KILL 200 python -m pytest tests/test_gpu_sharded_nccl.py -x -q 2>&1 | tail -1 | tee -a $out
