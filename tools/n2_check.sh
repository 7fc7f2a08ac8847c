#!/usr/bin/env bash
# 2-GPU check of the sharded path (gpurun --gpus 2): the staged probe under torchrun, the library-level tests, and a
# short C4 bench — every command under its own timeout, torchrun terminated (not killed) so that it reaps its ranks.
cd "$(dirname "$0")/.."
out=gpurun_out/n2_check.txt
timeout --kill-after=10 -s TERM 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/n2_probe.py 2>&1 | grep -E "equal to unsharded|done|Error|Timeout" | tee $out
timeout -s KILL 200 python -m pytest tests/test_gpu_sharded_nccl.py -x -q 2>&1 | tail -1 | tee -a $out
BBQ_BENCH_WATCHDOG=150 timeout --kill-after=10 -s TERM 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/n2_bench_c4.json 2> gpurun_out/n2_bench_c4.err
echo "bench c4 x2 rc=$?" | tee -a $out
python -c "
import json; d=json.loads(open('gpurun_out/n2_bench_c4.json').read().strip().splitlines()[-1]); print('c4 x2 value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'])" 2>&1 | tail -1 | tee -a $out
