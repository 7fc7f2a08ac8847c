cd /root/repo
out=gpurun_out/r02i_n2.txt
NCCL_DEBUG=WARN timeout -s KILL 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/n2_probe.py > $out 2>&1
echo "probe rc=$?" >> $out
grep -v "^frame\|^$\|Exception raised\|sendBytes\|should dump\|OMP_NUM\|\*\*\*\*" $out | tail -60
