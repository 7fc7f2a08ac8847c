cd /root/repo
V=build/variants
out=gpurun_out/r02g_ab.txt
L=better-binary-quantization_b200/libbbq_b200.so
cp $L $V/new.so
export TMO=90
echo "== 1M x 1024 COSINE, 1024 queries" | tee $out
for lib in $V/new.so $V/e8g1w32.so $V/e4g2w16.so $V/e4g2w32.so $V/e8g2w32.so; do DEBUG=0 WL=c4 EXTRA="--rows 1000000 --nq 1024" REPS=1 bash tools/ab_libs.sh $lib; done 2>&1 | tee -a $out
echo "== C3 (EUCLIDEAN)" | tee -a $out
for lib in $V/e8g1w32.so $V/e4g2w32.so; do DEBUG=0 WL=c3 REPS=1 bash tools/ab_libs.sh $lib; done 2>&1 | tee -a $out
for lib in $V/e4g2w32.so $V/e8g1w32.so; do
  cp $lib $L
  echo -n "$(basename $lib) parity (tensor-core tests): " | tee -a $out
  timeout -s KILL 150 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -x -q -k "mma or 4096" 2>&1 | tail -1 | tee -a $out
done
cp $V/new.so $L
