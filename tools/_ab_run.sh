cd /root/repo
V=build/variants
out=gpurun_out/r02n_e8g2.txt
L=better-binary-quantization_b200/libbbq_b200.so
cp $L $V/new.so
export TMO=60
echo "== 1M x 1024 COSINE, 1024 queries: e8g2 (640 threads) with parts of the epilogue switched off" | tee $out
for d in 0 2 1 3 4; do echo -n "debug=$d "; DEBUG=$d WL=c4 EXTRA="--rows 1000000 --nq 1024" REPS=1 bash tools/ab_libs.sh $V/e8g2.so; done 2>&1 | tee -a $out
echo -n "new (e8g1) debug=0 "; DEBUG=0 WL=c4 EXTRA="--rows 1000000 --nq 1024" REPS=1 bash tools/ab_libs.sh $V/new.so 2>&1 | tee -a $out
cp $V/new.so $L
