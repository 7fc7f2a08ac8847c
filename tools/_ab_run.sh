cd /root/repo
V=build/variants
out=gpurun_out/r02r_k1s.txt
L=better-binary-quantization_b200/libbbq_b200.so
cp $L $V/new.so
echo -n "popcount-path parity tests: " | tee $out
timeout -s KILL 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_recall.py -x -q -k "not mma and not full_size and not 4096" 2>&1 | tail -1 | tee -a $out
export TMO=90
for lib in $V/ref_9a9e1ac.so $V/new.so; do for w in c3q1 c2; do WL=$w STEPS=50 REPS=1 bash tools/ab_libs.sh $lib; done; done 2>&1 | tee -a $out
for lib in $V/ref_9a9e1ac.so $V/new.so; do echo -n "nq=4 "; WL=c3 EXTRA="--nq 4" STEPS=50 REPS=1 bash tools/ab_libs.sh $lib; done 2>&1 | tee -a $out
cp $V/new.so $L
