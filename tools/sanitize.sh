#!/usr/bin/env bash
# Sanitizer evidence.  ONE tool per invocation — and one invocation per gpurun
# call (B200_PROFILING.md: several compute-sanitizer tools in one call have left a GPU unusable):
#   gpurun --timeout 900 -- 'bash tools/sanitize.sh memcheck'     -> gpurun_out/sanitize_memcheck.log
#   gpurun --timeout 900 -- 'bash tools/sanitize.sh racecheck'    -> gpurun_out/sanitize_racecheck.log
#   gpurun --timeout 900 -- 'bash tools/sanitize.sh synccheck'
#   bash tools/sanitize.sh host                                    (CPU: ASan + UBSan build of the oracle, here)
# The program is first run plain (it must exit 0), then under the tool with its own timeout.
set -u
cd "$(dirname "$0")/.."
tool=${1:-memcheck}
mkdir -p gpurun_out
if [ "$tool" = host ]; then
  mkdir -p oracle/_san
  g++ -O1 -g -std=c++17 -ffp-contract=off -fno-fast-math -fPIC -fsanitize=address,undefined -fno-sanitize-recover=undefined \
      -shared -o oracle/_san/libbbq_oracle.so oracle/bbq_oracle.cpp || exit 1
  LD_PRELOAD="$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so)" ASAN_OPTIONS=detect_leaks=0 \
    BBQ_ORACLE_LIB=oracle/_san/libbbq_oracle.so python -m pytest -q -x tests/test_oracle_kat.py tests/test_golden_cpu.py \
    -p no:cacheprovider 2>&1 | tail -15
  exit ${PIPESTATUS[0]}
fi
export SAN_ROWS=${SAN_ROWS:-20000}
python tools/sanitize_case.py > gpurun_out/sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/sanitize_plain.log; exit 1; }
extra=""
[ "$tool" = memcheck ] && extra="--leak-check no"
timeout -s KILL ${SAN_TIMEOUT:-700} compute-sanitizer --tool "$tool" $extra --print-limit 50 --error-exitcode 7 \
  --log-file gpurun_out/sanitize_${tool}.log python tools/sanitize_case.py > gpurun_out/sanitize_${tool}.out 2>&1
rc=$?
echo "compute-sanitizer --tool $tool rc=$rc"
tail -5 gpurun_out/sanitize_${tool}.out
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|========= (Error|Warning|Race)" gpurun_out/sanitize_${tool}.log | sort | uniq -c | head -20
exit $rc
