#!/bin/bash
# Runs every gemm_selftest case in its own process (a trapped kernel cannot poison the rest).
# Usage: tests/native/run_selftest.sh [outfile]
cd "$(dirname "$0")/../.."
BIN=tests/native/gemm_selftest
OUT=${1:-gpurun_out/gemm_selftest.log}
mkdir -p "$(dirname "$OUT")"
: > "$OUT"
N=$($BIN list)
fail=0
for ((i=0;i<N;i++)); do
  timeout 120 $BIN $i >> "$OUT" 2>&1
  rc=$?
  if [ $rc -ne 0 ]; then fail=$((fail+1)); echo "case $i exit code $rc" >> "$OUT"; fi
done
echo "SELFTEST cases=$N failed=$fail" >> "$OUT"
timeout 300 $BIN perf >> "$OUT" 2>&1
grep -E "^(PASS|FAIL|SELFTEST|PERF|case)" "$OUT"
exit $fail
