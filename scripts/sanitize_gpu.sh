#!/usr/bin/env bash
# GPU box: compute-sanitizer (memcheck / racecheck / synccheck / initcheck) over the kernel parity tests -> gpurun_out/r2_sanitizer_<tool>.log
# usage: scripts/sanitize_gpu.sh "<tools>" "<pytest selection>" [per-tool timeout s]
tools="${1:-memcheck racecheck synccheck}"
sel="${2:-tests/test_kernels_gpu.py tests/test_prefill_tc05_gpu.py}"
tmo="${3:-500}"
mkdir -p gpurun_out
export PYTORCH_NO_CUDA_MEMORY_CACHING=1   # every torch tensor is its own cudaMalloc: out-of-bounds accesses cannot hide in the pool
for t in $tools; do
  log=gpurun_out/r2_sanitizer_$t.log
  extra=""
  [ "$t" = memcheck ] && extra="--leak-check no --report-api-errors no"
  [ "$t" = racecheck ] && extra="--racecheck-report all"
  echo "== compute-sanitizer --tool $t $extra : pytest $sel" > $log
  timeout $tmo compute-sanitizer --tool $t $extra --print-limit 40 --error-exitcode 0 \
    python -m pytest $sel -x -q -p no:cacheprovider >> $log 2>&1
  echo "exit $?" >> $log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit " $log | tail -4
done
