#!/usr/bin/env bash
# tuning helper (GPU box): prefill tc05 parity tests on the product build, then the C3 bench for each library variant
# usage: scripts/prefill_sweep.sh "<suffix list, '-' = product build>" ["<dtype list>"]
sfx="$1"; dts="${2:-bf16 f16}"
timeout 300 python -m pytest tests/test_prefill_tc05_gpu.py -x -q 2>&1 | tail -4
for s in $sfx; do
  [ "$s" = "-" ] && s=""
  for dt in $dts; do
    printf "== variant '%s' %s: " "$s" "$dt"
    TVMB200_LIB_SUFFIX="$s" timeout 120 python bench.py --workload prefill --dtype "$dt" --no-cpu 2>&1 | tail -1 |
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], 'TFLOP/s', d['ms_per_step'], 'ms', d['clocks'])"
  done
done
