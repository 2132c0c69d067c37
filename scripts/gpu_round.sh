#!/usr/bin/env bash
# GPU box (one GPU): full parity suite, both bench arms, ncu launch list of the headline + one `--set full` capture of the
# dominant kernel of every workload in the bench line -> gpurun_out/r2_*
mkdir -p gpurun_out
o=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $o/r2_pytest_gpu.log
timeout 400 python bench.py > $o/r2_bench_default.json 2> $o/r2_bench_default.err; tail -c 600 $o/r2_bench_default.json
timeout 300 python bench.py --impl reference > $o/r2_bench_reference.json 2> $o/r2_bench_reference.err; tail -c 400 $o/r2_bench_reference.json
P="--no-cpu --no-cupti --no-e2e"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $o/r2_launches_decode.csv python bench.py --steps 5 --warmup 3 --no-sub $P > /dev/null 2>&1
cap() {  # cap <name> <kernel regex> <launch-skip> <bench args...>
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$rx --launch-skip $skip -c 1 -f -o $o/r2_prof_$name python bench.py $P "$@" > /dev/null 2>&1
  ls -la $o/r2_prof_$name.ncu-rep 2>/dev/null | awk '{print $5, $9}'
}
cap decode_c2 '^decode_kernel' 6 --no-sub --steps 3 --warmup 3
cap decode_c5 '^decode_kernel' 6 --workload c5decode
cap append_c3 'transpose_append_kernel' 6 --workload append
cap rotary_append_c3 'split_rotary_warp_kernel' 6 --workload append
cap prefill_c3 'prefill_tc05_kernel' 6 --workload prefill
cap prefill_c5 'prefill_tc05_kernel' 7 --workload c5
ls -la $o | tail -12
