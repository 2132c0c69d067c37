#!/usr/bin/env bash
# GPU box: full parity suite, both bench arms, prefill benches, ncu launch list + full captures -> gpurun_out/
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
timeout 300 python bench.py > gpurun_out/bench_decode.json 2> gpurun_out/bench_decode.err; tail -c 1500 gpurun_out/bench_decode.json
timeout 300 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 600 gpurun_out/bench_reference.json
for dt in bf16 f16; do timeout 120 python bench.py --workload prefill --dtype $dt --no-cpu 2>/dev/null | tail -1 > gpurun_out/bench_prefill_$dt.json; cut -c1-120 gpurun_out/bench_prefill_$dt.json; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_decode.csv python bench.py --steps 5 --warmup 3 --no-cpu > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:prefill_tc05 --launch-skip 2 -c 1 -f -o gpurun_out/prof_prefill python bench.py --workload prefill --no-cpu --steps 2 --warmup 3 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:decode_kernel --launch-skip 4 -c 1 -f -o gpurun_out/prof_decode python bench.py --no-cpu --steps 3 --warmup 3 > /dev/null 2>&1
ls -la gpurun_out | tail -12
