#!/bin/bash
# AddressSanitizer run of the host cache (no GPU): builds a variant library whose host code is instrumented
# (tvm_b200/lib/libtvm_b200_asan.so, selected by TVMB200_LIB_SUFFIX), replays every fixture + the error / limits tests on
# planning-only caches and then runs 240 random programs (scripts/asan_fuzz_host_cache.py).  libstdc++ is preloaded next
# to libasan so that ASan's __cxa_throw interceptor finds the real function inside a Python process.
# Round 1 result: 43 tests + 240 programs + 14 boundary tests, no report; round 2 (with the disaggregation bookkeeping and the
# scratch reservation): 46 tests + 240 programs + 16 boundary tests, no report.  Remove tvm_b200/lib/*_asan* afterwards.
# The same recipe with TVMB200_LIB_SUFFIX=_ubsan, -fsanitize=undefined -fno-sanitize-recover=undefined and libubsan preloaded
# (UndefinedBehaviorSanitizer) was run as well: no report either.
set -e
cd "$(dirname "$0")/.."
export TVMB200_LIB_SUFFIX=_asan
export TVMB200_EXTRA_FLAGS="-Xcompiler -fsanitize=address -Xcompiler -fno-omit-frame-pointer -g"
python -m tvm_b200.build
ASAN=$(gcc -print-file-name=libasan.so)
STD=$(gcc -print-file-name=libstdc++.so.6)
export ASAN_OPTIONS=detect_leaks=0:halt_on_error=1:log_path=/tmp/asan_report
rm -f /tmp/asan_report*
LD_PRELOAD="$ASAN $STD" python -m pytest tests/test_host_cache_golden.py tests/test_zz_fuzz_gpu.py -q -m "not gpu" -x -s -p no:cacheprovider
LD_PRELOAD="$ASAN $STD" PYTHONPATH=oracle/ref_harness python scripts/asan_fuzz_host_cache.py
# the tvm-ffi boundary (argument marshalling of the packed functions and of the vm.builtin entries), CPU part
LD_PRELOAD="$ASAN $STD" python -m pytest tests/test_library_symbols.py tests/test_vm_builtins.py tests/test_rope_variant_angle.py -q -m "not gpu" -s -p no:cacheprovider
ls /tmp/asan_report* 2>/dev/null && { echo "ASan reports found"; exit 1; } || echo "no ASan report"
