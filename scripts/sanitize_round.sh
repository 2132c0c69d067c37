#!/usr/bin/env bash
# GPU box: the round's compute-sanitizer evidence.  memcheck + racecheck on the product library; synccheck on the
# TVMB200_SYNCCHECK build (identical kernels, except that the softmax threads of prefill_tc05_kernel also observe the PV_DONE
# phases they do not need -- synccheck reports an mbarrier phase nobody waited for as "missing wait"; see prefill_tc05.cu)
sel="tests/test_kernels_gpu.py tests/test_prefill_tc05_gpu.py tests/test_prefill_tc05_masks_gpu.py"
scripts/sanitize_gpu.sh "memcheck racecheck" "$sel" 600
# (the variant is selected by suffix AND flags: the test fixture rebuilds a library whose stamp does not match its flags)
TVMB200_LIB_SUFFIX=_sync TVMB200_EXTRA_FLAGS="-DTVMB200_SYNCCHECK=1" scripts/sanitize_gpu.sh "synccheck" "$sel" 600
