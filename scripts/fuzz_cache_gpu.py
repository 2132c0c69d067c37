#!/usr/bin/env python
"""GPU box: more seeds of tests/test_zz_fuzz_gpu.py (random programs through the C++ host cache + sm_100a kernels, the oracle
re-executing the recorded callbacks), with the tcgen05 prefill forced so that the tree / sliding / inline-RoPE routes of round 2
run on every prefill of every program.   usage: fuzz_cache_gpu.py [seeds per kind] [first seed]"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from tests import test_zz_fuzz_gpu as fz  # noqa: E402
from tvm_b200 import capi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
s0 = int(sys.argv[2]) if len(sys.argv) > 2 else 31000
capi.lib()
t0 = time.time()
done = 0
for impl in (2, 0):
    capi.set_prefill_impl(impl)
    for kind in [c[0] for c in fz.CASES]:
        for seed in range(s0, s0 + n):
            fz._run(kind, seed, 0)
            done += 1
    print(f"impl {impl}: {done} programs ok ({time.time() - t0:.0f} s), prefill paths (generic, tcgen05, pre-pass, split) = "
          f"{capi.prefill_path_counts()}", flush=True)
capi.set_prefill_impl(0)
print(f"ALL {done} RANDOM PROGRAMS AGREE WITH THE ORACLE (kinds {[c[0] for c in fz.CASES]}, seeds {s0}..{s0 + n - 1}, tcgen05 forced + auto)")
