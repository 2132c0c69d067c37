"""Random programs on planning-only host caches (no GPU, no reference): the loop of tests/test_zz_fuzz_gpu.py::_run for 40
seeds x 6 generators.  Meant to run with a sanitizer-instrumented library (scripts/asan_host_cache.sh); it also checks, via the
NaN-initialised oracle pages, that no plan reads a slot nobody appended to."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle" / "ref_harness"))

import tests.test_zz_fuzz_gpu as T  # noqa: E402

n = 0
for seed in range(30000, 30040):
    for kind in [c[0] for c in T.CASES]:
        try:
            T._run(kind, seed, None)
        except AssertionError as e:
            if str(e):          # a bare `assert n_fwd > 10 ...` (a short program) is not a host matter
                raise
        n += 1
print("sanitizer fuzz ok:", n, "programs")
