"""Live differential fuzz of the planning-only host cache against the REAL reference (oracle/ref_harness/fuzz_host.py):
runs only where the reference build of oracle/ref_harness/build_tvm.sh exists (this container; skipped on the GPU box,
which has no /root/reference).  Two fresh seeds per generator (7 generators) here; profiles/r1_fuzz_host_vs_reference.log holds a
450-program run."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
ENV = Path(os.environ.get("TVM_REF_SCRATCH", "/tmp/tvm_ref")) / "env.sh"


@pytest.mark.skipif(not (ENV.exists() and Path("/root/reference").exists()), reason="no reference build in this container")
def test_random_programs_match_the_reference_live(built_lib):
    seed = 5000 + (os.getpid() % 1000)  # different programs on every run; the seed is printed on failure
    r = subprocess.run(["bash", "-c", f"source {ENV} && {sys.executable} oracle/ref_harness/fuzz_host.py --seeds 2 --first {seed}"],
                       cwd=ROOT, capture_output=True, text=True, timeout=600)
    tail = "\n".join(line for line in (r.stdout + r.stderr).splitlines() if "arm_aprofile" not in line)[-3000:]
    assert r.returncode == 0, f"first seed {seed}:\n{tail}"
    assert "OK: 14 programs" in r.stdout, tail


@pytest.mark.skipif(not (ENV.exists() and Path("/root/reference").exists()), reason="no reference build in this container")
def test_library_loads_in_the_reference_runtime(built_lib):
    """INTEGRATION.md routes A / A2 inside the reference's own process (vendored tvm-ffi, libtvm_runtime, its vm.builtin
    registrations present): oracle/ref_harness/load_in_reference_runtime.py."""
    r = subprocess.run(["bash", "-c", f"source {ENV} && {sys.executable} oracle/ref_harness/load_in_reference_runtime.py"],
                       cwd=ROOT, capture_output=True, text=True, timeout=600)
    tail = "\n".join(line for line in (r.stdout + r.stderr).splitlines() if "arm_aprofile" not in line)[-3000:]
    assert r.returncode == 0 and "reference runtime ok" in r.stdout, tail
