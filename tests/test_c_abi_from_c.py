"""The drop-in boundary from plain C: include/tvm_b200.h and include/tvm_b200_cache.h must compile as strict C99 and a C
host (tests/c_abi/host_cache_plan.c -- no Python, no C++, no torch types) must be able to drive the host cache through the
shared library.  The cache is planning-only (device_id = -1), so no GPU is needed and no kernel is launched."""
import os
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_c99_host_drives_the_cache_through_the_c_abi(built_lib, tmp_path):
    exe = tmp_path / "host_cache_plan"
    libdir = built_lib.parent
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", f"-I{ROOT / 'include'}",
                           str(ROOT / "tests" / "c_abi" / "host_cache_plan.c"), f"-L{libdir}", "-ltvm_b200",
                           f"-Wl,-rpath,{libdir}", "-o", str(exe)])
    env = dict(os.environ)
    r = subprocess.run([str(exe)], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0, f"rc {r.returncode}\n{r.stdout}\n{r.stderr}"
    assert "c abi ok: tvm_b200" in r.stdout
