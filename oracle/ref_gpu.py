"""TEST / BENCH INFRASTRUCTURE ONLY -- client of oracle/ref_gpu_server.py (the reference's own runtime in a subprocess on
the GPU box).  `available()` is False where oracle/_ref/tvm_cuda was not packed (oracle/ref_harness/pack_ref_cuda.sh)."""
from __future__ import annotations

import json
import os
import subprocess
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF = HERE / "_ref" / "tvm_cuda"


def available(dtype: str | None = None) -> bool:
    ok = (REF / "lib" / "libtvm_runtime_extra.so").exists() and (REF / "py" / "tvm_ffi").is_dir()
    if dtype is not None:
        ok = ok and kernel_module(dtype).exists()
    return ok


def kernel_module(dtype: str, hq=32, hkv=8, d=128) -> Path:
    return HERE / "_ref" / f"ref_gpu_kernels_{dtype}_hq{hq}_hkv{hkv}_d{d}.so"


def _run(args, timeout):
    env = dict(os.environ)
    env.pop("PYTHONPATH", None)  # the server puts the reference's tvm-ffi first itself
    env["LD_LIBRARY_PATH"] = f"{REF / 'lib'}:{env.get('LD_LIBRARY_PATH', '')}"
    p = subprocess.run([sys.executable, str(HERE / "ref_gpu_server.py")] + args, capture_output=True, text=True,
                       timeout=timeout, env=env, cwd=str(HERE.parent))
    lines = [ln for ln in p.stdout.strip().splitlines() if ln.startswith("{")]
    if not lines:
        raise RuntimeError(f"reference server produced no result (rc {p.returncode}):\n{p.stdout[-2000:]}\n{p.stderr[-4000:]}")
    return json.loads(lines[-1])


def route_a(fixtures, timeout=900):
    """The reference's C++ cache around tvm_b200's callbacks over the named golden fixtures; returns the server's JSON."""
    return _run(["route_a", "--fixtures", ",".join(fixtures)], timeout)


def run_kernels(spec: dict, arrays: dict, timeout=600):
    """Run the reference's own GPU TIR kernels.  spec: see ref_gpu_server.run_kernels (module / tensors / calls / fetch /
    time); arrays: name -> numpy (16-bit floats as uint16 bit patterns).  Returns (json, dict of fetched arrays)."""
    import numpy as np

    with tempfile.TemporaryDirectory(prefix="refgpu_") as td:
        spec = dict(spec, inputs=os.path.join(td, "in.npz"), outputs=os.path.join(td, "out.npz"))
        np.savez(spec["inputs"], **arrays)
        sp = os.path.join(td, "spec.json")
        Path(sp).write_text(json.dumps(spec))
        res = _run(["kernels", "--spec", sp], timeout)
        out = {}
        if res.get("ok") and os.path.exists(spec["outputs"]):
            with np.load(spec["outputs"]) as z:
                out = {k: z[k] for k in z.files}
        return res, out
