"""Build the sm_100a shared library in-tree (tvm_b200/lib/libtvm_b200.so).

nvcc cross-compiles without a GPU; the built .so travels to the GPU box with the repo snapshot.
Usage: python -m tvm_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
LIBDIR = ROOT / "lib"
# tuning knobs: TVMB200_LIB_SUFFIX names a variant library (libtvm_b200<suffix>.so, also honoured by capi.py) and
# TVMB200_EXTRA_FLAGS adds -D... defines to it; the product build uses neither.
SUFFIX = os.environ.get("TVMB200_LIB_SUFFIX", "")
LIB = LIBDIR / f"libtvm_b200{SUFFIX}.so"
OBJDIR = ROOT / "lib" / f"obj{SUFFIX}"

SOURCES = [
    "core.cu",
    "page_kernels.cu",
    "decode.cu",
    "prefill_generic.cu",
    "prefill_tc05.cu",
    "prefill_prepass.cu",
    "prefill_api.cu",
    "kv_transfer.cu",
    "kv_cache_host.cc",
    "ffi_api.cc",
]

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _ffi_includes() -> list[str]:
    import tvm_ffi.libinfo as li

    incs = {li.find_include_path(), li.find_dlpack_include_path()}
    return [f"-I{p}" for p in sorted(incs)]


def _flags(verbose: bool) -> list[str]:
    f = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
         "--expt-relaxed-constexpr", "-Xptxas", "-warn-spills"]
    if verbose:
        f += ["-Xptxas", "-v"]
    f += os.environ.get("TVMB200_EXTRA_FLAGS", "").split()
    return f


def _digest(paths: list[Path], extra: str) -> str:
    h = hashlib.sha256(extra.encode())
    for p in sorted(paths):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    LIBDIR.mkdir(exist_ok=True)
    OBJDIR.mkdir(exist_ok=True)
    srcs = [CSRC / s for s in SOURCES if (CSRC / s).exists()]
    # tuning knob: TVMB200_SWAP_SRC="prefill_tc05.cu=scripts/ab/prefill_tc05_r1.cu" builds a variant with one source swapped
    # (A/B runs of a kernel against an earlier version on the same box)
    for spec in os.environ.get("TVMB200_SWAP_SRC", "").split():
        name, _, repl = spec.partition("=")
        srcs = [(ROOT.parent / repl) if s.name == name else s for s in srcs]
    hdrs = sorted(CSRC.glob("*.cuh")) + sorted((ROOT.parent / "include").glob("*.h"))
    flags = _flags(verbose) + ARCH + _ffi_includes()
    stamp = LIBDIR / f"build{SUFFIX}.stamp"
    digest = _digest(srcs + hdrs, " ".join(flags))
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == digest:
        return LIB

    def compile_one(src: Path) -> tuple[Path, str]:
        obj = OBJDIR / (src.stem + ".o")
        cmd = [NVCC, "-c", str(src), "-o", str(obj)] + flags
        if src.suffix == ".cc":
            cmd += ["-x", "cu"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        return obj, (r.stdout + r.stderr)

    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(compile_one, srcs))
    objs = [str(o) for o, _ in results]
    if verbose:
        for _, log in results:
            sys.stderr.write(log)
    link = [NVCC, "-shared", "-o", str(LIB)] + objs + ARCH + ["-cudart", "static", "-ldl", "-lpthread"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    stamp.write_text(digest)
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(p)
