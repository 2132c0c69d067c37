"""CPU timing of the reference's decode hot path -- TEST / BENCH INFRASTRUCTURE ONLY.

Used by bench.py's `cpu_baseline` leg and `--impl reference` arm.  Two back-ends:
  kind "reference": oracle/_ref/libref_kernels_*.so -- the reference's OWN CPU TIR PrimFuncs
      (`_attention_decode_cpu`, `_kv_cache_transpose_append`, `llama_rope_with_position_map`) compiled by
      the reference's own `c` target + gcc -O3 from /root/reference (recipe: oracle/ref_harness/), loaded
      through tvm-ffi.  These PrimFuncs have no T.parallel: they use 1 core.
  kind "port": the NumPy restatement in oracle/kernels.py (when oracle/_ref was not built).
"""
from __future__ import annotations

import os
import time
from pathlib import Path

import numpy as np

from . import kernels as ok

REF_DIR = Path(__file__).resolve().parent / "_ref"


def _ref_module(dtype="float16", hq=32, hkv=8, d=128):
    p = REF_DIR / f"ref_kernels_{dtype}_hq{hq}_hkv{hkv}_d{d}.so"
    if not p.exists():
        return None
    try:
        import ctypes

        import tvm_ffi

        shim = REF_DIR / "libref_shim.so"  # TVMBackendAllocWorkspace/FreeWorkspace (oracle/ref_shim.c)
        if shim.exists():
            ctypes.CDLL(str(shim), mode=ctypes.RTLD_GLOBAL)
        return tvm_ffi.load_module(str(p))
    except Exception:
        return None


def _decode_inputs(B, L, Hq, Hkv, D, dtype, seed=0):
    rng = np.random.default_rng(seed)
    page = 16
    ppseq = -(-L // page)
    nnz = B * ppseq
    pages = ok.round_dtype(rng.standard_normal((nnz + 1, 2, Hkv, page, D)).astype(np.float32), dtype)
    page_values = rng.permutation(nnz + 1).astype(np.int32)[:nnz]
    page_indptr = (np.arange(B + 1) * ppseq).astype(np.int32)
    length_info = np.full(B, ((L - 1) % page) + 1, np.int32)
    qkv = ok.round_dtype(rng.standard_normal((B, Hq + 2 * Hkv, D)).astype(np.float32), dtype)
    qpos = np.full(B, L - 1, np.int32)
    apos = (page_values.reshape(B, ppseq)[:, -1] * page + (L - 1) % page).astype(np.int32)
    return dict(pages=pages, page_values=page_values, page_indptr=page_indptr, length_info=length_info, qkv=qkv,
                qpos=qpos, apos=apos, kofs=np.zeros(B, np.int32))


def _step_bytes(B, L, Hq, Hkv, D, nnz):
    e = 2
    dec = B * L * Hkv * D * 2 * e + 2 * B * Hq * D * e + 4 * B * Hq + 4 * (nnz + 4 * B + 1)
    app = B * Hkv * D * 2 * 2 * 2 + 4 * B
    rot = 2 * B * (Hq + 2 * Hkv) * D * 2 + 4 * B
    return dec + app + rot


def _one_step_port(inp, Hq, Hkv, D, dtype, theta=5e5):
    q, k, v = ok.split_rotary(inp["qkv"], inp["qpos"], Hq, Hkv, 1, theta, 1.0, dtype)
    ok.transpose_append(inp["pages"], k, v, inp["apos"])
    return ok.attention_decode(q, inp["pages"], inp["page_indptr"], inp["page_values"], inp["length_info"],
                               inp["kofs"], inp["qpos"], 0, 1.0, theta, D ** -0.5, dtype)


def _one_step_ref(mod, t, Hq, Hkv, D, theta=5e5):
    mod["fused_rope"](t["qkv"], t["qpos"], t["q"], t["k"], t["v"], 1)
    mod["tir_kv_cache_transpose_append"](t["pages"], t["k"], t["v"], t["apos"])
    mod["batch_decode_paged_kv_cpu"](t["q"], t["pages"], t["page_indptr"], t["page_values"], t["length_info"],
                                     t["kofs"], t["qpos"], t["o"], t["lse"], 0, 1.0, theta, D ** -0.5)


def time_decode_steps(steps=3, warmup=1, B=1, L=4096, Hq=32, Hkv=8, D=128, budget_s=60.0, threads=None):
    """One step = split_rotary + append + decode of one layer on a slice of C2.  The reference's CPU PrimFuncs are
    serial, so the slice is `threads` independent B-sequence sub-batches (own pages / page tables), one per host
    thread (tvm-ffi releases the GIL during a call): all host cores work, each on the reference's own code."""
    import threading

    dtype = "float16"  # the reference's CPU path is tested in fp16/fp32 only (no bf16 CPU codegen)
    mod = _ref_module(dtype, Hq, Hkv, D)
    kind = "reference" if mod is not None else "port"
    ncpu = os.cpu_count() or 1
    T = max(1, min(int(threads or ncpu), 64 // B)) if kind == "reference" else 1
    inps = [_decode_inputs(B, L, Hq, Hkv, D, dtype, seed=i) for i in range(T)]
    nnz = inps[0]["page_values"].size
    if mod is not None:
        import torch

        def tensors(inp):
            t = {k_: torch.from_numpy(v_.astype(np.float16) if v_.dtype == np.float32 else v_) for k_, v_ in inp.items()}
            t["q"] = torch.zeros((B, Hq, D), dtype=torch.float16)
            t["k"] = torch.zeros((B, Hkv, D), dtype=torch.float16)
            t["v"] = torch.zeros((B, Hkv, D), dtype=torch.float16)
            t["o"] = torch.zeros((B, Hq, D), dtype=torch.float16)
            t["lse"] = torch.zeros((B, Hq), dtype=torch.float32)
            return t

        ts = [tensors(inp) for inp in inps]
        one = lambda i: _one_step_ref(mod, ts[i], Hq, Hkv, D)  # noqa: E731
    else:
        one = lambda i: _one_step_port(inps[i], Hq, Hkv, D, dtype)  # noqa: E731

    def run(n):
        """n steps on every sub-batch, the sub-batches in parallel; returns seconds"""
        def work(i):
            for _ in range(n):
                one(i)
        t0 = time.perf_counter()
        if T == 1:
            work(0)
        else:
            th = [threading.Thread(target=work, args=(i,)) for i in range(T)]
            for x in th:
                x.start()
            for x in th:
                x.join()
        return time.perf_counter() - t0

    t_one = run(max(warmup, 1)) / max(warmup, 1)
    steps = max(1, min(steps, int(budget_s / max(t_one, 1e-6))))  # bounded: the whole run ends within ~budget_s
    dt = run(steps) / steps
    bytes_ = T * _step_bytes(B, L, Hq, Hkv, D, nnz)
    return {"value": round(bytes_ / dt / 1e9, 4), "unit": "GB/s", "ms_per_step": round(dt * 1e3, 3), "steps": steps,
            "warmup": warmup, "cores": T if kind == "reference" else int(os.environ.get("OMP_NUM_THREADS", ncpu)),
            "kind": kind, "dtype": "f16",
            "sample": f"{T * B}/64 of the C2 batch ({T} threads x {B} seq x {L} ctx, {Hq}q/{Hkv}kv, D{D}, fp16), "
                      f"{steps} steps; host has {ncpu} cores"}


def time_decode(B=1, L=4096, Hq=32, Hkv=8, D=128, repeats=20):
    r = time_decode_steps(steps=max(1, repeats), warmup=1, B=B, L=L, Hq=Hq, Hkv=Hkv, D=D, budget_s=20.0)
    return {"value": r["value"], "unit": r["unit"], "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
