"""Shared helpers of the parity tests: synthetic paged caches and oracle<->torch conversions."""
from __future__ import annotations

import numpy as np

from oracle import kernels as ok

TORCH_DT = {}


def torch_dtype(name):
    import torch

    return {"float16": torch.float16, "bfloat16": torch.bfloat16}[name]


def to_dev(x: np.ndarray, dtype: str | None = None, device="cuda"):
    """numpy -> torch on device; float arrays are cast to the 16-bit dtype (values already rounded)."""
    import torch

    t = torch.from_numpy(np.ascontiguousarray(x))
    if dtype is not None:
        t = t.to(torch_dtype(dtype))
    return t.to(device)


def to_np(t) -> np.ndarray:
    import torch

    if t.dtype in (torch.float16, torch.bfloat16):
        return t.float().cpu().numpy()
    return t.cpu().numpy()


def rand16(rng, shape, dtype, dist="normal"):
    x = rng.standard_normal(shape) if dist == "normal" else rng.random(shape)
    return ok.round_dtype(x.astype(np.float32), dtype)


def make_paged_cache(rng, kv_lens, num_kv_heads, head_dim, dtype, page_size=16, extra_pages=3,
                     sliding=None, fill="normal"):
    """Random non-contiguous page table for sequences with `kv_lens` slots in use.

    sliding: optional list of (sliding_offset, sink) per sequence -> length_info [3,B] where
    slots-in-pages = kv_len_in_pages (kv_lens are then the slot counts; visible kv = slots - off + sink).
    Returns dict(pages, page_indptr, page_values, length_info) as numpy arrays.
    """
    npages = [(-(-L // page_size)) for L in kv_lens]
    total = sum(npages) + extra_pages
    perm = rng.permutation(total).astype(np.int32)
    page_indptr = np.zeros(len(kv_lens) + 1, np.int32)
    page_indptr[1:] = np.cumsum(npages)
    page_values = perm[: sum(npages)].copy()
    pages = rand16(rng, (total, 2, num_kv_heads, page_size, head_dim), dtype, fill)
    last = np.array([((L - 1) % page_size) + 1 if L > 0 else 0 for L in kv_lens], np.int32)
    if sliding is None:
        length_info = last
    else:
        length_info = np.stack([last, np.array([s[0] for s in sliding], np.int32),
                                np.array([s[1] for s in sliding], np.int32)]).astype(np.int32)
    return dict(pages=pages, page_indptr=page_indptr, page_values=page_values, length_info=length_info)


def assert_close(name, got, want, atol=2e-3, rtol=1e-2):
    """north_star tolerance: max-abs 2e-3 / rtol 1e-2 (|got-want| <= atol + rtol*|want|)."""
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    assert got.shape == want.shape, f"{name}: shape {got.shape} vs {want.shape}"
    err = np.abs(got - want)
    bound = atol + rtol * np.abs(want)
    bad = err > bound
    assert not bad.any(), (
        f"{name}: {int(bad.sum())}/{bad.size} elements out of tolerance; max abs err {err.max():.3e} "
        f"at {np.unravel_index(err.argmax(), err.shape)} (got {got.flat[err.argmax()]}, want {want.flat[err.argmax()]})"
    )
