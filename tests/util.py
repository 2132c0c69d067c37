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


# ---- DLPack views with a non-zero byte_offset --------------------------------------------------------------------
# The reference hands its callbacks views into ONE merged auxiliary buffer: `data` is the buffer's base pointer and
# `byte_offset` the 16-byte-aligned position of the array inside it (attn_utils.h:1027-1052); q / k / v are views of
# temp buffers the same way (paged_kv_cache.cc:1340-1345).  torch always exports byte_offset = 0, so the tests build the
# DLManagedTensor themselves (malloc'ed, released by libc free: no Python callback is alive at interpreter shutdown).
import ctypes as _ct


class _DLDevice(_ct.Structure):
    _fields_ = [("device_type", _ct.c_int32), ("device_id", _ct.c_int32)]


class _DLDataType(_ct.Structure):
    _fields_ = [("code", _ct.c_uint8), ("bits", _ct.c_uint8), ("lanes", _ct.c_uint16)]


class _DLTensor(_ct.Structure):
    _fields_ = [("data", _ct.c_void_p), ("device", _DLDevice), ("ndim", _ct.c_int32), ("dtype", _DLDataType),
                ("shape", _ct.POINTER(_ct.c_int64)), ("strides", _ct.POINTER(_ct.c_int64)), ("byte_offset", _ct.c_uint64)]


class _DLManagedTensor(_ct.Structure):
    _fields_ = [("dl_tensor", _DLTensor), ("manager_ctx", _ct.c_void_p), ("deleter", _ct.c_void_p)]


class DLView:
    """A DLPack producer for `shape` elements of (code, bits) at base_ptr + byte_offset, with the offset kept in
    DLTensor.byte_offset.  `owner` (the tensor that owns the memory) must outlive every consumer."""

    def __init__(self, base_ptr, byte_offset, shape, code, bits, device_type, device_id, owner=None):
        self.spec = (int(base_ptr), int(byte_offset), tuple(int(s) for s in shape), code, bits, device_type, device_id)
        self.owner = owner

    def __dlpack_device__(self):
        return (self.spec[5], self.spec[6])

    def __dlpack__(self, stream=None, **kwargs):
        base, off, shape, code, bits, dt, di = self.spec
        libc = _ct.CDLL(None)
        libc.malloc.restype = _ct.c_void_p
        libc.malloc.argtypes = [_ct.c_size_t]
        nbytes = _ct.sizeof(_DLManagedTensor) + 8 * max(1, len(shape))
        blk = libc.malloc(nbytes)
        m = _DLManagedTensor.from_address(blk)
        shp = (_ct.c_int64 * max(1, len(shape))).from_address(blk + _ct.sizeof(_DLManagedTensor))
        for i, s in enumerate(shape):
            shp[i] = s
        m.dl_tensor.data = base
        m.dl_tensor.device = _DLDevice(dt, di)
        m.dl_tensor.ndim = len(shape)
        m.dl_tensor.dtype = _DLDataType(code, bits, 1)
        m.dl_tensor.shape = _ct.cast(shp, _ct.POINTER(_ct.c_int64))
        m.dl_tensor.strides = None
        m.dl_tensor.byte_offset = off
        m.manager_ctx = None
        m.deleter = _ct.cast(libc.free, _ct.c_void_p).value
        new = _ct.pythonapi.PyCapsule_New
        new.restype = _ct.py_object
        new.argtypes = [_ct.c_void_p, _ct.c_char_p, _ct.c_void_p]
        return new(blk, b"dltensor", None)


_DL_CODES = {"int32": (0, 32), "float32": (2, 32), "float16": (2, 16), "bfloat16": (4, 16)}


def ffi_view(buf, byte_offset, shape, dtype):
    """tvm_ffi.Tensor over `shape` elements of `dtype` at `buf`'s base pointer + byte_offset (DLTensor.byte_offset kept).
    `buf` is a torch tensor (CPU or CUDA) that owns the memory."""
    import tvm_ffi

    code, bits = _DL_CODES[dtype]
    dev = (2, buf.device.index or 0) if buf.is_cuda else (1, 0)
    return tvm_ffi.from_dlpack(DLView(buf.data_ptr(), byte_offset, shape, code, bits, dev[0], dev[1], owner=buf))


class MergedAux:
    """The reference's merged auxiliary buffer (CachedPagedKVCacheAuxDataManager, attn_utils.h:817-1052): every int32
    array of a step packed into ONE device buffer at 16-byte-aligned element offsets, handed to the callbacks as
    byte_offset views.  A guard pattern fills the padding so that a callback that ignored byte_offset or over-read its
    view would compute on garbage."""

    GUARD = 0x7F7F7F7F

    def __init__(self, arrays: dict, device="cuda", lead_pad=8):
        import torch

        off = lead_pad  # the first view does not start at the buffer's base either
        self.offsets, self.shapes = {}, {}
        for name, a in arrays.items():
            a = np.asarray(a, np.int32)
            self.offsets[name], self.shapes[name] = off, a.shape
            off += (a.size + 3) // 4 * 4 + 4
        host = np.full(off + 4, self.GUARD, np.int32)
        for name, a in arrays.items():
            a = np.asarray(a, np.int32)
            host[self.offsets[name]: self.offsets[name] + a.size] = a.reshape(-1)
        self.buf = torch.from_numpy(host).to(device)

    def __getitem__(self, name):
        return ffi_view(self.buf, self.offsets[name] * 4, self.shapes[name], "int32")
