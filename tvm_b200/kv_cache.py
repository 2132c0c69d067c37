"""Python face of the host-side paged KV cache (C++: csrc/kv_cache_host.cc, C ABI: include/tvm_b200_cache.h).

Method names / argument meaning follow the reference's `vm.builtin.kv_state_*` and
`vm.builtin.attention_kv_cache_*` packed functions (src/runtime/vm/kv_state.cc:33-116) as the reference's own tests
call them (tests/python/relax/test_runtime_builtin_paged_attention_kv_cache_cpu.py:107-127), so scenario scripts read
the same.  All bookkeeping is in C++; this file only marshals arguments.
"""
from __future__ import annotations

import ctypes
import json

import numpy as np
from ctypes import POINTER, Structure, byref, c_char_p, c_double, c_int32, c_int64, c_void_p

from . import capi


class AttnKind:
    MHA = 0
    MHA_SLIDING = 3


class RopeMode:
    NONE = 0
    NORMAL = 1
    INLINE = 2


class _Config(Structure):
    _fields_ = [
        ("reserved_num_seqs", c_int64), ("total_token_capacity", c_int64), ("prefill_chunk_size", c_int64),
        ("page_size", c_int64), ("support_sliding_window", c_int32), ("layer_sliding_window_size", c_int64),
        ("layer_id_begin_offset", c_int64), ("num_layers", c_int64), ("num_qo_heads", c_int64),
        ("num_kv_heads", c_int64), ("head_dim", c_int64), ("attn_kinds", POINTER(c_int32)), ("rope_mode", c_int32),
        ("rotary_scale", c_double), ("rotary_theta", c_double), ("dtype", c_int32), ("device_id", c_int32),
    ]


_bound = False


def _lib():
    global _bound
    L = capi.lib()
    if not _bound:
        P, I32, I64 = c_void_p, c_int32, c_int64
        L.tvmb200_cache_create.argtypes = [POINTER(_Config), POINTER(P)]
        L.tvmb200_cache_destroy.argtypes = [P]
        L.tvmb200_cache_destroy.restype = None
        L.tvmb200_cache_clear.argtypes = [P]
        L.tvmb200_cache_add_sequence.argtypes = [P, I64]
        L.tvmb200_cache_remove_sequence.argtypes = [P, I64]
        L.tvmb200_cache_fork_sequence.argtypes = [P, I64, I64, I64]
        L.tvmb200_cache_popn.argtypes = [P, I64, I32]
        L.tvmb200_cache_begin_forward.argtypes = [P, P, P, I32, P, I32]
        L.tvmb200_cache_end_forward.argtypes = [P]
        L.tvmb200_cache_enable_sliding_window_for_seq.argtypes = [P, I64, I32, I32]
        L.tvmb200_cache_commit_accepted_token_tree_nodes.argtypes = [P, P, P, I32]
        L.tvmb200_cache_empty.argtypes = [P, POINTER(I32)]
        L.tvmb200_cache_get_num_available_pages.argtypes = [P, POINTER(I32)]
        L.tvmb200_cache_get_total_sequence_length.argtypes = [P, POINTER(I32)]
        L.tvmb200_cache_get_query_positions.argtypes = [P, POINTER(P), POINTER(I64), P]
        L.tvmb200_cache_attention_with_fused_qkv.argtypes = [P, I64, c_double, P, P, I64, P]
        L.tvmb200_cache_self_attention.argtypes = [P, I64, c_double, P, P, P, P, P, I64, P]
        L.tvmb200_cache_cross_attention.argtypes = [P, I64, c_double, P, P, P, I64, P]
        L.tvmb200_cache_attention_with_shared_kv.argtypes = [P, I64, c_double, P, P, P, P, I64, P]
        L.tvmb200_cache_merge_attn_output_inplace.argtypes = [P, P, P, P, P, I64, I64, I64, P]
        L.tvmb200_cache_debug_get_kv.argtypes = [P, I64, I64, I64, P, P, P]
        L.tvmb200_cache_pages.argtypes = [P, I64, POINTER(P), POINTER(I64)]
        L.tvmb200_cache_enable_kv_transfer.argtypes = [P, I32, I32, I32]
        L.tvmb200_cache_set_remote_pages.argtypes = [P, I32, I64, P]
        L.tvmb200_cache_disagg_prepare_recv.argtypes = [P, I64, I64, P, I64, POINTER(I64)]
        L.tvmb200_cache_disagg_mark_send.argtypes = [P, I64, I64, P, I64, I32]
        L.tvmb200_cache_shape.argtypes = [P, POINTER(I64)]
        L.tvmb200_cache_context.argtypes = [P, POINTER(P)]
        L.tvmb200_cache_set_trace.argtypes = [P, I32]
        L.tvmb200_cache_take_trace.argtypes = [P, POINTER(c_char_p)]
        _bound = True
    return L


class _I64Array:
    """int64 view of a Python sequence / numpy array for the C ABI (kept alive by the caller for the call's duration)."""

    __slots__ = ("arr", "_as_parameter_")

    def __init__(self, xs):
        self.arr = np.ascontiguousarray(xs, dtype=np.int64)
        self._as_parameter_ = ctypes.c_void_p(self.arr.__array_interface__["data"][0])


def _i64(xs):
    return _I64Array(xs)


class PagedKVCache:
    """`vm.builtin.paged_attention_kv_cache_create` without the callback arguments: the callbacks are the sm_100a
    kernels of this library.  device=None builds a planning-only cache (bookkeeping + call trace, no GPU)."""

    def __init__(self, *, reserved_num_seqs, total_token_capacity, prefill_chunk_size, page_size=16,
                 support_sliding_window=False, layer_sliding_window_size=None, num_layers, num_qo_heads, num_kv_heads,
                 head_dim, rope_mode=RopeMode.NORMAL, rotary_scale=1.0, rotary_theta=1e4, dtype="float16",
                 attn_kinds=None, layer_id_begin_offset=0, device=0):
        L = _lib()
        self._kinds = None
        cfg = _Config()
        cfg.reserved_num_seqs, cfg.total_token_capacity = reserved_num_seqs, total_token_capacity
        cfg.prefill_chunk_size, cfg.page_size = prefill_chunk_size, page_size
        cfg.support_sliding_window = int(bool(support_sliding_window))
        cfg.layer_sliding_window_size = layer_sliding_window_size or 0
        cfg.layer_id_begin_offset, cfg.num_layers = layer_id_begin_offset, num_layers
        cfg.num_qo_heads, cfg.num_kv_heads, cfg.head_dim = num_qo_heads, num_kv_heads, head_dim
        if attn_kinds is not None:
            self._kinds = (c_int32 * len(attn_kinds))(*attn_kinds)
            cfg.attn_kinds = ctypes.cast(self._kinds, POINTER(c_int32))
        cfg.rope_mode, cfg.rotary_scale, cfg.rotary_theta = int(rope_mode), rotary_scale, rotary_theta
        cfg.dtype = {"float16": capi.F16, "bfloat16": capi.BF16}[dtype]
        cfg.device_id = -1 if device is None else int(device)
        self.dtype, self.device = dtype, device
        self.num_layers, self.num_qo_heads, self.num_kv_heads, self.head_dim = num_layers, num_qo_heads, num_kv_heads, head_dim
        self._h = c_void_p()
        capi._check(L.tvmb200_cache_create(byref(cfg), byref(self._h)))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib().tvmb200_cache_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ---- vm.builtin.kv_state_* ----
    def clear(self):
        capi._check(_lib().tvmb200_cache_clear(self._h))

    def add_sequence(self, seq_id):
        capi._check(_lib().tvmb200_cache_add_sequence(self._h, seq_id))

    def remove_sequence(self, seq_id):
        capi._check(_lib().tvmb200_cache_remove_sequence(self._h, seq_id))

    def fork_sequence(self, parent_seq_id, child_seq_id, fork_pos=-1):
        capi._check(_lib().tvmb200_cache_fork_sequence(self._h, parent_seq_id, child_seq_id, fork_pos))

    def popn(self, seq_id, n):
        capi._check(_lib().tvmb200_cache_popn(self._h, seq_id, n))

    def begin_forward(self, seq_ids, append_lengths, token_tree_parent_ptr=None):
        if len(seq_ids) != len(append_lengths):
            raise capi.TvmB200Error(f"The seq_ids size ({len(seq_ids)}) and append_lengths size ({len(append_lengths)}) mismatch.")
        tree = _i64(token_tree_parent_ptr) if token_tree_parent_ptr is not None else None
        capi._check(_lib().tvmb200_cache_begin_forward(self._h, _i64(seq_ids), _i64(append_lengths), len(seq_ids), tree,
                                                       len(token_tree_parent_ptr) if token_tree_parent_ptr is not None else 0))

    def end_forward(self):
        capi._check(_lib().tvmb200_cache_end_forward(self._h))

    # ---- vm.builtin.attention_kv_cache_* ----
    def enable_sliding_window_for_seq(self, seq_id, sliding_window_size, attn_sink_size):
        capi._check(_lib().tvmb200_cache_enable_sliding_window_for_seq(self._h, seq_id, sliding_window_size, attn_sink_size))

    def commit_accepted_token_tree_nodes(self, seq_ids, leaf_indices):
        if len(seq_ids) != len(leaf_indices):
            raise capi.TvmB200Error("The given seq_ids and leaf_indices have different size.")
        capi._check(_lib().tvmb200_cache_commit_accepted_token_tree_nodes(self._h, _i64(seq_ids), _i64(leaf_indices), len(seq_ids)))

    def empty(self) -> bool:
        out = c_int32()
        capi._check(_lib().tvmb200_cache_empty(self._h, byref(out)))
        return bool(out.value)

    def get_num_available_pages(self) -> int:
        out = c_int32()
        capi._check(_lib().tvmb200_cache_get_num_available_pages(self._h, byref(out)))
        return out.value

    def get_total_sequence_length(self) -> int:
        out = c_int32()
        capi._check(_lib().tvmb200_cache_get_total_sequence_length(self._h, byref(out)))
        return out.value

    def attention_with_fused_qkv(self, layer_id, sm_scale, qkv, o):
        """qkv [n, Hq+2Hkv, D], o [n, Hq, D]: torch CUDA tensors (None for a planning-only cache)."""
        if self.device is None:
            capi._check(_lib().tvmb200_cache_attention_with_fused_qkv(self._h, layer_id, sm_scale, None, None, 1 << 40, None))
            return
        capi._check(_lib().tvmb200_cache_attention_with_fused_qkv(self._h, layer_id, sm_scale, capi._p(qkv), capi._p(o),
                                                                  qkv.shape[0], capi._stream(qkv)))

    def self_attention(self, layer_id, sm_scale, q, k, v, o, lse):
        """q / o [n, Hq, D], k / v [n, Hkv, D], lse [n, Hq] f32: the step's tokens against themselves (kv_state.cc:84-90)."""
        if self.device is None:
            capi._check(_lib().tvmb200_cache_self_attention(self._h, layer_id, sm_scale, None, None, None, None, None,
                                                            self._rows(q), None))
            return
        capi._check(_lib().tvmb200_cache_self_attention(self._h, layer_id, sm_scale, capi._p(q), capi._p(k), capi._p(v),
                                                        capi._p(o), capi._p(lse), q.shape[0], capi._stream(q)))

    def cross_attention(self, layer_id, sm_scale, q, o, lse):
        """q against the cached KV of `layer_id`, no causal mask (kv_state.cc:91-96)."""
        if self.device is None:
            capi._check(_lib().tvmb200_cache_cross_attention(self._h, layer_id, sm_scale, None, None, None, self._rows(q), None))
            return
        capi._check(_lib().tvmb200_cache_cross_attention(self._h, layer_id, sm_scale, capi._p(q), capi._p(o), capi._p(lse),
                                                         q.shape[0], capi._stream(q)))

    def attention_with_shared_kv(self, source_layer_id, sm_scale, q, current_k, current_v, o):
        """A layer without KV of its own attends over the K / V of `source_layer_id` (kv_state.cc:97-104)."""
        if self.device is None:
            capi._check(_lib().tvmb200_cache_attention_with_shared_kv(self._h, source_layer_id, sm_scale, None, None, None,
                                                                      None, self._rows(q), None))
            return
        capi._check(_lib().tvmb200_cache_attention_with_shared_kv(self._h, source_layer_id, sm_scale, capi._p(q),
                                                                  capi._p(current_k), capi._p(current_v), capi._p(o),
                                                                  q.shape[0], capi._stream(q)))

    def merge_attn_output_inplace(self, o_self, lse_self, o_cross, lse_cross):
        """(o_self, lse_self) <- merge with (o_cross, lse_cross); returns the pair like the reference (kv_state.cc:109-115)."""
        if self.device is None:
            n = self._rows(o_self)
            capi._check(_lib().tvmb200_cache_merge_attn_output_inplace(self._h, None, None, None, None, n, self.num_qo_heads,
                                                                       self.head_dim, None))
            return o_self, lse_self
        capi._check(_lib().tvmb200_cache_merge_attn_output_inplace(self._h, capi._p(o_self), capi._p(lse_self),
                                                                   capi._p(o_cross), capi._p(lse_cross), o_self.shape[0],
                                                                   o_self.shape[1], o_self.shape[2], capi._stream(o_self)))
        return o_self, lse_self

    @staticmethod
    def _rows(x):
        """Planning-only caches take the row count (an int) where a device cache takes the tensor."""
        return int(x) if isinstance(x, int) else int(x.shape[0])

    def debug_get_kv(self, seq_id, start_pos, end_pos, k_out=None, v_out=None):
        if self.device is None:
            capi._check(_lib().tvmb200_cache_debug_get_kv(self._h, seq_id, start_pos, end_pos, None, None, None))
            return
        capi._check(_lib().tvmb200_cache_debug_get_kv(self._h, seq_id, start_pos, end_pos, capi._p(k_out), capi._p(v_out),
                                                      capi._stream(k_out)))

    def get_query_positions(self):
        import torch

        ptr, n = c_void_p(), c_int64()
        st = c_void_p(torch.cuda.current_stream().cuda_stream)
        capi._check(_lib().tvmb200_cache_get_query_positions(self._h, byref(ptr), byref(n), st))
        return ptr.value, n.value

    # ---- introspection ----
    def context(self) -> "capi.Context":
        """The cache's own kernel-set context: `with cache.context(): capi.lib().tvmb200_set_rope_scaling(...)` changes
        the rope scaling of this cache only (it starts as a copy of the settings current when the cache was created)."""
        h = c_void_p()
        capi._check(_lib().tvmb200_cache_context(self._h, byref(h)))
        return capi.Context(h.value, owned=False)

    def pages_ptr(self, local_layer=0):
        ptr, n = c_void_p(), c_int64()
        capi._check(_lib().tvmb200_cache_pages(self._h, local_layer, byref(ptr), byref(n)))
        return ptr.value, n.value

    # ---- disaggregated prefill -> decode (vm.builtin.kv_cache_disagg_*; include/tvm_b200_cache.h) ----
    def enable_kv_transfer(self, local_tp_rank=0, num_pe=1, remote_num_kv_heads=None):
        capi._check(_lib().tvmb200_cache_enable_kv_transfer(self._h, local_tp_rank, num_pe,
                                                            remote_num_kv_heads or self.num_kv_heads))

    def set_remote_pages(self, pe, local_layer, peer_mapped_ptr):
        """peer_mapped_ptr: device pointer, valid on THIS cache's GPU, of PE `pe`'s page pool of the layer"""
        capi._check(_lib().tvmb200_cache_set_remote_pages(self._h, pe, local_layer, c_void_p(peer_mapped_ptr)))

    def disagg_prepare_recv(self, seq_id, append_length):
        """-> the reserved slots, run-length compressed [n, begin_1, length_1, ...] (what mark_send takes)"""
        cap = 2 * int(append_length) + 1
        buf, n = (c_int64 * cap)(), c_int64()
        capi._check(_lib().tvmb200_cache_disagg_prepare_recv(self._h, seq_id, append_length, buf, cap, byref(n)))
        return [int(buf[i]) for i in range(n.value)]

    def disagg_mark_send(self, seq_id, begin, compressed_remote_position_map, recver_pe_offset):
        a = _i64(compressed_remote_position_map)
        capi._check(_lib().tvmb200_cache_disagg_mark_send(self._h, seq_id, begin, a, len(compressed_remote_position_map),
                                                          recver_pe_offset))

    def set_trace(self, on=True):
        capi._check(_lib().tvmb200_cache_set_trace(self._h, int(on)))

    def take_trace(self):
        s = c_char_p()
        capi._check(_lib().tvmb200_cache_take_trace(self._h, byref(s)))
        return json.loads(s.value.decode())
