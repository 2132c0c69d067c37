"""ctypes binding of the C ABI declared in include/tvm_b200.h.

Tensors are passed as raw device pointers; the thin helpers below accept torch CUDA tensors (torch is
only the device-memory / stream plumbing) and launch on torch's current stream.  There is no CPU
fallback: a missing library raises at load time, a CPU tensor raises at call time.
"""
from __future__ import annotations

import os
import ctypes
import re
from ctypes import c_char_p, c_float, c_int, c_int32, c_int64, c_void_p
from pathlib import Path

ROOT = Path(__file__).resolve().parent
LIB_PATH = ROOT / "lib" / ("libtvm_b200" + os.environ.get("TVMB200_LIB_SUFFIX", "") + ".so")  # suffix: tuning variants only
HEADER_PATH = ROOT.parent / "include" / "tvm_b200.h"

F16, BF16 = 0, 1

_lib = None


class TvmB200Error(RuntimeError):
    pass


def declared_symbols() -> list[str]:
    """Every `tvmb200_*` entry point declared in include/*.h."""
    text = "\n".join(p.read_text() for p in sorted(HEADER_PATH.parent.glob("*.h")))
    return sorted(set(re.findall(r"TVMB200_API[^;]*?\b(tvmb200_\w+)\s*\(", text, flags=re.S)))


def lib() -> ctypes.CDLL:
    """Load tvm_b200/lib/libtvm_b200.so (built by `python -m tvm_b200.build`); fails loudly if missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise TvmB200Error(
            f"{LIB_PATH} is missing: the sm_100a kernels are not built (run `python -m tvm_b200.build`). "
            "There is no CPU fallback."
        )
    L = ctypes.CDLL(str(LIB_PATH), mode=ctypes.RTLD_GLOBAL)
    L.tvmb200_last_error.restype = c_char_p
    L.tvmb200_version.restype = c_char_p
    L.tvmb200_launch_count.restype = c_int64
    L.tvmb200_reserve_workspace.argtypes = [c_int, c_int64]
    L.tvmb200_reserve_workspace_stream.argtypes = [c_int, c_int64, c_void_p]
    L.tvmb200_context_create.argtypes = [ctypes.POINTER(c_void_p)]
    L.tvmb200_context_retain.argtypes = [c_void_p]
    L.tvmb200_context_retain.restype = None
    L.tvmb200_context_release.argtypes = [c_void_p]
    L.tvmb200_context_release.restype = None
    L.tvmb200_context_enter.argtypes = [c_void_p]
    L.tvmb200_context_enter.restype = c_void_p
    L.tvmb200_set_layer_sliding_window_size.argtypes = [c_int32]
    L.tvmb200_set_layer_sliding_window_size.restype = None
    L.tvmb200_kv_transfer.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int32,
                                      c_int32, c_int32, c_int32, c_int, c_void_p]
    L.tvmb200_kv_transfer_page_to_page.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32,
                                                   c_int32, c_int32, c_int32, c_int32, c_int, c_void_p]
    L.tvmb200_enable_peer_access.argtypes = [c_int32, c_int32]
    L.tvmb200_set_prefill_impl.argtypes = [c_int]
    L.tvmb200_set_prefill_impl.restype = None
    L.tvmb200_debug_prefill_path_counts.argtypes = [ctypes.POINTER(ctypes.c_int64)]
    L.tvmb200_debug_prefill_path_counts.restype = None
    L.tvmb200_set_prefill_prepass_cap.argtypes = [c_int64]
    L.tvmb200_set_prefill_prepass_cap.restype = None
    P, I32, I64, F = c_void_p, c_int32, c_int64, c_float
    L.tvmb200_transpose_append.argtypes = [P, P, P, P, I64, I64, I32, I32, I32, c_int, P]
    L.tvmb200_debug_get_kv.argtypes = [P, P, P, P, I64, I64, I64, I64, I32, I32, I32, c_int, P]
    L.tvmb200_copy_single_page.argtypes = [P, I64, I64, I64, I64, I32, I32, I32, c_int, P]
    L.tvmb200_compact_kv_copy.argtypes = [P, P, P, I32, I32, I64, I32, I32, I32, c_int, P]
    L.tvmb200_split_rotary.argtypes = [P, P, P, P, P, I64, I32, I32, I32, I32, I64, F, F, c_int, P]
    L.tvmb200_split_rotary_append.argtypes = [P, P, P, P, P, P, P, I64, I64, I32, I32, I32, I32, I32, I64, F, F, c_int, P]
    L.tvmb200_merge_state_inplace.argtypes = [P, P, P, P, I64, I32, I32, c_int, P]
    L.tvmb200_attention_decode.argtypes = [P, P, P, P, P, P, P, P, P, I32, I32, I64, I32, I32, I32, I32,
                                           c_int, c_int, F, F, F, c_int, P]
    L.tvmb200_attention_decode_fused_qkv.argtypes = [P, P, P, P, P, P, P, P, P, P, I32, I32, I64, I32, I32, I32, I32,
                                                     c_int, I64, F, F, F, c_int, P]
    L.tvmb200_attention_decode_gather.argtypes = [P, P, P, P, P, P, P, P, P, I32, I32, I64, I32, I32, I32, I32,
                                                  c_int, c_int, F, F, F, c_int, P, P, I32, I32, ctypes.c_uint32, P]
    L.tvmb200_wait_peer_flags.argtypes = [P, I32, ctypes.c_uint32, P]
    L.tvmb200_attention_decode_fused_qkv_gather.argtypes = [P, P, P, P, P, P, P, P, P, P, I32, I32, I64, I32, I32, I32, I32,
                                                            c_int, I64, F, F, F, c_int, P, P, I32, I32, ctypes.c_uint32, P]
    L.tvmb200_attention_prefill_paged.argtypes = [P, P, P, P, P, P, P, P, P, P, I32, I32, I32, I64, I32, I32,
                                                  I32, I32, c_int, I32, c_int, c_int, F, F, F, c_int, P]
    L.tvmb200_attention_prefill_ragged.argtypes = [P, P, P, P, P, P, P, P, P, I32, I32, I32, I32, I32, I32,
                                                   c_int, c_int, F, F, F, c_int, P]
    L.tvmb200_attention_prefill_tree_ragged.argtypes = [P, P, P, P, P, P, P, P, P, P, I32, I32, I32, I32, I32,
                                                        I32, c_int, F, F, F, c_int, P]
    L.tvmb200_attention_prefill_tree_paged.argtypes = [P, P, P, P, P, P, P, P, P, P, I32, I32, I32, I64, I32,
                                                       I32, I32, I32, c_int, F, F, F, P, P, c_int, P]
    _lib = L
    return L


def _check(rc: int) -> None:
    if rc != 0:
        raise TvmB200Error(lib().tvmb200_last_error().decode())


class Context:
    """A kernel-set context of include/tvm_b200.h (rope scaling, layer window, device scratch).  `with ctx:` makes it
    the calling thread's current context: setters and launches inside the block use it."""

    def __init__(self, handle=None, owned=True):
        if handle is None:
            h = c_void_p()
            _check(lib().tvmb200_context_create(ctypes.byref(h)))
            handle = h.value
        self.handle, self._owned, self._prev = handle, owned, []

    def __enter__(self):
        self._prev.append(lib().tvmb200_context_enter(c_void_p(self.handle)))
        return self

    def __exit__(self, *exc):
        lib().tvmb200_context_enter(c_void_p(self._prev.pop()))
        return False

    def __del__(self):
        if self._owned and self.handle:
            try:
                lib().tvmb200_context_release(c_void_p(self.handle))
            except Exception:  # interpreter shutdown
                pass


def set_prefill_impl(impl: int) -> None:
    """0 auto, 1 force the generic mma.sync kernel, 2 force the tcgen05 kernel where eligible."""
    lib().tvmb200_set_prefill_impl(impl)


def prefill_path_counts() -> tuple[int, int, int, int]:
    """prefill launches so far: (mma.sync kernel, tcgen05 kernel, tcgen05 kernel behind the gather / rotate pre-pass,
    tcgen05 launches that split their items' KV range)"""
    out = (ctypes.c_int64 * 4)()
    lib().tvmb200_debug_prefill_path_counts(out)
    return int(out[0]), int(out[1]), int(out[2]), int(out[3])


def set_prefill_prepass_cap(nbytes: int) -> None:
    lib().tvmb200_set_prefill_prepass_cap(nbytes)


def launch_count() -> int:
    return int(lib().tvmb200_launch_count())


# ---- torch-tensor helpers -------------------------------------------------------------------------
def _dt(t) -> int:
    import torch

    if t.dtype == torch.float16:
        return F16
    if t.dtype == torch.bfloat16:
        return BF16
    raise TvmB200Error(f"unsupported dtype {t.dtype}: the sm_100a kernels take float16 / bfloat16")


def _p(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise TvmB200Error("tensor is not on a CUDA device; tvm_b200 has no CPU fallback")
    if not t.is_contiguous():
        raise TvmB200Error("tensor must be compact row-major")
    return c_void_p(t.data_ptr())


def _stream(t):
    import torch

    return c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def transpose_append(pages, k, v, position_map):
    P, _, Hkv, page, D = pages.shape
    _check(lib().tvmb200_transpose_append(_p(pages), _p(k), _p(v), _p(position_map), k.shape[0], P, Hkv, page, D,
                                          _dt(pages), _stream(pages)))


def debug_get_kv(pages, position_map, k_out, v_out, layer_id):
    P, _, Hkv, page, D = pages.shape
    _check(lib().tvmb200_debug_get_kv(_p(pages), _p(position_map), _p(k_out), _p(v_out), layer_id, k_out.shape[0],
                                      k_out.shape[1], P, Hkv, page, D, _dt(pages), _stream(pages)))


def copy_single_page(pages, src, tgt, copy_length):
    P, _, Hkv, page, D = pages.shape
    _check(lib().tvmb200_copy_single_page(_p(pages), src, tgt, copy_length, P, Hkv, page, D, _dt(pages), _stream(pages)))


def compact_kv_copy(pages, copy_length_indptr, copy_src_dst_pos, batch_size):
    P, _, Hkv, page, D = pages.shape
    _check(lib().tvmb200_compact_kv_copy(_p(pages), _p(copy_length_indptr), _p(copy_src_dst_pos), batch_size,
                                         copy_src_dst_pos.shape[1], P, Hkv, page, D, _dt(pages), _stream(pages)))


def split_rotary(qkv, position_map, q, k, v, apply_rope, rope_scale, rope_theta, rotary_dim=0):
    _check(lib().tvmb200_split_rotary(_p(qkv), _p(position_map), _p(q), _p(k), _p(v), qkv.shape[0], q.shape[1],
                                      k.shape[1], qkv.shape[2], rotary_dim, apply_rope, rope_scale, rope_theta,
                                      _dt(qkv), _stream(qkv)))


def split_rotary_append(qkv, q_rope_position, append_position, q, k, v, pages, apply_rope, rope_scale, rope_theta,
                        rotary_dim=0):
    """f_split_rotary + f_transpose_append in one launch (bit-identical to the two calls)."""
    P, _, Hkv, page, D = pages.shape
    _check(lib().tvmb200_split_rotary_append(_p(qkv), _p(q_rope_position), _p(append_position), _p(q), _p(k), _p(v),
                                             _p(pages), qkv.shape[0], P, q.shape[1], Hkv, page, D, rotary_dim,
                                             apply_rope, rope_scale, rope_theta, _dt(qkv), _stream(qkv)))


ROPE_SCALING_KINDS = {"llama3": 1, "gptj": 2, "llama4": 3, "yarn": 5}


def set_rope_scaling(rope_scaling=None):
    """rope_scaling: None / {} (default frequencies) or the reference's dict (switch_rope_freq_func,
    position_embedding.py:257-299): {"rope_type": "llama3" | "llama4", "factor", "low_freq_factor", "high_freq_factor",
    "original_max_position_embeddings"}, {"rope_type": "gptj"}, or {"rope_type": "yarn", "factor",
    "original_max_position_embeddings", "beta_fast", "beta_slow"[, "inv_theta_log_scale"]}.  Process-wide, like the
    reference's build-time choice.  gptj / llama4 / yarn only exist in split_rotary (see include/tvm_b200.h)."""
    L = lib()
    F = ctypes.c_float
    L.tvmb200_set_rope_scaling.argtypes = [ctypes.c_int32, F, F, F, F]
    L.tvmb200_set_rope_scaling_yarn.argtypes = [F, F, F, F, F]
    if not rope_scaling or "rope_type" not in rope_scaling:
        _check(L.tvmb200_set_rope_scaling(0, 1.0, 0.0, 1.0, 1.0))
        return
    kind = ROPE_SCALING_KINDS.get(rope_scaling["rope_type"])
    if kind is None:
        raise TvmB200Error(f"rope_type {rope_scaling.get('rope_type')!r} is not implemented "
                           f"(default, {', '.join(ROPE_SCALING_KINDS)} are)")
    if kind == 2:
        _check(L.tvmb200_set_rope_scaling(2, 1.0, 0.0, 1.0, 1.0))
    elif kind == 5:
        _check(L.tvmb200_set_rope_scaling_yarn(float(rope_scaling["factor"]),
                                               float(rope_scaling["original_max_position_embeddings"]),
                                               float(rope_scaling["beta_fast"]), float(rope_scaling["beta_slow"]),
                                               float(rope_scaling.get("inv_theta_log_scale") or 0.0)))
    else:
        _check(L.tvmb200_set_rope_scaling(kind, float(rope_scaling["factor"]), float(rope_scaling["low_freq_factor"]),
                                          float(rope_scaling["high_freq_factor"]),
                                          float(rope_scaling["original_max_position_embeddings"])))


def get_rope_scaling_kind() -> int:
    L = lib()
    L.tvmb200_get_rope_scaling_kind.restype = ctypes.c_int32
    return int(L.tvmb200_get_rope_scaling_kind())


def merge_state_inplace(v, s, v_other, s_other):
    _check(lib().tvmb200_merge_state_inplace(_p(v), _p(s), _p(v_other), _p(s_other), v.shape[0], v.shape[1],
                                             v.shape[2], _dt(v), _stream(v)))


def attention_decode(q, pages, page_indptr, page_values, length_info, k_rope_pos_offset, q_rope_position, output,
                     lse, rotary_mode, rope_scale, rope_theta, sm_scale):
    P, _, Hkv, page, D = pages.shape
    _check(lib().tvmb200_attention_decode(
        _p(q), _p(pages), _p(page_indptr), _p(page_values), _p(length_info), _p(k_rope_pos_offset),
        _p(q_rope_position), _p(output), _p(lse), q.shape[0], page_values.shape[0], P, q.shape[1], Hkv, page, D,
        1 if length_info.dim() == 2 else 0, rotary_mode, rope_scale, rope_theta, sm_scale, _dt(pages), _stream(q)))


def attention_decode_fused_qkv(qkv, q_rope_position, append_position, pages, page_indptr, page_values, length_info,
                               k_rope_pos_offset, output, lse, apply_rope, rope_scale, rope_theta, sm_scale):
    """f_split_rotary + f_transpose_append + f_attention_decode of a one-token-per-sequence batch in one launch."""
    P, _, Hkv, page, D = pages.shape
    hq = qkv.shape[1] - 2 * Hkv
    _check(lib().tvmb200_attention_decode_fused_qkv(
        _p(qkv), _p(q_rope_position), _p(append_position), _p(pages), _p(page_indptr), _p(page_values), _p(length_info),
        _p(k_rope_pos_offset), _p(output), _p(lse), qkv.shape[0], page_values.shape[0], P, hq, Hkv, page, D,
        1 if length_info.dim() == 2 else 0, apply_rope, rope_scale, rope_theta, sm_scale, _dt(pages), _stream(qkv)))


def attention_decode_gather(q, pages, page_indptr, page_values, length_info, k_rope_pos_offset, q_rope_position, output,
                            lse, rotary_mode, rope_scale, rope_theta, sm_scale, peer_output_ptrs, peer_flag_ptrs, rank,
                            epoch):
    """attention_decode of this rank's KV-head shard + peer stores of its heads into every rank's gathered buffer.
    peer_output_ptrs / peer_flag_ptrs: lists of `world` integer device addresses (peer-mapped)."""
    P, _, Hkv, page, D = pages.shape
    world = len(peer_output_ptrs)
    outs = (c_void_p * world)(*[int(x) for x in peer_output_ptrs])
    flags = (c_void_p * world)(*[int(x) for x in peer_flag_ptrs])
    _check(lib().tvmb200_attention_decode_gather(
        _p(q), _p(pages), _p(page_indptr), _p(page_values), _p(length_info), _p(k_rope_pos_offset),
        _p(q_rope_position), _p(output), _p(lse), q.shape[0], page_values.shape[0], P, q.shape[1], Hkv, page, D,
        1 if length_info.dim() == 2 else 0, rotary_mode, rope_scale, rope_theta, sm_scale, _dt(pages),
        ctypes.cast(outs, c_void_p), ctypes.cast(flags, c_void_p), world, rank, epoch & 0xFFFFFFFF, _stream(q)))


def attention_decode_fused_qkv_gather(qkv, q_rope_position, append_position, pages, page_indptr, page_values, length_info,
                                      k_rope_pos_offset, output, lse, apply_rope, rope_scale, rope_theta, sm_scale,
                                      peer_output_ptrs, peer_flag_ptrs, rank, epoch):
    P, _, Hkv, page, D = pages.shape
    hq = qkv.shape[1] - 2 * Hkv
    world = len(peer_output_ptrs)
    outs = (c_void_p * world)(*[int(x) for x in peer_output_ptrs])
    flags = (c_void_p * world)(*[int(x) for x in peer_flag_ptrs])
    _check(lib().tvmb200_attention_decode_fused_qkv_gather(
        _p(qkv), _p(q_rope_position), _p(append_position), _p(pages), _p(page_indptr), _p(page_values), _p(length_info),
        _p(k_rope_pos_offset), _p(output), _p(lse), qkv.shape[0], page_values.shape[0], P, hq, Hkv, page, D,
        1 if length_info.dim() == 2 else 0, apply_rope, rope_scale, rope_theta, sm_scale, _dt(pages),
        ctypes.cast(outs, c_void_p), ctypes.cast(flags, c_void_p), world, rank, epoch & 0xFFFFFFFF, _stream(qkv)))


def _ptr_table(ptrs):
    return (c_void_p * len(ptrs))(*[c_void_p(int(x)) for x in ptrs])


def kv_transfer(remote_pages_ptrs, k, v, remote_position_map, remote_tp_group_pe_offset, remote_num_kv_heads, page_size,
                local_tp_rank=0):
    """nvshmem.KVTransfer over peer-mapped pointers: remote_pages_ptrs[pe] = device address of PE pe's page pool"""
    _check(lib().tvmb200_kv_transfer(_ptr_table(remote_pages_ptrs), _p(k), _p(v), _p(remote_position_map),
                                     _p(remote_tp_group_pe_offset), k.shape[0], k.shape[1], remote_num_kv_heads, page_size,
                                     k.shape[2], local_tp_rank, len(remote_pages_ptrs), _dt(k), _stream(k)))


def kv_transfer_page_to_page(remote_pages_ptrs, local_pages, remote_position_map, local_position_map,
                             remote_tp_group_pe_offset, remote_num_kv_heads, local_tp_rank=0):
    _, _, hkv, page, d = local_pages.shape
    _check(lib().tvmb200_kv_transfer_page_to_page(
        _ptr_table(remote_pages_ptrs), _p(local_pages), _p(remote_position_map), _p(local_position_map),
        _p(remote_tp_group_pe_offset), remote_position_map.shape[0], hkv, remote_num_kv_heads, page, d, local_tp_rank,
        len(remote_pages_ptrs), _dt(local_pages), _stream(local_pages)))


def enable_peer_access(device, peer):
    _check(lib().tvmb200_enable_peer_access(device, peer))


def wait_peer_flags(flags, world, epoch):
    _check(lib().tvmb200_wait_peer_flags(_p(flags), world, epoch & 0xFFFFFFFF, _stream(flags)))


def attention_prefill_paged(q, q_indptr, pages, page_indptr, page_values, length_info, k_rope_pos_offset,
                            q_rope_position, output, lse, causal, rotary_mode, rope_scale, rope_theta, sm_scale,
                            layer_sliding_window_size=0):
    P, _, Hkv, page, D = pages.shape
    _check(lib().tvmb200_attention_prefill_paged(
        _p(q), _p(q_indptr), _p(pages), _p(page_indptr), _p(page_values), _p(length_info), _p(k_rope_pos_offset),
        _p(q_rope_position), _p(output), _p(lse), q_indptr.shape[0] - 1, q.shape[0], page_values.shape[0], P,
        q.shape[1], Hkv, page, D, 1 if length_info.dim() == 2 else 0, layer_sliding_window_size, causal,
        rotary_mode, rope_scale, rope_theta, sm_scale, _dt(pages), _stream(q)))


def attention_prefill_ragged(q, q_indptr, k, v, kv_indptr, q_rope_position, k_rope_pos_offset, output, lse, causal,
                             rotary_mode, rope_scale, rope_theta, sm_scale):
    _check(lib().tvmb200_attention_prefill_ragged(
        _p(q), _p(q_indptr), _p(k), _p(v), _p(kv_indptr), _p(q_rope_position), _p(k_rope_pos_offset), _p(output),
        _p(lse), q_indptr.shape[0] - 1, q.shape[0], k.shape[0], q.shape[1], k.shape[1], q.shape[2], causal,
        rotary_mode, rope_scale, rope_theta, sm_scale, _dt(q), _stream(q)))


def attention_prefill_tree_ragged(q, q_indptr, k, v, kv_indptr, q_rope_position, mn_indptr, mask, output, lse,
                                  rotary_mode, rope_scale, rope_theta, sm_scale):
    _check(lib().tvmb200_attention_prefill_tree_ragged(
        _p(q), _p(q_indptr), _p(k), _p(v), _p(kv_indptr), _p(q_rope_position), _p(mn_indptr), _p(mask), _p(output),
        _p(lse), q_indptr.shape[0] - 1, q.shape[0], k.shape[0], q.shape[1], k.shape[1], q.shape[2], rotary_mode,
        rope_scale, rope_theta, sm_scale, _dt(q), _stream(q)))


def attention_prefill_tree_paged(q, q_indptr, pages, page_indptr, page_values, length_info, k_rope_pos_offset,
                                 q_rope_position, output, lse, rotary_mode, rope_scale, rope_theta, sm_scale,
                                 tree_order_indptr, tree_order):
    P, _, Hkv, page, D = pages.shape
    _check(lib().tvmb200_attention_prefill_tree_paged(
        _p(q), _p(q_indptr), _p(pages), _p(page_indptr), _p(page_values), _p(length_info), _p(k_rope_pos_offset),
        _p(q_rope_position), _p(output), _p(lse), q_indptr.shape[0] - 1, q.shape[0], page_values.shape[0], P,
        q.shape[1], Hkv, page, D, rotary_mode, rope_scale, rope_theta, sm_scale, _p(tree_order_indptr),
        _p(tree_order), _dt(pages), _stream(q)))
