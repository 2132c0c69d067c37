"""CPU oracle for the PagedKVCache kernel set -- TEST INFRASTRUCTURE ONLY.

This is a NumPy restatement of the reference's CPU TIR kernels (apache/tvm,
python/tvm/relax/frontend/nn/llm/*).  It is imported only by tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline leg, always as the *checker*; the product (tvm_b200/) never imports it.

Parity status: pinned against outputs of the real reference (its own C++ PagedAttentionKVCacheObj
driving its own CPU TIR PrimFuncs, built from /root/reference by oracle/ref_harness/) through the
fixtures in tests/golden/ -- see tests/test_oracle_golden.py.

Conventions: 16-bit tensors are carried as float32 arrays whose values are exactly representable
in the 16-bit type (use round_dtype); index arrays are int32.  Attention arithmetic is done in
float64 and rounded once at the end, i.e. this is the mathematically exact answer the reference's
float32 kernels approximate (its own tests use a float32 NumPy oracle at 1e-3,
tests/python/relax/test_runtime_builtin_paged_attention_kv_cache_cpu.py:530-535).
"""
from __future__ import annotations

import math

import numpy as np

NEG_INIT = -5e4  # running-max sentinel of every reference kernel (_decode_kernels.py:134)
LOG2E = math.log2(math.e)


# ------------------------------------------------------------------------------------------------
# dtype helpers
# ------------------------------------------------------------------------------------------------
def round_dtype(x: np.ndarray, dtype: str) -> np.ndarray:
    """Round float values to `dtype` (round-to-nearest-even) and return them as float32."""
    x = np.asarray(x, dtype=np.float32)
    if dtype == "float32":
        return x
    if dtype == "float16":
        return x.astype(np.float16).astype(np.float32)
    if dtype == "bfloat16":
        u = x.view(np.uint32).astype(np.uint64)
        rounded = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
        out = rounded.astype(np.uint32).view(np.float32)
        return np.where(np.isnan(x), x, out).astype(np.float32)
    raise ValueError(f"unsupported dtype {dtype}")


def to_bits16(x: np.ndarray, dtype: str) -> np.ndarray:
    """The 16-bit pattern (uint16) of values already representable in `dtype`."""
    x = np.asarray(x, dtype=np.float32)
    if dtype == "float16":
        return x.astype(np.float16).view(np.uint16)
    if dtype == "bfloat16":
        return (x.view(np.uint32) >> 16).astype(np.uint16)
    raise ValueError(dtype)


def from_bits16(b: np.ndarray, dtype: str) -> np.ndarray:
    b = np.asarray(b, dtype=np.uint16)
    if dtype == "float16":
        return b.view(np.float16).astype(np.float32)
    if dtype == "bfloat16":
        return (b.astype(np.uint32) << 16).view(np.float32)
    raise ValueError(dtype)


# ------------------------------------------------------------------------------------------------
# RoPE  (position_embedding.py:31-67 rope_freq_default, :500-521 _rope; _kernel_common.py:115-127)
# ------------------------------------------------------------------------------------------------
# Frequency scaling the reference bakes into its PrimFuncs at build time (`rope_scaling` dict -> switch_rope_freq_func,
# position_embedding.py:257-299).  None = rope_freq_default; {"rope_type": ...} selects rope_freq_gptj (:70-76),
# rope_freq_llama4 (:79-127), rope_freq_llama3 (:130-160) or rope_freq_yarn (:223-254) with the dict's parameters
# (longrope is not restated: the reference's own longrope PrimFunc does not build at this commit, so nothing could pin it).  Module state, like the kernels' tvmb200_set_rope_scaling: set_rope_scaling(None) restores the
# default.
_ROPE_SCALING = None
_ROPE_TYPES = ("llama3", "gptj", "llama4", "yarn")


def set_rope_scaling(rs):
    global _ROPE_SCALING
    if rs is not None and rs.get("rope_type") not in _ROPE_TYPES:
        raise ValueError(f"oracle: rope_type {rs.get('rope_type')!r} is not restated ({', '.join(_ROPE_TYPES)} are)")
    _ROPE_SCALING = dict(rs) if rs is not None else None


def rope_cos_sin(s: np.ndarray, rd: int, theta: float):
    """cos, sin [n, rd] (float32) of the rotation angle of element d at scaled position s[n], as the reference's
    rope_freq_* functions compute them (float32 arithmetic, float64 only inside cos / sin)."""
    f32 = np.float32
    rs = _ROPE_SCALING
    kind = rs.get("rope_type") if rs else None
    d = np.arange(rd)
    s = np.asarray(s, dtype=np.float32)
    if kind in ("gptj", "llama4"):
        e = 2 * (d // 2)
        if kind == "gptj":
            e = e % rd                                        # :72; llama4 has no modulo (:92)
        expo = e.astype(np.float32) / f32(rd)
    else:
        expo = ((d * 2) % rd).astype(np.float32) / f32(rd)
    denom = np.power(f32(theta), expo).astype(np.float32)     # [rd]
    if kind in (None, "gptj"):
        freq = (s[:, None] / denom[None, :]).astype(np.float32)
    elif kind in ("llama3", "llama4"):
        orig = (f32(1) / denom).astype(np.float32)
        if kind == "llama4" and rs["high_freq_factor"] == rs["low_freq_factor"]:
            wavelength = (f32(2 * np.pi) / orig).astype(np.float32)
            thr = f32(rs["original_max_position_embeddings"] / rs["low_freq_factor"])
            inv = np.where(wavelength > thr, (orig / f32(rs["factor"])).astype(np.float32), orig).astype(np.float32)
        else:
            # orig = 1/theta^e; smooth = clip(alpha*orig - beta, 0, 1); inv = (1 - smooth)*orig/factor + smooth*orig
            inv_diff = 1.0 / (rs["high_freq_factor"] - rs["low_freq_factor"])
            alpha = f32(rs["original_max_position_embeddings"] / (2 * np.pi) * inv_diff)
            beta = f32(rs["low_freq_factor"] * inv_diff)
            smooth = np.maximum(f32(0), np.minimum(f32(1), alpha * orig - beta)).astype(np.float32)
            inv = ((f32(1) - smooth) * orig * f32(1.0 / rs["factor"]) + smooth * orig).astype(np.float32)
        freq = (s[:, None] * inv[None, :]).astype(np.float32)
    elif kind == "yarn":
        # the correction range is in units of the element index d in [0, rd) and the ramp is NOT folded onto rd/2: the
        # two halves of a rotated pair get different angles, exactly as the reference computes them (:244-249)
        itls = rs.get("inv_theta_log_scale")
        if itls is None:
            itls = 1.0 / (2 * np.log(theta))                  # kv_cache.py:355-366
        orig_max = rs["original_max_position_embeddings"]
        low = max(rd * np.log(orig_max / (rs["beta_fast"] * 2 * np.pi)) * itls, 0)
        high = min(rd * np.log(orig_max / (rs["beta_slow"] * 2 * np.pi)) * itls, rd - 1)
        if low == high:
            high = high + 0.001
        freq_extra = (f32(1) / denom).astype(np.float32)
        freq_inter = (f32(1) / (f32(rs["factor"]) * denom)).astype(np.float32)
        ramp = ((d.astype(np.float32) - f32(low)) / f32(high - low)).astype(np.float32)
        mask = (f32(1) - np.maximum(np.minimum(ramp, f32(1)), f32(0))).astype(np.float32)
        inv = (freq_inter * (f32(1) - mask) + freq_extra * mask).astype(np.float32)
        freq = (s[:, None] * inv[None, :]).astype(np.float32)
    else:
        raise ValueError(kind)
    cos = np.cos(freq.astype(np.float64)).astype(np.float32)
    sin = np.sin(freq.astype(np.float64)).astype(np.float32)
    return cos, sin


def rope_rotate(x: np.ndarray, pos: np.ndarray, theta: float, scale: float, dtype: str,
                rotary_dim: int | None = None, interleaved: bool = False) -> np.ndarray:
    """x: [n, H, D] (values in dtype), pos: [n] int.  Returns rotated x rounded to dtype.

    out[d] = cos[d]*x[d] + sin[d]*partner(d); partner = (d<rd/2 ? -x[d+rd/2] : x[d-rd/2]), or, interleaved (the gptj
    layout of f_split_rotary, position_embedding.py:509-514), (d even ? -x[d+1] : x[d-1]).  The inline-RoPE helper of the
    attention kernels (_kernel_common.py:115-127) always pairs by halves.  cos/sin/products in float32 like the
    reference; the result is cast to dtype once.
    """
    x = np.asarray(x, dtype=np.float32)
    n, H, D = x.shape
    rd = D if rotary_dim is None else rotary_dim
    s = np.asarray(pos, dtype=np.float32) * np.float32(scale)  # [n]
    cos, sin = rope_cos_sin(s, rd, theta)
    cos, sin = cos[:, None, :], sin[:, None, :]
    xr = x[..., :rd]
    half = rd // 2
    if interleaved:
        rot = np.empty_like(xr)
        rot[..., 0::2] = -xr[..., 1::2]
        rot[..., 1::2] = xr[..., 0::2]
    else:
        rot = np.concatenate([-xr[..., half:], xr[..., :half]], axis=-1)
    out = x.copy()
    out[..., :rd] = round_dtype(cos * xr + sin * rot, dtype)
    return out


def split_rotary(qkv: np.ndarray, position_map: np.ndarray, num_q_heads: int, num_kv_heads: int,
                 apply_rope: int, theta: float, scale: float, dtype: str,
                 rotary_dim: int | None = None):
    """f_split_rotary = llama_rope_with_position_map (position_embedding.py:444-565); the gptj scaling pairs
    interleaved elements (:509-514), every other one pairs the two halves."""
    q = qkv[:, :num_q_heads].copy()
    k = qkv[:, num_q_heads:num_q_heads + num_kv_heads].copy()
    v = qkv[:, num_q_heads + num_kv_heads:].copy()
    gptj = bool(_ROPE_SCALING) and _ROPE_SCALING.get("rope_type") == "gptj"
    if apply_rope > 0:
        q = rope_rotate(q, position_map, theta, scale, dtype, rotary_dim, interleaved=gptj)
        k = rope_rotate(k, position_map, theta, scale, dtype, rotary_dim, interleaved=gptj)
    return q, k, v


# ------------------------------------------------------------------------------------------------
# page data movement (bit exact)  (_page_kernels.py)
# ------------------------------------------------------------------------------------------------
def transpose_append(pages: np.ndarray, k: np.ndarray, v: np.ndarray, position_map: np.ndarray) -> None:
    """_page_kernels.py:40-74.  pages [P,2,Hkv,page,D] mutated in place."""
    page_size = pages.shape[3]
    for t, pos in enumerate(np.asarray(position_map)):
        if pos == -1:
            continue
        pages[pos // page_size, 0, :, pos % page_size, :] = k[t]
        pages[pos // page_size, 1, :, pos % page_size, :] = v[t]


def debug_get_kv(pages: np.ndarray, position_map: np.ndarray):
    """_page_kernels.py:106-136 for one layer.  Returns k, v [seqlen, Hkv, D]."""
    page_size = pages.shape[3]
    pm = np.asarray(position_map)
    k = pages[pm // page_size, 0, :, pm % page_size, :]
    v = pages[pm // page_size, 1, :, pm % page_size, :]
    return k.copy(), v.copy()


def copy_single_page(pages: np.ndarray, src: int, tgt: int, copy_length: int) -> None:
    """_page_kernels.py:169-189."""
    pages[tgt, :, :, :copy_length, :] = pages[src, :, :, :copy_length, :]


def compact_kv_copy(pages: np.ndarray, copy_length_indptr: np.ndarray, copy_src_dst_pos: np.ndarray,
                    batch_size: int) -> None:
    """_page_kernels.py:235-263: per sequence, serially in list order, slot dst <- slot src."""
    page_size = pages.shape[3]
    for b in range(batch_size):
        for i in range(int(copy_length_indptr[b]), int(copy_length_indptr[b + 1])):
            s, d = int(copy_src_dst_pos[0, i]), int(copy_src_dst_pos[1, i])
            pages[d // page_size, :, :, d % page_size, :] = pages[s // page_size, :, :, s % page_size, :]


# ------------------------------------------------------------------------------------------------
# merge  (_decode_kernels.py:414-452)
# ------------------------------------------------------------------------------------------------
def merge_state_inplace(v: np.ndarray, s: np.ndarray, v_other: np.ndarray, s_other: np.ndarray, dtype: str):
    """Returns (v_new rounded to dtype, s_new float32); arithmetic in float32 like the reference."""
    s = np.asarray(s, np.float32)
    s_other = np.asarray(s_other, np.float32)
    s_max = np.maximum(s, s_other)
    a = np.exp2((s - s_max).astype(np.float64))
    b = np.exp2((s_other - s_max).astype(np.float64))
    scale = (a / (a + b))[..., None]
    other = (b / (a + b))[..., None]
    v_new = round_dtype(np.asarray(v, np.float64) * scale + np.asarray(v_other, np.float64) * other, dtype)
    s_new = (np.log2(a + b) + s_max).astype(np.float32)
    return v_new, s_new


# ------------------------------------------------------------------------------------------------
# attention core
# ------------------------------------------------------------------------------------------------
def _attend(q: np.ndarray, k: np.ndarray, v: np.ndarray, mask: np.ndarray, sm_scale: float, group: int):
    """q [n,Hq,D], k,v [m,Hkv,D], mask [n,m] bool (True = visible).  float64 math.

    Returns o [n,Hq,D] float64, lse [n,Hq] float64 (base 2).  Rows without any visible column give
    O = 0, lse = -5e4 (the reference's empty-KV result: m = -5e4, d = 1).
    """
    n, Hq, D = q.shape
    m = k.shape[0]
    o = np.zeros((n, Hq, D), np.float64)
    lse = np.full((n, Hq), NEG_INIT, np.float64)
    if n == 0 or m == 0:
        return o, lse
    qd = q.astype(np.float64)
    kd = np.repeat(k.astype(np.float64), group, axis=1)  # [m,Hq,D]
    vd = np.repeat(v.astype(np.float64), group, axis=1)
    s = np.einsum("nhd,mhd->hnm", qd, kd) * (sm_scale * LOG2E)
    s = np.where(mask[None, :, :], s, -np.inf)
    mx = s.max(axis=-1)  # [h,n]
    has = np.isfinite(mx)
    mx_safe = np.where(has, mx, 0.0)
    p = np.exp2(s - mx_safe[..., None])
    p = np.where(mask[None, :, :], p, 0.0)
    den = p.sum(axis=-1)
    den_safe = np.where(has, den, 1.0)
    oo = np.einsum("hnm,mhd->nhd", p / den_safe[..., None], vd)
    o[:] = oo
    lse_v = np.where(has, mx_safe + np.log2(den_safe), NEG_INIT)
    lse[:] = lse_v.T
    return o, lse


def _paged_kv_len(npages: int, b: int, length_info: np.ndarray, page_size: int) -> int:
    """_kernel_common.py:155-159 _get_kv_chunk_len (0 pages => 0)."""
    if npages == 0:
        return 0
    li = np.asarray(length_info)
    if li.ndim == 1:
        return (npages - 1) * page_size + int(li[b])
    return (npages - 1) * page_size + int(li[0, b]) - int(li[1, b]) + int(li[2, b])


def _gather_paged(pages, page_values, pg_beg, kv_len, b, length_info, h_all=True):
    """Rows 0..kv_len-1 of sequence b: K,V [kv_len,Hkv,D] (_get_seq_offset, _kernel_common.py:162-170)."""
    page_size = pages.shape[3]
    pos = np.arange(kv_len)
    li = np.asarray(length_info)
    if li.ndim == 2:
        sink, off = int(li[2, b]), int(li[1, b])
        slot = np.where(pos < sink, pos, pos - sink + off)
    else:
        slot = pos
    pg = np.asarray(page_values)[pg_beg + slot // page_size]
    k = pages[pg, 0, :, slot % page_size, :]
    v = pages[pg, 1, :, slot % page_size, :]
    return k, v


def attention_decode(q, pages, page_indptr, page_values, length_info, k_rope_pos_offset, q_rope_position,
                     rotary_mode, rope_scale, rope_theta, sm_scale, dtype):
    """_attention_decode_cpu (_decode_kernels.py:49-178).  q [B,Hq,D] -> (o [B,Hq,D] in dtype, lse [B,Hq] f32)."""
    B, Hq, D = q.shape
    Hkv = pages.shape[2]
    group = Hq // Hkv
    page_size = pages.shape[3]
    o = np.zeros((B, Hq, D), np.float32)
    lse = np.full((B, Hq), NEG_INIT, np.float32)
    for b in range(B):
        pb, pe = int(page_indptr[b]), int(page_indptr[b + 1])
        kv_len = _paged_kv_len(pe - pb, b, length_info, page_size)
        if kv_len <= 0:
            continue
        k, v = _gather_paged(pages, page_values, pb, kv_len, b, length_info)
        qb = q[b:b + 1]
        if rotary_mode == 1:
            qb = rope_rotate(qb, np.asarray(q_rope_position)[b:b + 1], rope_theta, rope_scale, dtype)
            k = rope_rotate(k, int(k_rope_pos_offset[b]) + np.arange(kv_len), rope_theta, rope_scale, dtype)
        ob, lb = _attend(qb, k, v, np.ones((1, kv_len), bool), sm_scale, group)
        o[b] = round_dtype(ob[0], dtype)
        lse[b] = lb[0].astype(np.float32)
    return o, lse


def _cross_mask(causal, qo_len, kv_len, sliding_window_size):
    """_causal_or_sliding_cross_mask (_kernel_common.py:130-144) as a [qo_len, kv_len] bool matrix."""
    row = np.arange(qo_len)[:, None]
    col = np.arange(kv_len)[None, :]
    if causal > 0 and sliding_window_size > 0:
        visible_past = np.maximum(sliding_window_size - row - 1, 0)
        return (col < kv_len) & (col >= np.maximum(kv_len - visible_past, 0))
    if causal > 0:
        return col < kv_len - qo_len + row + 1
    return np.broadcast_to(col < kv_len, (qo_len, kv_len))


def attention_prefill_paged(q, q_indptr, pages, page_indptr, page_values, length_info, k_rope_pos_offset,
                            q_rope_position, causal, rotary_mode, rope_scale, rope_theta, sm_scale, dtype,
                            sliding_window_size=0, tree_indptr=None, tree_order=None):
    """_attention_prefill_cpu (_prefill_kernels.py:54-214); with tree_* = tree_attn_with_paged_kv_cache_cpu
    (tree_attn.py:606-795).  sliding_window_size is only non-zero for the `_sliding_window` flavour."""
    n, Hq, D = q.shape
    Hkv = pages.shape[2]
    group = Hq // Hkv
    page_size = pages.shape[3]
    B = len(q_indptr) - 1
    o = np.zeros((n, Hq, D), np.float32)
    lse = np.full((n, Hq), NEG_INIT, np.float32)
    for b in range(B):
        qb0, qb1 = int(q_indptr[b]), int(q_indptr[b + 1])
        qo_len = qb1 - qb0
        pb, pe = int(page_indptr[b]), int(page_indptr[b + 1])
        kv_len = _paged_kv_len(pe - pb, b, length_info, page_size)
        if qo_len == 0 or kv_len <= 0:
            continue
        k, v = _gather_paged(pages, page_values, pb, kv_len, b, length_info)
        qb = q[qb0:qb1]
        if rotary_mode == 1:
            qb = rope_rotate(qb, np.asarray(q_rope_position)[qb0:qb1], rope_theta, rope_scale, dtype)
            k = rope_rotate(k, int(k_rope_pos_offset[b]) + np.arange(kv_len), rope_theta, rope_scale, dtype)
        if tree_order is not None:
            mask = _tree_mask(tree_indptr, tree_order, b, qo_len, kv_len)
        else:
            mask = _cross_mask(causal, qo_len, kv_len, sliding_window_size)
        ob, lb = _attend(qb, k, v, mask, sm_scale, group)
        o[qb0:qb1] = round_dtype(ob, dtype)
        lse[qb0:qb1] = lb.astype(np.float32)
    return o, lse


def _tree_mask(tree_indptr, tree_order, b, qo_len, kv_len):
    """_check_tree_order (tree_attn.py:48-65)."""
    t0, t1 = int(tree_indptr[b]), int(tree_indptr[b + 1])
    tlen = t1 - t0
    order = np.asarray(tree_order)[t0:t1]
    tree_start = kv_len - tlen
    mask = np.zeros((qo_len, kv_len), bool)
    mask[:, :max(tree_start, 0)] = True
    for r in range(qo_len):
        child = r + tlen - qo_len
        for c in range(max(tree_start, 0), kv_len):
            par = c - tree_start
            mask[r, c] = (order[child, 0] >= order[par, 0]) and (order[child, 0] < order[par, 1])
    return mask


def attention_prefill_ragged(q, q_indptr, k, v, kv_indptr, q_rope_position, k_rope_pos_offset, causal,
                             rotary_mode, rope_scale, rope_theta, sm_scale, dtype,
                             mn_indptr=None, tree_mask=None):
    """_attention_prefill_ragged_cpu (_prefill_kernels.py:677-791); with mn_indptr/tree_mask = tree_attn_cpu
    (tree_attn.py:68-261: K rope position is q_rope_position[kv row])."""
    n, Hq, D = q.shape
    Hkv = k.shape[1]
    group = Hq // Hkv
    B = len(q_indptr) - 1
    o = np.zeros((n, Hq, v.shape[2]), np.float32)
    lse = np.full((n, Hq), NEG_INIT, np.float32)
    for b in range(B):
        qb0, qb1 = int(q_indptr[b]), int(q_indptr[b + 1])
        kb0, kb1 = int(kv_indptr[b]), int(kv_indptr[b + 1])
        qo_len, kv_len = qb1 - qb0, kb1 - kb0
        if qo_len == 0 or kv_len == 0:
            continue
        qb, kb, vb = q[qb0:qb1], k[kb0:kb1], v[kb0:kb1]
        if rotary_mode == 1:
            qb = rope_rotate(qb, np.asarray(q_rope_position)[qb0:qb1], rope_theta, rope_scale, dtype)
            if tree_mask is not None:
                kpos = np.asarray(q_rope_position)[kb0:kb1]
            else:
                kpos = int(k_rope_pos_offset[b]) + np.arange(kv_len)
            kb = rope_rotate(kb, kpos, rope_theta, rope_scale, dtype)
        if tree_mask is not None:
            mask = _tree_mask(mn_indptr, tree_mask, b, qo_len, kv_len)
        else:
            mask = _cross_mask(causal, qo_len, kv_len, 0)
        ob, lb = _attend(qb, kb, vb, mask, sm_scale, group)
        o[qb0:qb1] = round_dtype(ob, dtype)
        lse[qb0:qb1] = lb.astype(np.float32)
    return o, lse


def _kv_transfer_target(local_h: int, remote_h: int, local_tp_rank: int, pe_offset: int, h: int):
    """(remote PE, remote kv head) of local kv head h -- src/runtime/extra/contrib/nvshmem/kv_transfer.cu:54-66."""
    if local_h <= remote_h:  # gather
        assert remote_h % local_h == 0
        gather = remote_h // local_h
        return pe_offset + local_tp_rank // gather, (local_tp_rank % gather) * local_h + h
    assert local_h % remote_h == 0  # scatter
    scatter = local_h // remote_h
    return pe_offset + local_tp_rank * scatter + h // remote_h, h % remote_h


def kv_transfer(remote_pages: list, k: np.ndarray, v: np.ndarray, remote_position_map: np.ndarray,
                remote_tp_group_pe_offset: np.ndarray, local_tp_rank: int = 0) -> None:
    """KVTransfer (kv_transfer.cu:38-83): remote_pages[pe] is PE pe's pool [P, 2, Hkv_remote, page, D], updated in place;
    k / v [ntokens, Hkv_local, D]; position -1 skips the token."""
    local_h = k.shape[1]
    for t, pos in enumerate(np.asarray(remote_position_map)):
        if pos == -1:
            continue
        for h in range(local_h):
            remote_h = next(x for x in remote_pages if x is not None).shape[2]
            pe, rh = _kv_transfer_target(local_h, remote_h, local_tp_rank, int(remote_tp_group_pe_offset[t]), h)
            page = remote_pages[pe].shape[3]
            remote_pages[pe][pos // page, 0, rh, pos % page] = k[t, h]
            remote_pages[pe][pos // page, 1, rh, pos % page] = v[t, h]


def kv_transfer_page_to_page(remote_pages: list, local_pages: np.ndarray, remote_position_map: np.ndarray,
                             local_position_map: np.ndarray, remote_tp_group_pe_offset: np.ndarray,
                             local_tp_rank: int = 0) -> None:
    """KVTransferPageToPage (kv_transfer.cu:84-130): rows already in the local pool [P, 2, Hkv_local, page, D]."""
    local_h, page = local_pages.shape[2], local_pages.shape[3]
    for t, (rpos, lpos) in enumerate(zip(np.asarray(remote_position_map), np.asarray(local_position_map))):
        if rpos == -1 or lpos == -1:
            continue
        for h in range(local_h):
            remote_h = next(x for x in remote_pages if x is not None).shape[2]
            pe, rh = _kv_transfer_target(local_h, remote_h, local_tp_rank, int(remote_tp_group_pe_offset[t]), h)
            for kv in (0, 1):
                remote_pages[pe][rpos // page, kv, rh, rpos % page] = local_pages[lpos // page, kv, h, lpos % page]
