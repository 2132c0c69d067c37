"""One-box multi-GPU partitioning of the attention hot path (one process per GPU, torch.distributed for plumbing).

The reference scales this path with Disco tensor parallelism: every worker owns `Hkv/tp` KV heads and `Hq/tp` query
heads, keeps identical page tables, and the per-head outputs are re-assembled with `runtime.disco.allgather` ->
ncclAllGather (src/runtime/extra/disco/nccl/nccl.cc:136-144; sharded attention test tests/python/disco/test_ccl.py:557-700).
KV pages never cross GPUs and no LSE is exchanged (heads, not keys, are split).  For decode the sequence batch can be
split instead (north_star); both are provided here.  NCCL all-gather concatenates on the outer axis, so the head
gather goes through a [tp, n, Hq/tp, D] buffer and a view-permute on the consumer side.
"""
from __future__ import annotations


def head_shard(num_qo_heads: int, num_kv_heads: int, tp: int, rank: int):
    """(q_head_begin, q_head_end, kv_head_begin, kv_head_end) of `rank`; KV-head groups stay intact."""
    if num_kv_heads % tp != 0:
        raise ValueError(f"num_kv_heads {num_kv_heads} is not divisible by tp {tp}: split the sequence batch instead")
    g = num_qo_heads // num_kv_heads
    kv_per = num_kv_heads // tp
    return rank * kv_per * g, (rank + 1) * kv_per * g, rank * kv_per, (rank + 1) * kv_per


def batch_shard(batch_size: int, world: int, rank: int):
    """[begin, end) of the sequences `rank` decodes; the remainder goes to the first ranks."""
    base, rem = divmod(batch_size, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_fused_qkv(qkv, num_qo_heads: int, num_kv_heads: int, tp: int, rank: int):
    """Slice a fused [n, Hq + 2 Hkv, D] tensor into the rank's [n, Hq/tp + 2 Hkv/tp, D] (q | k | v order kept)."""
    import torch

    q0, q1, k0, k1 = head_shard(num_qo_heads, num_kv_heads, tp, rank)
    hq, hkv = num_qo_heads, num_kv_heads
    return torch.cat([qkv[:, q0:q1], qkv[:, hq + k0:hq + k1], qkv[:, hq + hkv + k0:hq + hkv + k1]], dim=1).contiguous()


def all_gather_heads(o_local, group=None):
    """[n, Hq/tp, D] per rank -> [n, Hq, D] on every rank (one all-gather on the current stream)."""
    import torch
    import torch.distributed as dist

    tp = dist.get_world_size(group)
    n, h, d = o_local.shape
    buf = torch.empty((tp * n, h, d), dtype=o_local.dtype, device=o_local.device)
    dist.all_gather_into_tensor(buf, o_local.contiguous(), group=group)
    return buf.view(tp, n, h, d).permute(1, 0, 2, 3).reshape(n, tp * h, d)


def all_gather_batch(o_local, group=None):
    """[B/world, Hq, D] per rank -> [B, Hq, D] (equal shards)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    out = torch.empty((world * o_local.shape[0],) + tuple(o_local.shape[1:]), dtype=o_local.dtype, device=o_local.device)
    dist.all_gather_into_tensor(out, o_local.contiguous(), group=group)
    return out


class PeerHeadGather:
    """Re-assembly of the per-head decode outputs over NVLink peer memory, done BY the attention kernel
    (capi.attention_decode_gather) instead of an NCCL all-gather behind it.  One process per GPU; the gathered
    [n, Hq, D] buffer and a small flag array of every rank are allocated as torch symmetric memory so that each rank
    holds peer-mapped pointers to all of them.  `gathered` is double-buffered by epoch parity so a rank that runs ahead
    into step e+1 never overwrites a buffer a slower rank still reads for step e.  (The reference: Disco allgather ->
    ncclAllGather, src/runtime/extra/disco/nccl/nccl.cc:136-144.)"""

    def __init__(self, n, num_qo_heads_total, head_dim, dtype, device, group=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        group = group or dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.bufs, self.buf_ptrs = [], []
        for _ in range(2):
            t = symm_mem.empty((n, num_qo_heads_total, head_dim), dtype=dtype, device=device)
            h = symm_mem.rendezvous(t, group)
            self.bufs.append(t)
            self.buf_ptrs.append(list(h.buffer_ptrs))
        self.flags = symm_mem.empty((64,), dtype=torch.int32, device=device)
        self.flags.zero_()
        hf = symm_mem.rendezvous(self.flags, group)
        self.flag_ptrs = list(hf.buffer_ptrs)
        self.epoch = 0
        torch.cuda.synchronize(device)
        dist.barrier(group)

    def decode(self, capi, q, pages, page_indptr, page_values, length_info, k_rope_pos_offset, q_rope_position, output,
               lse, rotary_mode, rope_scale, rope_theta, sm_scale):
        """-> the gathered [n, Hq_total, D] tensor of this step (valid once the returned stream work has run)"""
        self.epoch += 1
        par = self.epoch & 1
        capi.attention_decode_gather(q, pages, page_indptr, page_values, length_info, k_rope_pos_offset,
                                     q_rope_position, output, lse, rotary_mode, rope_scale, rope_theta, sm_scale,
                                     self.buf_ptrs[par], self.flag_ptrs, self.rank, self.epoch)
        return self.bufs[par]  # the launch itself waits for every peer's epoch (include/tvm_b200.h)

    def decode_fused_qkv(self, capi, qkv, q_rope_position, append_position, pages, page_indptr, page_values, length_info,
                         k_rope_pos_offset, output, lse, apply_rope, rope_scale, rope_theta, sm_scale):
        """the whole decode step of this rank's shard (rotary + append + decode) + the peer gather: two launches"""
        self.epoch += 1
        par = self.epoch & 1
        capi.attention_decode_fused_qkv_gather(qkv, q_rope_position, append_position, pages, page_indptr, page_values,
                                               length_info, k_rope_pos_offset, output, lse, apply_rope, rope_scale,
                                               rope_theta, sm_scale, self.buf_ptrs[par], self.flag_ptrs, self.rank,
                                               self.epoch)
        return self.bufs[par]
