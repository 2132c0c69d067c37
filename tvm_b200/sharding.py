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
