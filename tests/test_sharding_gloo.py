"""World-size-2 gloo test of the multi-GPU partitioning (runs on CPU): each rank computes attention for its KV-head group
with the oracle (the checker standing in for the kernels), the per-head outputs are re-assembled with the same all-gather
helpers bench.py uses over NCCL, and must equal the unsharded result; same for the sequence-batch split."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import kernels as ok
        from tests.util import make_paged_cache, rand16
        from tvm_b200 import sharding
        from tvm_b200.kv_cache import PagedKVCache

        rng = np.random.default_rng(0)  # identical inputs on every rank
        hq, hkv, d, kv_lens = 8, 2, 64, [5, 33, 17, 40]
        B = len(kv_lens)
        c = make_paged_cache(rng, kv_lens, hkv, d, "float16")
        q = rand16(rng, (B, hq, d), "float16")
        zeros = np.zeros(B, np.int32)
        args = (c["page_indptr"], c["page_values"], c["length_info"], zeros, zeros, 0, 1.0, 1e4, d ** -0.5, "float16")
        full_o, _ = ok.attention_decode(q, c["pages"], *args)
        # --- KV-head-group sharding: pages and q sliced by heads, identical page tables ---
        q0, q1, k0, k1 = sharding.head_shard(hq, hkv, world, rank)
        o_loc, _ = ok.attention_decode(q[:, q0:q1], c["pages"][:, :, k0:k1], *args)
        got = sharding.all_gather_heads(torch.from_numpy(o_loc))
        assert np.array_equal(got.numpy(), full_o), "head-sharded reassembly differs"
        # fused qkv slicing keeps (q | k | v) order
        fused = torch.arange(3 * (hq + 2 * hkv) * 2).reshape(3, hq + 2 * hkv, 2)
        sl = sharding.shard_fused_qkv(fused, hq, hkv, world, rank)
        assert sl.shape[1] == (hq + 2 * hkv) // world and torch.equal(sl[:, : (q1 - q0)], fused[:, q0:q1])
        # --- sequence-batch split: every rank plans only its sequences on its own (planning-only) host cache ---
        b0, b1 = sharding.batch_shard(B, world, rank)
        cache = PagedKVCache(reserved_num_seqs=8, total_token_capacity=256, prefill_chunk_size=64, num_layers=1,
                             num_qo_heads=hq, num_kv_heads=hkv, head_dim=d, device=None)
        for sid in range(b0, b1):
            cache.add_sequence(sid)
        cache.begin_forward(list(range(b0, b1)), [kv_lens[s] for s in range(b0, b1)])
        assert cache.get_total_sequence_length() == sum(kv_lens[b0:b1])
        sub = [np.asarray(a)[b0:b1] if np.asarray(a).shape == (B,) else a for a in args[2:5]]
        pi = c["page_indptr"][b0:b1 + 1] - c["page_indptr"][b0]
        pv = c["page_values"][c["page_indptr"][b0]:c["page_indptr"][b1]]
        o_b, _ = ok.attention_decode(q[b0:b1], c["pages"], pi, pv, sub[0], sub[1], sub[2], 0, 1.0, 1e4, d ** -0.5, "float16")
        got_b = sharding.all_gather_batch(torch.from_numpy(o_b))
        assert np.array_equal(got_b.numpy(), full_o), "batch-split reassembly differs"
        ret[rank] = "ok"
    except Exception as e:  # surface the failure in the parent
        ret[rank] = repr(e)
    finally:
        dist.destroy_process_group()


def test_two_rank_head_and_batch_sharding(built_lib):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}, dict(ret)


def test_shard_arithmetic():
    from tvm_b200 import sharding

    assert sharding.head_shard(64, 8, 8, 3) == (24, 32, 3, 4)
    assert sharding.head_shard(32, 8, 2, 1) == (16, 32, 4, 8)
    with pytest.raises(ValueError):
        sharding.head_shard(32, 8, 3, 0)
    assert [sharding.batch_shard(10, 4, r) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
