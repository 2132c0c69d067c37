"""Replays a golden scenario program (tests/golden/kvcache_*.npz, captured from the reference by
oracle/ref_harness/gen_golden.py) on our host cache."""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"


def load(name):
    z = np.load(GOLDEN / f"kvcache_{name}.npz")
    meta = json.loads(bytes(z["meta"]).decode())
    return meta, z


def scenario_names():
    return sorted(p.stem[len("kvcache_"):] for p in GOLDEN.glob("kvcache_*.npz"))


# the reference's own scenario tests (test_runtime_builtin_paged_attention_kv_cache_cpu.py:706-1104); the other fixtures
# (randomised programs, per-layer sliding window, shared-KV, self / cross / merge entries) are replayed on the GPU by
# tests/test_zz_scenarios_gpu.py
BASE_SCENARIOS = ("fork", "prefill_and_decode", "prefill_and_decode_inline_rope_2layers", "remove_and_popn", "sliding_window",
                  "sliding_window_fork", "tree_attn", "unlimited_depth")


def base_scenario_names():
    return [n for n in scenario_names() if n in BASE_SCENARIOS]


# fixtures added late in the round: replayed on the GPU by the last test file (tests/test_zzz_*)
LATE_SCENARIOS = ("layer_offset",)


def extra_scenario_names():
    return [n for n in scenario_names() if n not in BASE_SCENARIOS and n not in LATE_SCENARIOS]


def late_scenario_names():
    return [n for n in scenario_names() if n in LATE_SCENARIOS]


def qkv_for(seed, num_layers, n, hq, hkv, d, dtype="float16"):
    """Same generator as oracle/ref_harness/gen_golden.py::qkv_for (inputs are not stored, only their seeds)."""
    rng = np.random.default_rng(seed)
    return rng.random((num_layers, n, hq + 2 * hkv, d), dtype=np.float32).astype(dtype)


def make_cache(cfg, device):
    from tvm_b200.kv_cache import PagedKVCache

    return PagedKVCache(reserved_num_seqs=cfg["reserved_nseq"], total_token_capacity=cfg["max_total_seq"],
                        prefill_chunk_size=cfg["prefill_chunk"], page_size=cfg["page_size"],
                        support_sliding_window=bool(cfg.get("support_sliding_window", False)), num_layers=cfg["num_layers"],
                        num_qo_heads=cfg["num_qo_heads"], num_kv_heads=cfg["num_kv_heads"], head_dim=cfg["head_dim"],
                        rope_mode=cfg["rope_mode"], rotary_scale=cfg["rope_scale"], rotary_theta=cfg["rope_theta"],
                        dtype=cfg["dtype"], device=device, attn_kinds=cfg.get("attn_kinds"),
                        layer_sliding_window_size=cfg.get("layer_sliding_window_size"),
                        layer_id_begin_offset=cfg.get("layer_begin", 0))


def q2_for(seed, num_layers, n, hq, d, dtype="float16"):
    """Same generator as oracle/ref_harness/gen_golden.py::q2_for (the shared-KV query of a step)."""
    rng = np.random.default_rng(seed + 500000)
    return rng.random((num_layers, n, hq, d), dtype=np.float32).astype(dtype)


def compare_trace(got, want, where):
    assert [g["fn"] for g in got] == [w["fn"] for w in want], f"{where}: callback sequence differs:\n got  {[g['fn'] for g in got]}\n want {[w['fn'] for w in want]}"
    for ci, (g, w) in enumerate(zip(got, want)):
        assert len(g["args"]) == len(w["args"]), f"{where} call {ci} {g['fn']}: arg count"
        for ai, (ga, wa) in enumerate(zip(g["args"], w["args"])):
            at = f"{where} call {ci} {g['fn']} arg {ai}"
            if "s" in wa:
                assert "s" in ga and float(ga["s"]) == float(wa["s"]), f"{at}: scalar {ga} != {wa}"
            else:
                assert ga["shape"] == wa["shape"], f"{at}: shape {ga['shape']} != {wa['shape']}"
                if wa["v"] is not None:
                    assert ga["v"] == wa["v"], f"{at}: int32 array differs\n got  {ga['v']}\n want {wa['v']}"


def replay(name, device, on_forward=None, on_kv=None, on_shared=None, on_split=None):
    """Runs the program; compares every callback trace bit-exactly; calls on_forward(idx, op, qkv, outs, golden_o),
    on_shared(idx, outs, golden_os), on_split(idx, dict of device results, npz) and on_kv(idx, k, v, golden_k, golden_v)
    with device results when a device is used."""
    meta, z = load(name)
    return replay_meta(name, meta, z, device, on_forward, on_kv, on_shared, on_split)


def replay_meta(name, meta, z, device, on_forward=None, on_kv=None, on_shared=None, on_split=None):
    """replay() on an already loaded (or freshly captured, oracle/ref_harness/fuzz_host.py) program."""
    cfg = meta["config"]
    cache = make_cache(cfg, device)
    cache.set_trace(True)
    L, hq, hkv, d = cfg["num_layers"], cfg["num_qo_heads"], cfg["num_kv_heads"], cfg["head_dim"]
    lb = cfg.get("layer_begin", 0)  # attention calls carry global layer ids
    for idx, (op, res) in enumerate(zip(meta["ops"], meta["results"])):
        k = op["op"]
        where = f"{name} op {idx} {op}"
        if k == "clear":
            cache.clear()
        elif k == "add":
            cache.add_sequence(op["seq"])
        elif k == "remove":
            cache.remove_sequence(op["seq"])
        elif k == "fork":
            cache.fork_sequence(op["parent"], op["child"], op["pos"])
        elif k == "popn":
            cache.popn(op["seq"], op["n"])
        elif k == "enable_sw":
            cache.enable_sliding_window_for_seq(op["seq"], op["window"], op["sink"])
        elif k == "commit":
            cache.commit_accepted_token_tree_nodes(op["seq_ids"], op["leaves"])
        elif k == "query":
            assert cache.empty() == res["empty"], where
            assert cache.get_num_available_pages() == res["num_available_pages"], where
            assert cache.get_total_sequence_length() == res["total_sequence_length"], where
        elif k == "debug_get_kv":
            if device is None:
                cache.debug_get_kv(op["seq"], op["start"], op["end"])
            else:
                import torch

                n = op["end"] - op["start"]
                tdt = torch.float16 if cfg["dtype"] == "float16" else torch.bfloat16
                kk = torch.zeros((L, n, hkv, d), dtype=tdt, device="cuda")
                vv = torch.zeros_like(kk)
                cache.debug_get_kv(op["seq"], op["start"], op["end"], kk, vv)
                torch.cuda.synchronize()
                if on_kv:
                    on_kv(idx, kk, vv, z[f"k_{idx}"], z[f"v_{idx}"])
        elif k == "debug_get_kv_rejected":
            import pytest

            from tvm_b200 import capi

            kk = None
            if device is not None:
                import torch

                kk = torch.zeros((L, 1, hkv, d), dtype=torch.float16 if cfg["dtype"] == "float16" else torch.bfloat16,
                                 device="cuda")
            with pytest.raises(capi.TvmB200Error, match="Only MHA is supported for DebugGetKV"):
                cache.debug_get_kv(op["seq"], 0, 1, kk, kk)
            cache.take_trace()  # the reference dumps the MHA layers in front of the offending one before it raises
            continue
        elif k == "forward":
            cache.begin_forward(op["seq_ids"], op["lens"], op["tree"])
            n = sum(op["lens"])
            shared = bool(op.get("shared"))
            if device is None:
                for layer in range(L):
                    cache.attention_with_fused_qkv(lb + layer, d ** -0.5, None, None)
                    if shared:
                        cache.attention_with_shared_kv(lb + layer, d ** -0.5, n, None, None, None)
            else:
                import torch

                qkv = qkv_for(op["seed"], L, n, hq, hkv, d, cfg["dtype"])
                q2 = q2_for(op["seed"], L, n, hq, d, cfg["dtype"]) if shared else None
                outs, shared_outs = [], []
                for layer in range(L):
                    tq = torch.from_numpy(qkv[layer]).cuda()
                    o = torch.full((n, hq, d), float("nan"), dtype=tq.dtype, device="cuda")
                    cache.attention_with_fused_qkv(lb + layer, d ** -0.5, tq, o)
                    outs.append(o)
                    if shared:
                        o2 = torch.full((n, hq, d), float("nan"), dtype=tq.dtype, device="cuda")
                        cache.attention_with_shared_kv(lb + layer, d ** -0.5, torch.from_numpy(q2[layer]).cuda(),
                                                       tq[:, hq:hq + hkv].contiguous(), tq[:, hq + hkv:].contiguous(), o2)
                        shared_outs.append(o2)
                torch.cuda.synchronize()
                if on_forward:
                    on_forward(idx, op, qkv, outs, z[f"o_{idx}"])
                if shared and on_shared:
                    on_shared(idx, shared_outs, z[f"os_{idx}"])
            cache.end_forward()
            assert cache.get_num_available_pages() == res["num_available_pages"], where
        elif k == "forward_split":
            cache.begin_forward(op["seq_ids"], op["lens"], None)
            n = sum(op["lens"])
            if device is None:
                for layer in range(L):
                    cache.self_attention(lb + layer, d ** -0.5, n, None, None, None, None)
                    cache.cross_attention(lb + layer, d ** -0.5, n, None, None)
                    cache.merge_attn_output_inplace(n, None, None, None)
            else:
                import torch

                qkv = qkv_for(op["seed"], L, n, hq, hkv, d, cfg["dtype"])
                got = {"o": [], "lse": [], "oself": [], "ocross": []}
                for layer in range(L):
                    tq = torch.from_numpy(qkv[layer]).cuda()
                    q, kk, vv = tq[:, :hq].contiguous(), tq[:, hq:hq + hkv].contiguous(), tq[:, hq + hkv:].contiguous()
                    o_self = torch.zeros((n, hq, d), dtype=tq.dtype, device="cuda")
                    lse_self = torch.full((n, hq), -5e4, dtype=torch.float32, device="cuda")
                    o_cross, lse_cross = torch.zeros_like(o_self), torch.full_like(lse_self, -5e4)
                    cache.self_attention(lb + layer, d ** -0.5, q, kk, vv, o_self, lse_self)
                    cache.cross_attention(lb + layer, d ** -0.5, q, o_cross, lse_cross)
                    got["oself"].append(o_self.clone())
                    got["ocross"].append(o_cross.clone())
                    ro, rl = cache.merge_attn_output_inplace(o_self, lse_self, o_cross, lse_cross)
                    assert ro is o_self and rl is lse_self
                    got["o"].append(o_self)
                    got["lse"].append(lse_self)
                torch.cuda.synchronize()
                if on_split:
                    on_split(idx, got, z)
            cache.end_forward()
            assert cache.get_num_available_pages() == res["num_available_pages"], where
        else:
            raise ValueError(k)
        compare_trace(cache.take_trace(), res["trace"], where)
    return cache
