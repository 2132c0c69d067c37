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
                        dtype=cfg["dtype"], device=device)


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


def replay(name, device, on_forward=None, on_kv=None):
    """Runs the program; compares every callback trace bit-exactly; calls on_forward(idx, op, qkv, outs, golden_o)
    and on_kv(idx, k, v, golden_k, golden_v) with device results when a device is used."""
    meta, z = load(name)
    cfg = meta["config"]
    cache = make_cache(cfg, device)
    cache.set_trace(True)
    L, hq, hkv, d = cfg["num_layers"], cfg["num_qo_heads"], cfg["num_kv_heads"], cfg["head_dim"]
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
        elif k == "forward":
            cache.begin_forward(op["seq_ids"], op["lens"], op["tree"])
            n = sum(op["lens"])
            if device is None:
                for layer in range(L):
                    cache.attention_with_fused_qkv(layer, d ** -0.5, None, None)
            else:
                import torch

                qkv = qkv_for(op["seed"], L, n, hq, hkv, d, cfg["dtype"])
                outs = []
                for layer in range(L):
                    tq = torch.from_numpy(qkv[layer]).cuda()
                    o = torch.full((n, hq, d), float("nan"), dtype=tq.dtype, device="cuda")
                    cache.attention_with_fused_qkv(layer, d ** -0.5, tq, o)
                    outs.append(o)
                torch.cuda.synchronize()
                if on_forward:
                    on_forward(idx, op, qkv, outs, z[f"o_{idx}"])
            cache.end_forward()
            assert cache.get_num_available_pages() == res["num_available_pages"], where
        else:
            raise ValueError(k)
        compare_trace(cache.take_trace(), res["trace"], where)
    return cache
