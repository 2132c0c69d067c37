"""TEST INFRASTRUCTURE ONLY -- live differential fuzzing of the host cache against the REFERENCE (no fixtures stored).

    source /tmp/tvm_ref/env.sh && python oracle/ref_harness/fuzz_host.py [--seeds N] [--first S]

For every seed, the randomised program generators of gen_golden.py (plain: prefill / decode / fork / popn / remove;
tree: 1-3 rounds of random token trees + commits; sliding: random windows + sinks, forks inside the sink; split: shared-KV
queries and self / cross / merge steps mixed into `plain`; layer_sliding: `plain` on caches with an MHA_SLIDING layer) are run on the
reference's own C++ PagedAttentionKVCacheObj (CPU TIR kernels) and, in the same process, on a planning-only tvm_b200 host
cache (ctypes, no GPU): the callback sequence with every int32 array and scalar, the page counts and the query results
must be identical.  With --oracle the captured callback traces are also re-executed by the NumPy oracle and compared with
the reference's attention outputs.  Prints one line per program and a summary; exits non-zero on the first difference."""
from __future__ import annotations

import argparse
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import gen_golden as gg  # noqa: E402

KINDS = {
    # name: (builder(seed), cache kwargs by seed parity)
    "plain": (gg.prog_random, [dict(rope_mode=1), dict(rope_mode=0), dict(rope_mode=2)]),
    "tree": (gg.prog_random_tree, [dict(rope_mode=1), dict(rope_mode=0)]),
    # popn across fork points (the sequence is re-created as a fork of itself)
    "deep_popn": (lambda seed: gg.prog_random(seed, deep_popn=True), [dict(rope_mode=1), dict(rope_mode=2)]),
    # token trees on sequences that fork and disappear between the phases
    "tree_forks": (lambda seed: gg.prog_random_tree(seed, forks=True), [dict(rope_mode=0), dict(rope_mode=1)]),
    "sliding": (gg.prog_random_sliding, [dict(rope_mode=2, support_sliding_window=True),
                                         dict(rope_mode=1, support_sliding_window=True)]),
    # attention_with_shared_kv behind every fused call + self / cross / merge steps (rope none or inline: the step's raw
    # k / v are the "current" k / v)
    "split": (lambda seed: gg.prog_random(seed, shared=True, split=True), [dict(rope_mode=0), dict(rope_mode=2)]),
    # per-layer sliding window (attn_kinds with MHA_SLIDING) under forks, popn and removals
    "layer_sliding": (gg.prog_random, [dict(rope_mode=0, num_layers=2, attn_kinds=[3, 0], layer_sliding_window_size=24),
                                       dict(rope_mode=1, num_layers=2, attn_kinds=[0, 3], layer_sliding_window_size=40)]),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=20)
    ap.add_argument("--first", type=int, default=1000)
    ap.add_argument("--only", default="", help="comma-separated generator names")
    ap.add_argument("--oracle", action="store_true", help="also re-execute the traces with the NumPy oracle")
    a = ap.parse_args()
    from tests.golden_replay import replay_meta

    kernels = {}  # compiled reference kernels by (num_layers, layer window): the other settings are run-time arguments
    n_prog = n_ops = n_calls = 0
    t0 = time.time()
    for seed in range(a.first, a.first + a.seeds):
        for kind, (builder, kws) in KINDS.items():
            if a.only and kind not in a.only.split(","):
                continue
            cfg = dict(gg.BASE)
            cfg.update(kws[seed % len(kws)])
            prog = builder(seed)
            name = f"fuzz_{kind}_{seed}"
            key = (cfg["num_layers"], cfg["layer_sliding_window_size"])
            meta, arrays = gg.capture(name, prog, cfg, kernels=kernels.get(key))
            kernels[key] = gg.capture.last_kernels
            replay_meta(name, meta, arrays, device=None)  # raises on the first differing array / scalar / count
            if a.oracle:
                check_oracle(name, meta, arrays)
            calls = sum(len(r["trace"]) for r in meta["results"])
            n_prog, n_ops, n_calls = n_prog + 1, n_ops + len(meta["ops"]), n_calls + calls
            print(f"{name}: {len(meta['ops'])} ops, {calls} callbacks identical", flush=True)
    print(f"OK: {n_prog} programs, {n_ops} cache operations, {n_calls} callbacks with every int32 array bit-identical "
          f"({time.time() - t0:.0f} s)")


def check_oracle(name, meta, z):
    import numpy as np

    from tests.golden_replay import q2_for, qkv_for
    from tests.test_oracle_golden import OracleMachine, _close

    cfg = meta["config"]
    m = OracleMachine(cfg)
    L, hq, hkv, d = cfg["num_layers"], cfg["num_qo_heads"], cfg["num_kv_heads"], cfg["head_dim"]
    for idx, (op, res) in enumerate(zip(meta["ops"], meta["results"])):
        if op["op"] == "clear":
            m = OracleMachine(cfg)
        elif op["op"] == "forward":
            n = sum(op["lens"])
            qkv = qkv_for(op["seed"], L, n, hq, hkv, d, cfg["dtype"])
            q2 = q2_for(op["seed"], L, n, hq, d, cfg["dtype"]) if op.get("shared") else None
            outs, shared_outs = m.run_forward(res["trace"], qkv, q2)
            for layer in range(L):
                want = z[f"o_{idx}"][layer].astype(np.float32)
                # the reference's CPU prefill returns NaN for a query farther than the layer window from every cached
                # key; those rows pin nothing
                fin = np.isfinite(want).all(axis=(1, 2))
                _close(f"{name} op {idx} layer {layer} O", np.asarray(outs[layer])[fin], want[fin])
                if q2 is not None:
                    want = z[f"os_{idx}"][layer].astype(np.float32)
                    fin = np.isfinite(want).all(axis=(1, 2))
                    _close(f"{name} op {idx} layer {layer} shared O", np.asarray(shared_outs[layer])[fin], want[fin])
        elif op["op"] == "forward_split":
            qkv = qkv_for(op["seed"], L, sum(op["lens"]), hq, hkv, d, cfg["dtype"])
            for layer, r in enumerate(m.run_split(res["trace"], qkv)):
                _close(f"{name} op {idx} layer {layer} merged O", r["o"], z[f"o_{idx}"][layer].astype(np.float32))
        elif op["op"] != "debug_get_kv_rejected":
            m.run_other(res["trace"])


if __name__ == "__main__":
    main()
