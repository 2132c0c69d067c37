#!/usr/bin/env python
"""GPU box: randomised kernel-level parity against the oracle (one-off confidence run, logged under profiles/).
decode: random batch sizes, sequence lengths (zeros, one-slot, page-aligned, split-KV sizes), head shapes incl. GQA groups of
16 / 32, sliding windows, inline RoPE, the fused step -- the balanced split-KV partition sees a different cut pattern each
time; prefill (tcgen05 forced): random ragged / paged batches, causal / none / tree / layer-window masks, inline RoPE and
sliding slots through the pre-pass, KV split.   usage: fuzz_kernels_gpu.py [cases] [seed]"""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from tests import test_kernels_gpu as tk  # noqa: E402
from tests.test_prefill_tc05_masks_gpu import _random_tree, _tree_arrays  # noqa: E402
from tvm_b200 import capi  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 150
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 9000
capi.lib()
t0 = time.time()
stats = {"decode": 0, "prefill_ragged": 0, "prefill_paged": 0}
SHAPES = [(32, 8), (8, 8), (32, 4), (8, 1), (16, 1), (64, 2), (16, 2), (4, 4)]
for case in range(N):
    rng = np.random.default_rng(seed0 + case)
    dtype = ["float16", "bfloat16"][case & 1]
    hq, hkv = SHAPES[int(rng.integers(0, len(SHAPES)))]

    def length():
        r = rng.random()
        if r < 0.1:
            return 0
        if r < 0.5:
            return int(rng.integers(1, 80))
        if r < 0.9:
            return int(rng.integers(80, 1500))
        return int(rng.integers(1500, 6000))

    kind = case % 3
    if kind == 0:
        B = int(rng.integers(1, 40))
        lens = [length() for _ in range(B)]
        mode = int(rng.integers(0, 3))
        if mode == 2:
            lens = [max(l, 40) for l in lens]
            slid = [(int(rng.integers(0, l // 2)) // 1, int(rng.integers(0, 8))) if rng.random() < 0.6 else (0, 0) for l in lens]
            slid = [(o, min(s, o)) for o, s in slid]
            tk._run_decode(capi, rng, lens, hq, hkv, 128, dtype, rotary_mode=1, sliding=slid)
        else:
            tk._run_decode(capi, rng, lens, hq, hkv, 128, dtype, rotary_mode=mode)
        stats["decode"] += 1
    else:
        if hq // hkv not in (1, 2, 4, 8, 16):
            hq, hkv = 32, 8
        capi.set_prefill_impl(2)
        try:
            B = int(rng.integers(1, 6))
            if kind == 1:
                q_lens = [max(1, length() // 4) for _ in range(B)]
                tree = rng.random() < 0.35
                if tree:
                    q_lens = [min(max(q, 2), 200) for q in q_lens]
                    trees = [_random_tree(rng, q) for q in q_lens]
                    tk._run_ragged(capi, rng, q_lens, q_lens, hq, hkv, 128, dtype, rotary_mode=int(rng.integers(0, 2)),
                                   tree=_tree_arrays(trees))
                else:
                    kv_lens = [q + int(rng.integers(0, 300)) for q in q_lens]
                    tk._run_ragged(capi, rng, q_lens, kv_lens, hq, hkv, 128, dtype, causal=int(rng.integers(0, 2)),
                                   rotary_mode=int(rng.integers(0, 2)))
                stats["prefill_ragged"] += 1
            else:
                q_lens = [int(rng.integers(1, 150)) for _ in range(B)]
                kv_lens = [q + length() for q in q_lens]
                if rng.random() < 0.3:  # few long contexts: the KV split
                    B = int(rng.integers(1, 3))
                    q_lens = q_lens[:B]
                    kv_lens = [q + int(rng.integers(2500, 9000)) for q in q_lens]
                r = rng.random()
                if r < 0.3:
                    trees = [_random_tree(rng, q) for q in q_lens]
                    tk._run_paged_prefill(capi, rng, q_lens, kv_lens, hq, hkv, 128, dtype, tree=_tree_arrays(trees))
                elif r < 0.55:
                    slid = [(int(rng.integers(0, max(1, (l - 1) // 2))), int(rng.integers(0, 6))) for l in kv_lens]
                    slid = [(o, min(s, o)) for o, s in slid]
                    causal = int(rng.integers(0, 2))
                    tk._run_paged_prefill(capi, rng, q_lens, kv_lens, hq, hkv, 128, dtype, causal=causal, rotary_mode=1,
                                          sliding=slid, layer_sws=int(rng.integers(2, 300)) if causal else 0)
                else:
                    tk._run_paged_prefill(capi, rng, q_lens, kv_lens, hq, hkv, 128, dtype, causal=int(rng.integers(0, 2)),
                                          rotary_mode=int(rng.integers(0, 2)), nan_tail=rng.random() < 0.5)
                stats["prefill_paged"] += 1
        finally:
            capi.set_prefill_impl(0)
    if (case + 1) % 25 == 0:
        print(f"{case + 1} cases ok ({time.time() - t0:.0f} s) {stats} paths (generic, tcgen05, pre-pass, split) {capi.prefill_path_counts()}", flush=True)
print(f"ALL {N} RANDOM CASES WITHIN TOLERANCE (seeds {seed0}..{seed0 + N - 1}): {stats}; prefill paths (generic, tcgen05, "
      f"pre-pass, split) = {capi.prefill_path_counts()}")
