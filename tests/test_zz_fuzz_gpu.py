"""Randomised programs on the GPU: the generators of oracle/ref_harness/gen_golden.py (plain prefill / decode / fork / popn /
remove; token trees with commits; sliding windows with sinks) drive the C++ host cache with the sm_100a kernels, and the
NumPy oracle re-executes the callbacks the cache recorded (same int32 arrays, same inputs): every attention output and
every debug_get_kv dump must agree.  The plans themselves are pinned against the reference by the CPU tests
(tests/test_host_cache_golden.py, oracle/ref_harness/fuzz_host.py); this closes the loop kernels <-> oracle off the
hand-written paths."""
import sys
from pathlib import Path

import numpy as np
import pytest

from tests.golden_replay import make_cache, qkv_for
from tests.test_oracle_golden import OracleMachine
from tests.util import assert_close, to_np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "oracle" / "ref_harness"))
import gen_golden as gg  # noqa: E402  (program generators only; nothing of the reference is imported)

CASES = [
    ("plain", gg.prog_random, dict(rope_mode=1)),
    ("plain_inline_rope", gg.prog_random, dict(rope_mode=2, num_layers=2)),
    ("tree", gg.prog_random_tree, dict(rope_mode=1)),
    ("sliding", gg.prog_random_sliding, dict(rope_mode=2, support_sliding_window=True)),
    ("deep_popn", lambda s: gg.prog_random(s, deep_popn=True), dict(rope_mode=1)),
    ("tree_forks", lambda s: gg.prog_random_tree(s, forks=True), dict(rope_mode=0)),
]
GPU_KINDS = ["plain", "plain_inline_rope", "tree", "sliding"]   # the ones that have run on a B200


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [7001])
@pytest.mark.parametrize("kind", GPU_KINDS)
def test_random_program_gpu_vs_oracle(built_lib, kind, seed):
    _run(kind, seed, 0)


@pytest.mark.parametrize("seed", [7001, 7002, 7003])
@pytest.mark.parametrize("kind", [c[0] for c in CASES])
def test_random_program_plan_is_executable_by_the_oracle(built_lib, kind, seed):
    """No GPU: the same loop on a planning-only cache -- the oracle must be able to execute every recorded callback and
    none of them may read a slot nobody appended to (the oracle's pages start as NaN)."""
    _run(kind, seed, None)


def _run(kind, seed, device):
    if device is not None:
        import torch

    builder, kw = next((b, k) for n, b, k in CASES if n == kind)
    cfg = dict(gg.BASE)
    cfg.update(kw)
    prog = builder(seed)
    cache = make_cache(cfg, device)
    cache.set_trace(True)
    m = OracleMachine(cfg)
    L, hq, hkv, d = cfg["num_layers"], cfg["num_qo_heads"], cfg["num_kv_heads"], cfg["head_dim"]
    n_fwd = n_dump = 0
    for idx, op in enumerate(prog.ops):
        k = op["op"]
        if k == "add":
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
            cache.get_num_available_pages()
        elif k == "debug_get_kv" and device is None:
            cache.debug_get_kv(op["seq"], op["start"], op["end"])
            for ok_k, ok_v in m.run_other(cache.take_trace()):
                assert np.isfinite(ok_k).all() and np.isfinite(ok_v).all(), f"{kind} {seed} op {idx}: dump of an unwritten slot"
            n_dump += 1
            continue
        elif k == "debug_get_kv":
            n = op["end"] - op["start"]
            kk = torch.zeros((L, n, hkv, d), dtype=torch.float16, device="cuda")
            vv = torch.zeros_like(kk)
            cache.debug_get_kv(op["seq"], op["start"], op["end"], kk, vv)
            torch.cuda.synchronize()
            dumps = m.run_other(cache.take_trace())
            for layer, (ok_k, ok_v) in enumerate(dumps):
                assert np.array_equal(to_np(vv)[layer], ok_v), f"{kind} {seed} op {idx}: V dump differs"
                if cfg["rope_mode"] == 1:   # K is cached rotated: the rotation's rounding may differ by an ulp
                    assert_close(f"{kind} {seed} op {idx} K", to_np(kk)[layer], ok_k, atol=2e-3, rtol=2e-3)
                else:
                    assert np.array_equal(to_np(kk)[layer], ok_k), f"{kind} {seed} op {idx}: K dump differs"
            n_dump += 1
            continue
        elif k == "forward":
            cache.begin_forward(op["seq_ids"], op["lens"], op["tree"])
            n = sum(op["lens"])
            qkv = qkv_for(op["seed"], L, n, hq, hkv, d, cfg["dtype"])
            outs = []
            for layer in range(L):
                if device is None:
                    cache.attention_with_fused_qkv(layer, d ** -0.5, None, None)
                    continue
                o = torch.full((n, hq, d), float("nan"), dtype=torch.float16, device="cuda")
                cache.attention_with_fused_qkv(layer, d ** -0.5, torch.from_numpy(qkv[layer]).cuda(), o)
                outs.append(o)
            cache.end_forward()
            if device is not None:
                torch.cuda.synchronize()
            want, _ = m.run_forward(cache.take_trace(), qkv, None)
            n_fwd += 1
            if device is None:
                for layer in range(L):
                    assert np.isfinite(np.asarray(want[layer])).all(), f"{kind} {seed} op {idx}: read of an unwritten slot"
                continue
            for layer in range(L):
                assert_close(f"{kind} {seed} op {idx} {op['seq_ids']} x {op['lens']} layer {layer} O", to_np(outs[layer]),
                             np.asarray(want[layer]))
            continue
        else:
            raise ValueError(k)
        m.run_other(cache.take_trace())   # copy_single_page / compact_copy issued by fork / commit
    assert n_fwd > 10 and n_dump > 0
