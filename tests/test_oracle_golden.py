"""Pins the CPU oracle (oracle/kernels.py) against the REAL reference: the callback traces captured from the reference
cache (tests/golden/kvcache_*.npz) are re-executed with the oracle kernels on the reference's exact int32 arguments, and
the attention outputs / cache dumps must match what the reference's own CPU TIR kernels produced (fp16).
Tolerance: the reference's own test tolerance for its kernels vs. its NumPy oracle, rtol = atol = 1e-3
(test_runtime_builtin_paged_attention_kv_cache_cpu.py:530-535) plus one fp16 ulp of output rounding."""
import numpy as np
import pytest

from oracle import kernels as ok
from tests.golden_replay import load, q2_for, qkv_for, scenario_names
from tests.util import assert_close


def _arr(a):
    return np.array(a["v"], np.int32).reshape(a["shape"])


def _close(name, got, want, atol=2e-3, rtol=2e-3):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    err = np.abs(got - want)
    assert (err <= atol + rtol * np.abs(want)).all(), f"{name}: max abs err {err.max():.3e}"


class OracleMachine:
    """Executes traced callbacks with the oracle kernels."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.dt = cfg["dtype"]
        L, hkv, d, ps = cfg["num_layers"], cfg["num_kv_heads"], cfg["head_dim"], cfg["page_size"]
        npages = (cfg["max_total_seq"] + ps - 1) // ps + 1 + (2 * cfg["reserved_nseq"] if cfg.get("support_sliding_window") else 0)
        # NaN, not zero: a plan that reads a slot nobody appended to (as the reference's own plans do for sliding-window
        # sequences spread over more than two blocks) must show up instead of comparing equal by luck
        self.pages = [np.full((npages, 2, hkv, ps, d), np.nan, np.float32) for _ in range(L)]
        self.theta, self.scale = cfg["rope_theta"], cfg["rope_scale"]

    def attend(self, calls, layer, q, k, v):
        """Executes a run of attention / merge callbacks (one AttentionInternal, SelfAttention or CrossAttention of the
        reference) on q and the step's k / v; returns (o, lse), both None when nothing was computed."""
        o = lse = tmp = None
        sw = self.cfg.get("layer_sliding_window_size") or 1024
        for c in calls:
            fn, a = c["fn"], c["args"]
            if fn == "prefill_ragged":
                o, lse = ok.attention_prefill_ragged(q, _arr(a[1]), k, v, _arr(a[4]), _arr(a[5]), _arr(a[6]), a[9]["s"],
                                                     a[10]["s"], a[11]["s"], a[12]["s"], a[13]["s"], self.dt)
            elif fn == "tree_ragged":
                o, lse = ok.attention_prefill_ragged(q, _arr(a[1]), k, v, _arr(a[4]), _arr(a[5]), None, 0, a[10]["s"],
                                                     a[11]["s"], a[12]["s"], a[13]["s"], self.dt, mn_indptr=_arr(a[6]),
                                                     tree_mask=_arr(a[7]))
            elif fn in ("prefill", "prefill_sliding_window", "decode", "decode_sliding_window", "tree_paged"):
                P = self.pages[layer]
                if fn.startswith("decode"):
                    r = ok.attention_decode(q, P, _arr(a[2]), _arr(a[3]), _arr(a[4]), _arr(a[5]), _arr(a[6]), a[9]["s"],
                                            a[10]["s"], a[11]["s"], a[12]["s"], self.dt)
                elif fn == "tree_paged":
                    r = ok.attention_prefill_paged(q, _arr(a[1]), P, _arr(a[3]), _arr(a[4]), _arr(a[5]), _arr(a[6]), _arr(a[7]),
                                                   0, a[10]["s"], a[11]["s"], a[12]["s"], a[13]["s"], self.dt,
                                                   tree_indptr=_arr(a[14]), tree_order=_arr(a[15]))
                else:
                    r = ok.attention_prefill_paged(q, _arr(a[1]), P, _arr(a[3]), _arr(a[4]), _arr(a[5]), _arr(a[6]), _arr(a[7]),
                                                   a[10]["s"], a[11]["s"], a[12]["s"], a[13]["s"], a[14]["s"], self.dt,
                                                   sliding_window_size=sw if fn.endswith("sliding_window") else 0)
                if o is None:
                    o, lse = r
                else:
                    tmp = r
            elif fn == "merge":
                o, lse = ok.merge_state_inplace(o, lse, tmp[0], tmp[1], self.dt)
            else:
                raise AssertionError(f"unexpected callback {fn} inside an attention")
        return o, lse

    def run_forward(self, trace, qkv, q2=None):
        """attention_with_fused_qkv per layer (split_rotary .. transpose_append), each optionally followed by the
        attention_with_shared_kv calls of the same layer on q2.  Returns (outs, shared_outs)."""
        hq, hkv = self.cfg["num_qo_heads"], self.cfg["num_kv_heads"]
        starts = [i for i, c in enumerate(trace) if c["fn"] == "split_rotary"] + [len(trace)]
        outs, shared_outs = [], []
        for layer in range(len(starts) - 1):
            calls = trace[starts[layer]:starts[layer + 1]]
            q, k, v = ok.split_rotary(qkv[layer].astype(np.float32), _arr(calls[0]["args"][1]), hq, hkv,
                                      calls[0]["args"][5]["s"], self.theta, self.scale, self.dt)
            ia = next(i for i, c in enumerate(calls) if c["fn"] == "transpose_append")
            if ia == 1:  # append before the attention: what follows is the fused attention (+ the same plan again on q2)
                ok.transpose_append(self.pages[layer], k, v, _arr(calls[ia]["args"][3]))
                rest = calls[2:]
                m = len(rest) // 2 if q2 is not None else len(rest)
                fused, shared = rest[:m], rest[m:]
                if q2 is not None:
                    assert [c["fn"] for c in fused] == [c["fn"] for c in shared]
                outs.append(self.attend(fused, layer, q, k, v)[0])
            else:        # attention first (self + cross), then the append, then the shared-KV query
                outs.append(self.attend(calls[1:ia], layer, q, k, v)[0])
                shared = calls[ia + 1:]
            if q2 is not None:
                # the step's raw k / v are the "current" k / v (rope is none or inline in the shared-KV scenarios)
                shared_outs.append(self.attend(shared, layer, q2[layer].astype(np.float32), k, v)[0])
            if ia != 1:
                ok.transpose_append(self.pages[layer], k, v, _arr(calls[ia]["args"][3]))
        return outs, shared_outs

    def run_split(self, trace, qkv):
        """self_attention, cross_attention, merge_attn_output_inplace per layer on the raw q, k, v of the step."""
        hq, hkv = self.cfg["num_qo_heads"], self.cfg["num_kv_heads"]
        L = self.cfg["num_layers"]
        self_fns = ("prefill_ragged", "tree_ragged")
        starts = [i for i, c in enumerate(trace) if c["fn"] in self_fns] + [len(trace)]
        assert len(starts) == L + 1
        res = []
        for layer in range(L):
            calls = trace[starts[layer]:starts[layer + 1]]
            assert calls[-1]["fn"] == "merge"
            x = qkv[layer].astype(np.float32)
            q, k, v = x[:, :hq], x[:, hq:hq + hkv], x[:, hq + hkv:]
            o_self, lse_self = self.attend(calls[:1], layer, q, k, v)
            o_cross, lse_cross = self.attend(calls[1:-1], layer, q, k, v)
            if o_cross is None:  # no cached page in the batch: the harness' initial values stay
                o_cross, lse_cross = np.zeros_like(o_self), np.full_like(lse_self, -5e4)
            o, lse = ok.merge_state_inplace(o_self.copy(), lse_self.copy(), o_cross, lse_cross, self.dt)
            res.append(dict(o=o, lse=lse, oself=o_self, ocross=o_cross))
        return res

    def run_other(self, trace):
        dumps, layer = [], 0
        for c in trace:
            fn, a = c["fn"], c["args"]
            if fn == "copy_single_page":
                ok.copy_single_page(self.pages[layer % len(self.pages)], a[1]["s"], a[2]["s"], a[3]["s"])
                layer += 1
            elif fn == "compact_copy":
                ok.compact_kv_copy(self.pages[layer % len(self.pages)], _arr(a[1]), _arr(a[2]), a[3]["s"])
                layer += 1
            elif fn == "debug_get_kv":
                dumps.append(ok.debug_get_kv(self.pages[a[4]["s"]], _arr(a[1])))
            else:
                raise AssertionError(f"unexpected callback {fn}")
        return dumps


@pytest.mark.parametrize("name", scenario_names())
def test_oracle_matches_reference_outputs(name):
    meta, z = load(name)
    cfg = meta["config"]
    m = OracleMachine(cfg)
    L, hq, hkv, d = cfg["num_layers"], cfg["num_qo_heads"], cfg["num_kv_heads"], cfg["head_dim"]
    checked = 0
    for idx, (op, res) in enumerate(zip(meta["ops"], meta["results"])):
        if op["op"] == "clear":
            m = OracleMachine(cfg)
        elif op["op"] == "forward":
            n = sum(op["lens"])
            qkv = qkv_for(op["seed"], L, n, hq, hkv, d, cfg["dtype"])
            q2 = q2_for(op["seed"], L, n, hq, d, cfg["dtype"]) if op.get("shared") else None
            outs, shared_outs = m.run_forward(res["trace"], qkv, q2)
            want = z[f"o_{idx}"].astype(np.float32)
            for layer in range(L):
                _close(f"{name} op {idx} layer {layer} O", outs[layer], want[layer])
                if q2 is not None:
                    _close(f"{name} op {idx} layer {layer} shared-KV O", shared_outs[layer], z[f"os_{idx}"][layer].astype(np.float32))
            checked += 1
        elif op["op"] == "forward_split":
            n = sum(op["lens"])
            qkv = qkv_for(op["seed"], L, n, hq, hkv, d, cfg["dtype"])
            for layer, r in enumerate(m.run_split(res["trace"], qkv)):
                for key in ("o", "oself", "ocross"):
                    _close(f"{name} op {idx} layer {layer} {key}", r[key], z[f"{key}_{idx}"][layer].astype(np.float32))
                _close(f"{name} op {idx} layer {layer} lse", r["lse"], z[f"lse_{idx}"][layer], atol=2e-3, rtol=1e-3)
            checked += 1
        elif op["op"] == "debug_get_kv_rejected":
            pass
        else:
            dumps = m.run_other(res["trace"])
            if op["op"] == "debug_get_kv":
                for layer, (kk, vv) in enumerate(dumps):
                    assert np.array_equal(vv, z[f"v_{idx}"][layer].astype(np.float32)), f"{name} op {idx}: V dump differs"
                    if cfg["rope_mode"] == 1:  # K was rotated by fused_rope in fp32 before the fp16 cast
                        _close(f"{name} op {idx} K", kk, z[f"k_{idx}"][layer].astype(np.float32), atol=1e-3, rtol=2e-3)
                    else:
                        assert np.array_equal(kk, z[f"k_{idx}"][layer].astype(np.float32)), f"{name} op {idx}: K dump differs"
    assert checked > 0


def test_oracle_llama3_rope_scaling_matches_the_reference():
    """tests/golden/rope_llama3.npz was produced by the reference's own fused_rope and _attention_decode_cpu (inline
    RoPE) built with rope_scaling = llama3 (oracle/ref_harness/gen_golden_rope.py)."""
    from pathlib import Path

    g = np.load(Path(__file__).parent / "golden" / "rope_llama3.npz")
    theta, scale, factor, low, high, orig = [float(x) for x in g["params"]]
    ok.set_rope_scaling({"rope_type": "llama3", "factor": factor, "low_freq_factor": low, "high_freq_factor": high,
                         "original_max_position_embeddings": orig})
    try:
        q, k, v = ok.split_rotary(g["qkv"].astype(np.float32), g["pos"], 8, 2, 1, theta, scale, "float16")
        assert np.array_equal(v, g["v"].astype(np.float32))
        # the angle pos * inv_freq is a float32: its ulp is 6e-8 * pos for the high-frequency dims (0.006 rad at
        # pos = 1e5), and powf differs by an ulp between libm and NumPy -- the slack grows with the position
        for i, pos in enumerate(g["pos"]):
            atol = 4e-3 + 3e-7 * float(pos)
            assert_close(f"q[{i}]", q[i], g["q"][i].astype(np.float32), atol=atol)
            assert_close(f"k[{i}]", k[i], g["k"][i].astype(np.float32), atol=atol)
        o, lse = ok.attention_decode(g["qd"].astype(np.float32), g["pages"].astype(np.float32), g["page_indptr"],
                                     g["page_values"], g["length_info"], g["kofs"], g["qpos"], 1, scale, theta, 128 ** -0.5,
                                     "float16")
        assert_close("decode O", o, g["o"].astype(np.float32), atol=6e-3)   # positions up to 5e4 inside (see above)
        assert_close("decode LSE", lse, g["lse"], atol=2e-2)
        # and the unscaled oracle must NOT match: the fixture really exercises the scaling
        ok.set_rope_scaling(None)
        q0, _, _ = ok.split_rotary(g["qkv"].astype(np.float32), g["pos"], 8, 2, 1, theta, scale, "float16")
        assert np.abs(q0 - g["q"].astype(np.float32)).max() > 0.1
    finally:
        ok.set_rope_scaling(None)


def _rope_variants():
    import json
    from pathlib import Path

    g = np.load(Path(__file__).parent / "golden" / "rope_variants.npz")
    return g, json.loads(bytes(g["meta"]).decode())


@pytest.mark.parametrize("name", ["gptj", "gptj_rd64", "llama4", "llama4_equal_factors", "yarn"])
def test_oracle_rope_variants_match_the_reference(name):
    """tests/golden/rope_variants.npz: the reference's own fused_rope built with rope_scaling = gptj / llama4 / yarn
    (oracle/ref_harness/gen_golden_rope_variants.py).  Same position-dependent slack as the llama3 fixture."""
    g, meta = _rope_variants()
    m = meta[name]
    want_q, want_k = g[f"{name}_q"].astype(np.float32), g[f"{name}_k"].astype(np.float32)
    ok.set_rope_scaling(m["rope_scaling"])
    try:
        q, k, v = ok.split_rotary(g[f"{name}_qkv"].astype(np.float32), g[f"{name}_pos"], 8, 2, 1, m["theta"], m["scale"],
                                  "float16", m["rotary_dim"])
    finally:
        ok.set_rope_scaling(None)
    assert np.array_equal(v, g[f"{name}_v"].astype(np.float32))
    for i, pos in enumerate(g[f"{name}_pos"]):
        atol = 4e-3 + 3e-7 * float(pos)
        assert_close(f"{name} q[{i}]", q[i], want_q[i], atol=atol)
        assert_close(f"{name} k[{i}]", k[i], want_k[i], atol=atol)
    assert np.array_equal(q[0], want_q[0]) and np.array_equal(q[1], want_q[1])  # small angles: bit exact
    # the default frequencies / pairing must NOT reproduce the fixture
    q0, _, _ = ok.split_rotary(g[f"{name}_qkv"].astype(np.float32), g[f"{name}_pos"], 8, 2, 1, m["theta"], m["scale"],
                               "float16", m["rotary_dim"])
    assert np.abs(q0 - want_q).max() > 0.1
