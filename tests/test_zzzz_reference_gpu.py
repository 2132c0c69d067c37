"""GPU box only: tvm_b200 next to the UNMODIFIED reference running on the same B200 (oracle/ref_gpu.py starts the
reference's own runtime -- libtvm_runtime{,_cuda,_extra} + its vendored tvm-ffi, built from /root/reference by
oracle/ref_harness/build_tvm_cuda.sh and packed into oracle/_ref/tvm_cuda -- in a subprocess).

  * route A: the reference's C++ PagedAttentionKVCacheObj constructed with tvm_b200's 13 packed callbacks
    (`vm.builtin.paged_attention_kv_cache_create`, paged_kv_cache.cc:2535-2639), driven through the scenario programs of
    the reference's own tests (tests/python/relax/test_runtime_builtin_paged_attention_kv_cache_{cpu,tir}.py scenario
    lists) and the randomised ones; outputs vs what the reference computed with its own kernels.
  * the reference's own GPU TIR kernels (`_attention_decode`, `_attention_prefill[_ragged]`, `tree_attn*`,
    `_merge_state_inplace`, `fused_rope`, append; _decode_kernels.py:181-411, _prefill_kernels.py:217-391, 795-923) built
    for sm_100a in float16 AND bfloat16: our kernels must land within max-abs 2e-3 / rtol 1e-2 of them on the same
    inputs -- the tolerance of the north_star is defined against exactly these, and bf16 exists in the reference only here.
"""
import numpy as np
import pytest

from oracle import kernels as ok
from oracle import ref_gpu
from tests import golden_replay as gr
from tests.util import assert_close, make_paged_cache, rand16, to_dev, to_np

pytestmark = pytest.mark.gpu

HQ, HKV, D = 32, 8, 128
SM = D ** -0.5
THETA = 5e5  # what oracle/ref_harness/emit_ref_gpu_kernels.py baked into the reference's fused_rope


def _need(dtype=None):
    if not ref_gpu.available(dtype):
        pytest.skip("oracle/_ref/tvm_cuda (the reference's CUDA runtime) is not packed: run oracle/ref_harness/pack_ref_cuda.sh")


def test_route_a_reference_cache_drives_our_callbacks(built_lib):
    _need()
    names = gr.scenario_names()
    res = ref_gpu.route_a(names, timeout=1500)
    assert res.get("tvm_ffi", "").startswith("0.1.14"), res
    bad = [f for f in res.get("fixtures", []) if not f.get("ok")]
    assert res.get("ok") and not bad, "route A failed:\n" + "\n".join(f"{f['name']}: {f.get('error')}\n  " + "\n  ".join(f.get("trace", [])) for f in bad) + str(res.get("error", ""))
    assert len(res["fixtures"]) == len(names) >= 22
    assert res["kernel_launches"] > 1000          # our kernels ran (no fallback exists in that process)
    roles = set(res["fixtures"][0]["callbacks"])
    used = {k for f in res["fixtures"] for k, v in f["callbacks"].items() if v > 0}
    assert len(roles) == 13 and used == roles, f"callback roles never called by the reference's cache: {roles - used}"


# ---------------------------------------------------------------------------------------------------------------------
def _bits(x, dtype):
    return ok.to_bits16(np.asarray(x, np.float32), dtype)


class Spec:
    """Collects tensors and calls for ONE reference-server run."""

    def __init__(self, dtype):
        self.dtype = dtype
        self.tensors, self.arrays, self.calls, self.fetch = {}, {}, [], []

    def inp(self, name, arr, kind=None):
        kind = kind or self.dtype
        if kind in ("float16", "bfloat16"):
            self.arrays[name] = _bits(arr, kind)
        else:
            self.arrays[name] = np.ascontiguousarray(arr)
        self.tensors[name] = {"dtype": kind, "init": "npz"}
        return name

    def out(self, name, shape, kind=None):
        self.tensors[name] = {"dtype": kind or self.dtype, "shape": list(shape), "init": "zeros"}
        self.fetch.append(name)
        return name

    def call(self, fn, *args):
        self.calls.append({"fn": fn, "args": list(args)})

    def run(self):
        res, out = ref_gpu.run_kernels({"module": ref_gpu.kernel_module(self.dtype).name, "tensors": self.tensors,
                                        "calls": self.calls, "fetch": self.fetch}, self.arrays)
        assert res.get("ok"), res
        dec = {}
        for k, v in out.items():
            kind = self.tensors[k]["dtype"]
            dec[k] = ok.from_bits16(v, kind) if kind in ("float16", "bfloat16") else v
        return dec


def _tree(rng, n):
    parent = [-1] + [int(rng.integers(-1 if k > 3 else 0, k)) for k in range(1, n)]
    children = [[] for _ in range(n)]
    roots = []
    for k, p in enumerate(parent):
        (roots if p == -1 else children[p]).append(k)
    iv = np.zeros((n, 2), np.int32)
    order = [0]

    def dfs(u):
        iv[u, 0] = order[0]
        order[0] += 1
        ub = iv[u, 0] + 1
        for ch in children[u]:
            ub = max(ub, dfs(ch))
        iv[u, 1] = ub
        return ub

    for r in roots:
        dfs(r)
    return iv


@pytest.mark.parametrize("dtype", ["float16", "bfloat16"])
def test_our_kernels_vs_the_references_gpu_tir_kernels(built_lib, dtype):
    import torch

    from tvm_b200 import capi

    _need(dtype)
    rng = np.random.default_rng(2026)
    sp = Spec(dtype)
    i32 = lambda x: to_dev(np.asarray(x, np.int32))  # noqa: E731
    ours = {}
    tdt = torch.float16 if dtype == "float16" else torch.bfloat16
    new = lambda *s, dt=None: torch.zeros(s, dtype=dt or tdt, device="cuda")  # noqa: E731

    # ---- fused_rope + append ----
    n = 70
    qkv = rand16(rng, (n, HQ + 2 * HKV, D), dtype)
    # positions < 4096: the rotation angle pos / theta^(2d/D) is an fp32 number, and at position 30000 one ulp of it is
    # 2e-3 rad -- the reference's GPU code and ours (and the reference's own CPU code) then differ by ~1e-2 on |x| ~ 4
    # purely through the order of the fp32 operations; below 4096 every side agrees to one ulp of the 16-bit output
    pos = rng.integers(0, 4096, n).astype(np.int32)
    P = 40
    pages0 = rand16(rng, (P, 2, HKV, 16, D), dtype)
    slots = rng.permutation(P * 16)[:n].astype(np.int32)
    slots[::9] = -1
    sp.inp("qkv", qkv), sp.inp("pos", pos, "int32"), sp.inp("pages0", pages0), sp.inp("slots", slots, "int32")
    sp.out("rq", (n, HQ, D)), sp.out("rk", (n, HKV, D)), sp.out("rv", (n, HKV, D))
    sp.call("fused_rope", "qkv", "pos", "rq", "rk", "rv", 1)
    sp.call("tir_kv_cache_transpose_append", "pages0", "rk", "rv", "slots")
    sp.fetch.append("pages0")
    capi.set_rope_scaling(None)
    q_, k_, v_ = new(n, HQ, D), new(n, HKV, D), new(n, HKV, D)
    capi.split_rotary(to_dev(qkv, dtype), i32(pos), q_, k_, v_, 1, 1.0, THETA)
    dp0 = to_dev(pages0, dtype)
    ours.update(rq=q_, rk=k_, rv=v_)

    # ---- decode: ragged lengths, an empty sequence, split-KV sizes; inline RoPE; sliding-window flavour ----
    def decode_case(tag, kv_lens, rotary, sliding=None):
        B = len(kv_lens)
        c = make_paged_cache(rng, kv_lens, HKV, D, dtype, sliding=sliding)
        q = rand16(rng, (B, HQ, D), dtype)
        kro = rng.integers(0, 64, B).astype(np.int32)
        qpos = (kro + np.array(kv_lens)).astype(np.int32)
        names = [sp.inp(f"{tag}_{k}", v, "int32" if v.dtype == np.int32 else None)
                 for k, v in [("q", q), ("pages", c["pages"]), ("ip", c["page_indptr"]), ("iv", c["page_values"]),
                              ("li", c["length_info"]), ("kro", kro), ("qpos", qpos)]]
        sp.out(f"{tag}_o", (B, HQ, D)), sp.out(f"{tag}_lse", (B, HQ), "float32")
        fn = "batch_decode_paged_kv_sliding_window" if sliding else "batch_decode_paged_kv"
        sp.call(fn, *names, f"{tag}_o", f"{tag}_lse", rotary, 1.0, THETA, SM)
        o, lse = new(B, HQ, D), new(B, HQ, dt=torch.float32)
        capi.attention_decode(to_dev(q, dtype), to_dev(c["pages"], dtype), i32(c["page_indptr"]), i32(c["page_values"]),
                              i32(c["length_info"]), i32(kro), i32(qpos), o, lse, rotary, 1.0, THETA, SM)
        ours[f"{tag}_o"], ours[f"{tag}_lse"] = o, lse

    decode_case("dec", [1, 16, 17, 700, 33, 2049, 4096], 0)
    decode_case("decr", [40, 7, 300], 1)
    decode_case("decs", [64, 100, 37], 1, sliding=[(0, 0), (19, 4), (5, 5)])

    # ---- ragged prefill (causal), generic-path and tcgen05-path sizes; inline RoPE ----
    def ragged_case(tag, lens, rotary, tree=None):
        B = len(lens)
        qi = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
        m = int(qi[-1])
        q, k, v = (rand16(rng, (m, h, D), dtype) for h in (HQ, HKV, HKV))
        kro = rng.integers(0, 32, B).astype(np.int32)
        qpos = np.concatenate([kro[b] + np.arange(lens[b]) for b in range(B)]).astype(np.int32)
        nq, nk, nv, nqi, nki, nqp, nkro = (sp.inp(f"{tag}_{k_}", v_, "int32" if v_.dtype == np.int32 else None)
                                           for k_, v_ in [("q", q), ("k", k), ("v", v), ("qi", qi), ("ki", qi.copy()),
                                                          ("qpos", qpos), ("kro", kro)])
        sp.out(f"{tag}_o", (m, HQ, D)), sp.out(f"{tag}_lse", (m, HQ), "float32")
        o, lse = new(m, HQ, D), new(m, HQ, dt=torch.float32)
        dq, dk, dv = to_dev(q, dtype), to_dev(k, dtype), to_dev(v, dtype)
        if tree is None:
            sp.call("batch_prefill_ragged_kv", nq, nqi, nk, nv, nki, nqp, nkro, f"{tag}_o", f"{tag}_lse", 1, rotary, 1.0,
                    THETA, SM)
            capi.attention_prefill_ragged(dq, i32(qi), dk, dv, i32(qi), i32(qpos), i32(kro), o, lse, 1, rotary, 1.0, THETA, SM)
        else:
            nmn, nmask = sp.inp(f"{tag}_mn", tree[0], "int32"), sp.inp(f"{tag}_mask", tree[1], "int32")
            sp.call("batch_tree_attn", nq, nqi, nk, nv, nki, nqp, nmn, nmask, f"{tag}_o", f"{tag}_lse", rotary, 1.0, THETA, SM)
            capi.attention_prefill_tree_ragged(dq, i32(qi), dk, dv, i32(qi), i32(qpos), i32(tree[0]), i32(tree[1]), o, lse,
                                               rotary, 1.0, THETA, SM)
        ours[f"{tag}_o"], ours[f"{tag}_lse"] = o, lse

    ragged_case("rag", [10, 20, 30, 40], 0)
    ragged_case("ragl", [257, 1, 300, 128], 0)
    ragged_case("ragr", [33, 5, 70], 1)
    sizes = [7, 64, 20, 1]
    ragged_case("tree", sizes, 0, tree=(np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32),
                                        np.concatenate([_tree(rng, s) for s in sizes]).astype(np.int32)))

    # ---- paged prefill: non-causal / causal, sliding flavour with inline RoPE, tree mask on the trailing columns ----
    def paged_case(tag, q_lens, kv_lens, causal, rotary=0, sliding=None, tree=None):
        B = len(q_lens)
        c = make_paged_cache(rng, kv_lens, HKV, D, dtype, sliding=sliding)
        qi = np.concatenate([[0], np.cumsum(q_lens)]).astype(np.int32)
        m = int(qi[-1])
        q = rand16(rng, (m, HQ, D), dtype)
        kro = rng.integers(0, 32, B).astype(np.int32)
        qpos = np.concatenate([kro[b] + kv_lens[b] - q_lens[b] + np.arange(q_lens[b]) for b in range(B)]).astype(np.int32)
        names = [sp.inp(f"{tag}_{k}", v, "int32" if v.dtype == np.int32 else None)
                 for k, v in [("q", q), ("qi", qi), ("pages", c["pages"]), ("ip", c["page_indptr"]), ("iv", c["page_values"]),
                              ("li", c["length_info"]), ("kro", kro), ("qpos", qpos)]]
        sp.out(f"{tag}_o", (m, HQ, D)), sp.out(f"{tag}_lse", (m, HQ), "float32")
        o, lse = new(m, HQ, D), new(m, HQ, dt=torch.float32)
        args = (to_dev(q, dtype), i32(qi), to_dev(c["pages"], dtype), i32(c["page_indptr"]), i32(c["page_values"]),
                i32(c["length_info"]), i32(kro), i32(qpos), o, lse)
        if tree is None:
            fn = "batch_prefill_paged_kv_sliding_window" if sliding else "batch_prefill_paged_kv"
            sp.call(fn, *names, f"{tag}_o", f"{tag}_lse", causal, rotary, 1.0, THETA, SM)
            capi.attention_prefill_paged(*args, causal, rotary, 1.0, THETA, SM, layer_sliding_window_size=1024 if sliding else 0)
        else:
            nti, nto = sp.inp(f"{tag}_ti", tree[0], "int32"), sp.inp(f"{tag}_to", tree[1], "int32")
            sp.call("tree_attn_paged_kv", *names, f"{tag}_o", f"{tag}_lse", rotary, 1.0, THETA, SM, nti, nto)
            capi.attention_prefill_tree_paged(*args, rotary, 1.0, THETA, SM, i32(tree[0]), i32(tree[1]))
        ours[f"{tag}_o"], ours[f"{tag}_lse"] = o, lse

    paged_case("pg0", [5, 64, 1, 130], [40, 64, 17, 450], 0)
    paged_case("pg1", [5, 64, 1, 130], [40, 64, 17, 450], 1)
    paged_case("pgl", [300, 260], [300 + 128, 260 + 517], 0)       # tcgen05 path on our side
    paged_case("pgs", [3, 20], [70, 120], 0, rotary=1, sliding=[(9, 2), (0, 0)])
    sizes = [7, 64, 20]
    paged_case("pgt", sizes, [7 + 33, 64 + 300, 20], 0,
               tree=(np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32),
                     np.concatenate([_tree(rng, s) for s in sizes]).astype(np.int32)))

    # ---- merge ----
    N = 90
    mv, mvo = rand16(rng, (N, HQ, D), dtype), rand16(rng, (N, HQ, D), dtype)
    ms = (rng.standard_normal((N, HQ)) * 3).astype(np.float32)
    mso = (rng.standard_normal((N, HQ)) * 3).astype(np.float32)
    mso[::5] = -5e4
    sp.inp("mv", mv), sp.inp("ms", ms, "float32"), sp.inp("mvo", mvo), sp.inp("mso", mso, "float32")
    sp.call("merge_state_inplace", "mv", "ms", "mvo", "mso")
    sp.fetch += ["mv", "ms"]
    dmv, dms = to_dev(mv, dtype), to_dev(ms)
    capi.merge_state_inplace(dmv, dms, to_dev(mvo, dtype), to_dev(mso))
    ours.update(mv=dmv, ms=dms)

    ref = sp.run()
    torch.cuda.synchronize()

    # append: feed OUR rotated k / v through OUR append and compare pages bit for bit with the reference's pipeline only
    # where its rotated k equals ours (RoPE rounding may differ by an ulp); V rows and untouched slots must be identical
    capi.transpose_append(dp0, ours["rk"], ours["rv"], i32(slots))
    torch.cuda.synchronize()
    got_pages, ref_pages = to_np(dp0), ref["pages0"]
    assert np.array_equal(got_pages[:, 1], ref_pages[:, 1]), "appended V pages differ from the reference's"
    assert_close("appended K pages", got_pages[:, 0], ref_pages[:, 0], atol=4e-3 if dtype == "float16" else 3.2e-2)
    assert np.array_equal(to_np(ours["rv"]), ref["rv"])
    for k in ("rq", "rk"):
        assert_close(k, to_np(ours[k]), ref[k], atol=4e-3 if dtype == "float16" else 3.2e-2)
    for k in sorted(ours):
        if k in ("rq", "rk", "rv"):
            continue
        assert_close(f"{dtype} {k}", to_np(ours[k]), ref[k])
