"""Dry run of tests/test_zzz_ref_kernels_gpu.py on the CPU: the oracle stands in for the CUDA kernels (a fake capi), so the
argument order / shapes of the test bodies are exercised end to end."""
import sys, types
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from oracle import kernels as ok, cpu_ref
import tests.util as U

# keep everything on the CPU
_orig_full, _orig_empty = torch.full, torch.empty
torch.full = lambda *a, **k: _orig_full(*a, **{**k, "device": "cpu"})
torch.empty = lambda *a, **k: _orig_empty(*a, **{**k, "device": "cpu"})
torch.cuda.synchronize = lambda *a, **k: None
U.to_dev = lambda x, dtype=None, device="cpu": (torch.from_numpy(np.ascontiguousarray(x)).to(U.torch_dtype(dtype)) if dtype else torch.from_numpy(np.ascontiguousarray(x)))
f = lambda t: t.float().numpy() if t.dtype in (torch.float16, torch.bfloat16) else t.numpy()
def put(o, lse, r):
    o.copy_(torch.from_numpy(r[0]).to(o.dtype)); lse.copy_(torch.from_numpy(np.asarray(r[1], np.float32)))
fake = types.SimpleNamespace()
def attention_prefill_paged(q, qi, pages, pip, piv, li, kofs, qpos, o, lse, causal, rot, scale, theta, sm, layer_sliding_window_size=0):
    assert (li.dim() == 2) == (layer_sliding_window_size != 0) or li.dim() == 1
    put(o, lse, ok.attention_prefill_paged(f(q), f(qi), f(pages), f(pip), f(piv), f(li), f(kofs), f(qpos), causal, rot, scale, theta, sm, "float16", sliding_window_size=layer_sliding_window_size))
def attention_decode(q, pages, pip, piv, li, kofs, qpos, o, lse, rot, scale, theta, sm):
    put(o, lse, ok.attention_decode(f(q), f(pages), f(pip), f(piv), f(li), f(kofs), f(qpos), rot, scale, theta, sm, "float16"))
def attention_prefill_tree_ragged(q, qi, k, v, ki, qpos, mn, mask, o, lse, rot, scale, theta, sm):
    put(o, lse, ok.attention_prefill_ragged(f(q), f(qi), f(k), f(v), f(ki), f(qpos), None, 0, rot, scale, theta, sm, "float16", mn_indptr=f(mn), tree_mask=f(mask)))
def attention_prefill_tree_paged(q, qi, pages, pip, piv, li, kofs, qpos, o, lse, rot, scale, theta, sm, ti, to):
    put(o, lse, ok.attention_prefill_paged(f(q), f(qi), f(pages), f(pip), f(piv), f(li), f(kofs), f(qpos), 0, rot, scale, theta, sm, "float16", tree_indptr=f(ti), tree_order=f(to)))
def attention_prefill_ragged(q, qi, k, v, ki, qpos, kofs, o, lse, causal, rot, scale, theta, sm):
    put(o, lse, ok.attention_prefill_ragged(f(q), f(qi), f(k), f(v), f(ki), f(qpos), f(kofs), causal, rot, scale, theta, sm, "float16"))
def merge_state_inplace(v, s, v2, s2):
    r = ok.merge_state_inplace(f(v).copy(), f(s).copy(), f(v2), f(s2), "float16"); put(v, s, r)
def split_rotary_append(qkv, qpos, apos, q, k, v, pages, apply, scale, theta, rotary_dim=0):
    hq, hkv = q.shape[1], k.shape[1]
    rq, rk, rv = ok.split_rotary(f(qkv), f(qpos), hq, hkv, apply, theta, scale, "float16")
    for t, r in ((q, rq), (k, rk), (v, rv)): t.copy_(torch.from_numpy(r).to(t.dtype))
    P = f(pages).copy(); ok.transpose_append(P, rk, rv, f(apos)); pages.copy_(torch.from_numpy(P).to(pages.dtype))
for fn in (attention_prefill_paged, attention_decode, attention_prefill_tree_ragged, attention_prefill_tree_paged, attention_prefill_ragged, merge_state_inplace, split_rotary_append):
    setattr(fake, fn.__name__, fn)
import tvm_b200
sys.modules["tvm_b200.capi"] = fake
tvm_b200.capi = fake
import tests.test_zzz_ref_kernels_gpu as T
T.to_dev = U.to_dev
T._i32 = lambda x: U.to_dev(np.asarray(x, np.int32))
mod = cpu_ref._ref_module("float16", 32, 8, 128)
T.test_decode_step_vs_the_reference_kernels(None, mod); print("decode step ok")
for rm in (0, 1):
    T.test_ragged_prefill_and_merge_vs_the_reference_kernels(None, mod, rm)
    T.test_sliding_window_flavours_vs_the_reference_kernels(None, mod, rm)
print("ragged/merge/sliding ok")
for c, r in [(0, 0), (1, 0), (0, 1), (1, 1)]:
    T.test_paged_prefill_vs_the_reference_kernels(None, mod, c, r)
print("paged ok")
T.test_tree_attention_vs_the_reference_kernels(None, mod); print("tree ok")
