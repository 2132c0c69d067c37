"""TEST INFRASTRUCTURE ONLY -- golden vectors for RoPE frequency scaling, produced by the REFERENCE itself.

Run inside the reference env:  source /tmp/tvm_ref/env.sh && python oracle/ref_harness/gen_golden_rope.py
Builds the reference's own `llama_rope_with_position_map` (f_split_rotary, position_embedding.py:444-565) and
`_attention_decode_cpu` with inline RoPE (_decode_kernels.py:49-178, rotary_mode = 1) for rope_scaling =
{"rope_type": "llama3", factor 8, low_freq_factor 1, high_freq_factor 4, original_max_position_embeddings 8192}
(the Llama-3.1 configuration; rope_freq_llama3, position_embedding.py:130-160) with the reference's `c` target, runs them
on seeded inputs and stores inputs + outputs in tests/golden/rope_llama3.npz (small: 5 tokens / 3 sequences)."""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tvm  # noqa: E402
import refenv  # noqa: E402,F401  (registers the exp2 lowering shim)
from tvm.relax.frontend.nn.llm.kv_cache import _attention_decode_cpu, llama_rope_with_position_map  # noqa: E402

RS = {"rope_type": "llama3", "factor": 8.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
      "original_max_position_embeddings": 8192}
THETA, SCALE, HQ, HKV, D, DT = 5e5, 1.0, 8, 2, 128, "float16"


def main():
    rope = llama_rope_with_position_map(THETA, SCALE, D, HQ, HKV, DT, RS).with_attr("global_symbol", "fused_rope")
    dec = _attention_decode_cpu(HKV, HQ, D, DT, False, RS).with_attr("global_symbol", "decode")
    lib = tvm.tirx.build(tvm.IRModule({"fused_rope": rope, "decode": dec}), target=tvm.target.Target("c"))
    path = os.path.join(tempfile.mkdtemp(), "rope_llama3.so")
    lib.export_library(path, options=["-O2", "-Dhalf=_Float16", "-lm"])
    mod = tvm.runtime.load_module(path)
    rng = np.random.default_rng(7)
    # ---- f_split_rotary: positions across the three llama3 regimes (high / smoothed / low frequency bands)
    n = 5
    qkv = rng.standard_normal((n, HQ + 2 * HKV, D)).astype(np.float16)
    pos = np.array([0, 1, 777, 9000, 100000], np.int32)
    t = lambda a: tvm.runtime.tensor(a)  # noqa: E731
    q, k, v = (t(np.zeros((n, h, D), np.float16)) for h in (HQ, HKV, HKV))
    mod["fused_rope"](t(qkv), t(pos), q, k, v, 1)
    # ---- decode with inline RoPE on a small paged cache (3 sequences)
    kv_lens = [5, 33, 40]
    B = len(kv_lens)
    npages = [-(-L // 16) for L in kv_lens]
    total = sum(npages) + 2
    pages = rng.standard_normal((total, 2, HKV, 16, D)).astype(np.float16)
    page_values = rng.permutation(total).astype(np.int32)[: sum(npages)]
    page_indptr = np.concatenate([[0], np.cumsum(npages)]).astype(np.int32)
    length_info = np.array([((L - 1) % 16) + 1 for L in kv_lens], np.int32)
    kofs = np.array([9000, 0, 50000], np.int32)
    qpos = (kofs + np.array(kv_lens) - 1).astype(np.int32)
    qd = rng.standard_normal((B, HQ, D)).astype(np.float16)
    o, lse = t(np.zeros((B, HQ, D), np.float16)), t(np.zeros((B, HQ), np.float32))
    mod["decode"](t(qd), t(pages), t(page_indptr), t(page_values), t(length_info), t(kofs), t(qpos), o, lse, 1, SCALE,
                  THETA, D ** -0.5)
    out = os.path.join(HERE, "..", "..", "tests", "golden", "rope_llama3.npz")
    np.savez_compressed(out, qkv=qkv, pos=pos, q=q.numpy(), k=k.numpy(), v=v.numpy(), pages=pages,
                        page_values=page_values, page_indptr=page_indptr, length_info=length_info, kofs=kofs, qpos=qpos,
                        qd=qd, o=o.numpy(), lse=lse.numpy(),
                        params=np.array([THETA, SCALE, RS["factor"], RS["low_freq_factor"], RS["high_freq_factor"],
                                         RS["original_max_position_embeddings"]], np.float64))
    print("wrote", os.path.abspath(out), os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
