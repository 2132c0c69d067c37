"""TEST INFRASTRUCTURE ONLY -- golden vectors for the remaining RoPE scalings of f_split_rotary, produced by the REFERENCE.

Run inside the reference env:  source /tmp/tvm_ref/env.sh && python oracle/ref_harness/gen_golden_rope_variants.py
Builds the reference's own `llama_rope_with_position_map` (position_embedding.py:444-667) for rope_scaling = gptj (also
with a partial rotary_dim), llama4 (smooth interpolation and the equal-factor threshold branch) and yarn, with the
reference's `c` target, runs each on seeded inputs and stores inputs + outputs in tests/golden/rope_variants.npz.

longrope is absent on purpose: at this commit the reference's own `fused_rope_longrope_scaling` cannot be built --
`_rope` tests `if ext_factors:` on a T.Buffer (position_embedding.py:502), which raises "Cannot use and / or / not
operator to Expr" in the TVMScript parser -- so there is no reference output to pin a longrope path against."""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tvm  # noqa: E402
import refenv  # noqa: E402,F401  (registers the exp2 lowering shim)
from tvm.relax.frontend.nn.llm.kv_cache import _prepare_yarn_rope_scaling, llama_rope_with_position_map  # noqa: E402

HQ, HKV, D, DT, SCALE = 8, 2, 128, "float16", 1.0
CASES = {
    # name: (rope_scaling, theta, rotary_dim)
    "gptj": ({"rope_type": "gptj"}, 1e4, None),
    "gptj_rd64": ({"rope_type": "gptj"}, 1e4, 64),
    "llama4": ({"rope_type": "llama4", "factor": 8.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
                "original_max_position_embeddings": 8192}, 5e5, None),
    "llama4_equal_factors": ({"rope_type": "llama4", "factor": 16.0, "low_freq_factor": 1.0, "high_freq_factor": 1.0,
                              "original_max_position_embeddings": 8192}, 5e5, None),
    "yarn": ({"rope_type": "yarn", "factor": 40.0, "original_max_position_embeddings": 4096, "beta_fast": 32,
              "beta_slow": 1}, 1e4, None),
}


def main():
    rng = np.random.default_rng(11)
    out = {}
    meta = {}
    t = lambda a: tvm.runtime.tensor(a)  # noqa: E731
    for name, (rs, theta, rd) in CASES.items():
        rs_built = _prepare_yarn_rope_scaling(rs, theta)
        fn = llama_rope_with_position_map(theta, SCALE, D, HQ, HKV, DT, rs_built, rd).with_attr("global_symbol", "fused_rope")
        lib = tvm.tirx.build(tvm.IRModule({"fused_rope": fn}), target=tvm.target.Target("c"))
        path = os.path.join(tempfile.mkdtemp(), f"rope_{name}.so")
        lib.export_library(path, options=["-O2", "-Dhalf=_Float16", "-lm"])
        mod = tvm.runtime.load_module(path)
        n = 5
        qkv = rng.standard_normal((n, HQ + 2 * HKV, D)).astype(np.float16)
        pos = np.array([0, 1, 777, 9000, 100000], np.int32)
        q, k, v = (t(np.zeros((n, h, D), np.float16)) for h in (HQ, HKV, HKV))
        mod["fused_rope"](t(qkv), t(pos), q, k, v, 1)
        out[f"{name}_qkv"], out[f"{name}_pos"] = qkv, pos
        out[f"{name}_q"], out[f"{name}_k"], out[f"{name}_v"] = q.numpy(), k.numpy(), v.numpy()
        meta[name] = {"rope_scaling": rs, "theta": theta, "rotary_dim": rd, "scale": SCALE}
        print(name, "ok")
    path = os.path.join(HERE, "..", "..", "tests", "golden", "rope_variants.npz")
    np.savez_compressed(path, meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), **out)
    print("wrote", os.path.abspath(path), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
