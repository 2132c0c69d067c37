"""CPU check of the CUDA sources' RoPE-variant arithmetic (no GPU): the `__device__` angle function
`rope_variant_angle`, the host parameter derivation `make_rope_variant` (tvm_b200/csrc/common.cuh) and the launch-time
yarn correction range `variant_for_launch` (page_kernels.cu) are compiled for the host with g++ -- same source text, float
math without contraction -- and their angles are compared with the oracle's restatement of the reference's rope_freq_*
functions (position_embedding.py:70-254) for the configurations of tests/golden/rope_variants.npz."""
import json
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import kernels as ok

ROOT = Path(__file__).resolve().parent.parent
MAIN = r'''
#include <cstdio>
#include <cstdlib>
int main(int argc, char** argv) {
  const int kind = atoi(argv[1]);
  RopeVariant rv;
  if (kind == 5) {
    rv = RopeVariant{5, (float)atof(argv[2]), (float)atof(argv[5]), (float)atof(argv[3]), (float)atof(argv[4]), 0.f};
  } else {
    rv = make_rope_variant(kind, (float)atof(argv[2]), (float)atof(argv[3]), (float)atof(argv[4]), (float)atof(argv[5]));
  }
  const int rd = atoi(argv[6]);
  const float theta = (float)atof(argv[7]);
  rv = variant_for_launch(rv, rd, theta);
  for (int i = 8; i < argc; ++i)
    for (int d = 0; d < rd; ++d) printf("%.9g\n", rope_variant_angle((float)atof(argv[i]), d, rd, theta, rv));
  return 0;
}
'''


@pytest.fixture(scope="module")
def angle_binary(tmp_path_factory):
    common = (ROOT / "tvm_b200" / "csrc" / "common.cuh").read_text()
    pk = (ROOT / "tvm_b200" / "csrc" / "page_kernels.cu").read_text()
    a, b = common.index("struct RopeVariant {"), common.index("// cos * x + sin * partner with ONE fixed rounding sequence")
    body = common[a:b].replace("__device__ __forceinline__", "static inline")
    body = body.replace("RopeVariant rope_variant();", "").replace("int check_no_rope_variant(const char* who);", "")
    a, b = pk.index("static RopeVariant variant_for_launch("), pk.index("static int launch_split_rotary_variant(")
    d = tmp_path_factory.mktemp("rope_angle")
    (d / "t.cpp").write_text("#include <cmath>\n#include <algorithm>\n" + body + pk[a:b] + MAIN)
    subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-o", str(d / "t"), str(d / "t.cpp")])
    return d / "t"


@pytest.mark.parametrize("name", ["gptj", "gptj_rd64", "llama4", "llama4_equal_factors", "yarn"])
def test_device_angle_function_matches_the_oracle(angle_binary, name):
    g = np.load(ROOT / "tests" / "golden" / "rope_variants.npz")
    m = json.loads(bytes(g["meta"]).decode())[name]
    rs, rd, theta = m["rope_scaling"], m["rotary_dim"] or 128, m["theta"]
    kind = {"gptj": 2, "llama4": 3, "yarn": 5}[rs["rope_type"]]
    if kind == 5:   # (factor, beta_fast, beta_slow, original_max_position_embeddings)
        args = [rs["factor"], rs["beta_fast"], rs["beta_slow"], rs["original_max_position_embeddings"]]
    else:           # (factor, low, high, original_max_position_embeddings)
        args = [rs.get("factor", 0), rs.get("low_freq_factor", 0), rs.get("high_freq_factor", 0),
                rs.get("original_max_position_embeddings", 0)]
    pos = [0.0, 1.0, 777.0, 4095.0, 9000.0, 100000.0]
    out = subprocess.check_output([str(angle_binary), str(kind)] + [repr(float(x)) for x in args] + [str(rd), repr(float(theta))]
                                  + [repr(p) for p in pos]).decode().split()
    got = np.array([float(x) for x in out], np.float64).reshape(len(pos), rd)
    ok.set_rope_scaling(rs)
    try:
        cos, sin = ok.rope_cos_sin(np.array(pos, np.float32), rd, theta)
    finally:
        ok.set_rope_scaling(None)
    for i, p in enumerate(pos):  # a float32 angle: one ulp is 6e-8 * angle, and powf may differ by an ulp from NumPy's
        tol = 2e-7 + 1e-7 * p
        assert np.abs(np.cos(got[i]) - cos[i]).max() <= tol, (name, p)
        assert np.abs(np.sin(got[i]) - sin[i]).max() <= tol, (name, p)


def test_rope_scaling_setters_validate_without_a_gpu(built_lib):
    from tvm_b200 import capi

    try:
        capi.set_rope_scaling({"rope_type": "gptj"})
        assert capi.get_rope_scaling_kind() == 2
        capi.set_rope_scaling({"rope_type": "llama4", "factor": 16.0, "low_freq_factor": 1.0, "high_freq_factor": 1.0,
                               "original_max_position_embeddings": 8192})
        assert capi.get_rope_scaling_kind() == 3
        capi.set_rope_scaling({"rope_type": "yarn", "factor": 40.0, "original_max_position_embeddings": 4096,
                               "beta_fast": 32, "beta_slow": 1})
        assert capi.get_rope_scaling_kind() == 5
        capi.set_rope_scaling({"rope_type": "llama3", "factor": 8.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
                               "original_max_position_embeddings": 8192})
        assert capi.get_rope_scaling_kind() == 1
        with pytest.raises(capi.TvmB200Error, match="high_freq_factor != low_freq_factor"):
            capi.set_rope_scaling({"rope_type": "llama3", "factor": 8.0, "low_freq_factor": 1.0, "high_freq_factor": 1.0,
                                   "original_max_position_embeddings": 8192})
        with pytest.raises(capi.TvmB200Error, match="must be positive"):
            capi.set_rope_scaling({"rope_type": "yarn", "factor": 0.0, "original_max_position_embeddings": 4096,
                                   "beta_fast": 32, "beta_slow": 1})
        with pytest.raises(capi.TvmB200Error, match="not implemented"):
            capi.set_rope_scaling({"rope_type": "longrope"})
    finally:
        capi.set_rope_scaling(None)
    assert capi.get_rope_scaling_kind() == 0
