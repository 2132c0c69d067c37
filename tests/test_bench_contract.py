"""bench.py's reference arm (`--impl reference`) runs without a GPU: the JSON line must carry the contract's keys, and the
CPU baseline must be the reference's own kernels (oracle/_ref) when they were built, the NumPy port otherwise."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "decode_attn_hbm_gbps" and d["unit"] == "GB/s"
    for key in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data", "config",
                "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["value"] > 0 and d["higher_is_better"] is True and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    if (ROOT / "oracle" / "_ref" / "ref_kernels_float16_hq32_hkv8_d128.so").exists():
        assert cb["kind"] == "reference"


def test_own_arm_fails_loudly_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1", "--warmup", "1", "--no-cpu"],
                       capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert r.returncode != 0 and not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]


def test_algorithmic_work_matches_baseline_md():
    """the byte / FLOP counts bench.py divides by are BASELINE.md section 3's (the judge recomputes `roofline.achieved` from
    them): C2 1.0749e9 B per layer-call, C3 5.50e11 FLOP + 268 MB of append traffic, C4 8.599e9 B, C5 decode 4.296e9 B;
    and the tree helper emits the reference's (dfs order, subtree end) rows (paged_kv_cache.cc:1900-1918)"""
    import bench

    c2 = bench.decode_bytes_of(64, 4096, 32, 8, 128)
    assert c2 == 1073741824 + 1048576 + 8192 + 4 * (16384 + 256 + 1) == 1074865156 and abs(c2 / 1.0749e9 - 1) < 1e-4
    assert abs(bench.causal_prefill_flops_of(16, 2048, 32, 128) / 5.50e11 - 1) < 1e-3
    assert abs(bench.append_bytes_of(32768, 8, 128) / 268e6 - 1) < 5e-3
    assert abs(bench.decode_bytes_of(256, 8192, 64, 8, 128) / 8.599e9 - 1) < 1e-3
    assert abs(bench.decode_bytes_of(32, 32768, 32, 8, 128) / 4.296e9 - 1) < 1e-3
    m = bench.dfs_tree_mask([-1, 0, 0, 1, 1, 2])          # 0 -> (1 -> 3, 4), (2 -> 5)
    assert m.tolist() == [[0, 6], [1, 4], [4, 6], [2, 3], [3, 4], [5, 6]]
