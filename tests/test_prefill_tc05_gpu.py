"""GPU parity of the tcgen05/TMEM prefill kernel (prefill_tc05.cu) against the CPU oracle, forced on for small
shapes the oracle finishes in seconds (auto dispatch only picks it for >= 2048 folded rows)."""
import numpy as np
import pytest

from tests.test_kernels_gpu import _run_paged_prefill, _run_ragged

pytestmark = pytest.mark.gpu
DTYPES = ["float16", "bfloat16"]


@pytest.fixture()
def tc05(built_lib):
    from tvm_b200 import capi

    capi.lib()
    capi.set_prefill_impl(2)
    yield capi
    capi.set_prefill_impl(0)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("causal", [1, 0])
def test_tc05_ragged_c1(tc05, dtype, causal):
    rng = np.random.default_rng(40)
    _run_ragged(tc05, rng, [10, 20, 30, 40], [10, 20, 30, 40], 32, 8, 128, dtype, causal=causal)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("hq,hkv", [(32, 8), (8, 8), (16, 1), (32, 4), (16, 8)])
def test_tc05_ragged_groups(tc05, dtype, hq, hkv):
    rng = np.random.default_rng(41)
    _run_ragged(tc05, rng, [65, 1, 300, 128], [65, 9, 300, 200], hq, hkv, 128, dtype, causal=1)


@pytest.mark.parametrize("dtype", DTYPES)
def test_tc05_ragged_long(tc05, dtype):
    rng = np.random.default_rng(42)
    _run_ragged(tc05, rng, [700, 513], [700, 900], 32, 8, 128, dtype, causal=1)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("q_scale,v_scale", [(2.0, 1.0), (4.0, 2.0), (8.0, 1.0)])
def test_tc05_peaked(tc05, dtype, q_scale, v_scale):
    """A few keys dominate every row (scores with std 2..8) and V is larger: the case where the precision of P shows
    (bf16 P would lose it: the kernel keeps P in fp16 and converts bf16 V tiles to fp16 in shared memory); must stay
    inside the 2e-3 / 1e-2 bar at every element."""
    rng = np.random.default_rng(46)
    _run_ragged(tc05, rng, [300, 77, 513], [300, 400, 600], 32, 8, 128, dtype, causal=1, q_scale=q_scale,
                v_scale=v_scale)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("causal", [0, 1])
def test_tc05_paged(tc05, dtype, causal):
    rng = np.random.default_rng(44)
    _run_paged_prefill(tc05, rng, [3, 17, 1, 64, 5, 200], [16, 18, 0 if not causal else 1, 300, 77, 1000], 32, 8, 128,
                       dtype, causal=causal)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("causal", [1, 0])
def test_tc05_work_queue_many_ragged_items(tc05, dtype, causal):
    """The persistent kernel walks a device-wide work queue with barriers reused across items: far more items than
    SMs, of wildly different lengths, including empty sequences, sequences without any visible KV tile and items whose
    second Q tile is absent -- every row against the oracle."""
    rng = np.random.default_rng(47)
    B = 150
    q_lens = [int(x) for x in rng.integers(0, 200, B)]
    q_lens[3] = 0
    q_lens[10] = 1
    q_lens[20] = 64
    q_lens[21] = 65
    kv_lens = [q + int(rng.integers(0, 300)) for q in q_lens]
    if not causal:
        kv_lens[5] = 0          # no KV at all: O = 0, LSE = -5e4
    _run_ragged(tc05, rng, q_lens, kv_lens, 8, 2, 128, dtype, causal=causal)


@pytest.mark.parametrize("dtype", DTYPES)
def test_tc05_work_queue_paged_many_items(tc05, dtype):
    rng = np.random.default_rng(48)
    B = 100
    q_lens = [int(x) for x in rng.integers(1, 150, B)]
    kv_lens = [int(x) for x in rng.integers(0, 400, B)]
    _run_paged_prefill(tc05, rng, q_lens, kv_lens, 8, 2, 128, dtype, causal=0)


def test_tc05_back_to_back_launches_reuse_the_work_counter(tc05):
    """The kernel must leave its work counter at zero: 20 launches in a row (no host sync in between) all agree."""
    import torch

    torch.manual_seed(1)
    n, hq, hkv, d = 3000, 8, 2, 128
    q = torch.randn(n, hq, d, device="cuda", dtype=torch.bfloat16)
    k = torch.randn(n, hkv, d, device="cuda", dtype=torch.bfloat16)
    v = torch.randn(n, hkv, d, device="cuda", dtype=torch.bfloat16)
    ip = torch.tensor([0, 700, 701, 2000, 3000], dtype=torch.int32, device="cuda")
    qpos = torch.zeros(n, dtype=torch.int32, device="cuda")
    kofs = torch.zeros(4, dtype=torch.int32, device="cuda")
    outs = []
    for _ in range(20):
        o = torch.empty_like(q)
        lse = torch.empty(n, hq, device="cuda", dtype=torch.float32)
        tc05.attention_prefill_ragged(q, ip, k, v, ip, qpos, kofs, o, lse, 1, 0, 1.0, 1e4, d ** -0.5)
        outs.append((o, lse))
    torch.cuda.synchronize()
    for o, lse in outs[1:]:
        assert torch.equal(o, outs[0][0]) and torch.equal(lse, outs[0][1])


@pytest.mark.parametrize("dtype", DTYPES)
def test_tc05_paged_uninitialised_tail(tc05, dtype):
    """Slots past kv_len in the last page hold NaN bit patterns (torch.empty-like pools): output must stay finite."""
    rng = np.random.default_rng(45)
    _run_paged_prefill(tc05, rng, [40, 7], [129, 17], 32, 8, 128, dtype, causal=0, nan_tail=True)


def test_auto_dispatch_uses_tc05_for_big_shapes(built_lib):
    """n*g >= 2048 rows -> tcgen05 kernel; result identical (within tolerance) to the forced-generic path."""
    import torch

    from tvm_b200 import capi

    capi.lib()
    torch.manual_seed(0)
    n, hq, hkv, d = 1024, 32, 8, 128
    q = torch.randn(n, hq, d, device="cuda", dtype=torch.bfloat16)
    k = torch.randn(n, hkv, d, device="cuda", dtype=torch.bfloat16)
    v = torch.randn(n, hkv, d, device="cuda", dtype=torch.bfloat16)
    ip = torch.tensor([0, 300, 1024], dtype=torch.int32, device="cuda")
    qpos = torch.zeros(n, dtype=torch.int32, device="cuda")
    kofs = torch.zeros(2, dtype=torch.int32, device="cuda")
    outs = []
    for impl in (1, 0):
        capi.set_prefill_impl(impl)
        o = torch.empty_like(q)
        lse = torch.empty(n, hq, device="cuda", dtype=torch.float32)
        capi.attention_prefill_ragged(q, ip, k, v, ip, qpos, kofs, o, lse, 1, 0, 1.0, 1e4, d ** -0.5)
        torch.cuda.synchronize()
        outs.append((o.float(), lse))
    capi.set_prefill_impl(0)
    assert torch.allclose(outs[0][0], outs[1][0], atol=2e-3, rtol=1e-2)
    assert torch.allclose(outs[0][1], outs[1][1], atol=2e-3, rtol=1e-2)
