"""GPU replay of the fixtures beyond the reference's own scenario tests (tests/test_cache_gpu_golden.py has those): the
randomised prefill / decode / fork / popn / remove programs, the per-layer sliding-window caches (attn_kinds with
MHA_SLIDING, ..._cpu.py:610-650), attention_with_shared_kv (..._cpu.py:654-703) and the self_attention /
cross_attention / merge_attn_output_inplace entries (kv_state.cc:84-115) -- callback traces bit-exact, outputs within the
north_star tolerance of what the reference's own cache + CPU kernels produced on the same inputs."""
import pytest

from tests.golden_replay import extra_scenario_names
from tests.test_cache_gpu_golden import run_scenario

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("impl", [0, 2])
@pytest.mark.parametrize("name", extra_scenario_names())
def test_extra_scenario_matches_reference(built_lib, name, impl):
    run_scenario(name, impl)
