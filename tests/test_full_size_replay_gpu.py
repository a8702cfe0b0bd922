"""BASELINE.json's configurations at FULL chain count, tied to the CPU oracle directly: the run uses in-kernel Philox
randomness and the FIFO transition scheduler under contention (more chains than resident groups); a few randomly
chosen chains are then replayed transition by transition in the oracle from their device pre-state with their own
Philox streams (dumped by lmc_rng_fill).  Same bar as tests/test_gpu_parity.py: integer / boolean statistics and the
number of uniforms consumed EXACT, float64 quantities to RTOL = 1e-9."""
import numpy as np
import pytest

from oracle import lmc_oracle as orc
from tests import parity_utils as pu

pytestmark = pytest.mark.gpu

RTOL = 1e-9


def _gauss(D, kind):
    from littlemcmc_b200 import _lib as L
    from littlemcmc_b200.engine import FusedTarget
    tau = 1 / 10 ** np.linspace(0, 4, D) if kind == "illcond" else 1 / (10 ** np.linspace(-0.5, 0.5, D)) ** 2
    return orc.diag_gaussian(tau), FusedTarget(L.TARGET_DIAG_GAUSSIAN, D, tau=tau)


def _funnel(D):
    from littlemcmc_b200 import _lib as L
    from littlemcmc_b200.engine import FusedTarget
    return orc.neal_funnel(D), FusedTarget(L.TARGET_FUNNEL, D, v_scale=3.0)


CONFIGS = {
    # name: (chains, D, target, max_treedepth, warm-up transitions)
    "headline_1024x1000": (1024, 1000, "gauss", 10, 40),
    "cfg2_1024x100": (1024, 100, "gauss", 10, 40),
    "cfg3_4096x1000_illcond": (4096, 1000, "illcond", 10, 30),
    "cfg4_8192x50_funnel_depth12": (8192, 50, "funnel", 12, 40),
}


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_sampled_chains_of_a_full_size_run_replay_in_the_oracle(name):
    C, D, kind, depth, warm = CONFIGS[name]
    f, tgt = _funnel(D) if kind == "funnel" else _gauss(D, kind)
    start = 2 * np.random.RandomState(5).rand(D) - 1          # one jittered start for all chains (sampling.py:584)
    results = pu.replay_sampled_chains(f, tgt, D, C, max_treedepth=depth, n_warm=warm, n_check=4, n_sample=8, seed=17,
                                       start=start)
    sizes = []
    for res in results:
        pu.assert_parity(res, rtol=RTOL)
        sizes += res.gpu_stats["tree_size"].ravel().tolist()
    print(name, "replayed tree sizes:", [int(s) for s in sizes])
    assert max(sizes) > 1
