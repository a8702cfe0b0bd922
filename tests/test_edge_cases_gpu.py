"""Edge cases through the C ABI: empty calls, one-dimensional chains, the largest supported dimension, ragged chain
counts (not a multiple of the chains-per-block or of the resident groups), argument validation."""
import ctypes as C

import numpy as np
import pytest

from tests import parity_utils as pu

pytestmark = pytest.mark.gpu


def _case(D, kind="nuts", n_chains=2, T=3, **kw):
    sigma = 10 ** np.linspace(-0.3, 0.3, D)
    case = dict(kind=kind, target="diag_gaussian", ndim=D, tau=1 / sigma**2, draws=1, tune=T - 1,
                start=np.full(D, 0.05), seeds=list(range(900, 900 + n_chains)), pot_adapt=1, pot_mean=np.zeros(D),
                pot_var=np.ones(D), pot_weight=10, max_treedepth=5, early_max_treedepth=4)
    case.update(kw)
    return case


def _parity(case):
    import torch
    ora = pu.oracle_run(case)
    q, var, wel, ad, st, ch = pu.gpu_run_transitionwise(case, ora)
    table = pu.NUTS_STATS if case["kind"] == "nuts" else pu.HMC_STATS
    from littlemcmc_b200 import _lib as L
    assert np.array_equal(st[:, :, L.STAT_N_UNIFORMS], ora["tapes"][2])
    for k in ("depth", "tree_size", "diverging") if case["kind"] == "nuts" else ("n_steps", "accepted", "diverging"):
        assert np.array_equal(st[:, :, table[k]], ora["stats"][k]), k
    np.testing.assert_allclose(q, ora["trace"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(var, ora["post"]["var"], rtol=1e-9, atol=1e-12)
    torch.cuda.synchronize()


@pytest.mark.parametrize("kind", ["nuts", "hmc"])
def test_one_dimensional_chains(kind):
    _parity(_case(1, kind=kind, n_chains=5, T=6))


@pytest.mark.parametrize("D", [2, 3, 63, 65, 511, 1025])
def test_dimensions_around_the_shape_boundaries(D):
    _parity(_case(D, n_chains=3, T=3))


def test_largest_supported_dimension():
    """8192 dimensions = 1024 threads x 4 pairs, the widest instantiated group; 8193 is refused, not mis-run."""
    from littlemcmc_b200 import _lib as L
    _parity(_case(8192, n_chains=2, T=2, max_treedepth=3, early_max_treedepth=3))
    lib = L.load()
    assert lib.lmc_workspace_bytes(L.KIND_NUTS, 4, 8193, 10, 0) == L.ERR_UNSUPPORTED


@pytest.mark.parametrize("n_chains", [1, 3, 5, 149, 445])
def test_ragged_chain_counts(n_chains):
    """Chain counts that are not multiples of the chains per block (4 warps) or of the resident groups."""
    import torch
    import littlemcmc_b200 as lmc
    D = 10
    target = lmc.targets.DiagGaussian(sigma=np.linspace(0.5, 2, D))
    kw = dict(model_ndim=D, draws=4, tune=6, start=np.full(D, 0.1), discard_tuned_samples=False)

    def step():   # an explicit potential: init_nuts would seed the Welford mean with a jitter drawn from seeds[0]
        return lmc.NUTS(target, D, potential=lmc.QuadPotentialDiagAdapt(D, np.zeros(D), np.ones(D), 10))
    trace, stats = lmc.sample(target, chains=n_chains, random_seed=list(range(n_chains)), step=step(), **kw)
    assert trace.shape == (n_chains, 10, D) and np.isfinite(trace).all()
    # every chain equals the same chain run alone (nothing leaks between neighbours)
    for c in (0, n_chains - 1):
        solo, _ = lmc.sample(target, chains=1, random_seed=[c], step=step(), **kw)
        assert np.array_equal(solo[0], trace[c])
    torch.cuda.synchronize()


def test_empty_calls_and_bad_arguments():
    import torch
    from littlemcmc_b200 import _lib as L
    from littlemcmc_b200 import engine
    lib = L.load()
    D = 6
    ch = engine.DeviceChains(4, D, "cuda:0")
    ch.reset_potential(np.ones(D), np.zeros(D), 10.0, 101)
    ch.reset_step_adapt(0.1)
    tgt = engine.FusedTarget(L.TARGET_DIAG_GAUSSIAN, D, tau=np.ones(D))
    params = dict(adapt_mass=1, adapt_step_size=1, target_accept=0.8, gamma=0.05, k=0.75, t0=10, Emax=1000.0)
    seeds = engine.seeds_tensor(np.arange(4), ch.device)
    q0 = ch.q.clone()
    tr, st = engine.run_transitions(L.KIND_NUTS, ch, tgt, n_trans=0, iter0=0, n_tune=0, params=params, seeds=seeds)
    torch.cuda.synchronize()
    assert tr.shape == (4, 0, D) and st.shape == (4, 0, L.NSTATS) and torch.equal(ch.q, q0)   # zero transitions: no-op
    a = L.SamplerArgs()
    assert lib.lmc_nuts_sample(C.byref(a)) == L.ERR_BADARG                     # wrong ABI version / null pointers
    assert lib.lmc_nuts_sample(None) == L.ERR_BADARG
    assert lib.lmc_workspace_bytes(L.KIND_NUTS, -1, D, 10, 0) == L.ERR_BADARG
    assert lib.lmc_workspace_bytes(L.KIND_NUTS, 4, D, 17, 0) == L.ERR_UNSUPPORTED  # max_treedepth > 16
    with pytest.raises(L.LmcError):
        engine.run_transitions(L.KIND_NUTS, ch, tgt, n_trans=1, iter0=0, n_tune=0, seeds=seeds,
                               params=dict(params, max_treedepth=40))
