"""Callback mode (csrc/lmc_callback.cu): the user's logp_dlogp_func evaluated OUTSIDE the kernels -- as one batched torch
op for all chains (targets.TorchBatched, eager or CUDA-graph) or as the reference's per-chain NumPy callable -- against
the CPU oracle on identical randomness, with the same protocol and the same bar as tests/test_gpu_parity.py: integer /
boolean statistics and the number of uniforms consumed EXACT, float64 quantities to RTOL = 1e-9."""
import numpy as np
import pytest

from tests import golden_cases as gc
from tests import parity_utils as pu

pytestmark = pytest.mark.gpu

RTOL = 1e-9


def test_device_loop_over_parallel_batches_of_chains():
    """engine.run_transitions_callback(split=K): K batches of chains on parallel branches of the loop body, one shared
    counter of running chains -- the same bits as one batch."""
    import torch
    from littlemcmc_b200 import _lib as L, engine
    case, _ = gc.load("nuts_diag_d37")
    D, Cn, T = int(case["ndim"]), 11, 12
    cb = pu.torch_callback(case, cuda_graph=True)
    params = pu.gpu_params(case)
    seeds = engine.seeds_tensor(np.arange(Cn) * 7919 + 3, "cuda:0")
    outs = []
    for split in (1, 3):
        ch = pu.gpu_chains(case, Cn)
        tr, st = engine.run_transitions_callback(L.KIND_NUTS, ch, cb, n_trans=T, iter0=0, n_tune=8, params=params,
                                                 seeds=seeds, cuda_graph=True, split=split)
        torch.cuda.synchronize()
        outs.append((tr.cpu().numpy(), st.cpu().numpy(), ch.adapt.cpu().numpy()))
    for a, b in zip(outs[0], outs[1]):
        assert np.array_equal(a, b)


def test_device_loop_transition_level_parity_deep_trees():
    """Depth-10 trees through the device-driven loop: ~1000 loop bodies per transition without the host."""
    res = pu.run_case_on_gpu_and_oracle("nuts_deep_d100", callback="torch-graph")
    pu.assert_parity(res, rtol=RTOL)


@pytest.mark.parametrize("name", gc.CASE_NAMES)
def test_callback_transition_level_parity(name):
    case, _ = gc.load(name)
    n = None if int(case["tune"]) + int(case["draws"]) <= 120 else 60     # bound the launch count of the long fixtures
    res = pu.run_case_on_gpu_and_oracle(name, n_trans=n, callback="torch")
    print(pu.parity_report(res))
    pu.assert_parity(res, rtol=RTOL)


@pytest.mark.parametrize("name", ["nuts_b1_d10", "nuts_static_d100"])
def test_callback_run_level_parity(name):
    """Whole runs chained on the device: every chain is at a different point of a different transition in every launch
    (the state machines are not in lock step), and the result still equals the reference's trace."""
    res = pu.run_case_on_gpu_and_oracle(name, chained=True, callback="torch")
    print(pu.parity_report(res))
    pu.assert_parity(res, rtol=RTOL)
    _, ref = gc.load(name)
    np.testing.assert_allclose(res.gpu_trace, ref["trace"], rtol=RTOL, atol=1e-12)


def test_callback_hmc_config1_statistics():
    """BASELINE config 1 through callback mode (HMC, 4 chains x 1000 transitions): all integer statistics exact."""
    res = pu.run_case_on_gpu_and_oracle("hmc_cfg1_d10", chained=True, callback="torch")
    _, ref = gc.load("hmc_cfg1_d10")
    assert int(res.gpu_stats["n_steps"].sum()) == int(ref["stat_n_steps"].sum()) == 5389
    # 1000 chained transitions with adaptation feedback: rounding differences of the logp summation compound (the
    # fused-kernel test allows 1e-6; here a near-zero coordinate needs the absolute term)
    pu.assert_parity(res, rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("mode", ["torch-graph", "torch-replay"])
@pytest.mark.parametrize("name", ["nuts_diag_d37", "nuts_funnel_d10", "hmc_static_d50"])
def test_cuda_graph_equals_eager(name, mode):
    """The device-driven loop (ONE graph launch: WHILE n_running > 0 { callback; advance } as a conditional node,
    lmc_callback_loop_*) and the replayed graphs of 8 iterations == the eager host loop, bit for bit."""
    eager = pu.run_case_on_gpu_and_oracle(name, n_trans=40, chained=True, callback="torch")
    graph = pu.run_case_on_gpu_and_oracle(name, n_trans=40, chained=True, callback=mode)
    assert np.array_equal(eager.gpu_trace, graph.gpu_trace)
    for k in eager.gpu_stats:
        assert np.array_equal(eager.gpu_stats[k], graph.gpu_stats[k], equal_nan=True), k
    assert np.array_equal(eager.gpu_adapt, graph.gpu_adapt)


@pytest.mark.parametrize("name", ["nuts_b1_d10", "hmc_static_d50"])
def test_numpy_callable_matches_reference(name):
    """The reference's own callback contract -- a per-chain NumPy callable -- driven through callback mode."""
    res = pu.run_case_on_gpu_and_oracle(name, n_trans=12, chained=True, callback="numpy")
    pu.assert_parity(res, rtol=RTOL)


def test_callback_and_fused_modes_agree():
    """Same chains, same randomness: the fused kernel and the callback state machine share the tree code; the only
    difference is the summation order inside logp, so decisions are identical and floats agree to 1e-9
    (short run: chained adaptive chains amplify last-bit differences, see tests/test_gpu_parity.py)"""
    a = pu.run_case_on_gpu_and_oracle("nuts_b1_d10", chained=True)
    b = pu.run_case_on_gpu_and_oracle("nuts_b1_d10", chained=True, callback="torch")
    for k in ("depth", "tree_size", "diverging"):
        assert np.array_equal(a.gpu_stats[k], b.gpu_stats[k]), k
    np.testing.assert_allclose(a.gpu_trace, b.gpu_trace, rtol=1e-9, atol=1e-12)


def test_sample_api_with_torch_callback_and_philox():
    """sample() with a TorchBatched callback built from a log density by autograd, in-kernel Philox randomness:
    shapes / dtypes of the reference contract and correct moments; chains finish at different launches."""
    import torch
    import littlemcmc_b200 as lmc
    D, chains = 6, 256
    sigma = torch.linspace(0.5, 2.0, D, dtype=torch.float64, device="cuda:0")
    target = lmc.targets.TorchBatched.from_logp(lambda q: -0.5 * ((q / sigma) ** 2).sum(1))
    step = lmc.NUTS(target, D, potential=lmc.QuadPotentialDiagAdapt(D, np.zeros(D), np.ones(D), 10))
    trace, stats = lmc.sample(target, D, draws=150, tune=150, step=step, chains=chains, start=np.zeros(D),
                              random_seed=7, progressbar=False)
    assert trace.shape == (chains, 150, D) and trace.dtype == np.float64
    for name, dtype in step.stats_dtypes[0].items():
        assert stats[name].shape == (chains, 150, 1) and stats[name].dtype == dtype, name
    np.testing.assert_allclose(trace.std((0, 1)), sigma.cpu().numpy(), rtol=0.05)
    assert abs(trace.mean()) < 0.05
    assert not stats["tune"].any() and stats["tree_size"].min() >= 1


def test_sample_api_with_reference_style_numpy_callback():
    """The reference's test model (tests/test_utils.py:19-28): 1-D normal whose logp comes back array-shaped."""
    import scipy.stats as sps
    import littlemcmc_b200 as lmc

    def logp_dlogp_func(x, loc=0, scale=1):
        return np.log(sps.norm.pdf(x, loc=loc, scale=scale)), -(x - loc) / scale

    for method in (lmc.NUTS, lmc.HamiltonianMC):
        step = method(logp_dlogp_func=logp_dlogp_func, model_ndim=1)
        trace, stats = lmc.sample(logp_dlogp_func, model_ndim=1, step=step, draws=40, tune=40, chains=3,
                                  start=np.zeros(1), random_seed=3)
        assert trace.shape == (3, 40, 1) and np.var(trace) > 0
        assert set(stats) == set(step.stats_dtypes[0])


def test_philox_callback_equals_tape_dump():
    """Callback mode with in-kernel Philox == callback mode fed the dump of the same streams (lmc_rng_fill)."""
    import torch
    from littlemcmc_b200 import _lib as L, engine
    case, _ = gc.load("nuts_diag_d37")
    D, Cn, T = int(case["ndim"]), 5, 20
    cb = pu.torch_callback(case)
    params = pu.gpu_params(case)
    params.update(max_treedepth=6, early_max_treedepth=5)
    seeds = engine.seeds_tensor(np.arange(Cn) * 104729 + 11, "cuda:0")
    outs = []
    for mode in ("philox", "tape"):
        ch = pu.gpu_chains(case, Cn)
        kw = dict(n_trans=T, iter0=0, n_tune=12, params=params)
        if mode == "philox":
            tr, st = engine.run_transitions_callback(L.KIND_NUTS, ch, cb, seeds=seeds, **kw)
        else:
            tr, st = engine.run_transitions_callback(L.KIND_NUTS, ch, cb, tapes=engine.rng_fill(seeds, D, 0, T, 2 ** 6 + 20), **kw)
        torch.cuda.synchronize()
        assert int(ch.status.abs().sum()) == 0
        outs.append((tr.cpu().numpy(), st.cpu().numpy()))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])


def test_callback_bad_initial_energy():
    import littlemcmc_b200 as lmc
    import torch
    target = lmc.targets.TorchBatched(lambda q: (-0.5 * (q * q).sum(1), -q))
    with pytest.raises(ValueError, match="Bad initial energy"):
        lmc.sample(target, model_ndim=4, draws=2, tune=2, chains=2, start=np.full(4, np.inf))
