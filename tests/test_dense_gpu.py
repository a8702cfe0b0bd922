"""Dense-mass mode (QuadPotentialFull / FullInv / FullAdapt, reference
quadpotential.py:390-615) on the GPU against the CPU oracle, which is pinned to the reference on the dense fixtures
(tests/test_oracle_golden.py::test_dense_oracle_matches_reference).

Bar (transition-level, as tests/test_gpu_parity.py): integer / boolean statistics and the count of uniforms consumed
exact; float64 quantities to RTOL = 1e-9.  Licence to differ: summation order of the matrix-vector products
(BLAS dgemv on the CPU, a warp-per-row reduction / cuBLAS on the GPU), the Cholesky / triangular-solve libraries,
and velocity(p + dt g) formed as velocity(p) + dt velocity(g).
"""
import numpy as np
import pytest

from tests import dense_utils as du
from tests import golden_cases as gc
from tests import parity_utils as pu

pytestmark = pytest.mark.gpu

RTOL = 1e-9


@pytest.mark.parametrize("name", gc.DENSE_CASE_NAMES)
def test_dense_transition_level_parity(name):
    from littlemcmc_b200 import _lib as L
    case, ref = gc.load(name)
    ora = du.oracle_run_dense(case)
    np.testing.assert_allclose(ora["post"]["q"], ref["trace"], rtol=1e-10)          # the oracle side IS the reference
    q, ad, st, pots, status = du.gpu_run_dense_transitionwise(case, ora)
    assert (status == 0).all()
    table = pu.NUTS_STATS if str(case["kind"]) == "nuts" else pu.HMC_STATS
    assert np.array_equal(st[:, :, L.STAT_N_UNIFORMS], ora["tapes"][2]), "uniform consumption differs"
    for k, v in ora["stats"].items():
        g, v = st[:, :, table[k]], np.asarray(v, dtype="d")
        if k in pu.EXACT:
            assert np.array_equal(g, v), k
        else:
            np.testing.assert_allclose(g, v, rtol=RTOL, atol=1e-12, err_msg=k)
    np.testing.assert_allclose(q, ora["post"]["q"], rtol=RTOL, atol=1e-12, err_msg="trace")
    np.testing.assert_allclose(ad, ora["post"]["adapt"], rtol=RTOL, atol=1e-12, err_msg="step-size state")
    if case["pot"] == "fulladapt":
        Cn, T = q.shape[:2]
        for t in range(T):
            for c in range(Cn):
                want = ora["post"]["pot"][c][t]
                for key in ("cov", "chol", "mean_fg", "raw_fg", "mean_bg", "raw_bg", "n_fg", "n_bg"):
                    np.testing.assert_allclose(pots[t][key][c], want[key], rtol=RTOL, atol=1e-12,
                                               err_msg="%s (chain %d, transition %d)" % (key, c, t))
                for key in ("n_samples", "previous_update", "window"):
                    assert int(pots[t][key][c]) == int(want[key]), (key, c, t)


@pytest.mark.parametrize("D,nrhs", [(9, 2), (64, 1), (257, 2), (1000, 2)])
def test_dense_matvec_kernel(D, nrhs):
    """lmc_dense_matvec against torch's float64 bmm, with and without a chain index list, odd and even n."""
    import ctypes as C
    import torch
    from littlemcmc_b200 import _lib as L
    lib = L.load()
    dev = torch.device("cuda", 0)
    Cn, lda, ld = 7, D + (D & 1), D + (D & 1)
    gen = torch.Generator(device=dev).manual_seed(D)
    A = torch.randn(Cn, D, lda, dtype=torch.float64, device=dev, generator=gen)
    x = torch.randn(Cn, nrhs, ld, dtype=torch.float64, device=dev, generator=gen)
    want = torch.einsum("cij,crj->cri", A[:, :, :D], x[:, :, :D])
    p = lambda t: None if t is None else C.c_void_p(t.data_ptr())  # noqa: E731
    for idx in (None, torch.tensor([5, 0, 3], dtype=torch.int32, device=dev)):
        y = torch.full((Cn, nrhs, ld), float("nan"), dtype=torch.float64, device=dev)
        n_idx = Cn if idx is None else idx.numel()
        L.check(lib.lmc_dense_matvec(p(idx), n_idx, p(A), D * lda, lda, D, ld, p(x), p(y), nrhs, None), "matvec")
        torch.cuda.synchronize()
        rows = list(range(Cn)) if idx is None else idx.tolist()
        np.testing.assert_allclose(y[rows][:, :, :D].cpu().numpy(), want[rows].cpu().numpy(), rtol=1e-12, atol=1e-12)
        others = [c for c in range(Cn) if c not in rows]
        assert torch.isnan(y[others]).all()                    # unlisted chains are not touched
    # one matrix shared by all chains (chain_stride 0)
    y = torch.zeros(Cn, nrhs, ld, dtype=torch.float64, device=dev)
    L.check(lib.lmc_dense_matvec(None, Cn, p(A[2]), 0, lda, D, ld, p(x), p(y), nrhs, None), "matvec")
    want0 = torch.einsum("ij,crj->cri", A[2, :, :D], x[:, :, :D])
    np.testing.assert_allclose(y[:, :, :D].cpu().numpy(), want0.cpu().numpy(), rtol=1e-12, atol=1e-12)


def test_dense_cov_update_kernel_matches_weighted_covariance():
    """lmc_dense_cov_update against the oracle's _WeightedCovariance (reference quadpotential.py:573-621), several
    samples in a row for a subset of chains."""
    import torch
    import littlemcmc_b200 as lmc
    from littlemcmc_b200 import engine
    from oracle import lmc_oracle as orc
    D, Cn = 11, 4
    rs = np.random.RandomState(3)
    mean0, cov0 = rs.randn(D), np.eye(D) * 2.0
    pot = lmc.QuadPotentialFullAdapt(D, mean0, cov0, 5, adaptation_window=4, adaptation_window_multiplier=2)
    ch = engine.DeviceChains(Cn, D, "cuda:0")
    pot._bind(ch)
    refs = [orc.FullAdaptPotential(D, mean0, cov0, 5, adaptation_window=4, adaptation_window_multiplier=2)
            for _ in range(Cn)]
    for step in range(14):
        rows = [c for c in range(Cn) if (step + c) % 3 != 0]         # a different subset of chains every time
        xs = rs.randn(Cn, D)
        ch.q[:, :D] = torch.as_tensor(xs, device=ch.device)
        pot._update_rows(torch.as_tensor(rows, device=ch.device), ch.q)
        for c in rows:
            refs[c].update(xs[c], True)
        for c in range(Cn):
            np.testing.assert_allclose(pot._cov_all[c, :, :D].cpu().numpy(), refs[c].cov, rtol=1e-12, atol=1e-14)
            np.testing.assert_allclose(pot._chol_all[c].cpu().numpy(), refs[c].chol, rtol=1e-10, atol=1e-12)
            assert pot._n_samples_all[c] == refs[c].n_samples and pot._window_all[c] == refs[c].adaptation_window
            assert pot._previous_update_all[c] == refs[c].previous_update
            np.testing.assert_allclose(pot._nsamp[c].cpu().numpy(), [refs[c].fg.n_samples, refs[c].bg.n_samples])


def test_equal_dense_and_random_dense():
    """reference tests/test_quadpotential.py:67-87 (dense / inverse forms agree) and :104-122 (random() covariance)."""
    import littlemcmc_b200 as lmc
    np.random.seed(42)
    for _ in range(3):
        cov = np.random.rand(5, 5)
        cov += cov.T
        cov += 10 * np.eye(5)
        inv = np.linalg.inv(cov)
        x = np.random.randn(5)
        pots = [lmc.quad_potential(cov, False), lmc.quad_potential(inv, True)]
        assert isinstance(pots[0], lmc.QuadPotentialFullInv) and isinstance(pots[1], lmc.QuadPotentialFull)
        v = np.linalg.solve(cov, x)
        e = 0.5 * x.dot(v)
        for pot in pots:
            np.testing.assert_allclose(pot.velocity(x), v, rtol=1e-10)
            np.testing.assert_allclose(pot.energy(x), e, rtol=1e-10)
            v_out = np.empty(5)
            np.testing.assert_allclose(pot.velocity_energy(x, v_out), e, rtol=1e-10)
            np.testing.assert_allclose(v_out, v, rtol=1e-10)
        for pot in (lmc.QuadPotentialFull(cov), lmc.QuadPotentialFullInv(inv)):
            cov_ = np.cov(np.array([pot.random() for _ in range(1000)]).T)
            assert np.allclose(cov_, inv, atol=0.1)


def test_init_nuts_adapt_full_and_sampling_recovers_covariance():
    """reference sampling.py:588-597 (init='adapt_full' / 'jitter+adapt_full') and an end-to-end run: 64 chains with a
    per-chain adapted dense mass matrix recover a strongly correlated Gaussian."""
    import torch
    import littlemcmc_b200 as lmc
    D = 6
    prec = du.spd(D, 5, cond=50.0)
    cov = np.linalg.inv(prec)
    target = du.torch_dense_gaussian(prec, torch.device("cuda", 0))
    for init in ("adapt_full", "jitter+adapt_full"):
        start, step = lmc.init_nuts(logp_dlogp_func=target, model_ndim=D, init=init, random_seed=3)
        assert isinstance(start, np.ndarray) and start.shape == (D,)
        assert isinstance(step, lmc.NUTS) and isinstance(step.potential, lmc.QuadPotentialFullAdapt)
    trace, stats = lmc.sample(target, D, draws=150, tune=250, init="adapt_full", chains=64, random_seed=11)
    assert trace.shape == (64, 150, D) and stats["depth"].shape == (64, 150, 1)
    flat = trace.reshape(-1, D)
    np.testing.assert_allclose(np.cov(flat.T), cov, atol=0.12 * np.abs(cov).max())
    assert stats["diverging"].sum() == 0


def test_dense_integrator_matches_oracle_and_is_reversible():
    """`step.integrator.compute_state / step` (reference integration.py:52-121) with dense potentials: one leapfrog
    against the oracle, and the reference's reversibility test (tests/test_hmc.py:23-40) with a dense matrix."""
    import littlemcmc_b200 as lmc
    from oracle import lmc_oracle as orc
    D = 12
    prec = du.spd(D, 7)
    cov = np.linalg.inv(du.spd(D, 8))
    f = orc.dense_gaussian(prec)
    rs = np.random.RandomState(0)
    for pot, opot in ((lmc.QuadPotentialFull(cov), orc.FullPotential(cov)),
                      (lmc.QuadPotentialFullInv(np.linalg.inv(cov)), orc.FullInvPotential(np.linalg.inv(cov)))):
        step = lmc.HamiltonianMC(logp_dlogp_func=f, model_ndim=D, potential=pot)
        q0, p0 = rs.randn(D), rs.randn(D)
        start = step.integrator.compute_state(q0, p0)
        ostart = orc.compute_state(f, opot, q0, p0)
        np.testing.assert_allclose(start.v, ostart.v, rtol=1e-11)
        np.testing.assert_allclose(start.energy, ostart.energy, rtol=1e-11)
        for eps in (0.1, -0.07):
            got, want = step.integrator.step(eps, start), orc.leapfrog(f, opot, eps, ostart)
            for name in ("q", "p", "v", "q_grad"):
                np.testing.assert_allclose(getattr(got, name), getattr(want, name), rtol=1e-10, atol=1e-12, err_msg=name)
            np.testing.assert_allclose(got.energy, want.energy, rtol=1e-10)
        for epsilon in (0.01, 0.1):
            for n_steps in (1, 2, 3, 4, 20):
                state = start
                for _ in range(n_steps):
                    state = step.integrator.step(epsilon, state)
                for _ in range(n_steps):
                    state = step.integrator.step(-epsilon, state)
                np.testing.assert_allclose(state.q, start.q, rtol=1e-5)
                np.testing.assert_allclose(state.p, start.p, rtol=1e-5)


def test_full_adapt_object_api_like_the_reference_tests():
    """reference tests/test_quadpotential.py:158-223 driven through the object-level API (`random`, `update`,
    `raise_ok`, `_cov`, `_previous_update`, `_adaptation_window`): test_full_adapt_sample_p, _update_window,
    _adaptation_window, _not_invertible."""
    import littlemcmc_b200 as lmc
    # sample_p: momenta drawn with inverse-covariance m_inv have covariance m (5 sigma of the Wishart spread)
    np.random.seed(4566)
    m = np.array([[3.0, -2.0], [-2.0, 4.0]])
    m_inv = np.linalg.inv(m)
    var = np.array([[2 * m[0, 0], m[1, 0] * m[1, 0] + m[1, 1] * m[0, 0]],
                    [m[0, 1] * m[0, 1] + m[1, 1] * m[0, 0], 2 * m[1, 1]]])
    n_samples = 1000
    pot = lmc.QuadPotentialFullAdapt(2, np.zeros(2), m_inv, 1)
    sample_cov = np.cov([pot.random() for _ in range(n_samples)], rowvar=0)
    assert np.all(np.abs(m - sample_cov) < 5 * np.sqrt(var / n_samples))
    # update_window: the matrix is refreshed every 50th update only
    np.random.seed(1123)
    init_cov = np.array([[1.0, 0.02], [0.02, 0.8]])
    pot = lmc.QuadPotentialFullAdapt(2, np.zeros(2), init_cov, 1, update_window=50)
    for _ in range(49):
        pot.update(np.random.randn(2), None, True)
    assert np.allclose(pot._cov, init_cov)
    pot.update(np.random.randn(2), None, True)
    assert not np.allclose(pot._cov, init_cov)
    # adaptation_window: the foreground estimator is swapped after `window` updates and the window doubles
    np.random.seed(8978)
    window = 10
    pot = lmc.QuadPotentialFullAdapt(2, np.zeros(2), np.eye(2), 1, adaptation_window=window)
    for _ in range(window + 1):
        pot.update(np.random.randn(2), None, True)
    assert pot._previous_update == window
    assert pot._adaptation_window == window * pot._adaptation_window_multiplier
    # not invertible: identical samples give a singular covariance; the error surfaces through raise_ok
    pot = lmc.QuadPotentialFullAdapt(2, np.zeros(2), np.eye(2), 0, adaptation_window=window)
    for _ in range(window + 1):
        pot.update(np.ones(2), None, True)
    with pytest.raises(ValueError):
        pot.raise_ok(None)


def test_held_updates_give_the_same_chains():
    """engine.DenseRun holds chains that ask for potential.update until enough of them wait (LMC_NEED_HOLD, one batched
    Cholesky per group instead of one per straggler).  Every chain only ever uses its own matrix, so the draws must not
    depend on the batching policy: serve-at-once (fraction 0), the default quarter, and wait-for-everybody (fraction 1)
    take the same tree decisions and agree to rounding -- not bit for bit: the batched Cholesky picks its algorithm by
    batch size, so a factor differs in the last bits with the company it is computed in (the run is kept short, the
    chaotic feedback of adaptation amplifies such differences like any other, tests/test_gpu_parity.py)."""
    import torch
    import littlemcmc_b200 as lmc
    D = 24
    prec = du.spd(D, 3)
    P = torch.as_tensor(prec, device="cuda")

    def fn(q):
        g = -(q @ P)
        return 0.5 * (q * g).sum(1), g

    target = lmc.targets.TorchBatched(fn)
    out = []
    for frac in (0.0, 0.25, 1.0):
        pot = lmc.QuadPotentialFullAdapt(D, np.zeros(D), np.eye(D), 10, adaptation_window=20)
        pot._update_batch_fraction = frac
        step = lmc.NUTS(target, D, potential=pot, max_treedepth=6)
        tr, st = lmc.sample(target, D, draws=2, tune=6, step=step, chains=32, start=np.zeros(D),
                            random_seed=list(range(32)), discard_tuned_samples=False)
        out.append((tr, st))
    for tr, st in out[1:]:
        assert np.array_equal(st["tree_size"], out[0][1]["tree_size"])
        assert np.array_equal(st["depth"], out[0][1]["depth"])
        np.testing.assert_allclose(tr, out[0][0], rtol=1e-8, atol=1e-10)
    assert out[0][1]["depth"].std() > 0, "the chains must not move in lock step for this test to mean anything"
