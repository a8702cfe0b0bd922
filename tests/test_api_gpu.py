"""The reference's own test-suite (tests/test_hmc.py, test_sampling.py, test_quadpotential.py), restated against this
package's drop-in API on the GPU.  Each test cites the reference test it mirrors."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _lmc():
    import littlemcmc_b200 as lmc
    return lmc


def test_leapfrog_reversible():
    """reference tests/test_hmc.py:23-40: n steps forward then n steps back return to the start (rtol 1e-5)."""
    lmc = _lmc()
    np.random.seed(42)
    for D in (1, 7, 100, 1000):
        target = lmc.targets.DiagGaussian(sigma=np.random.rand(D) + 0.5)
        scaling = np.random.rand(D) + 0.1
        step = lmc.HamiltonianMC(logp_dlogp_func=target, model_ndim=D, scaling=scaling)
        p = step.potential.random()
        q = np.random.randn(D)
        start = step.integrator.compute_state(q, p)
        for epsilon in [0.01, 0.1]:
            for n_steps in [1, 2, 3, 4, 20]:
                state = start
                for _ in range(n_steps):
                    state = step.integrator.step(epsilon, state)
                for _ in range(n_steps):
                    state = step.integrator.step(-epsilon, state)
                np.testing.assert_allclose(state.q, start.q, rtol=1e-5, atol=1e-9)
                np.testing.assert_allclose(state.p, start.p, rtol=1e-5, atol=1e-9)


def test_integrator_matches_numpy_arithmetic():
    """compute_state / step against the formulas of integration.py:52-121 written in NumPy, bit for bit on q, p, v, g."""
    lmc = _lmc()
    rs = np.random.RandomState(3)
    for D in (5, 64, 333, 1000, 2049):
        tau = rs.rand(D) + 0.2
        var = rs.rand(D) + 0.3
        target = lmc.targets.DiagGaussian(tau=tau)
        step = lmc.NUTS(target, D, potential=lmc.QuadPotentialDiag(var))
        q, p = rs.randn(4, D), rs.randn(4, D)
        s0 = step.integrator.compute_state(q, p)
        g0 = -(tau * q)
        assert np.array_equal(s0.q_grad, g0) and np.array_equal(s0.v, var * p)
        e0 = 0.5 * np.sum(p * (var * p), 1) - 0.5 * np.sum(q * g0, 1)
        np.testing.assert_allclose(s0.energy, e0, rtol=1e-13)
        eps = np.array([0.1, -0.2, 0.03, 0.5])
        s1 = step.integrator.step(eps, s0)
        dt = 0.5 * eps[:, None]
        ph = p + dt * g0
        qn = q + eps[:, None] * (var * ph)
        gn = -(tau * qn)
        pn = ph + dt * gn
        assert np.array_equal(s1.q, qn) and np.array_equal(s1.p, pn) and np.array_equal(s1.q_grad, gn)
        assert np.array_equal(s1.v, var * pn)
        np.testing.assert_allclose(s1.energy, 0.5 * np.sum(pn * (var * pn), 1) - 0.5 * np.sum(qn * gn, 1), rtol=1e-13)


def test_nuts_tuning():
    """reference tests/test_hmc.py:43-54."""
    lmc = _lmc()
    target = lmc.targets.StdNormal(1)
    step = lmc.NUTS(logp_dlogp_func=target, model_ndim=1)
    lmc.sample(target, model_ndim=1, draws=5, tune=5, step=step, chains=1, start=np.zeros(1), progressbar=False)
    assert not step.tune


def test_init_nuts():
    """reference tests/test_sampling.py:21-34: all four initialisers."""
    lmc = _lmc()
    target = lmc.targets.StdNormal(1)
    for init in ("auto", "adapt_diag", "jitter+adapt_diag", "adapt_full", "jitter+adapt_full"):
        start, step = lmc.init_nuts(logp_dlogp_func=target, model_ndim=1, init=init)
        assert isinstance(start, np.ndarray) and len(start) == 1
        assert isinstance(step, lmc.NUTS)
    with pytest.raises(ValueError):
        lmc.init_nuts(logp_dlogp_func=target, model_ndim=1, init="no such initialiser")


@pytest.mark.parametrize("method", ["HamiltonianMC", "NUTS"])
def test_sampling_runs_shapes_and_dtypes(method):
    """reference tests/test_sampling.py:37-88: the output-format contract."""
    lmc = _lmc()
    ndim, draws, tune, chains = 3, 2, 1, 2
    target = lmc.targets.StdNormal(ndim)
    step = getattr(lmc, method)(logp_dlogp_func=target, model_ndim=ndim)
    trace, stats = lmc.sample(target, model_ndim=ndim, step=step, draws=draws, tune=tune, chains=chains, cores=1,
                              start=np.zeros(ndim))
    assert trace.shape == (chains, draws, ndim) and trace.dtype == np.float64
    assert set(stats) == set(step.stats_dtypes[0])
    for name, dtype in step.stats_dtypes[0].items():
        assert stats[name].shape == (chains, draws, 1), name
        assert stats[name].dtype == dtype, name


def test_multichain_sampling_runs():
    """reference tests/test_sampling.py:91-100 (chains=4, cores=4 there: four processes; here: one launch)."""
    lmc = _lmc()
    target = lmc.targets.StdNormal(1)
    trace, stats = lmc.sample(target, model_ndim=1, draws=1, tune=1, chains=4, cores=4, progressbar=None)
    assert trace.shape == (4, 1, 1)


@pytest.mark.parametrize("method", ["HamiltonianMC", "NUTS"])
def test_recovers_1d_normal(method):
    """reference tests/test_sampling.py:103-130 (their tolerance is atol=1; with 64 chains we can afford 0.1)."""
    lmc = _lmc()
    target = lmc.targets.StdNormal(1)
    step = getattr(lmc, method)(logp_dlogp_func=target, model_ndim=1)
    trace, stats = lmc.sample(target, model_ndim=1, step=step, draws=1000, tune=1000, chains=64, start=np.zeros(1),
                              random_seed=1)
    assert np.allclose(np.mean(trace), 0, atol=0.1)
    assert np.allclose(np.std(trace), 1, atol=0.1)


def test_samples_not_all_same():
    """reference tests/test_sampling.py:133-140."""
    lmc = _lmc()
    target = lmc.targets.StdNormal(1)
    trace, _ = lmc.sample(target, model_ndim=1, draws=20, tune=20, chains=1, progressbar=None)
    assert np.var(trace) > 0


def test_reset_tuning():
    """reference tests/test_sampling.py:143-161."""
    lmc = _lmc()
    target = lmc.targets.StdNormal(1)
    tune, chains = 50, 2
    start, step = lmc.init_nuts(logp_dlogp_func=target, model_ndim=1)
    lmc.sample(target, model_ndim=1, draws=2, tune=tune, chains=chains, step=step, start=start, cores=1)
    assert step.potential._n_samples == tune
    assert step.step_adapt._count == tune + 1


def test_discard_and_blocks_are_consistent():
    """discard_tuned_samples slicing (sampling.py:473-476) and multi-block launches give the same chains."""
    lmc = _lmc()
    D = 11
    target = lmc.targets.DiagGaussian(sigma=np.linspace(0.5, 2, D))
    kw = dict(model_ndim=D, draws=30, tune=40, chains=6, start=np.full(D, 0.1), random_seed=[5, 6, 7, 8, 9, 10])
    t_all, s_all = lmc.sample(target, discard_tuned_samples=False, **kw)
    t_kept, s_kept = lmc.sample(target, discard_tuned_samples=True, block=7, **kw)
    assert t_all.shape == (6, 70, D) and t_kept.shape == (6, 30, D)
    assert np.array_equal(t_all[:, 40:], t_kept)
    for k in s_all:
        assert np.array_equal(s_all[k][:, 40:], s_kept[k]), k
    assert s_all["tune"][:, :40].all() and not s_all["tune"][:, 40:].any()


def test_bad_initial_energy_raises():
    """reference base_hmc.py:145-148."""
    lmc = _lmc()
    target = lmc.targets.DiagGaussian(sigma=np.ones(4))
    with pytest.raises(ValueError, match="Bad initial energy"):
        lmc.sample(target, model_ndim=4, draws=2, tune=2, chains=2, start=np.full(4, np.inf))


# ---- quadpotential (reference tests/test_quadpotential.py) -------------------------------------------------------------
def test_elemwise_posdef():
    """reference tests/test_quadpotential.py:21-24."""
    lmc = _lmc()
    from littlemcmc_b200.quadpotential import PositiveDefiniteError
    with pytest.raises(PositiveDefiniteError):
        lmc.quad_potential(np.array([-1.0, 2.0]), True)


def test_elemwise_velocity_and_energy():
    """reference tests/test_quadpotential.py:38-64."""
    lmc = _lmc()
    scaling = np.array([1.0, 2.0, 3.0])
    x = np.ones(3)
    pot = lmc.quad_potential(scaling, True)
    np.testing.assert_allclose(pot.velocity(x), scaling)
    np.testing.assert_allclose(pot.energy(x), 0.5 * scaling.sum())
    pot = lmc.quad_potential(scaling, False)
    np.testing.assert_allclose(pot.velocity(x), 1.0 / scaling)
    np.testing.assert_allclose(pot.energy(x), 0.5 * (1.0 / scaling).sum())
    v = np.zeros(3)
    np.testing.assert_allclose(pot.velocity_energy(x, v), 0.5 * (1.0 / scaling).sum())
    np.testing.assert_allclose(v, 1.0 / scaling)


def test_random_diag():
    """reference tests/test_quadpotential.py:90-101."""
    lmc = _lmc()
    d = np.arange(10) + 1
    np.random.seed(42)
    for pot in (lmc.quad_potential(d, True), lmc.quad_potential(1.0 / d, False)):
        samples = np.array([pot.random() for _ in range(1000)])
        np.testing.assert_allclose(np.std(samples, 0), np.sqrt(1.0 / d), atol=0.1)


def test_philox_mode_equals_tape_mode():
    """In-kernel Philox == tape mode fed with the dump of the same stream (lmc_rng_fill), bit for bit."""
    import torch
    from littlemcmc_b200 import _lib as L, engine
    D, Cn, T = 37, 5, 25
    sigma = np.linspace(0.5, 3, D)
    tgt = engine.FusedTarget(L.TARGET_DIAG_GAUSSIAN, D, tau=1 / sigma**2)
    params = dict(adapt_mass=1, adapt_step_size=1, target_accept=0.8, gamma=0.05, k=0.75, t0=10, Emax=1000.0,
                  max_treedepth=6, early_max_treedepth=5)
    seeds = engine.seeds_tensor(np.arange(Cn) * 7919 + 3, "cuda:0")
    outs = []
    for mode in ("philox", "tape"):
        ch = engine.DeviceChains(Cn, D, "cuda:0")
        ch.reset_potential(np.ones(D), np.zeros(D), 10.0, 101)
        ch.reset_step_adapt(0.25 / D**0.25)
        ch.set_position(np.full(D, 0.3))
        if mode == "philox":
            tr, st = engine.run_transitions(L.KIND_NUTS, ch, tgt, n_trans=T, iter0=0, n_tune=15, params=params, seeds=seeds)
        else:
            normals, uniforms = engine.rng_fill(seeds, D, 0, T, 2 ** 6 + 2 * 6 + 8)
            assert float(uniforms.min()) > 0 and float(uniforms.max()) < 1
            tr, st = engine.run_transitions(L.KIND_NUTS, ch, tgt, n_trans=T, iter0=0, n_tune=15, params=params,
                                            tapes=(normals, uniforms))
        torch.cuda.synchronize()
        assert int(ch.status.abs().sum()) == 0
        outs.append((tr.cpu().numpy(), st.cpu().numpy()))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    # and the normals really are standard normal
    normals, _ = engine.rng_fill(engine.seeds_tensor(np.arange(64), "cuda:0"), 1000, 0, 8, 4)
    z = normals.cpu().numpy().ravel()
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01


def test_per_draw_callback_and_interrupt():
    """reference sampling.py:303-308 (per-draw `callback(trace=, draw=)`) and :470-478 (KeyboardInterrupt returns the
    transitions sampled so far)."""
    lmc = _lmc()
    D, chains, tune, draws = 5, 3, 6, 9
    target = lmc.targets.DiagGaussian(sigma=np.linspace(0.5, 2, D))
    seen = []

    def cb(trace, draw):
        seen.append((draw.draw_idx, draw.chain, draw.tuning, draw.is_last, draw.point.copy(), draw.stats[0]["tree_size"]))

    kw = dict(model_ndim=D, draws=draws, tune=tune, chains=chains, start=np.full(D, 0.1), random_seed=[1, 2, 3],
              discard_tuned_samples=False, block=4)
    trace, stats = lmc.sample(target, callback=cb, **kw)
    assert len(seen) == chains * (tune + draws)
    assert [s[0] for s in seen] == sorted(s[0] for s in seen)            # draws arrive in order
    for idx, c, tuning, last, point, tree_size in seen:
        assert tuning == (idx < tune) and last == (idx == tune + draws - 1)
        assert np.array_equal(point, trace[c, idx])
        assert tree_size == stats["tree_size"][c, idx, 0]
    ref_trace, _ = lmc.sample(target, **kw)
    assert np.array_equal(trace, ref_trace)                              # the hook does not change the chains

    def stop_at_9(trace, draw):
        if draw.draw_idx == 9:
            raise KeyboardInterrupt

    part, pstats = lmc.sample(target, callback=stop_at_9, **kw)
    assert part.shape == (chains, 8, D) and pstats["depth"].shape == (chains, 8, 1)   # two finished blocks of 4
    assert np.array_equal(part, ref_trace[:, :8])


def test_step_rand_through_the_api():
    """reference base_hmc.py:154-155.  The hook sees the step sizes of all chains as one array (or, if it cannot take
    an array, one chain at a time) and the chain integrates with what it returns."""
    lmc = _lmc()
    D, chains = 7, 4
    target = lmc.targets.DiagGaussian(sigma=np.linspace(0.5, 2, D))
    calls = []

    def vec_hook(eps):
        calls.append(np.shape(eps))
        return eps * 0.5

    def scalar_hook(eps):
        return float(eps) * 0.5      # float() of a 4-vector raises: the per-chain fallback is used

    out = []
    for hook in (vec_hook, scalar_hook, None):
        pot = lmc.QuadPotentialDiagAdapt(D, np.zeros(D), np.ones(D), 10)
        step = lmc.NUTS(target, D, potential=pot, step_rand=hook)
        out.append(lmc.sample(target, D, draws=5, tune=10, step=step, chains=chains, start=np.full(D, 0.1),
                              random_seed=[4, 5, 6, 7], discard_tuned_samples=False))
    assert calls and all(c == (chains,) for c in calls) and len(calls) == 15
    assert np.array_equal(out[0][0], out[1][0])                          # both calling conventions agree
    assert not np.array_equal(out[0][0], out[2][0])                      # and the hook matters
    # halving every step size lengthens the trees
    assert out[0][1]["tree_size"].mean() > out[2][1]["tree_size"].mean()


def test_device_streams_are_the_documented_philox_function():
    """The per-chain random streams of the kernels (dumped by lmc_rng_fill; the sampler kernels consume exactly these:
    test_philox_mode_equals_tape_mode) are the function documented in lmc_device.cuh and restated in oracle/philox.py,
    which passes Random123's known-answer vectors on the CPU: uniforms bit for bit, Box-Muller normals to libm accuracy."""
    import torch
    from littlemcmc_b200 import engine
    from oracle import philox
    seeds = [0, 1, 12345, 2 ** 40 + 7, 2 ** 63 + 11]
    iter0, n_trans, D, u_stride = 3, 2, 37, 70
    st = engine.seeds_tensor(np.array(seeds, dtype=np.uint64), torch.device("cuda", 0))
    normals, uniforms = engine.rng_fill(st, D, iter0, n_trans, u_stride)
    normals, uniforms = normals.cpu().numpy(), uniforms.cpu().numpy()
    for c, seed in enumerate(seeds):
        for t in range(n_trans):
            assert np.array_equal(uniforms[c, t], philox.uniforms(seed, iter0 + t, u_stride)), (seed, t)
            np.testing.assert_allclose(normals[c, t], philox.normals(seed, iter0 + t, D), rtol=1e-13, atol=1e-14)


@pytest.mark.parametrize("D,chains,method", [(37, 300, "nuts"), (1000, 96, "nuts"), (600, 40, "nuts"), (20, 64, "hmc")])
@pytest.mark.parametrize("discard", [True, False])
def test_single_launch_host_trace_equals_blockwise_launches(D, chains, method, discard):
    """sample() with a host trace runs the whole run as ONE launch (the kernel skips discarded tuning draws and reports
    finished blocks of kept draws through device counters the copy engine follows): bit-identical to one launch per
    block, for the chunked warp kernel, the lean and register kernels and HMC, with a block that does not divide the run."""
    import littlemcmc_b200 as lmc
    sigma = 10 ** np.linspace(-0.5, 0.5, D)
    tgt = lmc.targets.DiagGaussian(sigma=sigma)
    out = []
    for single in (True, False):
        cls = lmc.NUTS if method == "nuts" else lmc.HamiltonianMC
        step = cls(tgt, D, potential=lmc.QuadPotentialDiagAdapt(D, np.zeros(D), np.ones(D), 10))
        out.append(lmc.sample(tgt, D, draws=23, tune=31, step=step, chains=chains, start=np.zeros(D), random_seed=5,
                              discard_tuned_samples=discard, block=7, single_launch=single, progressbar=False))
        assert step.iter_count == 54 and not step.tune
    (t1, s1), (t2, s2) = out
    assert t1.shape == (chains, 23 if discard else 54, D)
    assert np.array_equal(t1, t2)
    for k in s1:
        assert np.array_equal(s1[k], s2[k]), k
