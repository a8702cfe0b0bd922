"""Pin the oracle (oracle/lmc_oracle.py) to the unmodified reference's outputs (tests/golden/*.npz).

Same legacy-MT19937 seeds on both sides, so this also pins the order in which randomness is consumed.
Tolerances: integer/bool statistics exact; continuous quantities rtol 1e-12 (BLAS ddot summation order is
the only licence to differ; in practice almost everything is bit-identical).
"""
import numpy as np
import pytest

from tests import golden_cases as gc
from oracle import lmc_oracle as orc

RTOL = 1e-12


@pytest.mark.parametrize("name", gc.CASE_NAMES)
def test_oracle_matches_reference(name):
    case, ref = gc.load(name)
    trace, stats = gc.run_oracle(case)
    assert trace.shape == ref["trace"].shape
    for k, v in stats.items():
        r = ref["stat_" + k]
        if k in gc.EXACT_STATS:
            assert np.array_equal(v, r), k
        else:
            np.testing.assert_allclose(v, r, rtol=RTOL, atol=1e-300, err_msg=k)
    np.testing.assert_allclose(trace, ref["trace"], rtol=RTOL, atol=1e-300)


@pytest.mark.parametrize("name", gc.CASE_NAMES)
def test_tape_replay_is_identical(name):
    """Recording the consumed randomness and replaying it through TapeRNG gives bit-identical chains
    (this is the mechanism the GPU parity tests rely on)."""
    case, _ = gc.load(name)
    if int(case["draws"]) + int(case["tune"]) > 300:
        pytest.skip("long case; covered by the shorter ones")
    trace, stats, (normals, uniforms, n_u) = gc.run_oracle(case, record=True)
    for c in range(trace.shape[0]):
        smp = orc.Sampler(gc.target_fn(case)(), int(case["ndim"]),
                          orc.DiagPotential(int(case["ndim"]), **gc.potential_kw(case)),
                          kind=str(case["kind"]), **gc.sampler_kw(case))
        tr, st = orc.sample_chain(smp, case["start"], int(case["draws"]), int(case["tune"]),
                                  orc.TapeRNG(normals[c], uniforms[c]))
        assert np.array_equal(tr, trace[c])
        for k in st:
            assert np.array_equal(st[k], stats[k][c]), k


def test_final_adaptation_state_matches_reference():
    case, ref = gc.load("nuts_diag_d37")
    seeds = [int(s) for s in case["seeds"]]
    for c, seed in enumerate(seeds):
        smp = orc.Sampler(gc.target_fn(case)(), int(case["ndim"]),
                          orc.DiagPotential(int(case["ndim"]), **gc.potential_kw(case)), kind="nuts",
                          **gc.sampler_kw(case))
        orc.sample_chain(smp, case["start"], int(case["draws"]), int(case["tune"]), np.random.RandomState(seed))
        np.testing.assert_allclose(smp.pot.var, ref["final_var"][c], rtol=RTOL)
        sa = smp.step_adapt
        np.testing.assert_allclose([sa.log_step, sa.log_bar, sa.hbar, sa.count, sa.mu],
                                   ref["final_step_adapt"][c], rtol=RTOL)
        assert smp.pot.n_samples == int(ref["final_n_samples"][c]) == int(case["tune"])


@pytest.mark.parametrize("name", gc.DENSE_CASE_NAMES)
def test_dense_oracle_matches_reference(name):
    """Dense potentials (reference quadpotential.py:390-615): QuadPotentialFull / FullInv / FullAdapt.  The matrix
    products run through BLAS on both sides (dgemv / dtrsv / dpotrf), hence the slightly wider float tolerance."""
    case, ref = gc.load(name)
    trace, stats, pots = gc.run_oracle_dense(case)
    assert trace.shape == ref["trace"].shape
    for k, v in stats.items():
        r = ref["stat_" + k]
        if k in gc.EXACT_STATS:
            assert np.array_equal(v, r), k
        else:
            np.testing.assert_allclose(v, r, rtol=1e-10, atol=1e-300, err_msg=k)
    np.testing.assert_allclose(trace, ref["trace"], rtol=1e-10, atol=1e-300)
    for c, (pot, smp) in enumerate(pots):
        sa = smp.step_adapt
        np.testing.assert_allclose([sa.log_step, sa.log_bar, sa.hbar, sa.count, sa.mu], ref["final_step_adapt"][c],
                                   rtol=1e-10)
        if case["pot"] == "fulladapt":
            np.testing.assert_allclose(pot.cov, ref["final_cov"][c], rtol=1e-10)
            assert pot.n_samples == int(ref["final_n_samples"][c]) == int(case["tune"])
            assert pot.adaptation_window == int(ref["final_window"][c])


def test_oracle_weighted_covariance_is_the_sample_covariance(ndim=10, seed=5432):
    """reference tests/test_quadpotential.py:119-156 on the oracle's `_WeightedCovariance` restatement."""
    np.random.seed(seed)
    L = np.random.randn(ndim, ndim)
    L[np.triu_indices_from(L, 1)] = 0.0
    L[np.diag_indices_from(L)] = np.exp(L[np.diag_indices_from(L)])
    cov = np.dot(L, L.T)
    mean = np.random.randn(ndim)
    samples = np.random.multivariate_normal(mean, cov, size=100)
    est = orc.WelfordCov(ndim)
    for s in samples:
        est.add_sample(s, 1)
    assert np.allclose(est.mean, samples.mean(0)) and np.allclose(est.current_covariance(), np.cov(samples, rowvar=0))
    est2 = orc.WelfordCov(ndim, samples[:10].mean(0), np.cov(samples[:10], rowvar=0, bias=True), 10)
    for s in samples[10:]:
        est2.add_sample(s, 1)
    assert np.allclose(est2.mean, samples.mean(0)) and np.allclose(est2.current_covariance(), np.cov(samples, rowvar=0))
