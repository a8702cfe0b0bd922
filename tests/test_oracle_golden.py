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


def test_philox_restatement_known_answers():
    """Random123's known-answer vectors for philox4x32_10 (kat_vectors: counter, key -> output)."""
    from oracle import philox
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = philox.philox4x32_10(ctr, key)
        assert tuple(int(np.asarray(x).reshape(-1)[0]) for x in got) == want
    u = philox.uniforms(12345, 7, 1000)
    assert ((u > 0) & (u < 1)).all() and abs(u.mean() - 0.5) < 0.03
    z = philox.normals(12345, 7, 1001)
    assert z.shape == (1001,) and abs(z.mean()) < 0.15 and abs(z.std() - 1) < 0.1
