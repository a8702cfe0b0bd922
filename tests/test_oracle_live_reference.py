"""The oracle against the UNMODIFIED reference run live, on cases no fixture holds (CPU tier only).

tests/golden/*.npz pin the oracle to outputs the reference produced once; this runs both side by side on freshly drawn
problem definitions (dimension, scales, start, seeds from a seeded generator), so the pin does not rest on seventeen
hand-picked cases alone.  The reference is the copy `oracle/build_ref.py` installs from /root/reference into oracle/_ref
(git-ignored; skipped where it is absent).  Same bar as test_oracle_golden.py: integer statistics exact, floats 1e-12.
"""
import os
import sys

import numpy as np
import pytest

from oracle import lmc_oracle as orc
from tests import golden_cases as gc

REF_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF_DIR, "littlemcmc", "__init__.py")),
                                reason="oracle/_ref (the installed reference) is not present")


def _reference():
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import littlemcmc as ref
    assert os.path.abspath(ref.__file__).startswith(REF_DIR), ref.__file__
    return ref


def _run_reference(ref, f, case):
    """One chain at a time through the reference's own sample() (sampling.py:35-222), float64 potentials."""
    D = int(case["ndim"])
    traces, stats_all = [], []
    for seed in case["seeds"]:
        pot = ref.QuadPotentialDiagAdapt(D, np.array(case["pot_mean"], dtype="d"), np.array(case["pot_var"], dtype="d"),
                                         float(case["pot_weight"]), dtype="float64")
        cls = ref.NUTS if case["kind"] == "nuts" else ref.HamiltonianMC
        step = cls(logp_dlogp_func=f, model_ndim=D, potential=pot, **gc.sampler_kw(case))
        trace, stats = ref.sample(f, D, draws=int(case["draws"]), tune=int(case["tune"]), step=step, chains=1, cores=1,
                                  start=np.array(case["start"], dtype="d"), progressbar=False,
                                  random_seed=[int(seed)], discard_tuned_samples=False)
        traces.append(trace[0])
        stats_all.append({k: np.asarray(v[0, :, 0], dtype="d") for k, v in stats.items()})
    return np.stack(traces), {k: np.stack([s[k] for s in stats_all]) for k in stats_all[0]}


def _fresh_cases():
    rs = np.random.RandomState(20261017)
    cases = []
    for kind, target, D, extra in (("nuts", "diag_gaussian", 7, dict(max_treedepth=10, early_max_treedepth=8)),
                                   ("nuts", "diag_gaussian", 23, dict(max_treedepth=4, early_max_treedepth=6)),
                                   ("nuts", "funnel", 5, dict(max_treedepth=6, early_max_treedepth=5, Emax=50.0)),
                                   ("hmc", "diag_gaussian", 4, dict(path_length=1.5, max_steps=64))):
        sigma = np.exp(rs.uniform(-1.0, 1.0, D))
        case = dict(kind=kind, target=target, ndim=D, tau=1.0 / sigma ** 2, draws=5, tune=14,
                    start=rs.normal(size=D) * 0.3, seeds=[int(s) for s in rs.randint(1, 2 ** 30, size=2)],
                    pot_adapt=1, pot_mean=np.zeros(D), pot_var=np.exp(rs.uniform(-0.3, 0.3, D)), pot_weight=10.0)
        case.update(extra)
        cases.append(case)
    return cases


@pytest.mark.parametrize("idx", range(4))
def test_oracle_matches_the_live_reference_on_fresh_cases(idx):
    ref = _reference()
    case = _fresh_cases()[idx]
    f = gc.target_fn(case)()
    trace_ref, stats_ref = _run_reference(ref, f, case)
    trace, stats = gc.run_oracle(case)
    assert trace.shape == trace_ref.shape
    np.testing.assert_allclose(trace, trace_ref, rtol=1e-12, atol=1e-300)
    for k, v in stats.items():
        if k not in stats_ref:
            continue
        if k in gc.EXACT_STATS:
            assert np.array_equal(np.asarray(v, dtype="d"), stats_ref[k]), k
        else:
            np.testing.assert_allclose(np.asarray(v, dtype="d"), stats_ref[k], rtol=1e-12, atol=1e-300, err_msg=k)
    size = stats.get("tree_size", stats.get("n_steps"))
    assert size is not None and float(np.sum(size)) > 0


@pytest.mark.parametrize("pot", ["full", "fullinv", "fulladapt"])
def test_dense_oracle_matches_the_live_reference_on_a_fresh_case(pot):
    """Dense potentials (reference quadpotential.py:390-615) on a freshly drawn 6-dimensional correlated Gaussian."""
    ref = _reference()
    rs = np.random.RandomState(777 + len(pot))
    n = 6
    A = rs.normal(size=(n, n))
    cov = A @ A.T / n + np.eye(n) * 0.5
    prec = np.linalg.inv(cov)
    case = dict(kind="nuts", target="dense_gaussian", ndim=n, prec=prec, draws=4, tune=16, start=rs.normal(size=n) * 0.2,
                seeds=[int(s) for s in rs.randint(1, 2 ** 30, size=2)], pot=pot,
                pot_matrix=(prec if pot == "fullinv" else np.eye(n) if pot == "fulladapt" else cov),
                pot_mean=np.zeros(n), pot_weight=5.0, adaptation_window=7, adaptation_window_multiplier=2.0,
                max_treedepth=6, early_max_treedepth=5)
    f = gc.target_fn(case)()
    traces, stats_all = [], []
    for seed in case["seeds"]:
        if pot == "full":
            p = ref.QuadPotentialFull(np.array(case["pot_matrix"]), dtype="float64")
        elif pot == "fullinv":
            p = ref.QuadPotentialFullInv(np.array(case["pot_matrix"]), dtype="float64")
        else:
            p = ref.QuadPotentialFullAdapt(n, np.zeros(n), np.array(case["pot_matrix"]), case["pot_weight"],
                                           adaptation_window=7, adaptation_window_multiplier=2.0, dtype="float64")
        step = ref.NUTS(logp_dlogp_func=f, model_ndim=n, potential=p, **gc.sampler_kw(case))
        trace, stats = ref.sample(f, n, draws=4, tune=16, step=step, chains=1, cores=1, start=np.array(case["start"]),
                                  progressbar=False, random_seed=[int(seed)], discard_tuned_samples=False)
        traces.append(trace[0])
        stats_all.append({k: np.asarray(v[0, :, 0], dtype="d") for k, v in stats.items()})
    trace_ref = np.stack(traces)
    trace, stats, _ = gc.run_oracle_dense(case)
    np.testing.assert_allclose(trace, trace_ref, rtol=1e-10, atol=1e-300)   # Cholesky / solves: LAPACK vs LAPACK
    for k in gc.EXACT_STATS:
        if k in stats:
            assert np.array_equal(np.asarray(stats[k], dtype="d"), np.stack([s[k] for s in stats_all])), k
