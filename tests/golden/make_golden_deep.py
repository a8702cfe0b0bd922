"""Generate the DEEP-TREE fixtures (tests/golden/nuts_deep_*.npz, nuts_early_gt_max_d20.npz) with the UNMODIFIED reference.

BASELINE.json's configurations run max_treedepth 10 and 12, but a well-tuned sampler rarely builds trees deeper
than 8, so round 1's fixtures never exercised stack levels 8..11, proposal slots 9..12 or the all-global part of the
kernels' scratch.  These cases force deep trees with a small FIXED step size (adapt_step_size=False) and
early_max_treedepth = max_treedepth, a handful of transitions each (<= 4095 leapfrogs per transition: seconds for
the reference).  One more case has early_max_treedepth (8, the reference default) ABOVE max_treedepth (5), which is
legal in the reference (nuts.py:205-208) and makes tuning trees deeper than the post-tuning cap.

Run in the build container only:   python tests/golden/make_golden_deep.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

from make_golden import lmc, run_reference, save  # noqa: E402  (asserts the reference comes from /root/reference)
from oracle.lmc_oracle import diag_gaussian, neal_funnel  # noqa: E402  (target densities only)

assert lmc.__file__.startswith("/root/reference"), lmc.__file__


def deep_case(name, D, target, tau, step_scale, max_depth, seeds, tune, draws, start, early=None, Emax=1000.0,
              pot_var=None):
    early = max_depth if early is None else early
    pot_var = np.ones(D) if pot_var is None else pot_var
    case = dict(kind="nuts", target=target, ndim=D, tau=tau, draws=draws, tune=tune, start=start, seeds=seeds,
                pot_adapt=0, pot_mean=np.zeros(D), pot_var=pot_var, pot_weight=0, max_treedepth=max_depth,
                early_max_treedepth=early, adapt_step_size=0, step_scale=step_scale, Emax=Emax)
    f = diag_gaussian(tau) if target == "diag_gaussian" else neal_funnel(D)
    out = run_reference(f, D, "nuts", draws, tune, start, seeds, dict(adapt=False, var=pot_var),
                        dict(max_treedepth=max_depth, early_max_treedepth=early, adapt_step_size=False,
                             step_scale=step_scale, Emax=Emax))
    print("   %s depths: %s  tree sizes: %s  diverging: %d" % (
        name, out["stat_depth"].astype(int).tolist(), out["stat_tree_size"].astype(int).tolist(),
        int(out["stat_diverging"].sum())))
    save(name, case, out)
    return out


def main():
    rs = np.random.RandomState(2024)
    # ---- headline shape: D=1000 diagonal Gaussian, depth up to 12 ------------------------------------------------------
    D = 1000
    sigma = 10 ** np.linspace(-0.5, 0.5, D)
    out = deep_case("nuts_deep_d1000", D, "diag_gaussian", 1 / sigma**2, step_scale=0.011, max_depth=12, seeds=[71, 72],
                    tune=2, draws=2, start=0.3 * rs.randn(D))
    assert out["stat_depth"].max() >= 12
    # ---- cfg2 shape: D=100, one warp per chain, depth up to 10 -------------------------------------------------------------
    D = 100
    sigma = 10 ** np.linspace(-0.5, 0.5, D)
    out = deep_case("nuts_deep_d100", D, "diag_gaussian", 1 / sigma**2, step_scale=0.036, max_depth=10, seeds=[73, 74, 75],
                    tune=2, draws=3, start=0.3 * rs.randn(D))
    assert out["stat_depth"].max() >= 10
    # ---- cfg4's exact shape: D=50 funnel, max_treedepth 12 ---------------------------------------------------------------------
    D = 50
    out = deep_case("nuts_deep_funnel_d50", D, "funnel", np.zeros(0), step_scale=0.0017, max_depth=12,
                    seeds=[76, 77, 78, 79], tune=2, draws=2, start=np.concatenate([[0.5], 0.5 * rs.randn(D - 1)]))
    assert out["stat_depth"].max() >= 12
    # ---- D=10 funnel, Emax small enough that the energy error of a long trajectory trips it (divergence at depth 5..8) -----------------------------------------------------------------
    D = 10
    out = deep_case("nuts_deep_funnel_d10", D, "funnel", np.zeros(0), step_scale=0.02, max_depth=12,
                    seeds=[81, 82, 83, 84], tune=3, draws=3, start=np.concatenate([[-1.0], 0.3 * rs.randn(D - 1)]),
                    Emax=0.0005)
    assert out["stat_depth"].max() >= 8 and out["stat_diverging"].sum() >= 8     # divergences inside deep subtrees
    # ---- early_max_treedepth (8) > max_treedepth (5): tuning trees deeper than the post-tuning cap -------------------------------
    D = 20
    sigma = 10 ** np.linspace(-0.5, 0.5, D)
    out = deep_case("nuts_early_gt_max_d20", D, "diag_gaussian", 1 / sigma**2, step_scale=0.05, max_depth=5,
                    early=8, seeds=[91, 92, 93], tune=5, draws=5, start=0.3 * rs.randn(D))
    assert out["stat_depth"][:, :5].max() >= 7 and out["stat_depth"][:, 5:].max() <= 5
    # ---- dense potential (QuadPotentialFull), depth up to 10: the dense state machine's deeper stack levels -------------------
    import make_golden_dense as mgd
    from tests.dense_utils import spd
    D = 12
    prec = spd(D, 1)
    cov = np.linalg.inv(prec) * (1 + 0.2 * np.cos(np.arange(D)))[:, None] * (1 + 0.2 * np.cos(np.arange(D)))[None, :]
    case = dict(kind="nuts", target="dense_gaussian", ndim=D, prec=prec, draws=2, tune=2, start=np.full(D, 0.1),
                seeds=[95, 96], pot="full", pot_matrix=cov, max_treedepth=10, early_max_treedepth=10,
                adapt_step_size=0, step_scale=0.012)
    out = mgd.run_reference(case, dict(max_treedepth=10, early_max_treedepth=10, adapt_step_size=False, step_scale=0.012))
    print("   dense_full_nuts_deep_d12 depths:", out["stat_depth"].astype(int).tolist())
    assert out["stat_depth"].max() >= 9
    mgd.save("dense_full_nuts_deep_d12", case, out)


if __name__ == "__main__":
    main()
