"""Generate tests/golden/*.npz by running the UNMODIFIED reference (eigenfoo/littlemcmc @ 2b5dd87).

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

The reference is imported from /root/reference with a stand-in for the missing `fastprogress` package
(tests/golden/_stubs).  Nothing in the reference is patched: randomness is its own process-global legacy
MT19937 stream seeded per chain through `random_seed=[...]` (reference sampling.py:131-134,496-497), which
NumPy keeps frozen across versions, so `numpy.random.RandomState(seed)` reproduces it anywhere.  Potentials
are built with dtype="float64" (SURVEY.md A.2-1).  Each fixture stores the case definition, the seeds and
the reference's outputs (trace, every sampler statistic, final mass-matrix variance and step-size state).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "_stubs"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import littlemcmc as lmc  # noqa: E402  (the reference)
from oracle.lmc_oracle import diag_gaussian, neal_funnel  # noqa: E402  (target densities only)

assert lmc.__file__.startswith("/root/reference"), lmc.__file__


def run_reference(f, ndim, kind, draws, tune, start, seeds, pot_kw, step_kw):
    """Chains run one at a time so the per-chain final adaptation state can be captured."""
    traces, stats_all, finals = [], [], []
    for seed in seeds:
        if pot_kw["adapt"]:
            pot = lmc.QuadPotentialDiagAdapt(
                ndim, np.array(pot_kw["initial_mean"], dtype="d"), np.array(pot_kw["var"], dtype="d"),
                pot_kw["initial_weight"], adaptation_window=pot_kw.get("adaptation_window", 101),
                dtype="float64")
        else:
            pot = lmc.QuadPotentialDiag(np.array(pot_kw["var"], dtype="d"), dtype="float64")
        cls = lmc.NUTS if kind == "nuts" else lmc.HamiltonianMC
        step = cls(logp_dlogp_func=f, model_ndim=ndim, potential=pot, **step_kw)
        trace, stats = lmc.sample(f, ndim, draws=draws, tune=tune, step=step, chains=1, cores=1,
                                  start=np.array(start, dtype="d"), progressbar=False,
                                  random_seed=[int(seed)], discard_tuned_samples=False)
        traces.append(trace[0])
        stats_all.append({k: v[0, :, 0] for k, v in stats.items()})
        var = pot._var if pot_kw["adapt"] else pot.v
        sa = step.step_adapt
        finals.append((np.array(var, dtype="d"),
                       np.array([sa._log_step, sa._log_bar, sa._hbar, sa._count, sa._mu], dtype="d"),
                       getattr(pot, "_n_samples", 0)))
    out = {"trace": np.stack(traces)}
    for k in stats_all[0]:
        out["stat_" + k] = np.stack([s[k] for s in stats_all]).astype("d")
    out["final_var"] = np.stack([f_[0] for f_ in finals])
    out["final_step_adapt"] = np.stack([f_[1] for f_ in finals])
    out["final_n_samples"] = np.array([f_[2] for f_ in finals], dtype="d")
    return out


def save(name, case, out):
    flat = {"case_" + k: np.asarray(v) for k, v in case.items()}
    flat.update(out)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **flat)
    print("%-28s %8.1f KB  leapfrogs=%d" % (name, os.path.getsize(path) / 1e3,
          int(out.get("stat_tree_size", out.get("stat_n_steps")).sum())))


def main():
    # ---- B.1: the survey's NUTS smoke case, cross-checked against SURVEY.md appendix B here ----------
    D = 10
    sigma = np.linspace(0.5, 2, D)
    case = dict(kind="nuts", target="diag_gaussian", ndim=D, tau=1 / sigma**2, draws=3, tune=5,
                start=np.full(D, 0.1), seeds=[11, 12], pot_adapt=1, pot_mean=np.zeros(D), pot_var=np.ones(D),
                pot_weight=10, max_treedepth=10, early_max_treedepth=8)
    out = run_reference(diag_gaussian(case["tau"]), D, "nuts", 3, 5, case["start"], case["seeds"],
                        dict(adapt=True, initial_mean=np.zeros(D), var=np.ones(D), initial_weight=10), {})
    # SURVEY.md B.1 values (the survey used g=-q/sigma^2; ours is -(tau*q): equal to ~1e-15, so compare loosely)
    assert out["stat_tree_size"][0, :3].tolist() == [15, 1, 15] and out["stat_depth"][1, 0] == 6
    assert abs(out["trace"][0, 0, 0] - (-0.746964253525814)) < 1e-9
    assert abs(out["trace"][1, 7, 0] - (-0.867558841064338)) < 1e-9
    assert abs(out["stat_step_size_bar"][0, 4] - 0.662239562978267) < 1e-9
    save("nuts_b1_d10", case, out)

    # ---- B.2 / BASELINE config 1: HMC, 4 chains, D=10 isotropic Gaussian, path_length 2, 500+500 --------
    D = 10
    case = dict(kind="hmc", target="diag_gaussian", ndim=D, tau=np.ones(D), draws=500, tune=500,
                start=np.zeros(D), seeds=[101, 102, 103, 104], pot_adapt=1, pot_mean=np.zeros(D),
                pot_var=np.ones(D), pot_weight=10, path_length=2.0, max_steps=1024)
    out = run_reference(diag_gaussian(case["tau"]), D, "hmc", 500, 500, case["start"], case["seeds"],
                        dict(adapt=True, initial_mean=np.zeros(D), var=np.ones(D), initial_weight=10),
                        dict(path_length=2.0))
    assert int(out["stat_n_steps"].sum()) == 5389, out["stat_n_steps"].sum()
    assert abs(out["trace"][0, 0, 0] - 0.753565613928047) < 1e-12
    assert out["stat_n_steps"][3, :5].tolist() == [4, 1, 3, 1, 2]
    save("hmc_cfg1_d10", case, out)

    # ---- NUTS, odd D, long enough to cross the 101-draw adaptation window and iter_count 200 -----------
    D = 37
    rs = np.random.RandomState(7)
    sigma = 10 ** np.linspace(-0.5, 0.5, D)
    start = 2 * rs.rand(D) - 1
    case = dict(kind="nuts", target="diag_gaussian", ndim=D, tau=1 / sigma**2, draws=20, tune=215,
                start=start, seeds=[21, 22, 23], pot_adapt=1, pot_mean=start, pot_var=np.ones(D),
                pot_weight=10, max_treedepth=10, early_max_treedepth=8)
    out = run_reference(diag_gaussian(case["tau"]), D, "nuts", 20, 215, start, case["seeds"],
                        dict(adapt=True, initial_mean=start, var=np.ones(D), initial_weight=10), {})
    save("nuts_diag_d37", case, out)

    # ---- NUTS, D=100 (cfg2's shape), static QuadPotentialDiag, no step-size adaptation ----------------
    D = 100
    sigma = 10 ** np.linspace(-0.5, 0.5, D)
    case = dict(kind="nuts", target="diag_gaussian", ndim=D, tau=1 / sigma**2, draws=12, tune=12,
                start=np.full(D, 0.05), seeds=[31, 32], pot_adapt=0, pot_mean=np.zeros(D), pot_var=sigma**2,
                pot_weight=0, max_treedepth=6, early_max_treedepth=4, adapt_step_size=0, step_scale=1.2)
    out = run_reference(diag_gaussian(case["tau"]), D, "nuts", 12, 12, case["start"], case["seeds"],
                        dict(adapt=False, var=sigma**2),
                        dict(max_treedepth=6, early_max_treedepth=4, adapt_step_size=False, step_scale=1.2))
    save("nuts_static_d100", case, out)

    # ---- NUTS, D=1000 ill-conditioned (cfg3's shape, kappa = 1e4) ---------------------------------------
    D = 1000
    sig2 = 10 ** np.linspace(0, 4, D)
    case = dict(kind="nuts", target="diag_gaussian", ndim=D, tau=1 / sig2, draws=4, tune=14,
                start=np.zeros(D), seeds=[41], pot_adapt=1, pot_mean=np.zeros(D), pot_var=np.ones(D),
                pot_weight=10, max_treedepth=10, early_max_treedepth=8)
    out = run_reference(diag_gaussian(case["tau"]), D, "nuts", 4, 14, case["start"], case["seeds"],
                        dict(adapt=True, initial_mean=np.zeros(D), var=np.ones(D), initial_weight=10), {})
    save("nuts_illcond_d1000", case, out)

    # ---- NUTS on Neal's funnel (cfg4's density): divergences, deep trees -------------------------------
    D = 10
    case = dict(kind="nuts", target="funnel", ndim=D, tau=np.zeros(0), draws=60, tune=80,
                start=np.zeros(D), seeds=[51, 52, 53, 54], pot_adapt=1, pot_mean=np.zeros(D), pot_var=np.ones(D),
                pot_weight=10, max_treedepth=7, early_max_treedepth=5, Emax=50.0)
    out = run_reference(neal_funnel(D), D, "nuts", 60, 80, case["start"], case["seeds"],
                        dict(adapt=True, initial_mean=np.zeros(D), var=np.ones(D), initial_weight=10),
                        dict(max_treedepth=7, early_max_treedepth=5, Emax=50.0))
    print("   funnel divergences:", int(out["stat_diverging"].sum()),
          " max depth hits:", int((out["stat_depth"] == 7).sum()))
    save("nuts_funnel_d10", case, out)

    # ---- HMC, static potential, D=50, forced divergences through a huge step ---------------------------
    D = 50
    sigma = 10 ** np.linspace(-1, 0, D)
    case = dict(kind="hmc", target="diag_gaussian", ndim=D, tau=1 / sigma**2, draws=30, tune=30,
                start=np.full(D, 0.01), seeds=[61, 62], pot_adapt=0, pot_mean=np.zeros(D), pot_var=np.ones(D),
                pot_weight=0, path_length=3.0, max_steps=16, step_scale=0.9, Emax=1000.0)
    out = run_reference(diag_gaussian(case["tau"]), D, "hmc", 30, 30, case["start"], case["seeds"],
                        dict(adapt=False, var=np.ones(D)),
                        dict(path_length=3.0, max_steps=16, step_scale=0.9))
    print("   hmc divergences:", int(out["stat_diverging"].sum()), " accepted:", int(out["stat_accepted"].sum()))
    save("hmc_static_d50", case, out)


if __name__ == "__main__":
    main()
