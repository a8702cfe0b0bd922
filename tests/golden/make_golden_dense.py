"""Generate tests/golden/dense_*.npz: the UNMODIFIED reference (eigenfoo/littlemcmc @ 2b5dd87) with its dense
potentials QuadPotentialFull / QuadPotentialFullInv / QuadPotentialFullAdapt (reference quadpotential.py:390-615),
dtype="float64", on a correlated Gaussian target.  Same conventions as make_golden.py (build container only).

    python tests/golden/make_golden_dense.py

Every chain gets a FRESH potential object: the reference never resets QuadPotentialFullAdapt between chains (its
reset() is the base-class no-op, quadpotential.py:138-140), so with one object chain 2 would start from chain 1's
adapted matrix; the GPU path gives every chain the initial matrix, which is what a fresh object per chain does.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "_stubs"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import littlemcmc as lmc  # noqa: E402  (the reference)
from oracle.lmc_oracle import dense_gaussian  # noqa: E402  (target density only)
from tests.dense_utils import spd  # noqa: E402

assert lmc.__file__.startswith("/root/reference"), lmc.__file__


def make_potential(case):
    n = int(case["ndim"])
    if case["pot"] == "full":
        return lmc.QuadPotentialFull(np.array(case["pot_matrix"]), dtype="float64")
    if case["pot"] == "fullinv":
        return lmc.QuadPotentialFullInv(np.array(case["pot_matrix"]), dtype="float64")
    return lmc.QuadPotentialFullAdapt(n, np.array(case["pot_mean"], dtype="d"), np.array(case["pot_matrix"]),
                                      case["pot_weight"], adaptation_window=int(case["adaptation_window"]),
                                      adaptation_window_multiplier=float(case["adaptation_window_multiplier"]),
                                      dtype="float64")


def run_reference(case, step_kw):
    f = dense_gaussian(case["prec"])
    n = int(case["ndim"])
    traces, stats_all, finals = [], [], []
    for seed in case["seeds"]:
        pot = make_potential(case)
        cls = lmc.NUTS if case["kind"] == "nuts" else lmc.HamiltonianMC
        step = cls(logp_dlogp_func=f, model_ndim=n, potential=pot, **step_kw)
        trace, stats = lmc.sample(f, n, draws=int(case["draws"]), tune=int(case["tune"]), step=step, chains=1, cores=1,
                                  start=np.array(case["start"], dtype="d"), progressbar=False,
                                  random_seed=[int(seed)], discard_tuned_samples=False)
        traces.append(trace[0])
        stats_all.append({k: v[0, :, 0] for k, v in stats.items()})
        sa = step.step_adapt
        cov = getattr(pot, "_cov", None)
        finals.append((np.zeros((n, n)) if cov is None else np.array(cov, dtype="d"),
                       np.array([sa._log_step, sa._log_bar, sa._hbar, sa._count, sa._mu], dtype="d"),
                       getattr(pot, "_n_samples", 0), getattr(pot, "_adaptation_window", 0)))
    out = {"trace": np.stack(traces)}
    for k in stats_all[0]:
        out["stat_" + k] = np.stack([s[k] for s in stats_all]).astype("d")
    out["final_cov"] = np.stack([f_[0] for f_ in finals])
    out["final_step_adapt"] = np.stack([f_[1] for f_ in finals])
    out["final_n_samples"] = np.array([f_[2] for f_ in finals], dtype="d")
    out["final_window"] = np.array([f_[3] for f_ in finals], dtype="d")
    return out


def save(name, case, out):
    flat = {"case_" + k: np.asarray(v) for k, v in case.items()}
    flat.update(out)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **flat)
    print("%-28s %8.1f KB  leapfrogs=%d" % (name, os.path.getsize(path) / 1e3,
          int(out.get("stat_tree_size", out.get("stat_n_steps")).sum())))


def main():
    # ---- NUTS, static dense covariance close to the target's (QuadPotentialFull) -----------------------------------
    D = 12
    prec = spd(D, 1)
    cov = np.linalg.inv(prec) * (1 + 0.2 * np.cos(np.arange(D)))[:, None] * (1 + 0.2 * np.cos(np.arange(D)))[None, :]
    case = dict(kind="nuts", target="dense_gaussian", ndim=D, prec=prec, draws=10, tune=15,
                start=np.full(D, 0.1), seeds=[71, 72, 73], pot="full", pot_matrix=cov, max_treedepth=8,
                early_max_treedepth=6)
    save("dense_full_nuts_d12", case, run_reference(case, dict(max_treedepth=8, early_max_treedepth=6)))

    # ---- NUTS, static dense inverse covariance (QuadPotentialFullInv) ------------------------------------------------
    case = dict(kind="nuts", target="dense_gaussian", ndim=D, prec=prec, draws=10, tune=15,
                start=np.full(D, -0.2), seeds=[81, 82], pot="fullinv", pot_matrix=np.linalg.inv(cov), max_treedepth=8,
                early_max_treedepth=6)
    save("dense_fullinv_nuts_d12", case, run_reference(case, dict(max_treedepth=8, early_max_treedepth=6)))

    # ---- HMC, static dense covariance ---------------------------------------------------------------------------
    case = dict(kind="hmc", target="dense_gaussian", ndim=D, prec=prec, draws=15, tune=15,
                start=np.full(D, 0.3), seeds=[91, 92], pot="full", pot_matrix=cov, path_length=2.0, max_steps=64)
    save("dense_full_hmc_d12", case, run_reference(case, dict(path_length=2.0, max_steps=64)))

    # ---- NUTS, adapted dense mass matrix, short windows so that two window switches happen (20, then 40) ----------
    D = 9
    prec = spd(D, 2, cond=12.0)
    rs = np.random.RandomState(3)
    start = 2 * rs.rand(D) - 1
    case = dict(kind="nuts", target="dense_gaussian", ndim=D, prec=prec, draws=8, tune=72,
                start=start, seeds=[101, 102], pot="fulladapt", pot_matrix=np.eye(D), pot_mean=start, pot_weight=10,
                adaptation_window=20, adaptation_window_multiplier=2, max_treedepth=8, early_max_treedepth=6)
    out = run_reference(case, dict(max_treedepth=8, early_max_treedepth=6))
    print("   final windows:", out["final_window"], " n_samples:", out["final_n_samples"])
    save("dense_fulladapt_nuts_d9", case, out)


if __name__ == "__main__":
    main()
