"""Load the committed reference fixtures (tests/golden/*.npz) and rebuild each case for the oracle."""
import glob
import os

import numpy as np

from oracle import lmc_oracle as orc

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_ALL = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
CASE_NAMES = [n for n in _ALL if not n.startswith("dense_")]          # diagonal potentials (make_golden.py)
DENSE_CASE_NAMES = [n for n in _ALL if n.startswith("dense_")]        # dense potentials (make_golden_dense.py)

_SAMPLER_KEYS = ("max_treedepth", "early_max_treedepth", "Emax", "path_length", "max_steps", "step_scale",
                 "adapt_step_size", "target_accept")


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    case = {k[5:]: z[k] for k in z.files if k.startswith("case_")}
    ref = {k: z[k] for k in z.files if not k.startswith("case_")}
    case = {k: (v.item() if v.ndim == 0 else v) for k, v in case.items()}
    return case, ref


def target_fn(case):
    if case["target"] == "dense_gaussian":
        return lambda: orc.dense_gaussian(case["prec"])
    if case["target"] == "diag_gaussian":
        return lambda: orc.diag_gaussian(case["tau"])
    if case["target"] == "funnel":
        return lambda: orc.neal_funnel(int(case["ndim"]))
    raise KeyError(case["target"])


def potential_kw(case):
    if int(case["pot_adapt"]):
        return dict(var=case["pot_var"], initial_mean=case["pot_mean"], initial_weight=float(case["pot_weight"]),
                    adapt=True)
    return dict(var=case["pot_var"], adapt=False)


def dense_potential(case):
    """A fresh oracle potential for one chain of a dense case."""
    if case["pot"] == "full":
        return orc.FullPotential(case["pot_matrix"])
    if case["pot"] == "fullinv":
        return orc.FullInvPotential(case["pot_matrix"])
    return orc.FullAdaptPotential(int(case["ndim"]), case["pot_mean"], case["pot_matrix"], float(case["pot_weight"]),
                                  adaptation_window=int(case["adaptation_window"]),
                                  adaptation_window_multiplier=float(case["adaptation_window_multiplier"]))


def run_oracle_dense(case, record=False):
    """Every chain with a fresh potential (see tests/golden/make_golden_dense.py).  -> trace [C,T,D], stats, pots"""
    D, kind = int(case["ndim"]), str(case["kind"])
    traces, stats_all, pots = [], [], []
    for s in case["seeds"]:
        pot = dense_potential(case)
        smp = orc.Sampler(target_fn(case)(), D, pot, kind=kind, **sampler_kw(case))
        tr, st = orc.sample_chain(smp, case["start"], int(case["draws"]), int(case["tune"]),
                                  np.random.RandomState(int(s)))
        traces.append(tr)
        stats_all.append(st)
        pots.append((pot, smp))
    return np.stack(traces), {n: np.stack([s[n] for s in stats_all]) for n in stats_all[0]}, pots


def sampler_kw(case):
    kw = {k: case[k] for k in _SAMPLER_KEYS if k in case}
    for k in ("max_treedepth", "early_max_treedepth", "max_steps"):
        if k in kw:
            kw[k] = int(kw[k])
    if "adapt_step_size" in kw:
        kw["adapt_step_size"] = bool(kw["adapt_step_size"])
    return kw


def run_oracle(case, record=False):
    return orc.run_chains(target_fn(case), int(case["ndim"]), str(case["kind"]), int(case["draws"]),
                          int(case["tune"]), case["start"], [int(s) for s in case["seeds"]],
                          potential=potential_kw(case), record=record, **sampler_kw(case))


EXACT_STATS = ("depth", "tree_size", "diverging", "tune", "n_steps", "accepted")
