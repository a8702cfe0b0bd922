"""Load the committed reference fixtures (tests/golden/*.npz) and rebuild each case for the oracle."""
import glob
import os

import numpy as np

from oracle import lmc_oracle as orc

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASE_NAMES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))

_SAMPLER_KEYS = ("max_treedepth", "early_max_treedepth", "Emax", "path_length", "max_steps", "step_scale",
                 "adapt_step_size", "target_accept")


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    case = {k[5:]: z[k] for k in z.files if k.startswith("case_")}
    ref = {k: z[k] for k in z.files if not k.startswith("case_")}
    case = {k: (v.item() if v.ndim == 0 else v) for k, v in case.items()}
    return case, ref


def target_fn(case):
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
