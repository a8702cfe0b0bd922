"""Shared machinery of the GPU parity tests: run a golden case through the CPU oracle (recording the randomness it
consumes) and through the CUDA path in tape mode with the very same numbers, then compare."""
from dataclasses import dataclass
from typing import Dict

import numpy as np

from oracle import lmc_oracle as orc
from tests import golden_cases as gc

NUTS_STATS = {"depth": 0, "tree_size": 1, "mean_tree_accept": 2, "energy": 3, "energy_error": 4,
              "max_energy_error": 5, "model_logp": 6, "diverging": 7, "tune": 8, "step_size": 9, "step_size_bar": 10}
HMC_STATS = {"n_steps": 0, "path_length": 1, "accept": 2, "energy": 3, "energy_error": 4, "accepted": 5,
             "model_logp": 6, "diverging": 7, "tune": 8, "step_size": 9, "step_size_bar": 10}
EXACT = ("depth", "tree_size", "diverging", "tune", "n_steps", "accepted")


@dataclass
class ParityResult:
    kind: str
    gpu_trace: np.ndarray
    cpu_trace: np.ndarray
    gpu_stats: Dict[str, np.ndarray]
    cpu_stats: Dict[str, np.ndarray]
    gpu_var: np.ndarray
    cpu_var: np.ndarray
    gpu_adapt: np.ndarray
    cpu_adapt: np.ndarray
    n_uniforms_gpu: np.ndarray
    n_uniforms_cpu: np.ndarray
    status: np.ndarray


def truncate_case(case, n_trans):
    if n_trans is None:
        return case
    case = dict(case)
    tune = min(int(case["tune"]), n_trans)
    case["tune"], case["draws"] = tune, n_trans - tune
    return case


def oracle_run(case):
    """-> trace [C,T,D], stats, (normals, uniforms, n_uniforms), final var [C,D], final adapt scalars [C,5]."""
    D, kind = int(case["ndim"]), str(case["kind"])
    traces, stats_all, tapes, finals = [], [], [], []
    for s in case["seeds"]:
        rng = orc.TapeRecorder(np.random.RandomState(int(s)))
        smp = orc.Sampler(gc.target_fn(case)(), D, orc.DiagPotential(D, **gc.potential_kw(case)), kind=kind,
                          **gc.sampler_kw(case))
        tr, st = orc.sample_chain(smp, case["start"], int(case["draws"]), int(case["tune"]), rng)
        traces.append(tr)
        stats_all.append(st)
        tapes.append(rng)
        sa = smp.step_adapt
        finals.append((smp.pot.var.copy(), np.array([sa.log_step, sa.log_bar, sa.hbar, sa.count, sa.mu], dtype="d")))
    width = max(max(max(len(u) for u in t.uniforms) for t in tapes), 1)
    parts = [t.tapes(pad_to=width) for t in tapes]
    stats = {n: np.stack([s[n] for s in stats_all]) for n in stats_all[0]}
    return (np.stack(traces), stats, tuple(np.stack([p[i] for p in parts]) for i in range(3)),
            np.stack([f[0] for f in finals]), np.stack([f[1] for f in finals]))


def gpu_params(case):
    kw = gc.sampler_kw(case)
    return dict(adapt_mass=int(case["pot_adapt"]), adapt_step_size=int(kw.get("adapt_step_size", True)),
                target_accept=kw.get("target_accept", 0.8), gamma=0.05, k=0.75, t0=10,
                Emax=kw.get("Emax", 1000.0), max_treedepth=kw.get("max_treedepth", 10),
                early_max_treedepth=kw.get("early_max_treedepth", 8), path_length=kw.get("path_length", 2.0),
                max_steps=kw.get("max_steps", 1024))


def gpu_target(case):
    from littlemcmc_b200 import _lib as L
    from littlemcmc_b200.engine import FusedTarget
    if case["target"] == "diag_gaussian":
        return FusedTarget(L.TARGET_DIAG_GAUSSIAN, int(case["ndim"]), tau=case["tau"])
    return FusedTarget(L.TARGET_FUNNEL, int(case["ndim"]), v_scale=3.0)


def gpu_chains(case, n_chains, device="cuda:0"):
    from littlemcmc_b200.engine import DeviceChains
    D = int(case["ndim"])
    ch = DeviceChains(n_chains, D, device)
    ch.reset_potential(case["pot_var"], case["pot_mean"], float(case["pot_weight"]) if int(case["pot_adapt"]) else 0.0,
                       101)
    kw = gc.sampler_kw(case)
    ch.reset_step_adapt(kw.get("step_scale", 0.25) / D ** 0.25)
    ch.set_position(case["start"])
    return ch


def gpu_run(case, tapes, chunks=1, knobs=None, device="cuda:0"):
    import torch
    from littlemcmc_b200 import _lib as L
    from littlemcmc_b200 import engine
    normals, uniforms, _ = tapes
    Cn, T, D = normals.shape
    kind = L.KIND_NUTS if str(case["kind"]) == "nuts" else L.KIND_HMC
    ch = gpu_chains(case, Cn, device)
    tgt = gpu_target(case)
    params = gpu_params(case)
    bounds = np.linspace(0, T, chunks + 1).astype(int)
    traces, stats = [], []
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        if hi == lo:
            continue
        tr, st = engine.run_transitions(kind, ch, tgt, n_trans=int(hi - lo), iter0=int(lo), n_tune=int(case["tune"]),
                                        params=params, tapes=(normals[:, lo:hi], uniforms[:, lo:hi]), knobs=knobs)
        traces.append(tr)
        stats.append(st)
    torch.cuda.synchronize()
    trace = torch.cat(traces, 1).cpu().numpy()
    st = torch.cat(stats, 1).cpu().numpy()
    return trace, st, ch


def run_case_on_gpu_and_oracle(name, n_trans=None, chunks=1, knobs=None, device="cuda:0") -> ParityResult:
    from littlemcmc_b200 import _lib as L
    case, _ = gc.load(name)
    case = truncate_case(case, n_trans)
    cpu_trace, cpu_stats, tapes, cpu_var, cpu_adapt = oracle_run(case)
    trace, st, ch = gpu_run(case, tapes, chunks=chunks, knobs=knobs, device=device)
    table = NUTS_STATS if str(case["kind"]) == "nuts" else HMC_STATS
    gpu_stats = {n: st[:, :, i] for n, i in table.items()}
    return ParityResult(str(case["kind"]), trace, cpu_trace, gpu_stats, cpu_stats,
                        ch.var[:, : ch.ndim].cpu().numpy(), cpu_var,
                        ch.adapt[:, :5].cpu().numpy(), cpu_adapt, st[:, :, L.STAT_N_UNIFORMS], tapes[2],
                        ch.status.cpu().numpy())


def assert_parity(res: ParityResult, rtol=1e-9, atol=1e-12):
    """Bar: integer / boolean statistics and the number of uniforms consumed are exact; float64 quantities agree
    to `rtol` (the only licence to differ is the summation order of dot products and libm ulps)."""
    assert (res.status == 0).all(), res.status
    assert np.array_equal(res.n_uniforms_gpu, res.n_uniforms_cpu), "uniform consumption differs"
    for k, v in res.cpu_stats.items():
        g = res.gpu_stats[k]
        if k in EXACT:
            assert np.array_equal(g, np.asarray(v, dtype="d")), k
        else:
            np.testing.assert_allclose(g, v, rtol=rtol, atol=atol, err_msg=k)
    np.testing.assert_allclose(res.gpu_trace, res.cpu_trace, rtol=rtol, atol=atol)
    np.testing.assert_allclose(res.gpu_var, res.cpu_var, rtol=rtol, atol=atol)
    np.testing.assert_allclose(res.gpu_adapt, res.cpu_adapt, rtol=rtol, atol=atol)
