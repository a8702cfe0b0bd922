"""Shared machinery of the GPU parity tests: run a golden case through the CPU oracle (recording the randomness it
consumes) and through the CUDA path in tape mode with the very same numbers, then compare."""
from dataclasses import dataclass
from typing import Dict

import numpy as np

from oracle import lmc_oracle as orc
from tests import golden_cases as gc

NUTS_STATS = {"depth": 0, "tree_size": 1, "mean_tree_accept": 2, "energy": 3, "energy_error": 4,
              "max_energy_error": 5, "model_logp": 6, "diverging": 7, "tune": 8, "step_size": 9, "step_size_bar": 10,
              "reached_max_treedepth": 12}
HMC_STATS = {"n_steps": 0, "path_length": 1, "accept": 2, "energy": 3, "energy_error": 4, "accepted": 5,
             "model_logp": 6, "diverging": 7, "tune": 8, "step_size": 9, "step_size_bar": 10}
EXACT = ("depth", "tree_size", "diverging", "tune", "n_steps", "accepted", "reached_max_treedepth")


@dataclass
class ParityResult:
    kind: str
    gpu_trace: np.ndarray          # [C, T, D] position after each transition
    cpu_trace: np.ndarray
    gpu_stats: Dict[str, np.ndarray]   # name -> [C, T]
    cpu_stats: Dict[str, np.ndarray]
    gpu_var: np.ndarray            # [C, T, D] mass-matrix variance after each transition
    cpu_var: np.ndarray
    gpu_adapt: np.ndarray          # [C, T, 9] adaptation scalars after each transition (LMC_ADAPT_* order)
    cpu_adapt: np.ndarray
    gpu_welford: np.ndarray        # [C, T, 4, D] mean_fg, rawvar_fg, mean_bg, rawvar_bg after each transition
    cpu_welford: np.ndarray
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


def _snapshot(smp, q):
    """Everything one transition reads: position, potential arrays, adaptation scalars (LMC_ADAPT_* order)."""
    pot, sa = smp.pot, smp.step_adapt
    scal = np.array([sa.log_step, sa.log_bar, sa.hbar, sa.count, sa.mu, pot.fg.w_sum, pot.bg.w_sum, pot.n_samples,
                     pot.adaptation_window], dtype="d")
    wel = np.stack([pot.fg.mean, pot.fg.raw_var, pot.bg.mean, pot.bg.raw_var])
    return np.array(q, dtype="d"), pot.var.copy(), wel, scal


def oracle_run(case):
    """Run every chain of `case` through the CPU oracle with the reference's MT19937 stream, recording the randomness
    consumed and the full sampler state before and after every transition.

    Returns dict: trace [C,T,D], stats{name: [C,T]}, tapes (normals [C,T,D], uniforms [C,T,U], n_uniforms [C,T]),
    pre/post: q [C,T,D], var [C,T,D], welford [C,T,4,D], adapt [C,T,9].
    """
    D, kind = int(case["ndim"]), str(case["kind"])
    T, tune = int(case["tune"]) + int(case["draws"]), int(case["tune"])
    # + the flag behind NUTS._reached_max_treedepth (nuts.py:218-220), which the kernels report per transition
    names = orc.NUTS_STAT_NAMES + ("reached_max_treedepth",) if kind == "nuts" else orc.HMC_STAT_NAMES
    recs, tapes, stats_all = [], [], []
    for s in case["seeds"]:
        rng = orc.TapeRecorder(np.random.RandomState(int(s)))
        smp = orc.Sampler(gc.target_fn(case)(), D, orc.DiagPotential(D, **gc.potential_kw(case)), kind=kind,
                          step_rand=case.get("step_rand"), **gc.sampler_kw(case))
        # sampling.py:503-513, spelled out so the state can be snapshotted around each _astep
        smp.tune = bool(tune)
        smp.reset_tuning()
        q = np.array(case["start"], dtype="d")
        pre, post, st = [], [], {n: np.zeros(T) for n in names}
        for i in range(T):
            if i == 0:
                smp.iter_count = 0
            if i == tune:
                smp.tune = False
            pre.append(_snapshot(smp, q))
            q, sd = smp.astep(q, rng)
            post.append(_snapshot(smp, q))
            for n in names:
                st[n][i] = sd[n]
        recs.append((pre, post))
        stats_all.append(st)
        tapes.append(rng)
    width = max(max(max(len(u) for u in t.uniforms) for t in tapes), 1)
    parts = [t.tapes(pad_to=width) for t in tapes]

    def stack(which, field):
        return np.stack([np.stack([snap[field] for snap in r[which]]) for r in recs])

    out = dict(stats={n: np.stack([s[n] for s in stats_all]) for n in names},
               tapes=tuple(np.stack([p[i] for p in parts]) for i in range(3)))
    for w, nm in ((0, "pre"), (1, "post")):
        out[nm] = dict(q=stack(w, 0), var=stack(w, 1), welford=stack(w, 2), adapt=stack(w, 3))
    out["trace"] = out["post"]["q"]
    return out


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


def torch_callback(case, device="cuda:0", cuda_graph=False):
    """The case's target density as a batched torch op (targets.TorchBatched): the same formulas as the oracle's NumPy
    callables (oracle/lmc_oracle.py diag_gaussian / neal_funnel), evaluated for all chains at once on the device."""
    import torch
    from littlemcmc_b200.targets import TorchBatched
    D = int(case["ndim"])
    if case["target"] == "diag_gaussian":
        tau = torch.as_tensor(np.asarray(case["tau"], dtype="d"), device=device)

        def fn(q):
            g = -(tau * q)
            return 0.5 * (q * g).sum(1), g
        return TorchBatched(fn, cuda_graph=cuda_graph)
    inv_s2, half_nm1 = 1.0 / 9.0, 0.5 * (D - 1)

    def fn(q):
        v, x = q[:, 0], q[:, 1:]
        S = (x * x).sum(1)
        ev = torch.exp(-v)
        hs = 0.5 * ev * S
        g = torch.empty_like(q)
        g[:, 1:] = -(ev[:, None] * x)
        g[:, 0] = -(v * inv_s2) + hs - half_nm1
        return -(0.5 * v * v * inv_s2) - hs - half_nm1 * v, g
    return TorchBatched(fn, cuda_graph=cuda_graph)


def _launch(case, ch, tgt, callback, **kw):
    """Fused kernel (callback None) or callback mode."""
    from littlemcmc_b200 import engine
    if callback is None:
        return engine.run_transitions(_kind(case), ch, tgt, **kw)
    kw.pop("knobs", None)
    return engine.run_transitions_callback(_kind(case), ch, callback, cuda_graph=getattr(callback, "cuda_graph", False),
                                           **kw)


def _kind(case):
    from littlemcmc_b200 import _lib as L
    return L.KIND_NUTS if str(case["kind"]) == "nuts" else L.KIND_HMC


def _read_state(ch):
    import torch
    D = ch.ndim
    wel = torch.stack([ch.mean_fg[:, :D], ch.rawvar_fg[:, :D], ch.mean_bg[:, :D], ch.rawvar_bg[:, :D]], 1)
    return ch.q[:, :D].cpu().numpy(), ch.var[:, :D].cpu().numpy(), wel.cpu().numpy(), ch.adapt[:, :9].cpu().numpy()


def gpu_run_chained(case, tapes, chunks=1, knobs=None, device="cuda:0", callback=None, target=None):
    """The whole run on the GPU, state carried on the device between `chunks` launches (run-level)."""
    import torch
    from littlemcmc_b200 import engine
    normals, uniforms, _ = tapes
    Cn, T, D = normals.shape
    ch = gpu_chains(case, Cn, device)
    tgt, params = target or gpu_target(case), gpu_params(case)
    bounds = np.linspace(0, T, chunks + 1).astype(int)
    traces, stats = [], []
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        if hi == lo:
            continue
        tr, st = _launch(case, ch, tgt, callback, n_trans=int(hi - lo), iter0=int(lo), n_tune=int(case["tune"]),
                         params=params, tapes=(normals[:, lo:hi], uniforms[:, lo:hi]), knobs=knobs)
        traces.append(tr)
        stats.append(st)
    torch.cuda.synchronize()
    return torch.cat(traces, 1).cpu().numpy(), torch.cat(stats, 1).cpu().numpy(), ch


def gpu_run_transitionwise(case, ora, knobs=None, device="cuda:0", callback=None, target=None):
    """Transition-level protocol: before EVERY transition the device state of every chain is set to
    the oracle's state before that transition, so both sides see identical (q0, var, step-size state, Welford state,
    normals, uniform tape) and only one transition's arithmetic is compared -- differences cannot compound through
    the (chaotic) adaptation feedback."""
    import torch
    from littlemcmc_b200 import _lib as L
    from littlemcmc_b200 import engine
    normals, uniforms, _ = ora["tapes"]
    Cn, T, D = normals.shape
    ch = gpu_chains(case, Cn, device)
    tgt, params = target or gpu_target(case), gpu_params(case)
    dev = ch.device
    up = lambda x: torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float64, device=dev)  # noqa: E731
    pre = {k: up(v) for k, v in ora["pre"].items()}
    normals_d, uniforms_d = up(normals), up(uniforms)
    out_q, out_var, out_wel, out_ad, out_st = [], [], [], [], []
    for t in range(T):
        ch.q[:, :D] = pre["q"][:, t]
        ch.var[:, :D] = pre["var"][:, t]
        for i, buf in enumerate((ch.mean_fg, ch.rawvar_fg, ch.mean_bg, ch.rawvar_bg)):
            buf[:, :D] = pre["welford"][:, t, i]
        ch.adapt[:, :9] = pre["adapt"][:, t]
        extra = {}
        if case.get("step_rand") is not None:   # base_hmc.py:154-155: the hook's output replaces current()
            tuning = t < int(case["tune"]) and bool(params["adapt_step_size"])
            cur = np.exp(ora["pre"]["adapt"][:, t, L.ADAPT_LOG_STEP if tuning else L.ADAPT_LOG_BAR])
            extra["step_size_override"] = np.array([case["step_rand"](float(e)) for e in cur])
        _, st = _launch(case, ch, tgt, callback, n_trans=1, iter0=t, n_tune=int(case["tune"]), params=params,
                        tapes=(normals_d[:, t:t + 1], uniforms_d[:, t:t + 1]), knobs=knobs, **extra)
        q, var, wel, ad = _read_state(ch)
        out_q.append(q); out_var.append(var); out_wel.append(wel); out_ad.append(ad)  # noqa: E702
        out_st.append(st[:, 0].cpu().numpy())
    return (np.stack(out_q, 1), np.stack(out_var, 1), np.stack(out_wel, 1), np.stack(out_ad, 1),
            np.stack(out_st, 1), ch)


def run_case_on_gpu_and_oracle(name, n_trans=None, knobs=None, device="cuda:0", chained=False,
                               chunks=1, callback=None, overrides=None, target=None) -> ParityResult:
    """`callback`: None = fused kernels; "torch" / "torch-graph" / "torch-replay" = callback mode with the case's density
    as a batched torch op (eager host loop / device-driven WHILE graph / replayed graphs); "numpy" = callback mode with
    the oracle's per-chain NumPy callable."""
    from littlemcmc_b200 import _lib as L
    case, _ = gc.load(name)
    case = truncate_case(case, n_trans)
    if overrides:                       # e.g. max_treedepth=3, step_rand=callable: variations on a committed fixture
        case = dict(case, **overrides)
    ora = oracle_run(case)
    table = NUTS_STATS if str(case["kind"]) == "nuts" else HMC_STATS
    if callback in ("torch", "torch-graph", "torch-replay"):
        callback = torch_callback(case, device, cuda_graph={"torch": False, "torch-graph": True,
                                                            "torch-replay": "replay"}[callback])
    elif callback == "numpy":
        callback = gc.target_fn(case)()
    if chained:
        trace, st, ch = gpu_run_chained(case, ora["tapes"], chunks=chunks, knobs=knobs, device=device,
                                        callback=callback, target=target)
        q, var, wel, ad = _read_state(ch)
        # only the final adaptation state is observable in a chained run
        return ParityResult(str(case["kind"]), trace, ora["trace"], {n: st[:, :, i] for n, i in table.items()},
                            ora["stats"], var[:, None], ora["post"]["var"][:, -1:], ad[:, None],
                            ora["post"]["adapt"][:, -1:], wel[:, None], ora["post"]["welford"][:, -1:],
                            st[:, :, L.STAT_N_UNIFORMS], ora["tapes"][2], ch.status.cpu().numpy())
    q, var, wel, ad, st, ch = gpu_run_transitionwise(case, ora, knobs=knobs, device=device, callback=callback,
                                                     target=target)
    return ParityResult(str(case["kind"]), q, ora["trace"], {n: st[:, :, i] for n, i in table.items()}, ora["stats"],
                        var, ora["post"]["var"], ad, ora["post"]["adapt"], wel, ora["post"]["welford"],
                        st[:, :, L.STAT_N_UNIFORMS], ora["tapes"][2], ch.status.cpu().numpy())


# statistics describing the END of a trajectory that was rejected as divergent: the integration there is unstable
# by definition, so last-bit differences in the step size are amplified exponentially along it
_DIVERGENT_SENSITIVE = ("energy", "energy_error", "max_energy_error", "model_logp")


def assert_parity(res: ParityResult, rtol=1e-9, atol=1e-12, rtol_divergent=1e-3):
    """Bar: integer / boolean statistics and the number of uniforms consumed are exact; float64 quantities agree
    to `rtol`.  For transitions flagged `diverging` the end-of-trajectory energies are compared with
    `rtol_divergent` only (see _DIVERGENT_SENSITIVE)."""
    assert (res.status == 0).all(), res.status
    assert np.array_equal(res.n_uniforms_gpu, res.n_uniforms_cpu), "uniform consumption differs"
    div = np.asarray(res.cpu_stats["diverging"], dtype=bool)
    for k, v in res.cpu_stats.items():
        g, v = res.gpu_stats[k], np.asarray(v, dtype="d")
        if k in EXACT:
            assert np.array_equal(g, v), k
        elif k in _DIVERGENT_SENSITIVE:
            np.testing.assert_allclose(g[~div], v[~div], rtol=rtol, atol=atol, err_msg=k)
            np.testing.assert_allclose(g[div], v[div], rtol=rtol_divergent, atol=atol, err_msg=k + " (divergent)",
                                       equal_nan=True)
        else:
            np.testing.assert_allclose(g, v, rtol=rtol, atol=atol, err_msg=k)
    np.testing.assert_allclose(res.gpu_trace, res.cpu_trace, rtol=rtol, atol=atol, err_msg="trace")
    np.testing.assert_allclose(res.gpu_var, res.cpu_var, rtol=rtol, atol=atol, err_msg="var")
    np.testing.assert_allclose(res.gpu_adapt, res.cpu_adapt, rtol=rtol, atol=atol, err_msg="adapt scalars")
    np.testing.assert_allclose(res.gpu_welford, res.cpu_welford, rtol=rtol, atol=atol, err_msg="welford")


def parity_report(res: ParityResult):
    """Human-readable max relative differences (diagnostics)."""
    def rel(a, b):
        with np.errstate(all="ignore"):
            r = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
        r = np.where(np.isfinite(r), r, np.where(np.equal(a, b) | (np.isnan(a) & np.isnan(b)), 0.0, np.inf))
        return float(np.max(r)) if r.size else 0.0
    lines = ["uniform counts equal: %s" % np.array_equal(res.n_uniforms_gpu, res.n_uniforms_cpu)]
    for k, v in res.cpu_stats.items():
        g, v = res.gpu_stats[k], np.asarray(v, dtype="d")
        lines.append("  %-18s %s  max rel %.2e" % (k, "EXACT" if np.array_equal(g, v) else "differs", rel(g, v)))
    for nm, a, b in (("trace", res.gpu_trace, res.cpu_trace), ("var", res.gpu_var, res.cpu_var),
                     ("adapt", res.gpu_adapt, res.cpu_adapt), ("welford", res.gpu_welford, res.cpu_welford)):
        lines.append("  %-18s max rel %.2e" % (nm, rel(a, b)))
    return "\n".join(lines)


# ---- replay of chains sampled from a FULL-SIZE device run -------------------------------------------------------------------
def oracle_sampler_from_state(f, D, q, var, wel, scal, *, iter_count, tune, **sampler_kw):
    """An oracle Sampler whose potential / step-size state is the snapshot (`_snapshot` order) read back from the
    device for one chain: the inverse of `_snapshot`, used to replay one transition of a chain picked out of a
    full-size run."""
    pot = orc.DiagPotential(D, var=np.ones(D), initial_mean=np.zeros(D), initial_weight=10.0, adapt=True)
    pot.var = np.array(var, dtype="d")
    pot.stds = np.sqrt(pot.var)
    pot.inv_stds = 1.0 / pot.stds
    pot.fg = orc.Welford(np.array(wel[0], dtype="d"), np.array(wel[1], dtype="d"), float(scal[5]))
    pot.bg = orc.Welford(np.array(wel[2], dtype="d"), np.array(wel[3], dtype="d"), float(scal[6]))
    pot.n_samples, pot.adaptation_window = int(scal[7]), int(scal[8])
    smp = orc.Sampler(f, D, pot, kind="nuts", **sampler_kw)
    sa = smp.step_adapt
    sa.log_step, sa.log_bar, sa.hbar, sa.count, sa.mu = (float(scal[0]), float(scal[1]), float(scal[2]),
                                                         int(scal[3]), float(scal[4]))
    smp.tune, smp.iter_count = bool(tune), int(iter_count)
    return smp


def replay_sampled_chains(f, fused_target, D, n_chains, *, max_treedepth, n_warm, n_check, n_sample=8, seed=0,
                          start=None, knobs=None, device="cuda:0"):
    """Full-size run on the GPU with in-kernel Philox randomness and the FIFO scheduler under contention; after `n_warm`
    tuning transitions, `n_check` more are launched one at a time and `n_sample` randomly chosen chains are replayed in
    the CPU oracle from their device pre-state, with their own Philox streams dumped by lmc_rng_fill.
    -> list of ParityResult (one per checked transition)."""
    import torch
    from littlemcmc_b200 import _lib as L
    from littlemcmc_b200 import engine
    rs = np.random.RandomState(seed)
    ch = engine.DeviceChains(n_chains, D, device)
    ch.reset_potential(np.ones(D), np.zeros(D), 10.0, 101)
    ch.reset_step_adapt(0.25 / D ** 0.25)
    ch.set_position(np.zeros(D) if start is None else start)
    params = dict(adapt_mass=1, adapt_step_size=1, target_accept=0.8, gamma=0.05, k=0.75, t0=10, Emax=1000.0,
                  max_treedepth=max_treedepth, early_max_treedepth=8)
    seeds_np = rs.randint(2 ** 30, size=n_chains)
    seeds = engine.seeds_tensor(seeds_np, device)
    big = 10 ** 9
    if n_warm:
        engine.run_transitions(L.KIND_NUTS, ch, fused_target, n_trans=n_warm, iter0=0, n_tune=big, params=params,
                               seeds=seeds, knobs=knobs)
    sel = np.sort(rs.choice(n_chains, size=n_sample, replace=False))
    sel_t = torch.as_tensor(sel, device=device)
    u_stride = (1 << max_treedepth) + max_treedepth + 8

    def read():
        wel = torch.stack([ch.mean_fg[sel_t, :D], ch.rawvar_fg[sel_t, :D], ch.mean_bg[sel_t, :D],
                           ch.rawvar_bg[sel_t, :D]], 1)
        return (ch.q[sel_t, :D].cpu().numpy(), ch.var[sel_t, :D].cpu().numpy(), wel.cpu().numpy(),
                ch.adapt[sel_t, :9].cpu().numpy())

    out = []
    for t in range(n_check):
        it = n_warm + t
        pre = read()
        _, st = engine.run_transitions(L.KIND_NUTS, ch, fused_target, n_trans=1, iter0=it, n_tune=big, params=params,
                                       seeds=seeds, knobs=knobs)
        torch.cuda.synchronize()
        post = read()
        st = st[sel_t, 0].cpu().numpy()
        normals, uniforms = engine.rng_fill(seeds[sel_t], D, it, 1, u_stride)
        normals, uniforms = normals.cpu().numpy(), uniforms.cpu().numpy()
        assert int(st[:, L.STAT_N_UNIFORMS].max()) <= u_stride
        names = orc.NUTS_STAT_NAMES + ("reached_max_treedepth",)
        cpu_stats = {n: np.zeros((n_sample, 1)) for n in names}
        cpu_q, cpu_var, cpu_wel, cpu_ad, n_u = [], [], [], [], []
        for j in range(n_sample):
            smp = oracle_sampler_from_state(f, D, pre[0][j], pre[1][j], pre[2][j], pre[3][j], iter_count=it, tune=True,
                                            max_treedepth=max_treedepth, early_max_treedepth=8)
            rng = orc.TapeRNG(normals[j], uniforms[j])
            q, sd = smp.astep(pre[0][j], rng)
            snap = _snapshot(smp, q)
            cpu_q.append(snap[0]); cpu_var.append(snap[1]); cpu_wel.append(snap[2]); cpu_ad.append(snap[3])  # noqa: E702
            n_u.append(rng._k)
            for n in names:
                cpu_stats[n][j, 0] = sd[n]
        out.append(ParityResult(
            "nuts", post[0][:, None], np.stack(cpu_q)[:, None], {n: st[:, i:i + 1] for n, i in NUTS_STATS.items()},
            cpu_stats, post[1][:, None], np.stack(cpu_var)[:, None], post[3][:, None], np.stack(cpu_ad)[:, None],
            post[2][:, None], np.stack(cpu_wel)[:, None], st[:, L.STAT_N_UNIFORMS][:, None],
            np.array(n_u, dtype="d")[:, None], ch.status[sel_t].cpu().numpy()))
    return out
