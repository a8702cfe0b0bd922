"""Parity machinery for the dense-mass mode (lmc_dense_* entry points, quadpotential_dense.py): the committed dense
fixtures are run through the CPU oracle, recording the randomness it consumes and the full sampler + potential state
around every transition; the CUDA path then runs every transition from the oracle's exact pre-state in tape mode
(transition-level protocol)."""
import numpy as np

from oracle import lmc_oracle as orc
from tests import golden_cases as gc


def spd(n, seed, cond=30.0):
    """A well-conditioned random SPD matrix with strong off-diagonal structure (the fixtures' precision matrices)."""
    rs = np.random.RandomState(seed)
    q, _ = np.linalg.qr(rs.randn(n, n))
    ev = np.exp(np.linspace(0, np.log(cond), n)) / np.sqrt(cond)
    m = (q * ev) @ q.T
    return 0.5 * (m + m.T)


def _pot_snapshot(pot):
    if not getattr(pot, "adapt", False):
        return None
    return dict(cov=pot.cov.copy(), chol=pot.chol.copy(), mean_fg=pot.fg.mean.copy(), raw_fg=pot.fg.raw_cov.copy(),
                n_fg=pot.fg.n_samples, mean_bg=pot.bg.mean.copy(), raw_bg=pot.bg.raw_cov.copy(), n_bg=pot.bg.n_samples,
                n_samples=pot.n_samples, previous_update=pot.previous_update, window=pot.adaptation_window)


def oracle_run_dense(case):
    """-> dict(stats{name: [C,T]}, tapes(normals [C,T,D], uniforms [C,T,U], n_uniforms [C,T]),
               pre/post: q [C,T,D], adapt [C,T,5], pot: list[C][T] of potential snapshots (None for static ones))"""
    D, kind = int(case["ndim"]), str(case["kind"])
    T, tune = int(case["tune"]) + int(case["draws"]), int(case["tune"])
    names = orc.NUTS_STAT_NAMES + ("reached_max_treedepth",) if kind == "nuts" else orc.HMC_STAT_NAMES
    out = dict(pre=dict(q=[], adapt=[], pot=[]), post=dict(q=[], adapt=[], pot=[]))
    tapes, stats_all = [], []
    for s in case["seeds"]:
        rng = orc.TapeRecorder(np.random.RandomState(int(s)))
        pot = gc.dense_potential(case)
        smp = orc.Sampler(gc.target_fn(case)(), D, pot, kind=kind, **gc.sampler_kw(case))
        smp.tune = bool(tune)
        smp.reset_tuning()
        q = np.array(case["start"], dtype="d")
        st = {n: np.zeros(T) for n in names}
        rec = {w: dict(q=[], adapt=[], pot=[]) for w in ("pre", "post")}

        def snap(w):
            sa = smp.step_adapt
            rec[w]["q"].append(np.array(q, dtype="d"))
            rec[w]["adapt"].append(np.array([sa.log_step, sa.log_bar, sa.hbar, sa.count, sa.mu], dtype="d"))
            rec[w]["pot"].append(_pot_snapshot(pot))
        for i in range(T):
            if i == 0:
                smp.iter_count = 0
            if i == tune:
                smp.tune = False
            snap("pre")
            q, sd = smp.astep(q, rng)
            snap("post")
            for n in names:
                st[n][i] = sd[n]
        for w in ("pre", "post"):
            out[w]["q"].append(np.stack(rec[w]["q"]))
            out[w]["adapt"].append(np.stack(rec[w]["adapt"]))
            out[w]["pot"].append(rec[w]["pot"])
        stats_all.append(st)
        tapes.append(rng)
    width = max(max(max(len(u) for u in t.uniforms) for t in tapes), 1)
    parts = [t.tapes(pad_to=width) for t in tapes]
    out["tapes"] = tuple(np.stack([p[i] for p in parts]) for i in range(3))
    out["stats"] = {n: np.stack([s[n] for s in stats_all]) for n in names}
    for w in ("pre", "post"):
        out[w]["q"], out[w]["adapt"] = np.stack(out[w]["q"]), np.stack(out[w]["adapt"])
    return out


def make_gpu_potential(case):
    import littlemcmc_b200 as lmc
    n = int(case["ndim"])
    if case["pot"] == "full":
        return lmc.QuadPotentialFull(case["pot_matrix"])
    if case["pot"] == "fullinv":
        return lmc.QuadPotentialFullInv(case["pot_matrix"])
    return lmc.QuadPotentialFullAdapt(n, case["pot_mean"], case["pot_matrix"], float(case["pot_weight"]),
                                      adaptation_window=int(case["adaptation_window"]),
                                      adaptation_window_multiplier=float(case["adaptation_window_multiplier"]))


def torch_dense_gaussian(prec, device):
    import torch
    from littlemcmc_b200.targets import TorchBatched
    P = torch.as_tensor(np.asarray(prec, dtype="d"), device=device)

    def fn(q):
        g = -(q @ P.mT)
        return 0.5 * (q * g).sum(1), g
    return TorchBatched(fn)


def _inject_potential(pot, snaps):
    """Set the per-chain state of a bound QuadPotentialFullAdapt to the oracle's snapshots (one per chain)."""
    import torch
    if snaps[0] is None:
        return
    D, dev = pot._n, pot._dev
    up = lambda k: torch.as_tensor(np.stack([s[k] for s in snaps]), dtype=torch.float64, device=dev)  # noqa: E731
    pot._cov_all[:, :, :D] = up("cov")
    pot._chol_all[:] = up("chol")
    pot._raw_fg[:, :, :D], pot._raw_bg[:, :, :D] = up("raw_fg"), up("raw_bg")
    pot._mean_fg[:, :D], pot._mean_bg[:, :D] = up("mean_fg"), up("mean_bg")
    pot._nsamp[:, 0], pot._nsamp[:, 1] = up("n_fg"), up("n_bg")
    pot._n_samples_all[:] = [s["n_samples"] for s in snaps]
    pot._previous_update_all[:] = [s["previous_update"] for s in snaps]
    pot._window_all[:] = [s["window"] for s in snaps]


def _read_potential(pot):
    if not pot._adaptive:
        return None
    D = pot._n
    return dict(cov=pot._cov_all[:, :, :D].cpu().numpy(), chol=pot._chol_all.cpu().numpy(),
                mean_fg=pot._mean_fg[:, :D].cpu().numpy(), raw_fg=pot._raw_fg[:, :, :D].cpu().numpy(),
                n_fg=pot._nsamp[:, 0].cpu().numpy(), mean_bg=pot._mean_bg[:, :D].cpu().numpy(),
                raw_bg=pot._raw_bg[:, :, :D].cpu().numpy(), n_bg=pot._nsamp[:, 1].cpu().numpy(),
                n_samples=pot._n_samples_all.copy(), previous_update=pot._previous_update_all.copy(),
                window=pot._window_all.copy())


def gpu_run_dense_transitionwise(case, ora, device="cuda:0"):
    """Every transition from the oracle's pre-state.  -> (q [C,T,D], adapt [C,T,5], stats [C,T,NSTATS],
    pots: list[T] of per-chain potential state dicts, status)"""
    import torch
    from littlemcmc_b200 import engine
    from tests import parity_utils as pu
    normals, uniforms, _ = ora["tapes"]
    Cn, T, D = normals.shape
    ch = engine.DeviceChains(Cn, D, device)
    kw = gc.sampler_kw(case)
    ch.reset_step_adapt(kw.get("step_scale", 0.25) / D ** 0.25)
    pot = make_gpu_potential(case)
    pot._bind(ch)
    cb = torch_dense_gaussian(case["prec"], ch.device)
    params = dict(adapt_mass=int(pot._adaptive), adapt_step_size=int(kw.get("adapt_step_size", True)),
                  target_accept=kw.get("target_accept", 0.8), gamma=0.05, k=0.75, t0=10, Emax=kw.get("Emax", 1000.0),
                  max_treedepth=kw.get("max_treedepth", 10), early_max_treedepth=kw.get("early_max_treedepth", 8),
                  path_length=kw.get("path_length", 2.0), max_steps=kw.get("max_steps", 1024))
    up = lambda x: torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float64, device=ch.device)  # noqa: E731
    pre_q, pre_ad = up(ora["pre"]["q"]), up(ora["pre"]["adapt"])
    normals_d, uniforms_d = up(normals), up(uniforms)
    kind = pu._kind(case)
    out_q, out_ad, out_st, out_pot = [], [], [], []
    for t in range(T):
        ch.q[:, :D] = pre_q[:, t]
        ch.adapt[:, :5] = pre_ad[:, t]
        _inject_potential(pot, [ora["pre"]["pot"][c][t] for c in range(Cn)])
        _, st = engine.run_transitions_dense(kind, ch, cb, pot, n_trans=1, iter0=t, n_tune=int(case["tune"]),
                                             params=params, tapes=(normals_d[:, t:t + 1], uniforms_d[:, t:t + 1]))
        torch.cuda.synchronize()
        out_q.append(ch.q[:, :D].cpu().numpy())
        out_ad.append(ch.adapt[:, :5].cpu().numpy())
        out_st.append(st[:, 0].cpu().numpy())
        out_pot.append(_read_potential(pot))
    return np.stack(out_q, 1), np.stack(out_ad, 1), np.stack(out_st, 1), out_pot, ch.status.cpu().numpy()
