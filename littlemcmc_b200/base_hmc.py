"""Common machinery of the HMC step methods: mirror of reference base_hmc.py:28-230 with chains as a tensor dimension.

A step method owns `n_chains` chains on one GPU (engine.DeviceChains).  `_astep` keeps the reference's per-transition
contract; `_run` is the batched entry the driver uses: `n` transitions of every chain in one kernel launch, including
momentum draws, both adaptations, the trace and the statistics.
"""
from collections import namedtuple

import numpy as np
import torch

from . import _lib as L
from . import engine, integration, step_sizes
from .quadpotential import QuadPotentialDiagAdapt, quad_potential
from .report import SamplerWarning, WarningType
from .targets import fused_descriptor

HMCStepData = namedtuple("HMCStepData", "end, accept_stat, divergence_info, stats")
DivergenceInfo = namedtuple("DivergenceInfo", "message, exec_info, state")


class BaseHMC:
    _kind = None          # L.KIND_NUTS / L.KIND_HMC
    _stat_columns = {}    # stat name -> column of the kernel's stats rows

    def __init__(self, logp_dlogp_func, model_ndim, scaling, is_cov, potential, target_accept, Emax, adapt_step_size,
                 step_scale, gamma, k, t0, step_rand):
        """Same arguments and defaults as reference base_hmc.py:31-126."""
        self._logp_dlogp_func = logp_dlogp_func
        self.adapt_step_size = adapt_step_size
        self.Emax = Emax
        self.iter_count = 0
        self.model_ndim = int(model_ndim)
        self.step_size = step_scale / (model_ndim ** 0.25)            # base_hmc.py:102
        self.target_accept = target_accept
        self.step_adapt = step_sizes.DualAverageAdaptation(self.step_size, target_accept, gamma, k, t0)
        self.tune = True
        if scaling is None and potential is None:                     # base_hmc.py:109-113
            potential = QuadPotentialDiagAdapt(model_ndim, np.zeros(model_ndim), np.ones(model_ndim), 10)
        if scaling is not None and potential is not None:
            raise ValueError("Cannot specify both `potential` and `scaling`.")
        self.potential = potential if potential is not None else quad_potential(np.asarray(scaling), is_cov)
        self.integrator = integration.GpuLeapfrogIntegrator(self.potential, self._logp_dlogp_func)
        # base_hmc.py:154-155: `step_size = self.step_rand(step_size)` before every transition.  Here the hook gets the
        # step sizes of ALL chains as one float64 array [n_chains] (a function written for a scalar that uses NumPy
        # arithmetic works unchanged; one that does not is called once per chain).  A hook forces one launch per
        # transition, because it is host code that must run between transitions.
        if step_rand is not None and not callable(step_rand):
            raise TypeError("step_rand must be callable")
        self._step_rand = step_rand
        self._warnings = []
        self._samples_after_tune = 0
        self._num_divs_sample = 0
        self._chains = None
        self._seeds = None
        self._knobs = {}

    # ---- device binding ----------------------------------------------------------------------------------------------
    def _bind(self, n_chains, device=None, seeds=None):
        """Allocate the device state of `n_chains` chains and attach potential / step-size adaptation to it."""
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else "cpu"
        if (self._chains is None or self._chains.n_chains != n_chains
                or self._chains.device != torch.device(device)):
            self._chains = engine.DeviceChains(n_chains, self.model_ndim, device)
            self.potential._bind(self._chains)
            self.step_adapt._bind(self._chains)
        if seeds is None and self._seeds is None:
            seeds = np.random.randint(2 ** 30, size=n_chains)
        if seeds is not None:
            seeds = np.asarray(seeds)
            if seeds.shape != (n_chains,):
                raise ValueError("need one seed per chain")
            self._seeds = engine.seeds_tensor(seeds, self._chains.device)
        return self._chains

    def _fused_target(self):
        fused = fused_descriptor(self._logp_dlogp_func)
        if fused is not None and fused.ndim != self.model_ndim:
            raise ValueError("target has %d dimensions, step method %d" % (fused.ndim, self.model_ndim))
        return fused

    def _params(self):
        sa, pot = self.step_adapt, self.potential
        return dict(adapt_mass=int(pot._adaptive), adapt_step_size=int(bool(self.adapt_step_size)),
                    window_multiplier=getattr(pot, "adaptation_window_multiplier", 1.0),
                    target_accept=sa._target, gamma=sa._gamma, k=sa._k, t0=sa._t0, Emax=self.Emax)

    # ---- batched driver entry ----------------------------------------------------------------------------------------
    def _run(self, n_trans, n_tune, tapes=None, trace=None, stats=None, events=None, trace_skip=0, progress=None,
             progress_block=0):
        """`n_trans` transitions of every chain starting at `self.iter_count`; transitions with index < `n_tune` tune.
        Returns device tensors (trace [C, n_trans, D], stats [C, n_trans, NSTATS]).

        A target with a fused descriptor (targets.DiagGaussian / NealFunnel) runs whole transitions inside one kernel
        launch (asynchronous).  Any other callback -- a targets.TorchBatched device op, or the reference's per-chain
        NumPy callable -- runs in callback mode: the callback is evaluated for all chains between two launches
        (engine.CallbackRun); that path returns when every chain has finished."""
        fused = self._fused_target()
        if self._step_rand is not None and n_trans > 1:
            return self._run_with_step_rand(n_trans, n_tune, tapes, trace, stats)
        override = None
        if self._step_rand is not None:
            override = self._apply_step_rand(self.iter_count < n_tune)
        self._last_override = override
        common = dict(n_trans=n_trans, iter0=self.iter_count, n_tune=n_tune, params=self._params(), seeds=self._seeds,
                      tapes=tapes, trace=trace, stats=stats, step_size_override=override)
        if getattr(self.potential, "_dense", False):
            # dense mass matrix: velocity / momentum draw / matrix update are batched operations next to the gradient
            cb = self._logp_dlogp_func
            if fused is not None:            # a built-in density: evaluate it as one torch op over the chains that ask
                key = str(self._chains.device)
                if getattr(self, "_dense_cb_key", None) != key:
                    self._dense_cb, self._dense_cb_key = cb.torch_batched(self._chains.device), key
                cb = self._dense_cb
            if events is not None:
                events[0].record()
            tr, st = engine.run_transitions_dense(self._kind, self._chains, cb, self.potential, **common)
            if events is not None:
                events[1].record()
        elif fused is not None:
            tr, st = engine.run_transitions(self._kind, self._chains, fused, knobs=self._knobs, events=events,
                                            trace_skip=trace_skip, progress=progress, progress_block=progress_block,
                                            **common)
        else:
            graph = getattr(self._logp_dlogp_func, "cuda_graph", False)
            if events is not None:
                events[0].record()
            tr, st = engine.run_transitions_callback(self._kind, self._chains, self._logp_dlogp_func, cuda_graph=graph,
                                                     **common)
            if events is not None:
                events[1].record()
        self.iter_count += n_trans
        return tr, st

    def _apply_step_rand(self, tuning):
        """Current step size of every chain (step_sizes.py:58-69) passed through the user's hook (base_hmc.py:154-155)."""
        eps = self.step_adapt.current_all(tuning and bool(self.adapt_step_size)).cpu().numpy()
        out = None
        try:
            out = np.asarray(self._step_rand(eps), dtype="d")
        except Exception:           # a hook written for Python scalars only
            out = None
        if out is None or out.shape != eps.shape:
            out = np.array([float(self._step_rand(float(e))) for e in eps], dtype="d")
        return torch.as_tensor(out, device=self._chains.device)

    def _run_with_step_rand(self, n_trans, n_tune, tapes, trace, stats):
        """One launch per transition, the hook evaluated on the host in between."""
        Cn, D, dev = self._chains.n_chains, self.model_ndim, self._chains.device
        if trace is None:
            trace = torch.empty(Cn, n_trans, D, dtype=torch.float64, device=dev)
        if stats is None:
            stats = torch.empty(Cn, n_trans, L.NSTATS, dtype=torch.float64, device=dev)
        for t in range(n_trans):
            tp = None if tapes is None else (tapes[0][:, t:t + 1], tapes[1][:, t:t + 1])
            st1 = torch.empty(Cn, 1, L.NSTATS, dtype=torch.float64, device=dev)
            self._run(1, n_tune, tapes=tp, trace=trace[:, t:t + 1], stats=st1)
            stats[:, t:t + 1] = st1
        return trace, stats

    def _check_status(self):
        status = self._chains.status
        bad = (status & L.STATUS_BAD_INITIAL_ENERGY) != 0
        if bool(bad.any()):
            idx = bad.nonzero().flatten().tolist()
            raise ValueError("Bad initial energy in chain(s) %s. The model might be misspecified." % idx[:8])
        if bool(((status & L.STATUS_TAPE_EXHAUSTED) != 0).any()):
            raise L.LmcError("uniform tape exhausted")

    def _account(self, stats_dev, n_tune_in_block):
        """Host bookkeeping the reference does per transition (base_hmc.py:164-183, nuts.py:218-220)."""
        st = stats_dev
        post = st[:, n_tune_in_block:, :]
        if post.shape[1]:
            # counted over ALL chains, like the divergence / max-treedepth counters they are compared with (the
            # reference's sequential path accumulates every chain into the same step object, base_hmc.py:164-183)
            self._samples_after_tune += post.shape[0] * post.shape[1]
            self._num_divs_sample += int(post[:, :, L.STAT_DIVERGING].sum().item())
            self.step_adapt._tuned_stats.extend(post[-1, :, L.STAT_ACCEPT].cpu().tolist())

    def stats_dict(self, stats_dev, chain=None):
        """Kernel stats rows -> {name: array} with the reference's names."""
        return {n: stats_dev[..., c] for n, c in self._stat_columns.items()}

    # ---- the reference's per-transition API --------------------------------------------------------------------------
    def _astep(self, q0):
        """One transition (reference base_hmc.py:140-190).  `q0`: [D] (one chain) or [C, D]."""
        q0a = np.asarray(q0.cpu() if torch.is_tensor(q0) else q0, dtype="d")
        one_d = q0a.ndim == 1
        n_chains = 1 if one_d else q0a.shape[0]
        if self._chains is None or self._chains.n_chains != n_chains:
            self._bind(n_chains)
        self._chains.set_position(q0a)
        n_tune = self.iter_count + 1 if self.tune else 0
        # base_hmc.py:152-156: step.step_size is the step size THIS transition integrates with (current() before the
        # dual-averaging update, after the step_rand hook); last chain, as the reference's sequential path leaves it
        eps_used = float(self.step_adapt.current_all(bool(self.tune and self.adapt_step_size))[-1].item())
        tr, st = self._run(1, n_tune)
        self._check_status()
        self._account(st, 1 if self.tune else 0)
        self.step_size = eps_used if self._last_override is None else float(self._last_override[-1].item())
        sd = {n: st[:, 0, c].cpu().numpy().astype(self.stats_dtypes[0][n]) for n, c in self._stat_columns.items()}
        if one_d:
            sd = {n: v[0] for n, v in sd.items()}
            return tr[0, 0].cpu().numpy(), [sd]
        return tr[:, 0].cpu().numpy(), [sd]

    def stop_tuning(self):
        if hasattr(self, "tune"):
            self.tune = False

    def reset_tuning(self, start=None):
        """reference base_hmc.py:192-195."""
        self.step_adapt.reset()
        self.reset(start=None)

    def reset(self, start=None):
        """reference base_hmc.py:197-200."""
        self.tune = True
        self.potential.reset()

    def warnings(self):
        """reference base_hmc.py:202-230."""
        warnings = list(self._warnings)
        n_divs = self._num_divs_sample
        message = ""
        if n_divs and self._samples_after_tune == n_divs:
            message = "The chain contains only diverging samples. The model is probably misspecified."
        elif n_divs == 1:
            message = "There was 1 divergence after tuning. Increase `target_accept` or reparameterize."
        elif n_divs > 1:
            message = ("There were %s divergences after tuning. Increase `target_accept` or reparameterize." % n_divs)
        if message:
            warnings.append(SamplerWarning(WarningType.DIVERGENCES, message, "error", None, None, None))
        warnings.extend(self.step_adapt.warnings())
        return warnings
