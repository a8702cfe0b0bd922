"""Leapfrog integrator on the GPU: drop-in for reference integration.CpuLeapfrogIntegrator (integration.py:34-121).

`compute_state(q, p)` and `step(epsilon, state)` keep the reference's signatures and `State` record, accept one chain
(`[D]`) or a batch (`[C, D]`), NumPy or torch, and run through the C ABI:

* fused targets (targets.DiagGaussian / NealFunnel): lmc_compute_state / lmc_leapfrog_step, gradient inside the kernel;
* any other callback: lmc_leapfrog_half1 -> callback -> lmc_leapfrog_half2 (integration.py:105-112, 115, 116-119).

The samplers do not call these per leapfrog (that Python seam is what the reference pays ~60 us for); they launch
whole transitions (base_hmc.py).  This class is the operator slot `step.integrator` of the reference (base_hmc.py:122)
and what the reversibility test of the reference exercises (tests/test_hmc.py:23-40).
"""
import ctypes as C
from collections import namedtuple

import numpy as np
import torch

from . import _lib as L
from .engine import evaluate_callback, padded_ld
from .targets import fused_descriptor

State = namedtuple("State", "q, p, v, q_grad, energy, model_logp")


class IntegrationError(RuntimeError):
    """Numerical errors during leapfrog integration (reference integration.py:28)."""


def _vp(t):
    return C.c_void_p(t.data_ptr())


class GpuLeapfrogIntegrator:
    def __init__(self, potential, logp_dlogp_func, device=None):
        self._potential = potential
        self._logp_dlogp_func = logp_dlogp_func
        self._device = device

    # -- helpers ---------------------------------------------------------------------------------------------------
    def _dev(self):
        if self._device is not None:
            return torch.device(self._device)
        ch = getattr(self._potential, "_chains", None)
        return ch.device if ch is not None else torch.device("cuda", torch.cuda.current_device())

    def _rows(self, x, n_chains=None):
        """-> (device tensor [C, ld], was_1d, was_numpy)"""
        was_np = not torch.is_tensor(x)
        t = torch.as_tensor(np.asarray(x, dtype="d") if was_np else x, dtype=torch.float64, device=self._dev())
        one_d = t.ndim == 1
        if one_d:
            t = t.unsqueeze(0)
        Cn, D = t.shape
        out = torch.zeros(Cn, padded_ld(D), dtype=torch.float64, device=t.device)
        out[:, :D] = t
        return out, one_d, was_np

    def _var_rows(self, Cn, D):
        ch = getattr(self._potential, "_chains", None)
        if ch is not None and ch.n_chains == Cn and ch.ndim == D and Cn > 1:
            return ch.var, ch.ld
        v = torch.zeros(1, padded_ld(D), dtype=torch.float64, device=self._dev())
        v[0, :D] = torch.as_tensor(np.asarray(self._potential._current_var(), dtype="d"), device=v.device)
        return v, 0

    @staticmethod
    def _out(t, D, one_d, was_np):
        t = t[:, :D]
        if one_d:
            t = t[0]
        return t.cpu().numpy() if was_np else t

    def _callback(self, q_rows, D):
        """Evaluate a non-fused callback for every chain: -> (logp [C], grad [C, ld]) on the device."""
        logp, grad = evaluate_callback(self._logp_dlogp_func, q_rows[:, :D])
        g = torch.zeros_like(q_rows)
        g[:, :D] = grad
        return logp.contiguous(), g

    # -- dense potentials: velocity is a matrix product (quadpotential_dense.py), the rest is elementwise ---------------
    def _dense_velocity(self, rows, D):
        """rows: [C, ld] -> velocity of every row, [C, ld].  Shared-matrix potentials serve any number of rows; a
        per-chain adapted potential serves its bound chains row by row, or a single vector with the last chain's matrix."""
        pot = self._potential
        dev = rows.device
        if getattr(pot, "_dev", None) != dev:
            pot._to_device(dev, pot._chains.n_chains if pot._chains is not None else 1)
        Cn, ld = rows.shape
        per_chain = getattr(pot, "_nc", None)
        if per_chain is not None and Cn != per_chain:
            out = torch.zeros_like(rows)
            for c in range(Cn):
                out[c, :D] = pot._velocity_one(rows[c, :D].contiguous())
            return out
        x = torch.zeros(Cn, 2, ld, dtype=torch.float64, device=dev)
        x[:, 0] = rows
        v = torch.zeros_like(x)
        pot._velocity_rows(None, x, v)
        return v[:, 0].contiguous()

    def _dense_compute_state(self, q, p):
        qr, one_d, was_np = self._rows(q)
        pr, _, _ = self._rows(p)
        D = (np.asarray(q).shape if was_np else q.shape)[-1]
        logp, g = self._callback(qr, D)
        v = self._dense_velocity(pr, D)
        energy = 0.5 * (pr * v).sum(1) - logp                                       # integration.py:63-65
        o = lambda t: self._out(t, D, one_d, was_np)  # noqa: E731
        sc = lambda t: (t[0].item() if one_d else (t.cpu().numpy() if was_np else t))  # noqa: E731
        return State(q, p, o(v), o(g), sc(energy), sc(logp))

    def _dense_step(self, epsilon, state):
        qr, one_d, was_np = self._rows(state.q)
        pr, _, _ = self._rows(state.p)
        gr, _, _ = self._rows(state.q_grad)
        Cn = qr.shape[0]
        D = (np.asarray(state.q).shape if was_np else state.q.shape)[-1]
        eps = torch.as_tensor(np.broadcast_to(np.asarray(epsilon, dtype="d"), (Cn,)).copy(), device=qr.device) \
            if not torch.is_tensor(epsilon) else epsilon.to(torch.float64).expand(Cn).contiguous()
        dt = (0.5 * eps)[:, None]
        p_half = pr + dt * gr                                                       # integration.py:108
        q_new = qr + eps[:, None] * self._dense_velocity(p_half, D)                 # :111-112
        logp, g_new = self._callback(q_new, D)                                      # :115
        p_new = p_half + dt * g_new                                                 # :116
        v_new = self._dense_velocity(p_new, D)                                      # :118
        energy = 0.5 * (p_new * v_new).sum(1) - logp                                # :119-120
        o = lambda t: self._out(t, D, one_d, was_np)  # noqa: E731
        sc = lambda t: (t[0].item() if one_d else (t.cpu().numpy() if was_np else t))  # noqa: E731
        return State(o(q_new), o(p_new), o(v_new), o(g_new), sc(energy), sc(logp))

    # -- the reference API -------------------------------------------------------------------------------------------
    def compute_state(self, q, p):
        """reference integration.py:52-66."""
        if getattr(self._potential, "_dense", False):
            return self._dense_compute_state(q, p)
        lib = L.load()
        qr, one_d, was_np = self._rows(q)
        pr, _, _ = self._rows(p)
        Cn, ld = qr.shape
        D = (np.asarray(q).shape if was_np else q.shape)[-1]
        var, vstride = self._var_rows(Cn, D)
        v, g = torch.empty_like(qr), torch.zeros_like(qr)
        energy = torch.empty(Cn, dtype=torch.float64, device=qr.device)
        logp = torch.empty_like(energy)
        st = C.c_void_p(torch.cuda.current_stream(qr.device).cuda_stream)
        fused = fused_descriptor(self._logp_dlogp_func)
        with torch.cuda.device(qr.device):
            if fused is not None:
                tgt = fused.c_struct(qr.device)
                L.check(lib.lmc_compute_state(C.byref(tgt), Cn, D, ld, _vp(qr), _vp(pr), _vp(var), vstride, _vp(v),
                                              _vp(g), _vp(energy), _vp(logp), st), "lmc_compute_state")
            else:
                logp, g = self._callback(qr, D)
                zero = torch.zeros(Cn, dtype=torch.float64, device=qr.device)
                # half2 with a zero step: p unchanged, v = var*p, energy = 0.5 p.v - logp
                L.check(lib.lmc_leapfrog_half2(Cn, D, ld, _vp(zero), None, _vp(pr), _vp(v), _vp(torch.zeros_like(qr)),
                                               _vp(logp), _vp(var), vstride, _vp(energy), st), "lmc_leapfrog_half2")
        o = lambda t: self._out(t, D, one_d, was_np)  # noqa: E731
        sc = lambda t: (t[0].item() if one_d else (t.cpu().numpy() if was_np else t))  # noqa: E731
        return State(q, p, o(v), o(g), sc(energy), sc(logp))

    def step(self, epsilon, state, out=None):
        """reference integration.py:68-121 (epsilon may be negative; scalar or one value per chain)."""
        if getattr(self._potential, "_dense", False):
            return self._dense_step(epsilon, state)
        lib = L.load()
        qr, one_d, was_np = self._rows(state.q)
        pr, _, _ = self._rows(state.p)
        gr, _, _ = self._rows(state.q_grad)
        Cn, ld = qr.shape
        D = (np.asarray(state.q).shape if was_np else state.q.shape)[-1]
        var, vstride = self._var_rows(Cn, D)
        eps = torch.as_tensor(np.broadcast_to(np.asarray(epsilon, dtype="d"), (Cn,)).copy(), device=qr.device) \
            if not torch.is_tensor(epsilon) else epsilon.to(torch.float64).expand(Cn).contiguous()
        v = torch.empty_like(qr)
        energy = torch.empty(Cn, dtype=torch.float64, device=qr.device)
        logp = torch.empty_like(energy)
        st = C.c_void_p(torch.cuda.current_stream(qr.device).cuda_stream)
        fused = fused_descriptor(self._logp_dlogp_func)
        with torch.cuda.device(qr.device):
            if fused is not None:
                tgt = fused.c_struct(qr.device)
                g_new = torch.empty_like(qr)
                L.check(lib.lmc_leapfrog_step(C.byref(tgt), Cn, D, ld, _vp(eps), _vp(qr), _vp(pr), _vp(gr), _vp(var),
                                              vstride, _vp(qr), _vp(pr), _vp(v), _vp(g_new), _vp(energy), _vp(logp),
                                              st), "lmc_leapfrog_step")
            else:
                L.check(lib.lmc_leapfrog_half1(Cn, D, ld, _vp(eps), None, _vp(qr), _vp(pr), _vp(gr), _vp(var), vstride,
                                               st), "lmc_leapfrog_half1")
                logp, g_new = self._callback(qr, D)
                L.check(lib.lmc_leapfrog_half2(Cn, D, ld, _vp(eps), None, _vp(pr), _vp(v), _vp(g_new), _vp(logp),
                                               _vp(var), vstride, _vp(energy), st), "lmc_leapfrog_half2")
        o = lambda t: self._out(t, D, one_d, was_np)  # noqa: E731
        sc = lambda t: (t[0].item() if one_d else (t.cpu().numpy() if was_np else t))  # noqa: E731
        return State(o(qr), o(pr), o(v), o(g_new), sc(energy), sc(logp))
