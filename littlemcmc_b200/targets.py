"""Target densities.

Every target here is a plain ``logp_dlogp_func``: calling it with a NumPy vector ``q[D]`` returns ``(logp, dlogp[D])``
exactly like the callbacks the reference takes (base_hmc.py:34, integration.py:40), so the same object can be handed
to the reference sampler and to this one.  In addition it carries a ``fused`` descriptor; when a step method sees one
it evaluates the density INSIDE the CUDA kernels (no per-leapfrog callback), which is the throughput path.
Arbitrary callables are supported through :class:`TorchBatched` / the NumPy adapter (callback mode).
"""
import numpy as np

from . import _lib as L
from .engine import FusedTarget


class DiagGaussian:
    """logp(q) = -1/2 sum_i q_i^2 / sigma_i^2.  g = -(tau * q), logp = 0.5 * q.g with tau = 1/sigma^2."""

    def __init__(self, sigma=None, tau=None):
        if (sigma is None) == (tau is None):
            raise ValueError("give exactly one of sigma / tau")
        self.tau = np.asarray(1.0 / np.asarray(sigma, dtype="d") ** 2 if tau is None else tau, dtype="d")
        if self.tau.ndim != 1:
            raise ValueError("sigma / tau must be one-dimensional")
        self.ndim = self.tau.shape[0]
        self.fused = FusedTarget(L.TARGET_DIAG_GAUSSIAN, self.ndim, tau=self.tau)

    def __call__(self, q):
        g = -(self.tau * q)
        return 0.5 * np.dot(q, g), g

    def torch_batched(self, device, cuda_graph=False):
        """The same density as a batched torch op (callback mode instead of the fused kernel)."""
        import torch
        tau = torch.as_tensor(self.tau, dtype=torch.float64, device=device)

        def fn(q):
            g = -(tau * q)
            return 0.5 * (q * g).sum(1), g
        return TorchBatched(fn, cuda_graph=cuda_graph)


class StdNormal(DiagGaussian):
    """Isotropic standard normal (BASELINE config 1)."""

    def __init__(self, ndim):
        super().__init__(tau=np.ones(int(ndim)))


class NealFunnel:
    """q[0] = v ~ N(0, v_scale^2); q[1:] | v ~ N(0, e^v)   (BASELINE config 4)."""

    def __init__(self, ndim, v_scale=3.0):
        self.ndim, self.v_scale = int(ndim), float(v_scale)
        self.fused = FusedTarget(L.TARGET_FUNNEL, self.ndim, v_scale=self.v_scale)

    def __call__(self, q):
        inv_s2 = 1.0 / (self.v_scale * self.v_scale)
        half_nm1 = 0.5 * (self.ndim - 1)
        v, x = q[0], q[1:]
        S = np.dot(x, x)
        with np.errstate(over="ignore", invalid="ignore"):
            ev = np.exp(-v)
            g = np.empty_like(q)
            g[1:] = -(ev * x)
            hs = 0.5 * ev * S
            g[0] = -(v * inv_s2) + hs - half_nm1
            logp = -(0.5 * v * v * inv_s2) - hs - half_nm1 * v
        return logp, g

    def torch_batched(self, device=None, cuda_graph=False):
        """The same density as a batched torch op (callback mode instead of the fused kernel)."""
        import torch
        inv_s2, half_nm1 = 1.0 / (self.v_scale * self.v_scale), 0.5 * (self.ndim - 1)

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


class TorchBatched:
    """Marks a batched device callback: ``fn(q: torch.Tensor[C, D] float64 cuda) -> (logp[C], grad[C, D])``.

    It is evaluated on the current CUDA stream between two launches of the sampler's state-machine kernel (callback
    mode, csrc/lmc_callback.cu): one call per leapfrog step for ALL chains, instead of the reference's one call per
    leapfrog step per chain (integration.py:115).  ``cuda_graph=True`` (= ``"device"``) captures (callback + kernel) in
    a CUDA graph and runs it as the body of a WHILE conditional node: the whole run is ONE graph launch that loops on the
    device until every chain has finished -- no host round trip per gradient.  ``cuda_graph="replay"``: graphs of 8
    iterations replayed from the host.  Either way the callback must be capture-safe (no host syncs, no data-dependent
    shapes)."""

    def __init__(self, fn, cuda_graph=False):
        self.fn = fn
        if cuda_graph not in (False, True, "device", "replay"):
            raise ValueError("cuda_graph must be False, True, 'device' or 'replay'")
        self.cuda_graph = cuda_graph

    def __call__(self, q):
        return self.fn(q)

    @classmethod
    def from_logp(cls, logp_fn, cuda_graph=False):
        """Build the callback from a scalar-per-chain log density ``logp_fn(q[C, D]) -> logp[C]`` with autograd
        (chains are independent, so the gradient of ``logp.sum()`` is every chain's own gradient)."""
        import torch

        def fn(q):
            with torch.enable_grad():
                x = q.detach().requires_grad_(True)
                lp = logp_fn(x)
                (g,) = torch.autograd.grad(lp.sum(), x)
            return lp.detach(), g
        return cls(fn, cuda_graph=cuda_graph)


def fused_descriptor(logp_dlogp_func):
    return getattr(logp_dlogp_func, "fused", None)
