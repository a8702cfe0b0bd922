"""Diagonal quadpotentials (mass matrices).  Mirror of the diagonal half of reference quadpotential.py:33-387.

The objects below are host-side descriptors: constructor arguments, validation and the single-vector convenience
methods of the reference API (`velocity`, `energy`, `velocity_energy`, `random`).  When a step method binds them to
`n_chains` chains their arrays become rows of [n_chains, ld] device tensors (engine.DeviceChains) and
`velocity`/`energy`/`update` for the sampler run inside the CUDA kernels.  Everything is float64: the reference's
float32 default (quadpotential.py:175-176) only injects rounding noise into the first state of each draw
 and is not reproduced; `dtype` is accepted for signature compatibility.

Dense potentials (QuadPotentialFull / FullInv / FullAdapt) live in quadpotential_dense.py.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L

__all__ = ["quad_potential", "QuadPotentialDiag", "QuadPotentialDiagAdapt", "QuadPotentialFull",
           "QuadPotentialFullInv", "QuadPotentialFullAdapt", "PositiveDefiniteError", "isquadpotential"]


class PositiveDefiniteError(ValueError):
    def __init__(self, msg, idx):
        super().__init__(msg)
        self.idx, self.msg = idx, msg

    def __str__(self):
        return "Scaling is not positive definite: %s. Check indexes %s." % (self.msg, self.idx)


def partial_check_positive_definite(C_):
    """reference quadpotential.py:69-78."""
    d = C_ if C_.ndim == 1 else np.diag(C_)
    (bad,) = np.nonzero(np.logical_or(np.isnan(d), d <= 0))
    if len(bad):
        raise PositiveDefiniteError("Simple check failed. Diagonal contains negatives", bad)


def quad_potential(C_, is_cov):
    """reference quadpotential.py:33-66: build a potential from a scaling vector or matrix."""
    C_ = np.asarray(C_)
    partial_check_positive_definite(C_)
    if C_.ndim == 1:
        return QuadPotentialDiag(C_ if is_cov else 1.0 / C_)
    return QuadPotentialFull(C_) if is_cov else QuadPotentialFullInv(C_)


class QuadPotential:
    """Base class (reference quadpotential.py:93-140)."""

    _adaptive = False

    def update(self, sample, grad, tune):
        pass

    def raise_ok(self, vmap=None):
        return None

    def reset(self):
        pass

    # ---- device plumbing -------------------------------------------------------------------------------------
    _chains = None

    def _bind(self, chains):
        self._chains = chains
        self.reset()

    def _current_var(self):
        raise NotImplementedError

    # ---- single-vector API of the reference, evaluated by the CUDA library -----------------------------------
    def _velocity_energy(self, x):
        """v = var * x and 0.5 x.v through lmc_leapfrog_half2 with a zero step (integration.py:118-119)."""
        lib = L.load()
        x = np.asarray(x, dtype="d")
        n = x.shape[0]
        dev = self._chains.device if self._chains is not None else torch.device("cuda", torch.cuda.current_device())
        ld = n + (n & 1)
        buf = torch.zeros(6, ld, dtype=torch.float64, device=dev)  # p, v, g(=0), var, [eps, logp, energy]
        buf[0, :n] = torch.as_tensor(x, device=dev)
        buf[3, :n] = torch.as_tensor(np.asarray(self._current_var(), dtype="d"), device=dev)
        scal = torch.zeros(3, dtype=torch.float64, device=dev)
        p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
        with torch.cuda.device(dev):
            L.check(lib.lmc_leapfrog_half2(1, n, ld, p(scal[0:1]), None, p(buf[0]), p(buf[1]), p(buf[2]), p(scal[1:2]),
                                           p(buf[3]), 0, p(scal[2:3]),
                                           C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
                    "lmc_leapfrog_half2")
        return buf[1, :n].cpu().numpy(), float(scal[2].item())

    def velocity(self, x, out=None):
        v, _ = self._velocity_energy(x)
        if out is not None:
            out[:] = v
            return out
        return v

    def energy(self, x, velocity=None):
        if velocity is not None:
            return 0.5 * float(np.dot(x, velocity))
        return self._velocity_energy(x)[1]

    def velocity_energy(self, x, v_out):
        v, e = self._velocity_energy(x)
        v_out[:] = v
        return e

    def random(self):
        """One momentum draw with NumPy's global stream, like the reference (quadpotential.py:221-224, 374-376).
        The sampler itself draws momenta inside the kernel (per-chain Philox streams)."""
        var = np.asarray(self._current_var(), dtype="d")
        return (1.0 / np.sqrt(var)) * np.random.normal(size=var.shape[0])


def isquadpotential(value):
    return isinstance(value, QuadPotential)


class QuadPotentialDiag(QuadPotential):
    """Static diagonal covariance `v` (reference quadpotential.py:346-387)."""

    def __init__(self, v, dtype=None):
        self.dtype = "float64"
        v = np.asarray(v, dtype="d")
        self.v = v
        self.s = v ** 0.5
        self.inv_s = 1.0 / self.s

    def _current_var(self):
        return self.v

    def reset(self):
        if self._chains is not None:
            self._chains.reset_potential(self.v, np.zeros_like(self.v), 0.0, 101)


class QuadPotentialDiagAdapt(QuadPotential):
    """Diagonal mass matrix adapted from the running sample variance (reference quadpotential.py:148-291)."""

    _adaptive = True

    def __init__(self, n, initial_mean, initial_diag=None, initial_weight=0, adaptation_window=101,
                 adaptation_window_multiplier=1, dtype=None):
        initial_mean = np.asarray(initial_mean)
        if initial_diag is not None and np.asarray(initial_diag).ndim != 1:
            raise ValueError("Initial diagonal must be one-dimensional.")
        if initial_mean.ndim != 1:
            raise ValueError("Initial mean must be one-dimensional.")
        if initial_diag is not None and len(initial_diag) != n:
            raise ValueError("Wrong shape for initial_diag: expected %s got %s" % (n, len(initial_diag)))
        if len(initial_mean) != n:
            raise ValueError("Wrong shape for initial_mean: expected %s got %s" % (n, len(initial_mean)))
        if initial_diag is None:
            initial_diag, initial_weight = np.ones(n), 1
        self.dtype = "float64"
        self._n = int(n)
        self._initial_mean = np.array(initial_mean, dtype="d")
        self._initial_diag = np.array(initial_diag, dtype="d")
        self._initial_weight = initial_weight
        self.adaptation_window = int(adaptation_window)
        self.adaptation_window_multiplier = float(adaptation_window_multiplier)

    def reset(self):
        """reference quadpotential.py:195-204, applied to every bound chain."""
        if self._chains is not None:
            self._chains.reset_potential(self._initial_diag, self._initial_mean, float(self._initial_weight),
                                         self.adaptation_window)

    # views of the LAST chain's state under the reference's attribute names
    def _current_var(self):
        if self._chains is None:
            return self._initial_diag
        return self._chains.var[-1, : self._n].cpu().numpy()

    _var = property(_current_var)
    _stds = property(lambda self: np.sqrt(self._current_var()))
    _inv_stds = property(lambda self: 1.0 / np.sqrt(self._current_var()))

    @property
    def _n_samples(self):
        return 0 if self._chains is None else int(self._chains.adapt[-1, L.ADAPT_NSAMPLES].item())

    def var_all(self):
        """[n_chains, n] device view of every chain's adapted variance."""
        return self._chains.var[:, : self._n]


from .quadpotential_dense import QuadPotentialFull, QuadPotentialFullAdapt, QuadPotentialFullInv  # noqa: E402,F401
