"""Dense quadpotentials (mass matrices): mirror of reference quadpotential.py:390-615 with chains as a tensor dimension.

`QuadPotentialFull(cov)` and `QuadPotentialFullInv(A)` hold ONE static matrix shared by every chain; their velocity /
momentum operations over all chains are plain library GEMM-shaped calls (cuBLAS / cuSOLVER through torch).
`QuadPotentialFullAdapt` holds one covariance, one Cholesky factor and two Welford covariance estimators PER CHAIN
(`[n_chains, n, n]` device tensors): its velocity is the hand-written batched matrix-vector kernel `lmc_dense_matvec`
(HBM-bound: 8 n^2 bytes per chain and leapfrog) and its update the `lmc_dense_cov_update` kernel followed by a batched
Cholesky factorisation.

The sampler drives these objects through three batched hooks (engine.DenseRun, include/lmc_b200.h lmc_dense_args):
`_velocity_rows`, `_momentum_rows`, `_update_rows`, each for the subset of chains that asked for it.  The single-vector
methods of the reference API (`velocity`, `energy`, `velocity_energy`, `random`) run on the device as well and refer
to the LAST chain's matrix, which is what the reference object holds after sequential sampling.  Everything is float64
(`dtype` is accepted for signature compatibility).

One deliberate difference: the reference never resets a QuadPotentialFullAdapt between chains (its `reset` is the
base-class no-op, quadpotential.py:138-140), so sequential chains inherit the previous chain's adapted matrix; here
every chain starts from the initial matrix, like QuadPotentialDiagAdapt does in the reference.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L
from .quadpotential import QuadPotential


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _even(n):
    return int(n) + (int(n) & 1)


class _DenseBase(QuadPotential):
    _dense = True
    _adaptive = False

    def _device(self):
        if self._chains is not None:
            return self._chains.device
        if not torch.cuda.is_available():
            raise L.LmcError("littlemcmc_b200 runs on CUDA devices only; there is no CPU fallback")
        return torch.device("cuda", torch.cuda.current_device())

    def _bind(self, chains):
        self._chains = chains
        self._to_device(chains.device, chains.n_chains)
        self.reset()

    def _to_device(self, device, n_chains):
        raise NotImplementedError

    # ---- single-vector API of the reference ----------------------------------------------------------------------
    def _single(self, x):
        dev = self._device()
        if getattr(self, "_dev", None) != dev:
            self._to_device(dev, self._chains.n_chains if self._chains is not None else 1)
        return torch.as_tensor(np.asarray(x, dtype="d"), device=dev)

    def velocity(self, x, out=None):
        v = self._velocity_one(self._single(x)).cpu().numpy()
        if out is not None:
            out[:] = v
            return out
        return v

    def energy(self, x, velocity=None):
        if velocity is None:
            velocity = self.velocity(x)
        return 0.5 * float(np.dot(np.asarray(x, dtype="d"), velocity))

    def velocity_energy(self, x, v_out):
        self.velocity(x, out=v_out)
        return 0.5 * float(np.dot(np.asarray(x, dtype="d"), v_out))

    def random(self):
        """One momentum draw with NumPy's global stream, like the reference; the sampler draws its own."""
        n = self._single(np.random.normal(size=self._n))
        return self._momentum_one(n).cpu().numpy()

    __call__ = random


class QuadPotentialFull(_DenseBase):
    """Static dense covariance (reference quadpotential.py:430-468): velocity = cov @ x, random = chol^-T n."""

    # engine.DenseRun serves every chain's row (no gather / scatter) when at least this fraction of the chains asks.
    # Measured at 1024 chains x 1000 dimensions: gathering wins at any fraction below 1 (0.31 s against 0.35 s)
    _all_rows_above = 1.0

    def __init__(self, cov, dtype=None):
        self.dtype = "float64"
        self._cov_host = np.array(cov, dtype="d", copy=True)
        if self._cov_host.ndim != 2 or self._cov_host.shape[0] != self._cov_host.shape[1]:
            raise ValueError("cov must be a square matrix")
        self._n = len(self._cov_host)
        self._dev = None

    def _to_device(self, device, n_chains):
        self._dev = device
        self._cov = torch.as_tensor(self._cov_host, device=device)
        self._chol = torch.linalg.cholesky(self._cov)                    # quadpotential.py:446
        # the matrix never changes: its momentum draws chol^-T n are served as one GEMM with the explicit inverse
        # factor (chains ask for them a few at a time; a triangular solve costs ~0.4 ms per call whatever the count)
        eye = torch.eye(self._n, dtype=torch.float64, device=device)
        self._chol_inv = torch.linalg.solve_triangular(self._chol, eye, upper=False)

    def _velocity_one(self, x):
        return self._cov @ x

    def _momentum_one(self, n):
        return torch.linalg.solve_triangular(self._chol.mT, n[:, None], upper=True)[:, 0]

    # batched hooks: x_eval / v_eval [C, 2, ld], n_eval / p0_eval [C, ld]; idx: int64 device tensor or None (= all)
    def _velocity_rows(self, idx, x_eval, v_eval):
        D = self._n
        xs = x_eval[:, :, :D] if idx is None else x_eval[idx][:, :, :D]
        vs = xs @ self._cov.mT                                           # np.dot(cov, x) for every row (:449-451)
        if idx is None:
            v_eval[:, :, :D] = vs
        else:
            v_eval[idx, :, :D] = vs

    def _momentum_rows(self, idx, n_eval, p0_eval):
        D = self._n
        ns = n_eval[:, :D] if idx is None else n_eval[idx][:, :D]
        ps = ns @ self._chol_inv                      # rows of solve_triangular(chol.T, n) (:455-456): n^T chol^-1
        if idx is None:
            p0_eval[:, :D] = ps
        else:
            p0_eval[idx, :D] = ps

    def _update_rows(self, idx, q):
        pass


class QuadPotentialFullInv(_DenseBase):
    """Static dense inverse covariance A (reference quadpotential.py:390-427): velocity = A^-1 x, random = L n."""

    _all_rows_above = 1.0

    def __init__(self, A, dtype=None):
        self.dtype = "float64"
        self._A_host = np.array(A, dtype="d", copy=True)
        if self._A_host.ndim != 2 or self._A_host.shape[0] != self._A_host.shape[1]:
            raise ValueError("A must be a square matrix")
        self._n = len(self._A_host)
        self._dev = None

    def _to_device(self, device, n_chains):
        self._dev = device
        self._A = torch.as_tensor(self._A_host, device=device)
        self.L = torch.linalg.cholesky(self._A)                          # quadpotential.py:405
        # static matrix: cho_solve((L, True), x) for many rows = one GEMM with A^-1 formed once from the factor
        self._A_inv = torch.cholesky_inverse(self.L)

    def _velocity_one(self, x):
        return torch.cholesky_solve(x[:, None], self.L)[:, 0]

    def _momentum_one(self, n):
        return self.L @ n

    def _velocity_rows(self, idx, x_eval, v_eval):
        D = self._n
        xs = x_eval[:, :, :D] if idx is None else x_eval[idx][:, :, :D]
        vs = xs @ self._A_inv.mT                                         # rows of cho_solve((L, True), x) (:409)
        if idx is None:
            v_eval[:, :, :D] = vs
        else:
            v_eval[idx, :, :D] = vs

    def _momentum_rows(self, idx, n_eval, p0_eval):
        D = self._n
        ns = n_eval[:, :D] if idx is None else n_eval[idx][:, :D]
        ps = ns @ self.L.mT                                              # np.dot(L, n) (:416-417)
        if idx is None:
            p0_eval[:, :D] = ps
        else:
            p0_eval[idx, :D] = ps

    def _update_rows(self, idx, q):
        pass


class QuadPotentialFullAdapt(_DenseBase):
    """Dense mass matrix adapted from the running sample covariance (reference quadpotential.py:471-570), one matrix
    per chain."""

    _adaptive = True
    # engine.DenseRun holds chains that ask for their update until this fraction of the running chains waits for one
    # (one batched Cholesky costs the same for 8 matrices as for 256); 0 = serve every request at once
    _update_batch_fraction = 0.25

    def __init__(self, n, initial_mean, initial_cov=None, initial_weight=0, adaptation_window=101,
                 adaptation_window_multiplier=2, update_window=1, dtype=None):
        initial_mean = np.asarray(initial_mean)
        if initial_cov is not None and np.asarray(initial_cov).ndim != 2:
            raise ValueError("Initial covariance must be two-dimensional.")
        if initial_mean.ndim != 1:
            raise ValueError("Initial mean must be one-dimensional.")
        if initial_cov is not None and np.asarray(initial_cov).shape != (n, n):
            raise ValueError("Wrong shape for initial_cov: expected %s got %s" % (n, np.asarray(initial_cov).shape))
        if len(initial_mean) != n:
            raise ValueError("Wrong shape for initial_mean: expected %s got %s" % (n, len(initial_mean)))
        if initial_cov is None:                                          # :500-502
            initial_cov, initial_weight = np.eye(n), 1
        self.dtype = "float64"
        self._n = int(n)
        self._initial_mean = np.array(initial_mean, dtype="d")
        self._initial_cov = np.array(initial_cov, dtype="d")
        self._initial_weight = float(initial_weight)
        self._initial_window = int(adaptation_window)
        self._adaptation_window_multiplier = float(adaptation_window_multiplier)
        self._update_window = int(update_window)
        self._chol_error = None
        self._dev = None

    def _to_device(self, device, n_chains):
        self._dev, self._nc = device, int(n_chains)
        D, lda, ld = self._n, _even(self._n), _even(self._n)
        z = lambda *s: torch.zeros(*s, dtype=torch.float64, device=device)  # noqa: E731
        self._lda, self._ld = lda, ld
        self._cov_all, self._chol_all = z(self._nc, D, lda), z(self._nc, D, D)
        self._raw_fg, self._raw_bg = z(self._nc, D, lda), z(self._nc, D, lda)
        self._mean_fg, self._mean_bg = z(self._nc, ld), z(self._nc, ld)
        self._nsamp = z(self._nc, 2)
        self._work = None                              # gather / factor buffers of _update_rows, allocated on first use
        self._reset_state()

    def _reset_state(self):
        D, dev = self._n, self._dev
        cov0 = torch.as_tensor(self._initial_cov, device=dev)
        self._cov_all.zero_()
        self._cov_all[:, :, :D] = cov0
        self._chol_all[:] = torch.linalg.cholesky(cov0)                  # :509
        self._raw_fg.zero_()
        self._raw_fg[:, :, :D] = cov0 * self._initial_weight             # _WeightedCovariance.__init__ (:600)
        self._raw_bg.zero_()                                             # eye * n_samples(0)
        self._mean_fg.zero_()
        self._mean_fg[:, :D] = torch.as_tensor(self._initial_mean, device=dev)
        self._mean_bg.zero_()
        self._nsamp[:, 0] = self._initial_weight
        self._nsamp[:, 1] = 0.0
        # host-side per-chain counters (:513-518)
        self._n_samples_all = np.zeros(self._nc, dtype=np.int64)
        self._previous_update_all = np.zeros(self._nc, dtype=np.int64)
        self._window_all = np.full(self._nc, self._initial_window, dtype=np.int64)
        self._chol_error = None

    def reset(self):
        """Every bound chain back to the initial matrix (see the module docstring for how this differs from the
        reference's no-op)."""
        if self._dev is not None:
            self._reset_state()

    # views of the LAST chain under the reference's attribute names
    _cov = property(lambda self: self._cov_all[-1, :, :self._n].cpu().numpy())
    _chol = property(lambda self: self._chol_all[-1].cpu().numpy())
    _n_samples = property(lambda self: int(self._n_samples_all[-1]))
    _adaptation_window = property(lambda self: int(self._window_all[-1]))
    _previous_update = property(lambda self: int(self._previous_update_all[-1]))

    def _velocity_one(self, x):
        return self._cov_all[-1, :, :self._n] @ x

    def _momentum_one(self, n):
        return torch.linalg.solve_triangular(self._chol_all[-1].mT, n[:, None], upper=True)[:, 0]

    def _velocity_rows(self, idx, x_eval, v_eval):
        """v_eval[c, r] = cov_c @ x_eval[c, r] for the listed chains: the hand-written batched matvec (in place, no
        gather: every listed chain's 8 n^2-byte matrix is streamed from HBM exactly once for both vectors)."""
        lib = L.load()
        idx32 = None if idx is None else idx.to(torch.int32)
        n_idx = self._nc if idx is None else int(idx.numel())
        stream = C.c_void_p(torch.cuda.current_stream(self._dev).cuda_stream)
        L.check(lib.lmc_dense_matvec(_ptr(idx32), n_idx, _ptr(self._cov_all), self._n * self._lda, self._lda, self._n,
                                     x_eval.shape[2], _ptr(x_eval), _ptr(v_eval), 2, stream), "lmc_dense_matvec")
        from . import engine
        engine.LAUNCH_COUNT["kernels"] += 1
        if idx32 is not None:
            idx32.record_stream(torch.cuda.current_stream(self._dev))

    def _momentum_rows(self, idx, n_eval, p0_eval):
        D = self._n
        ns = n_eval[:, :D] if idx is None else n_eval[idx][:, :D]
        if idx is None:
            chol = self._chol_all
        else:                                   # gathered into the long-lived work buffer (see _update_rows)
            if self._work is None:
                self._work = (torch.empty_like(self._cov_all), torch.empty_like(self._chol_all),
                              torch.empty(self._nc, dtype=torch.int32, device=self._dev))
            chol = torch.index_select(self._chol_all, 0, idx, out=self._work[1][:int(idx.numel())])
        ps = torch.linalg.solve_triangular(chol.mT, ns[:, :, None], upper=True)[:, :, 0]   # :455-456 per chain
        if idx is None:
            p0_eval[:, :D] = ps
        else:
            p0_eval[idx, :D] = ps

    def _update_rows(self, idx, q):
        """potential.update(sample, grad, tune=True) for the listed chains (reference :528-554).  `idx`: int64 device
        tensor; `q`: [n_chains, ld] positions."""
        lib = L.load()
        D = self._n
        idx_h = idx.cpu().numpy()
        delta = self._n_samples_all[idx_h] - self._previous_update_all[idx_h]
        due = (delta + 1) % self._update_window == 0                     # :540
        stream = C.c_void_p(torch.cuda.current_stream(self._dev).cuda_stream)
        for flag in (True, False):
            sel = idx[torch.as_tensor(due == flag, device=idx.device)]
            if sel.numel() == 0:
                continue
            sel32 = sel.to(torch.int32)
            L.check(lib.lmc_dense_cov_update(_ptr(sel32), int(sel32.numel()), D, q.shape[1], self._lda, _ptr(q),
                                             _ptr(self._mean_fg), _ptr(self._raw_fg), _ptr(self._mean_bg),
                                             _ptr(self._raw_bg), _ptr(self._nsamp),
                                             _ptr(self._cov_all) if flag else None, stream), "lmc_dense_cov_update")
            sel32.record_stream(torch.cuda.current_stream(self._dev))
            if flag:                                                      # _update_from_weightvar (:520-526)
                # gather -> factor -> scatter through two work buffers that live as long as the potential: the batch
                # changes from call to call, and fresh [k, D, D] temporaries of a new size every time are a cudaMalloc
                # (and eventually a synchronising cudaFree) each -- measured 22 ms per call instead of 8
                k = int(sel.numel())
                if self._work is None:
                    self._work = (torch.empty_like(self._cov_all), torch.empty_like(self._chol_all),
                                  torch.empty(self._nc, dtype=torch.int32, device=self._dev))
                wa, wl, winfo = (w[:k] for w in self._work)
                torch.index_select(self._cov_all, 0, sel, out=wa)
                chol, info = torch.linalg.cholesky_ex(wa[:, :, :D], out=(wl, winfo))
                ok = (info == 0) & torch.isfinite(chol.sum(dim=(1, 2)))      # LinAlgError / ValueError in scipy (:524)
                if bool(ok.all()):
                    self._chol_all.index_copy_(0, sel, chol)
                else:
                    self._chol_error = "Cholesky failed for chain(s) %s" % sel[~ok].tolist()[:8]
                    self._chol_all[sel[ok]] = chol[ok]
        switch = idx_h[delta >= self._window_all[idx_h]]                 # :545-552
        if switch.size:
            sw = torch.as_tensor(switch, device=self._dev)
            self._mean_fg[sw] = self._mean_bg[sw]
            self._raw_fg[sw] = self._raw_bg[sw]
            self._nsamp[sw, 0] = self._nsamp[sw, 1]
            self._mean_bg[sw] = 0.0
            self._raw_bg[sw] = 0.0
            self._nsamp[sw, 1] = 0.0
            self._previous_update_all[switch] = self._n_samples_all[switch]
            self._window_all[switch] = (self._window_all[switch] * self._adaptation_window_multiplier).astype(np.int64)
        self._n_samples_all[idx_h] += 1

    def update(self, sample, grad, tune):
        """reference quadpotential.py:528-554 for ONE sample: the reference's object-level API (its tests drive the
        potential directly).  Applies to the last chain (the only one when the potential is not bound to a sampler);
        the sampler itself updates all chains that finished a tuning transition at once (`_update_rows`)."""
        if not tune:
            return
        x = self._single(sample)                      # (moves the state to the device on first use)
        q = torch.zeros(self._nc, self._ld, dtype=torch.float64, device=self._dev)
        q[-1, :self._n] = x
        self._update_rows(torch.tensor([self._nc - 1], dtype=torch.int64, device=self._dev), q)

    def raise_ok(self, vmap=None):
        """reference quadpotential.py:556-559."""
        if self._chol_error is not None:
            raise ValueError("{0}".format(self._chol_error))
