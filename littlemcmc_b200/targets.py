"""Target densities.

Every target here is a plain ``logp_dlogp_func``: calling it with a NumPy vector ``q[D]`` returns ``(logp, dlogp[D])``
exactly like the callbacks the reference takes (base_hmc.py:34, integration.py:40), so the same object can be handed
to the reference sampler and to this one.  In addition it carries a ``fused`` descriptor; when a step method sees one
it evaluates the density INSIDE the CUDA kernels (no per-leapfrog callback), which is the throughput path.
Arbitrary callables are supported through :class:`TorchBatched` / the NumPy adapter (callback mode).
"""
import numpy as np

from . import _lib as L
from .engine import FusedTarget, UserFusedTarget


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
        neg_tau = torch.as_tensor(-self.tau, dtype=torch.float64, device=device)

        zero = torch.zeros(1, 1, 1, dtype=torch.float64, device=device)

        def fn(q):
            # two kernels per evaluation (callback mode is bound by launches per gradient, not by bytes):
            # g = q * (-tau) [bitwise -(tau * q)]; logp = 0.5 * <q, g> per chain as ONE strided-batched GEMM (alpha = 0.5)
            g = q * neg_tau
            return torch.baddbmm(zero, q.unsqueeze(1), g.unsqueeze(2), beta=0, alpha=0.5).view(-1), g
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

        def c(x):
            return torch.full((1, 1), float(x), dtype=torch.float64, device=device)
        zero, zero3, c_h, c_nh = c(0.0), torch.zeros(1, 1, 1, dtype=torch.float64, device=device), c(half_nm1), c(-half_nm1)

        def fn(q):
            # ten small kernels per evaluation instead of the ~20 of the literal transcription: every line is ONE kernel
            # (addcmul / add-with-alpha fold the constants; the scalar-per-chain quantities are [C, 1] columns)
            v, x = q[:, :1], q[:, 1:]
            ev = torch.exp(-v)                                                        # (2 kernels) e^-v
            g = torch.addcmul(zero, q, ev, value=-1.0)                                # -(ev * q): columns 1.. are final
            S = torch.baddbmm(zero3, x.unsqueeze(1), x.unsqueeze(2), beta=0).view(-1, 1)   # sum_i x_i^2
            nhs = torch.addcmul(zero, ev, S, value=-0.5)                              # -hs = -(ev * S) / 2
            r = torch.add(c_h, v, alpha=0.5 * inv_s2)                                 # v / (2 s^2) + (n-1)/2
            logp = torch.addcmul(nhs, v, r, value=-1.0)                               # -hs - v^2 / (2 s^2) - (n-1)/2 v
            s1 = torch.add(nhs, v, alpha=inv_s2)                                      # -hs + v / s^2
            torch.sub(c_nh, s1, out=g[:, :1])                                         # g_0 = hs - v / s^2 - (n-1)/2
            return logp.view(-1), g
        return TorchBatched(fn, cuda_graph=cuda_graph)


class CudaTarget:
    """A user-written density evaluated INSIDE the fused sampler kernels (no callback, no launch per gradient).

    ``source`` is CUDA C++ defining ``struct <type_name>`` with the protocol of the built-in targets
    (csrc/lmc_device.cuh): ``static constexpr int kPre``; ``pre<G,NP>(lane, D, q, out)`` partial sums the gradient needs
    (if any); ``grad<G,NP>(lane, D, ldh, q, g, pre)`` writes this thread's gradient pairs and returns its partial of the
    log-density sum; ``finish(sum, pre)`` the log density.  Its only data member is ``const double* params``; ``params``
    is uploaded once.  Optionally the source specialises ``lmc::StageTraits<T>`` (csrc/lmc_device.cuh) to have a
    per-dimension parameter vector kept in shared memory by the chunked warp kernel (``ElementwiseTarget`` does so for a
    single parameter); the host reads what the compiled kernel stages from the module, nothing to declare.  The sampler kernel is compiled for sm_100a at first use (NVRTC) and cached on disk.
    ``numpy_fn`` (optional) is the same density as a reference-style callable ``q[D] -> (logp, dlogp[D])``
    (base_hmc.py:34), so that the object also drives the reference / the per-step integrator API."""

    def __init__(self, source, type_name, ndim, params=None, numpy_fn=None):
        self.ndim = int(ndim)
        self.fused = UserFusedTarget(source, type_name, ndim, params)
        self._numpy_fn = numpy_fn

    def __call__(self, q):
        if self._numpy_fn is None:
            raise NotImplementedError("this CudaTarget was built without a NumPy callable (numpy_fn=...)")
        return self._numpy_fn(q)


_ELEMENTWISE_TEMPLATE = """
struct %(name)s {
  static constexpr int kPre = 0;
  %(member)s
  template <int G, int NP>
  __device__ __forceinline__ void pre(int, int, const double2 (&)[NP], double (&)[2]) const {}
  template <int G, int NP>
  __device__ __forceinline__ double grad(int lane, int D, int ldh, const double2 (&q_)[NP], double2 (&g_)[NP],
                                         const double (&)[2]) const {
    double part = 0.0;
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      const int j = lane + k * G;
      // parameter loads as selects and NO branch around the element code (like the built-in targets): the compiler hoists
      // the read-only loads out of the leapfrog loop; an element beyond ndim is evaluated at q = 0 with zero parameters and
      // its results are discarded by selects (whatever they are, NaN included)
%(loads)s
      const bool in_x = 2 * j < D, in_y = 2 * j + 1 < D;
      double2 gk;
      {  // the gradient of an element is in scope of its `logp` expression as `g`
        const double q = q_[k].x;
%(bind_x)s
        const double g = (%(grad)s);
        const double l = (%(logp)s);
        gk.x = in_x ? g : 0.0;
        part += in_x ? l : 0.0;
      }
      {
        const double q = q_[k].y;
%(bind_y)s
        const double g = (%(grad)s);
        const double l = (%(logp)s);
        gk.y = in_y ? g : 0.0;
        part += in_y ? l : 0.0;
      }
      g_[k] = gk;
    }
    return part;
  }
  __device__ __forceinline__ double finish(double sum, const double (&)[2]) const { return sum; }
};
"""


# One parameter vector: the chunked warp kernel may keep it in shared memory (csrc/lmc_device.cuh: StageTraits; it does
# where the slot has a spare KB, i.e. 65..128 dimensions) -- the same element code reading the staged copy.
_STAGE_TRAITS_TEMPLATE = """
namespace lmc {
template <>
struct StageTraits< ::%(name)s> {
  static constexpr int kVecs = 1;
  using Staged = ::%(name)s_S;
  template <int NP>
  static __device__ __forceinline__ Staged make(const ::%(name)s& t, double2* s_lane, int lane, int ldh) {
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      const int j = lane + k * 32;
      s_lane[k * 32] = j < ldh ? __ldg(reinterpret_cast<const double2*>(t.params) + j) : make_double2(0.0, 0.0);
    }
    return Staged{s_lane};
  }
};
}  // namespace lmc
"""


class ElementwiseTarget(CudaTarget):
    """Separable density ``logp(q) = sum_i f(q_i; theta_i)`` from two C expressions: ``logp`` = f and ``grad`` = df/dq,
    written in terms of ``q`` and the names of ``params`` (a dict name -> array[ndim] or scalar); ``logp`` may also use
    ``g``, the value of the ``grad`` expression at that element (``g`` is therefore not available as a parameter name).
    Example (the built-in diagonal Gaussian): ``ElementwiseTarget(D, logp="0.5 * q * g", grad="-(tau * q)",
    params={"tau": tau})``.
    The same expressions are evaluated with NumPy for the reference-style callable (exp, log, sqrt, tanh ... map to
    numpy's).  Multiply-adds are not contracted (-fmad=false), as in NumPy."""

    def __init__(self, ndim, logp, grad, params=None):
        import hashlib
        params = {k: np.broadcast_to(np.asarray(v, dtype="d"), (int(ndim),)).copy() for k, v in (params or {}).items()}
        if "g" in params or "q" in params:
            raise ValueError("`q` and `g` are the element and its gradient inside the expressions, not parameter names")
        names = sorted(params)
        ld = int(ndim) + (int(ndim) & 1)
        blob = np.zeros((max(1, len(names)), ld))
        for i, n in enumerate(names):
            blob[i, :ndim] = params[n]
        name = "LmcElementwise_" + hashlib.sha256(repr((logp, grad, names)).encode()).hexdigest()[:12]
        loads = "\n".join("      const double2 %s_2 = j < ldh ? __ldg(reinterpret_cast<const double2*>(params + %d * 2 * (size_t)ldh) + j)"
                          " : make_double2(0.0, 0.0);" % (n, i) for i, n in enumerate(names))
        bind = lambda c: "\n".join("          const double %s = %s_2.%s;" % (n, n, c) for n in names)  # noqa: E731
        member = "const double* params;  // [n_params][ld] rows, ld = ndim rounded up to even, padding 0"
        src = _ELEMENTWISE_TEMPLATE % dict(name=name, member=member, loads=loads, bind_x=bind("x"), bind_y=bind("y"),
                                           grad=grad, logp=logp)
        if len(names) == 1:
            src += _ELEMENTWISE_TEMPLATE % dict(
                name=name + "_S", member="const double2* stage;  // this lane's [NP][32] pairs of the staged parameter",
                loads="      const double2 %s_2 = stage[k * G];" % names[0], bind_x=bind("x"), bind_y=bind("y"), grad=grad,
                logp=logp)
            src += _STAGE_TRAITS_TEMPLATE % dict(name=name)
        env = {k: getattr(np, k) for k in ("exp", "log", "sqrt", "tanh", "sin", "cos", "log1p", "expm1", "fabs")}

        def numpy_fn(q):
            scope = dict(env, q=np.asarray(q, dtype="d"), **params)
            gval = np.asarray(eval(grad, {"__builtins__": {}}, scope), dtype="d") * np.ones(int(ndim))
            return float(np.sum(eval(logp, {"__builtins__": {}}, dict(scope, g=gval)))), gval
        super().__init__(src, name, ndim, params=blob, numpy_fn=numpy_fn)
        self.logp_expr, self.grad_expr, self.params = logp, grad, params


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
