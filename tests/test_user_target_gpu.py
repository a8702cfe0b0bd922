"""User-written target densities compiled into the fused sampler kernels at run time (NVRTC; include/lmc_b200.h:
lmc_user_kernel_build / lmc_user_sample; littlemcmc_b200.targets.CudaTarget / ElementwiseTarget) -- the fast path for
the reference's "any logp_dlogp_func" contract (base_hmc.py:34, integration.py:62,115).  Same oracle protocol and bar as
tests/test_gpu_parity.py: integer / boolean statistics and uniform counts EXACT, float64 quantities to RTOL = 1e-9."""
import numpy as np
import pytest

from tests import golden_cases as gc
from tests import parity_utils as pu

pytestmark = pytest.mark.gpu

RTOL = 1e-9

# Neal's funnel written against the public Target protocol (a density whose gradient needs sums over the chain's
# dimensions: kPre = 2).  Parameters: params[0] = 1 / v_scale^2, params[1] = (D - 1) / 2.
USER_FUNNEL = r"""
struct UserFunnel {
  static constexpr int kPre = 2;   // pre[0] = sum_{i>=1} q_i^2, pre[1] = q_0
  const double* params;
  template <int G, int NP>
  __device__ void pre(int lane, int D, const double2 (&q)[NP], double (&out)[2]) const {
    double S = 0.0;
    for (int k = 0; k < NP; ++k) {
      const int j = lane + k * G;
      if (j > 0) S = fma(q[k].x, q[k].x, S);
      S = fma(q[k].y, q[k].y, S);
    }
    out[0] = S;
    out[1] = (lane == 0) ? q[0].x : 0.0;
  }
  template <int G, int NP>
  __device__ double grad(int lane, int D, int, const double2 (&q)[NP], double2 (&g)[NP], const double (&pre)[2]) const {
    const double ev = exp(-pre[1]);
    for (int k = 0; k < NP; ++k) {
      const int j = lane + k * G;
      g[k] = make_double2(-lmc::mul_rn(ev, q[k].x), -lmc::mul_rn(ev, q[k].y));
      if (2 * j >= D) g[k].x = 0.0;
      if (2 * j + 1 >= D) g[k].y = 0.0;
    }
    if (lane == 0) {
      const double hs = lmc::mul_rn(lmc::mul_rn(0.5, ev), pre[0]);
      g[0].x = lmc::add_rn(lmc::add_rn(-lmc::mul_rn(pre[1], params[0]), hs), -params[1]);
    }
    return 0.0;
  }
  __device__ double finish(double, const double (&pre)[2]) const {
    const double v = pre[1], ev = exp(-v);
    const double hs = lmc::mul_rn(lmc::mul_rn(0.5, ev), pre[0]);
    const double a = -lmc::mul_rn(lmc::mul_rn(lmc::mul_rn(0.5, v), v), params[0]);
    return lmc::add_rn(lmc::add_rn(a, -hs), -lmc::mul_rn(params[1], v));
  }
};
"""


def _user_gauss(case, logp="0.5 * q * (-(tau * q))"):
    import littlemcmc_b200 as lmc
    return lmc.targets.ElementwiseTarget(int(case["ndim"]), logp=logp, grad="-(tau * q)",
                                         params={"tau": np.asarray(case["tau"], dtype="d")})


@pytest.mark.parametrize("name", ["nuts_diag_d37", "nuts_static_d100"])
def test_user_gaussian_logp_in_terms_of_the_gradient(name):
    """`g` inside the `logp` expression is the element's gradient; D = 37 exercises the half-filled last pair."""
    case, _ = gc.load(name)
    tgt = _user_gauss(case, logp="0.5 * q * g")
    q = np.linspace(-1, 1, int(case["ndim"]))
    lp, g = tgt(q)
    lp0, g0 = gc.target_fn(case)()(q)
    assert np.array_equal(g, g0) and abs(lp - lp0) <= 1e-12 * abs(lp0)
    res = pu.run_case_on_gpu_and_oracle(name, target=tgt.fused)
    pu.assert_parity(res, rtol=RTOL)


@pytest.mark.parametrize("name", ["nuts_b1_d10", "nuts_diag_d37", "nuts_static_d100", "nuts_deep_d100",
                                  "nuts_early_gt_max_d20", "nuts_illcond_d1000", "hmc_static_d50", "hmc_cfg1_d10"])
def test_user_gaussian_matches_the_oracle(name):
    """The diagonal Gaussian written by a user as two expressions: chunked warp kernel (D <= 256), register kernel
    (D = 1000) and the HMC kernels, transition level."""
    case, _ = gc.load(name)
    n = None if int(case["tune"]) + int(case["draws"]) <= 120 else 40
    tgt = _user_gauss(case)
    q = np.linspace(-1, 1, int(case["ndim"]))
    lp, g = tgt(q)                                            # the same object is a reference-style NumPy callable
    lp0, g0 = gc.target_fn(case)()(q)
    assert np.array_equal(g, g0) and abs(lp - lp0) <= 1e-12 * abs(lp0)
    res = pu.run_case_on_gpu_and_oracle(name, n_trans=n, target=tgt.fused)
    print(pu.parity_report(res))
    pu.assert_parity(res, rtol=RTOL)


@pytest.mark.parametrize("name", ["nuts_funnel_d10", "nuts_deep_funnel_d50", "nuts_deep_funnel_d10"])
def test_user_funnel_with_pre_sums_matches_the_oracle(name):
    import littlemcmc_b200 as lmc
    case, _ = gc.load(name)
    D = int(case["ndim"])
    tgt = lmc.targets.CudaTarget(USER_FUNNEL, "UserFunnel", D, params=[1.0 / 9.0, 0.5 * (D - 1)])
    res = pu.run_case_on_gpu_and_oracle(name, target=tgt.fused)
    pu.assert_parity(res, rtol=RTOL)


def test_user_target_through_sample_api():
    """sample() with a non-Gaussian separable density (logistic: logp = -q - 2 log(1 + e^-q), variance pi^2 / 3 per
    coordinate) in-kernel Philox randomness: moments, dtypes, and the cached kernel is reused."""
    import littlemcmc_b200 as lmc
    D, chains = 24, 512
    tgt = lmc.targets.ElementwiseTarget(D, logp="-(s * q) - 2.0 * log1p(exp(-(s * q)))",
                                        grad="s * (2.0 / (1.0 + exp(s * q)) - 1.0)", params={"s": np.full(D, 2.0)})
    step = lmc.NUTS(tgt, D)
    trace, stats = lmc.sample(tgt, D, draws=120, tune=200, step=step, chains=chains, start=np.zeros(D), random_seed=3,
                              progressbar=False)
    assert trace.shape == (chains, 120, D) and np.isfinite(trace).all()
    np.testing.assert_allclose(trace.var((0, 1)), np.pi ** 2 / 3 / 4.0, rtol=0.06)     # scale 1/s = 1/2
    assert abs(trace.mean()) < 0.02 and abs(stats["mean_tree_accept"].mean() - 0.8) < 0.06
    assert len(tgt.fused._kernels) == 1


def test_user_source_that_does_not_compile_reports_the_compiler_log():
    import littlemcmc_b200 as lmc
    from littlemcmc_b200 import _lib as L
    tgt = lmc.targets.ElementwiseTarget(8, logp="-0.5 * q * q", grad="-q +")
    with pytest.raises(L.LmcError, match="error"):
        lmc.sample(tgt, 8, draws=2, tune=2, chains=2, start=np.zeros(8), step=lmc.NUTS(tgt, 8))
