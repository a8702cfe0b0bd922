"""littlemcmc_b200.diagnostics: the chain-moments kernel against NumPy, split R-hat and ESS
against loop transcriptions of their definitions, and their behaviour on chains with known properties."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _split(x):
    half = x.shape[1] // 2
    return np.concatenate([x[:, :half], x[:, half:2 * half]], 0)


def rhat_ref(x):
    """x: [chains, draws] -> split R-hat (BDA3 11.4)."""
    s = _split(x)
    n = s.shape[1]
    W = s.var(1, ddof=1).mean()
    B_over_n = s.mean(1).var(ddof=1)
    return np.sqrt((W * (n - 1) / n + B_over_n) / W)


def ess_ref(x):
    """x: [chains, draws] -> Stan's ESS on split chains, written as plain loops."""
    s = _split(x)
    m, n = s.shape
    c = s - s.mean(1, keepdims=True)
    acov = np.array([[np.dot(c[j, :n - t], c[j, t:]) / n for t in range(n)] for j in range(m)])
    mean_var = acov[:, 0].mean() * n / (n - 1.0)
    var_plus = mean_var * (n - 1.0) / n + (s.mean(1).var(ddof=1) if m > 1 else 0.0)
    rho = 1.0 - (mean_var - acov.mean(0)) / var_plus
    rho[0] = 1.0
    n_pairs = max(1, min((n - 2) // 2 if n > 4 else 1, n // 2))
    P = [rho[2 * k] + rho[2 * k + 1] for k in range(n_pairs)]
    K = 1
    while K < n_pairs and P[K] >= 0:
        K += 1
    for k in range(1, K):
        P[k] = min(P[k], P[k - 1])
    extra = rho[2 * K] if (K < n_pairs and rho[2 * K] > 0) else 0.0
    tau = -1.0 + 2.0 * sum(P[:K]) + extra
    return min(m * n / tau, m * n * np.log10(m * n))


@pytest.mark.parametrize("shape,n_seg", [((3, 50, 7), 2), ((5, 33, 130), 3), ((2, 1001, 1), 8), ((64, 16, 257), 4)])
def test_chain_moments_kernel(shape, n_seg):
    import torch
    from littlemcmc_b200 import diagnostics as dg
    rs = np.random.RandomState(1)
    x = rs.randn(*shape) * rs.rand(shape[2]) * 5 + 100.0 * rs.randn(shape[2])      # large means: the pivot matters
    Cn, T, D = shape
    for view in ("contiguous", "strided"):
        t = torch.as_tensor(x, device="cuda")
        if view == "strided":                      # a [C, T, D] window of a bigger trace (sample()'s blocks)
            big = torch.zeros(Cn, T + 5, D + 3, dtype=torch.float64, device="cuda")
            big[:, 2:2 + T, :D] = t
            t = big[:, 2:2 + T, :D]
        mean, m2, counts = dg.chain_moments(t, n_seg)
        ln = T // n_seg
        for s in range(n_seg):
            seg = x[:, s * ln:(T if s == n_seg - 1 else (s + 1) * ln)]
            assert counts[s].item() == seg.shape[1]
            np.testing.assert_allclose(mean[:, s].cpu().numpy(), seg.mean(1), rtol=1e-13)
            np.testing.assert_allclose(m2[:, s].cpu().numpy(), ((seg - seg.mean(1, keepdims=True)) ** 2).sum(1),
                                       rtol=1e-10, atol=1e-12)


def test_rhat_and_ess_match_their_definitions():
    from littlemcmc_b200 import diagnostics as dg
    rs = np.random.RandomState(2)
    Cn, T, D = 6, 201, 5
    x = np.empty((Cn, T, D))
    phi = np.array([0.0, 0.3, 0.6, 0.9, -0.5])
    e = rs.randn(Cn, T, D)
    x[:, 0] = e[:, 0]
    for t in range(1, T):
        x[:, t] = phi * x[:, t - 1] + np.sqrt(1 - phi**2) * e[:, t]
    x[:, :, 1] += np.arange(Cn)[:, None] * 0.5               # chains that disagree in dimension 1
    r = dg.rhat(x).cpu().numpy()
    n_eff = dg.ess(x).cpu().numpy()
    for d in range(D):
        np.testing.assert_allclose(r[d], rhat_ref(x[:, :, d]), rtol=1e-10)
        np.testing.assert_allclose(n_eff[d], ess_ref(x[:, :, d]), rtol=1e-8)
    assert r[1] > 1.2 and abs(r[0] - 1) < 0.05
    s = dg.summary(x)
    np.testing.assert_allclose(s["mean"].cpu().numpy(), x[:, :200].reshape(-1, D).mean(0), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(s["sd"].cpu().numpy(), x[:, :200].reshape(-1, D).std(0, ddof=1), rtol=1e-12)


def test_ess_of_known_processes_and_of_a_real_run():
    """iid draws: ESS ~ chains * draws; AR(1) with coefficient phi: ESS ~ N (1 - phi) / (1 + phi); and NUTS draws of a
    1024-chain run straight from the device."""
    import littlemcmc_b200 as lmc
    from littlemcmc_b200 import diagnostics as dg
    rs = np.random.RandomState(3)
    Cn, T = 64, 1000
    e = rs.randn(Cn, T, 2)
    x = e.copy()
    phi = 0.8
    for t in range(1, T):
        x[:, t, 1] = phi * x[:, t - 1, 1] + np.sqrt(1 - phi**2) * e[:, t, 1]
    n_eff = dg.ess(x).cpu().numpy()
    N = Cn * T
    assert 0.85 * N < n_eff[0] < 1.25 * N
    assert 0.75 < n_eff[1] / (N * (1 - phi) / (1 + phi)) < 1.3
    D = 20
    target = lmc.targets.DiagGaussian(sigma=np.linspace(0.5, 3, D))
    trace, stats = lmc.sample(target, D, draws=200, tune=300, chains=1024, random_seed=5, return_device=True)
    r = dg.rhat(trace).cpu().numpy()
    n_eff = dg.ess(trace).cpu().numpy()
    assert (np.abs(r - 1) < 0.01).all(), r
    assert (n_eff > 0.3 * 1024 * 200).all(), n_eff           # NUTS on a Gaussian: nearly independent (or antithetic) draws
    sd = dg.summary(trace)["sd"].cpu().numpy()
    np.testing.assert_allclose(sd, np.linspace(0.5, 3, D), rtol=0.03)
