"""CPU-only checks: the C-ABI library loads and exports every symbol include/lmc_b200.h declares, and the host logic
(seeding, validation, shape selection) behaves like the reference's.  No compute calls (no GPU here)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from littlemcmc_b200 import _lib
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    from littlemcmc_b200 import _lib
    header = open(os.path.join(ROOT, "include", "lmc_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(lmc_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.lmc_abi_version() == _lib.ABI_VERSION


def test_struct_layout_matches_header(lib):
    """sizeof(lmc_sampler_args) as nvcc lays it out == the ctypes mirror (checked through a bad-arg call that must
    come back as LMC_ERR_BADARG, not crash, and by the field offsets being naturally aligned)."""
    import ctypes as C
    from littlemcmc_b200 import _lib as L
    a = L.SamplerArgs()
    assert C.sizeof(L.Target) == 24 and C.sizeof(L.Rng) == 40
    for name, _ in L.SamplerArgs._fields_:
        off = getattr(L.SamplerArgs, name).offset
        size = getattr(L.SamplerArgs, name).size
        assert off % min(size, 8) == 0, name
    a.abi_version = L.ABI_VERSION + 1
    assert lib.lmc_nuts_sample(C.byref(a)) == L.ERR_BADARG
    assert lib.lmc_hmc_sample(None) == L.ERR_BADARG


def test_workspace_bytes_is_host_only(lib):
    from littlemcmc_b200 import _lib as L
    n = lib.lmc_workspace_bytes(L.KIND_NUTS, 1024, 1000, 10, 0)
    assert n > 0 and n % 16 == 0
    assert lib.lmc_workspace_bytes(L.KIND_NUTS, 1024, 1000, 17, 0) == L.ERR_UNSUPPORTED
    assert lib.lmc_workspace_bytes(L.KIND_NUTS, 8, 10 ** 6, 10, 0) == L.ERR_UNSUPPORTED
    assert lib.lmc_workspace_bytes(L.KIND_HMC, 8, 10, 10, 0) > 0


def test_seed_resolution_matches_reference_semantics():
    """sampling.py:131-138: int seed -> np.random.seed + one randint(2**30) per chain; list -> truncated."""
    from littlemcmc_b200.sampling import _resolve_seeds
    np.random.seed(123)
    expect = [int(np.random.randint(2 ** 30)) for _ in range(4)]
    assert _resolve_seeds(123, 4) == expect
    assert _resolve_seeds([9, 8, 7, 6, 5], 3) == [9, 8, 7]
    with pytest.raises(TypeError):
        _resolve_seeds(1.5, 2)


def test_init_nuts_start_is_the_reference_start():
    """init_nuts reseeds with the first chain seed and draws 2*rand(D)-1 (sampling.py:574-584)."""
    import littlemcmc_b200 as lmc
    tgt = lmc.targets.StdNormal(5)
    start, step = lmc.init_nuts(tgt, 5, init="jitter+adapt_diag", random_seed=[77, 78])
    np.random.seed(77)
    assert np.array_equal(start, 2 * np.random.rand(5) - 1)
    assert step.potential._initial_weight == 10 and np.array_equal(step.potential._initial_mean, start)
    assert step.step_size == 0.25 / 5 ** 0.25
    s2, _ = lmc.init_nuts(tgt, 5, init="adapt_diag")
    assert np.array_equal(s2, np.zeros(5))
    with pytest.raises(ValueError):
        lmc.init_nuts(tgt, 5, init="nope")
    with pytest.raises(TypeError):
        lmc.init_nuts(tgt, 5, init=3)


def test_potential_validation():
    import littlemcmc_b200 as lmc
    from littlemcmc_b200.quadpotential import PositiveDefiniteError
    with pytest.raises(PositiveDefiniteError):
        lmc.quad_potential(np.array([1.0, -1.0]), True)
    with pytest.raises(ValueError):
        lmc.QuadPotentialDiagAdapt(3, np.zeros(2))
    with pytest.raises(ValueError):
        lmc.QuadPotentialDiagAdapt(3, np.zeros(3), np.ones((3, 3)))
    with pytest.raises(ValueError):
        lmc.NUTS(lmc.targets.StdNormal(2), 2, scaling=np.ones(2), potential=lmc.QuadPotentialDiag(np.ones(2)))
    pot = lmc.QuadPotentialDiagAdapt(3, np.zeros(3))          # initial_diag None -> ones, weight 1 (:178-180)
    assert pot._initial_weight == 1 and np.array_equal(pot._initial_diag, np.ones(3))


def test_targets_are_reference_style_callbacks():
    """The fused targets are plain logp_dlogp_func callables; gradients checked by finite differences."""
    import littlemcmc_b200 as lmc
    rs = np.random.RandomState(0)
    for tgt, D in ((lmc.targets.DiagGaussian(sigma=rs.rand(6) + 0.5), 6), (lmc.targets.NealFunnel(6), 6)):
        q = rs.randn(D) * 0.5
        lp, g = tgt(q)
        for i in range(D):
            e = np.zeros(D); e[i] = 1e-6  # noqa: E702
            fd = (tgt(q + e)[0] - tgt(q - e)[0]) / 2e-6
            assert abs(fd - g[i]) < 1e-5 * max(1, abs(g[i]))


def test_torch_batched_densities_match_the_numpy_callables():
    """The library's torch-op densities (two kernels for the Gaussian, nine for the funnel: constants folded into
    baddbmm / addcmul / add(alpha=)) evaluate the same functions as the reference-style NumPy callables."""
    import torch
    import littlemcmc_b200 as lmc
    rs = np.random.RandomState(1)
    for tgt, D in ((lmc.targets.DiagGaussian(tau=rs.rand(11) + 0.5), 11), (lmc.targets.NealFunnel(9), 9),
                   (lmc.targets.NealFunnel(2), 2)):
        f = tgt.torch_batched("cpu")
        q = torch.as_tensor(rs.randn(7, D) * 0.7)
        q[:, 0] = torch.linspace(-5.0, 3.0, 7)
        lp, g = f(q)
        assert lp.shape == (7,) and g.shape == (7, D) and g.is_contiguous()
        for i in range(7):
            lp0, g0 = tgt(q[i].numpy())
            assert abs(lp[i].item() - lp0) <= 1e-13 * max(1.0, abs(lp0))
            np.testing.assert_allclose(g[i].numpy(), g0, rtol=1e-13, atol=0)


def test_elementwise_target_expressions_and_generated_source():
    """ElementwiseTarget: `g` inside `logp` is the element's gradient; `q` / `g` are not parameter names; the generated
    Target type is branch-free (select-guarded parameter loads outside any `if`: hoistable out of the leapfrog loop)."""
    import littlemcmc_b200 as lmc
    tau = np.linspace(0.5, 2.0, 7)
    a = lmc.targets.ElementwiseTarget(7, logp="0.5 * q * g", grad="-(tau * q)", params={"tau": tau})
    b = lmc.targets.ElementwiseTarget(7, logp="0.5 * q * (-(tau * q))", grad="-(tau * q)", params={"tau": tau})
    q = np.linspace(-1.0, 1.0, 7)
    (la, ga), (lb, gb), (l0, g0) = a(q), b(q), lmc.targets.DiagGaussian(tau=tau)(q)
    assert np.array_equal(ga, g0) and np.array_equal(gb, g0) and abs(la - l0) < 1e-15 and abs(lb - l0) < 1e-15
    src = a.fused.source
    assert "if (" not in src and "? __ldg(" in src and "in_x ? g : 0.0" in src
    with pytest.raises(ValueError, match="parameter names"):
        lmc.targets.ElementwiseTarget(3, logp="q * g", grad="-q", params={"g": np.ones(3)})


def test_scheduler_protocol_model():
    """A model of the transition scheduler of the fused kernels (csrc/lmc_sampler.cuh: ticket / peek / take / push with the
    consumed-slot handshake) under random interleavings, including pushers that stall between taking their ticket and
    publishing the entry (the late store must not overwrite the entry one ring later): every (chain, transition)
    runs exactly once, in order per chain, and every group terminates."""
    import random

    def run(n_chains, n_trans, n_groups, seed):
        rnd = random.Random(seed)
        total = n_chains * n_trans
        ring = [(i + 1, i, 0) for i in range(n_chains)]      # (ticket + 1, chain, t); None = consumed
        head, tail = 0, n_chains
        last_t = [-1] * n_chains
        state = [("idle",)] * n_groups
        finished = [False] * n_groups
        steps = 0
        while not all(finished):
            steps += 1
            assert steps < 400 * total + 10000, "scheduler model does not terminate"
            g = rnd.randrange(n_groups)
            if finished[g]:
                continue
            st = state[g]
            if st[0] == "idle":                               # sched_ticket
                tk, head = head, head + 1
                if tk >= total:
                    finished[g] = True
                else:
                    state[g] = ("take", tk)
            elif st[0] == "take":                             # sched_take: spin until the entry of this ticket is there
                e = ring[st[1] % n_chains]
                if e is not None and e[0] == st[1] + 1:
                    ring[st[1] % n_chains] = None             # consumed
                    state[g] = ("run", e[1], e[2])
            elif st[0] == "run":                              # one transition; the chain's transitions stay ordered
                _, c, t = st
                assert last_t[c] == t - 1
                last_t[c] = t
                if t + 1 < n_trans:
                    tk, tail = tail, tail + 1                 # push ticket taken ...
                    state[g] = ("push", tk, c, t + 1)
                else:
                    state[g] = ("idle",)
            else:                                             # ... entry published only into a consumed slot (CAS)
                _, tk, c, t = st
                if rnd.random() < 0.3:                        # the pusher stalls for a while
                    continue
                if ring[tk % n_chains] is None:
                    ring[tk % n_chains] = (tk + 1, c, t)
                    state[g] = ("idle",)
        assert last_t == [n_trans - 1] * n_chains and tail == total

    for n_chains, n_trans, n_groups in ((1, 9, 3), (2, 7, 5), (5, 6, 2), (7, 5, 7), (16, 4, 3)):
        for seed in range(4):
            run(n_chains, n_trans, n_groups, seed)


def test_no_cpu_fallback():
    """Binding chains to a non-CUDA device must fail loudly."""
    from littlemcmc_b200 import _lib as L, engine
    with pytest.raises(L.LmcError):
        engine.DeviceChains(2, 3, "cpu")


def test_dense_potential_validation_and_factory():
    """reference quadpotential.py:33-66 (quad_potential dispatch), :484-499 (QuadPotentialFullAdapt argument checks);
    host-side only: no device is touched until a step method binds the potential."""
    import littlemcmc_b200 as lmc
    from littlemcmc_b200.quadpotential import PositiveDefiniteError
    cov = np.array([[2.0, 0.3], [0.3, 1.0]])
    assert isinstance(lmc.quad_potential(cov, True), lmc.QuadPotentialFull)
    assert isinstance(lmc.quad_potential(cov, False), lmc.QuadPotentialFullInv)
    assert isinstance(lmc.quad_potential(np.array([1.0, 2.0]), True), lmc.QuadPotentialDiag)
    with pytest.raises(PositiveDefiniteError):
        lmc.quad_potential(np.array([[1.0, 0.0], [0.0, -1.0]]), True)
    with pytest.raises(ValueError, match="two-dimensional"):
        lmc.QuadPotentialFullAdapt(2, np.zeros(2), np.ones(2), 1)
    with pytest.raises(ValueError, match="one-dimensional"):
        lmc.QuadPotentialFullAdapt(2, np.zeros((2, 1)), np.eye(2), 1)
    with pytest.raises(ValueError, match="Wrong shape for initial_cov"):
        lmc.QuadPotentialFullAdapt(3, np.zeros(3), np.eye(2), 1)
    with pytest.raises(ValueError, match="Wrong shape for initial_mean"):
        lmc.QuadPotentialFullAdapt(2, np.zeros(3), np.eye(2), 1)
    pot = lmc.QuadPotentialFullAdapt(2, np.zeros(2))              # initial_cov None -> identity with weight 1 (:500-502)
    assert pot._initial_weight == 1 and np.array_equal(pot._initial_cov, np.eye(2))
    for init in ("adapt_full", "jitter+adapt_full"):
        start, step = lmc.init_nuts(lmc.targets.StdNormal(3), 3, init=init, random_seed=4)
        assert isinstance(step.potential, lmc.QuadPotentialFullAdapt) and start.shape == (3,)


def test_arviz_export_dicts():
    """The reference cookbook's `arviz_from_littlemcmc` layout (docs/tutorials/framework_cookbook.rst:199-205)."""
    import littlemcmc_b200 as lmc
    trace = np.arange(2 * 5 * 3, dtype="d").reshape(2, 5, 3)
    stats = {"depth": np.ones((2, 5, 1), dtype=np.int64), "diverging": np.zeros((2, 5, 1), dtype=bool)}
    posterior, sample_stats = lmc.interop.to_arviz_dict(trace, stats)
    assert posterior["x"].shape == (2, 5, 3) and np.array_equal(posterior["x"], trace)
    assert sample_stats["depth"].shape == (2, 5) and sample_stats["depth"].dtype == np.int64
    assert sample_stats["diverging"].dtype == bool
    with pytest.raises(ValueError):
        lmc.interop.to_arviz_dict(trace[0], stats)
    with pytest.raises(ValueError):
        lmc.interop.to_arviz_dict(trace, {"depth": np.ones((2, 4, 1))})
    try:
        import arviz  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError, match="arviz"):
            lmc.interop.arviz_from_littlemcmc(trace, stats)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to the GPU arm): one JSON line with the contract's
    keys, the unmodified reference (oracle/_ref, when built: this container) or else its oracle port timed on the host
    cores, zero transfer bytes, and the same `config` the GPU arm prints."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-trans", "6", "--workload", "cfg2"], capture_output=True, text=True,
                         timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in line, key
    assert line["impl"] == "reference" and line["metric"].startswith("leapfrog-steps/sec") and line["value"] > 0
    have_ref = os.path.exists(os.path.join(root, "oracle", "_ref", "littlemcmc", "__init__.py"))
    assert line["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")
    assert line["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    sys.path.insert(0, root)
    import bench
    assert line["config"] == bench.static_config("cfg2", 1)      # identical to the GPU arm's description of the workload
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["e2e"]["value"] == line["value"] and line["config"]["workload"].startswith("cfg2")
    # ranks other than 0 print nothing and exit 0 (torchrun launches the arm on every rank)
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1"],
                         capture_output=True, text=True, timeout=120, cwd=root, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_bench_gpu_arm_fails_loudly_without_a_gpu():
    """No CUDA device => the GPU arm exits non-zero and prints no result line (there is no CPU fallback to time)."""
    import subprocess
    import sys
    import torch
    if torch.cuda.is_available():
        pytest.skip("this check is for machines without a GPU")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "1", "--warmup", "0", "--no-cpu"],
                         capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode != 0
    assert '"value"' not in out.stdout


def test_ctypes_mirrors_have_the_c_layout(tmp_path):
    """sizeof / offsetof of every struct in include/lmc_b200.h as a C compiler lays them out == the ctypes mirrors in
    littlemcmc_b200/_lib.py (the header is plain C: it must compile with gcc, no CUDA headers)."""
    import ctypes as C
    import subprocess
    from littlemcmc_b200 import _lib as L
    structs = {"lmc_target": L.Target, "lmc_rng": L.Rng, "lmc_sampler_args": L.SamplerArgs,
               "lmc_callback_args": L.CallbackArgs, "lmc_dense_args": L.DenseArgs}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "lmc_b200.h"', "int main(void) {"]
    for cname, cls in structs.items():
        lines.append('  printf("%s sizeof %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('  printf("%s %s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines += ['  printf("abi %d nstats %d adapt_stride %d\\n", LMC_ABI_VERSION, LMC_NSTATS, LMC_ADAPT_STRIDE);',
              "  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                   check=True, capture_output=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    seen = 0
    for row in out:
        parts = row.split()
        if len(parts) == 3 and parts[1] == "sizeof":
            assert C.sizeof(structs[parts[0]]) == int(parts[2]), row
            seen += 1
        elif len(parts) == 3 and parts[0] in structs:
            assert getattr(structs[parts[0]], parts[1]).offset == int(parts[2]), row
            seen += 1
        elif parts and parts[0] == "abi":
            assert (int(parts[1]), int(parts[3]), int(parts[5])) == (L.ABI_VERSION, L.NSTATS, L.ADAPT_STRIDE)
            seen += 1
    assert seen > 70
