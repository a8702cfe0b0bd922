"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on identical randomness.

Protocol.  The oracle replays the reference's own legacy-MT19937 stream (it is bit-identical to
the unmodified reference on every fixture: tests/test_oracle_golden.py) and records the numbers it consumed and its
full state around every transition.  The CUDA path then runs in tape mode:

* transition-level (the parity bar): every transition starts from the oracle's exact pre-state, so one transition's
  arithmetic is compared at a time.  Integer / boolean statistics and the count of uniforms consumed must be EXACT;
  float64 quantities must agree to RTOL = 1e-9 (measured: ~1e-13; the licence to differ is the summation order of
  dot products and 1-ulp libm differences in exp/log).
* run-level (chained on the device): Markov chains with adaptation are chaotic -- perturbing the step size by ONE ulp
  inside the CPU oracle itself moves the trace by 5e-8 within five transitions on the ill-conditioned fixture and
  flips tree decisions on the 235-transition one -- so chained runs are compared on short benign fixtures only, and
  otherwise through GPU-vs-GPU bit-identity properties (chunking, launch shapes).
"""
import numpy as np
import pytest

from tests import golden_cases as gc
from tests import parity_utils as pu

pytestmark = pytest.mark.gpu

RTOL = 1e-9


@pytest.mark.parametrize("name", gc.CASE_NAMES)
def test_transition_level_parity(name):
    res = pu.run_case_on_gpu_and_oracle(name)
    print(pu.parity_report(res))
    pu.assert_parity(res, rtol=RTOL)
    # the oracle side of this comparison IS the reference: same seeds, bit-identical trace
    _, ref = gc.load(name)
    assert np.array_equal(res.cpu_trace, ref["trace"])


@pytest.mark.parametrize("name", ["nuts_b1_d10", "nuts_static_d100"])
def test_run_level_parity_short_runs(name):
    res = pu.run_case_on_gpu_and_oracle(name, chained=True)
    print(pu.parity_report(res))
    pu.assert_parity(res, rtol=RTOL)
    _, ref = gc.load(name)
    np.testing.assert_allclose(res.gpu_trace, ref["trace"], rtol=RTOL, atol=1e-12)


def test_run_level_config1_hmc_statistics():
    """BASELINE config 1 (HMC, 4 chains x 1000 transitions) chained on the device: every integer statistic of all
    4000 transitions matches the reference; continuous ones to 1e-6 (drift of 1000 chained transitions)."""
    res = pu.run_case_on_gpu_and_oracle("hmc_cfg1_d10", chained=True)
    print(pu.parity_report(res))
    _, ref = gc.load("hmc_cfg1_d10")
    assert int(res.gpu_stats["n_steps"].sum()) == int(ref["stat_n_steps"].sum()) == 5389
    pu.assert_parity(res, rtol=1e-6)


@pytest.mark.parametrize("name", ["nuts_diag_d37", "hmc_static_d50", "nuts_funnel_d10"])
def test_chunked_calls_are_bit_identical(name):
    """A run split over several launches (state carried in the device buffers) equals one launch bit for bit."""
    one = pu.run_case_on_gpu_and_oracle(name, n_trans=60, chained=True)
    many = pu.run_case_on_gpu_and_oracle(name, n_trans=60, chained=True, chunks=7)
    assert np.array_equal(one.gpu_trace, many.gpu_trace)
    for k in one.gpu_stats:
        assert np.array_equal(one.gpu_stats[k], many.gpu_stats[k], equal_nan=True), k
    assert np.array_equal(one.gpu_var, many.gpu_var)
    assert np.array_equal(one.gpu_adapt, many.gpu_adapt)


@pytest.mark.parametrize("group", [32, 64, 128, 256, 512])
def test_every_group_shape_agrees(group):
    """D=100 forced through each threads-per-chain shape, with the tree scratch in shared memory and all-global."""
    for smem in (-1, 0):
        res = pu.run_case_on_gpu_and_oracle("nuts_static_d100", knobs=dict(group=group, smem_vecs=smem))
        pu.assert_parity(res, rtol=RTOL)


@pytest.mark.parametrize("group", [128, 256, 512, 1024])
def test_d1000_shapes(group):
    res = pu.run_case_on_gpu_and_oracle("nuts_illcond_d1000", knobs=dict(group=group))
    print(pu.parity_report(res))
    pu.assert_parity(res, rtol=RTOL)


def test_fixtures_reach_the_depths_the_baseline_configs_run():
    """BASELINE configs run max_treedepth 10 and 12: the committed reference fixtures must reach those depths (stack
    levels 8..11, proposal slots 9..12, the all-global part of the scratch), and one must have tuning trees deeper
    than max_treedepth (early_max_treedepth > max_treedepth, legal in the reference: nuts.py:205-208)."""
    depth = {n: gc.load(n)[1]["stat_depth"].max() for n in gc.CASE_NAMES if n.startswith("nuts")}
    assert depth["nuts_deep_d1000"] == 12 and depth["nuts_deep_funnel_d50"] == 12 and depth["nuts_deep_d100"] == 10
    case, ref = gc.load("nuts_early_gt_max_d20")
    assert int(case["max_treedepth"]) == 5 and ref["stat_depth"][:, :int(case["tune"])].max() == 8


@pytest.mark.parametrize("smem", [-1, 0, 3])
@pytest.mark.parametrize("group", [0, 32, 64, 256])
def test_deep_trees_every_scratch_placement(group, smem):
    """Depth-10 trees at D=100 through one warp per chain, a CTA per chain, with the scratch in shared memory (as much
    as fits), all-global, and split after three vectors."""
    res = pu.run_case_on_gpu_and_oracle("nuts_deep_d100", knobs=dict(group=group, smem_vecs=smem))
    pu.assert_parity(res, rtol=RTOL)


@pytest.mark.parametrize("callback", [None, "torch"])
def test_reached_max_treedepth_flag(callback):
    """The for/else of nuts.py:212-220 (tree loop exhausted without divergence or U-turn, what
    NUTS._reached_max_treedepth counts) with a tree depth cap low enough to be hit often."""
    res = pu.run_case_on_gpu_and_oracle("nuts_diag_d37", n_trans=40, callback=callback,
                                        overrides=dict(max_treedepth=2, early_max_treedepth=2))
    pu.assert_parity(res, rtol=RTOL)
    hits = res.cpu_stats["reached_max_treedepth"]
    assert 0 < hits.sum() < hits.size, "the case must exercise both outcomes"
    # a transition that stopped at the cap by a U-turn has depth == cap but is NOT counted (statistics alone cannot
    # tell the two apart, the kernel's flag can)
    at_cap = (res.cpu_stats["depth"] == 2) & (res.cpu_stats["diverging"] == 0)
    assert (at_cap & (hits == 0)).any()


@pytest.mark.parametrize("callback", [None, "torch"])
def test_step_rand_hook(callback):
    """BaseHMC.step_rand (base_hmc.py:154-155): the step size every transition integrates with is the hook's output,
    while dual averaging keeps adapting its own state."""
    jitter = lambda eps: eps * 0.8 + 0.01   # noqa: E731  (deterministic so both sides see the same numbers)
    res = pu.run_case_on_gpu_and_oracle("nuts_diag_d37", n_trans=30, callback=callback, overrides=dict(step_rand=jitter))
    pu.assert_parity(res, rtol=RTOL)
    plain = pu.run_case_on_gpu_and_oracle("nuts_diag_d37", n_trans=30, callback=callback)
    assert not np.array_equal(res.gpu_trace, plain.gpu_trace)


@pytest.mark.parametrize("group", [-64, -128])
@pytest.mark.parametrize("name", ["nuts_illcond_d1000", "nuts_diag_d37", "nuts_static_d100", "nuts_funnel_d10", "nuts_b1_d10",
                                  "nuts_deep_d1000", "nuts_deep_funnel_d50", "nuts_early_gt_max_d20"])
def test_lean_kernel_parity(name, group):
    """The lean NUTS kernel (lmc_sampler_lean.cuh: only q, p, grad in registers; selected with a negative group knob)
    against the oracle, transition level, with its scratch in shared memory and all-global."""
    for smem in (-1, 0):
        res = pu.run_case_on_gpu_and_oracle(name, knobs=dict(group=group, smem_vecs=smem))
        pu.assert_parity(res, rtol=RTOL)


@pytest.mark.parametrize("name", ["nuts_diag_d37", "nuts_funnel_d10"])
def test_lean_kernel_equals_default_kernel_on_chained_runs(name):
    """Same group width => same reduction tree => bit-identical chained runs (adaptation included)."""
    a = pu.run_case_on_gpu_and_oracle(name, n_trans=80, chained=True, knobs=dict(group=64))
    b = pu.run_case_on_gpu_and_oracle(name, n_trans=80, chained=True, knobs=dict(group=-64))
    assert np.array_equal(a.gpu_trace, b.gpu_trace)
    for k in a.gpu_stats:
        assert np.array_equal(a.gpu_stats[k], b.gpu_stats[k], equal_nan=True), k
    assert np.array_equal(a.gpu_var, b.gpu_var) and np.array_equal(a.gpu_adapt, b.gpu_adapt)


WARP_CASES = ["nuts_b1_d10", "nuts_diag_d37", "nuts_static_d100", "nuts_funnel_d10", "nuts_deep_d100",
              "nuts_deep_funnel_d50", "nuts_deep_funnel_d10", "nuts_early_gt_max_d20"]


@pytest.mark.parametrize("chunk", [4, 8, 16])
@pytest.mark.parametrize("name", WARP_CASES)
def test_warp_kernel_parity(name, chunk):
    """The chunked warp-per-chain NUTS kernel (lmc_sampler_warp.cuh; the default up to 256 dimensions, forced here with
    group=1) for every chunk length, with the trajectory vectors in the global workspace and in shared memory."""
    for smem in (0, 3):
        res = pu.run_case_on_gpu_and_oracle(name, knobs=dict(group=1, chunk=chunk, smem_vecs=smem))
        pu.assert_parity(res, rtol=RTOL)
    # fewer resident warps than chains: the FIFO scheduler instead of the sticky chain -> warp assignment
    res = pu.run_case_on_gpu_and_oracle(name, knobs=dict(group=1, chunk=chunk, max_slots=1))
    pu.assert_parity(res, rtol=RTOL)


CTA_CASES = ["nuts_illcond_d1000", "nuts_deep_d1000", "nuts_diag_d37", "nuts_static_d100", "nuts_funnel_d10",
             "nuts_deep_d100", "nuts_deep_funnel_d50", "nuts_deep_funnel_d10", "nuts_early_gt_max_d20", "nuts_b1_d10"]


@pytest.mark.parametrize("chunk", [2, 4])
@pytest.mark.parametrize("name", CTA_CASES)
def test_cta_kernel_parity(name, chunk):
    """The chunked CTA-per-chain NUTS kernel (lmc_sampler_cta.cuh: 128 threads per chain, the default for 513..1024
    dimensions; forced here with group=2 for every fixture) for both chunk lengths, with the tree scratch in the global
    workspace and as much of it in shared memory as fits, sticky and through the FIFO scheduler."""
    for smem in (0, -1):
        res = pu.run_case_on_gpu_and_oracle(name, knobs=dict(group=2, chunk=chunk, smem_vecs=smem))
        pu.assert_parity(res, rtol=RTOL)
    # fewer resident CTAs than chains: the FIFO scheduler instead of the sticky chain -> CTA assignment
    res = pu.run_case_on_gpu_and_oracle(name, knobs=dict(group=2, chunk=chunk, max_slots=1))
    pu.assert_parity(res, rtol=RTOL)


def test_cta_kernel_occupancy_variants():
    """chunk of 2 with three / four resident CTAs per SM (group 203 / 204): same results"""
    for group in (203, 204):
        res = pu.run_case_on_gpu_and_oracle("nuts_illcond_d1000", knobs=dict(group=group, chunk=2))
        pu.assert_parity(res, rtol=RTOL)


@pytest.mark.parametrize("name", ["nuts_diag_d37", "nuts_funnel_d10"])
def test_warp_kernel_agrees_with_the_one_warp_register_kernel(name):
    """Chained adaptive runs of the chunked kernel and of the round-1 one-warp kernel (group=32): same decisions, floats
    to rounding (the two sum their dot products in different orders), for as long as the chaotic feedback allows."""
    a = pu.run_case_on_gpu_and_oracle(name, n_trans=12, chained=True, knobs=dict(group=1))
    b = pu.run_case_on_gpu_and_oracle(name, n_trans=12, chained=True, knobs=dict(group=32))
    for k in ("depth", "tree_size", "diverging"):
        assert np.array_equal(a.gpu_stats[k], b.gpu_stats[k]), k
    np.testing.assert_allclose(a.gpu_trace, b.gpu_trace, rtol=1e-8, atol=1e-11)


def test_chained_adaptive_run_tracks_the_reference_until_chaos_takes_over():
    """A long chained run with both adaptations (235 transitions, nuts_diag_d37) is chaotic: a last-bit difference in a
    dot product is amplified through the step-size / mass-matrix feedback until a tree decision flips, after which the
    chains are different (equally valid) chains -- the same happens inside the CPU oracle when its step size is nudged by
    one ulp.  The test states the number: the device run must follow the reference EXACTLY (tree sizes) and to 1e-6 (draws)
    for at least the first 25 transitions of every chain, and it reports where the first decision actually flips."""
    res = pu.run_case_on_gpu_and_oracle("nuts_diag_d37", chained=True)
    same = res.gpu_stats["tree_size"] == res.cpu_stats["tree_size"]
    first_flip = [int(np.argmin(row)) if not row.all() else row.size for row in same]
    print("first transition whose tree size differs from the reference, per chain:", first_flip, "of", same.shape[1])
    assert min(first_flip) >= 25
    n = min(first_flip)
    np.testing.assert_allclose(res.gpu_trace[:, :25], res.cpu_trace[:, :25], rtol=1e-6, atol=1e-9)
    assert n <= same.shape[1]
