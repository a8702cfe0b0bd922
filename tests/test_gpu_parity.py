"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on identical randomness, and
against the committed reference fixtures."""
import numpy as np
import pytest

from tests import golden_cases as gc
from tests import parity_utils as pu

pytestmark = pytest.mark.gpu

RTOL = 1e-9  # float64 tolerance of the north star's "stated fp64 tolerance"; integer statistics are exact


@pytest.mark.parametrize("name", gc.CASE_NAMES)
def test_cuda_matches_oracle_and_reference(name):
    res = pu.run_case_on_gpu_and_oracle(name)
    pu.assert_parity(res, rtol=RTOL)
    # and against the unmodified reference's own outputs (the oracle consumed the reference's MT19937 stream)
    _, ref = gc.load(name)
    np.testing.assert_allclose(res.gpu_trace, ref["trace"], rtol=RTOL, atol=1e-12)
    for k in res.gpu_stats:
        r = ref["stat_" + k]
        if k in pu.EXACT:
            assert np.array_equal(res.gpu_stats[k], r), k
        else:
            np.testing.assert_allclose(res.gpu_stats[k], r, rtol=RTOL, atol=1e-12, err_msg=k)
    np.testing.assert_allclose(res.gpu_var, ref["final_var"], rtol=RTOL, atol=1e-12)


@pytest.mark.parametrize("name", ["nuts_diag_d37", "hmc_static_d50"])
def test_chunked_calls_are_bit_identical(name):
    """A run split over several launches (state carried in the device buffers) equals one launch bit for bit."""
    one = pu.run_case_on_gpu_and_oracle(name, n_trans=60)
    many = pu.run_case_on_gpu_and_oracle(name, n_trans=60, chunks=7)
    assert np.array_equal(one.gpu_trace, many.gpu_trace)
    for k in one.gpu_stats:
        assert np.array_equal(one.gpu_stats[k], many.gpu_stats[k]), k
    assert np.array_equal(one.gpu_var, many.gpu_var)


@pytest.mark.parametrize("group", [32, 64, 128, 256, 512])
def test_every_group_shape_agrees(group):
    """D=100 forced through each threads-per-chain shape (and all-global / all-shared scratch placement)."""
    for smem in (-1, 0):
        res = pu.run_case_on_gpu_and_oracle("nuts_static_d100", knobs=dict(group=group, smem_vecs=smem))
        pu.assert_parity(res, rtol=RTOL)


@pytest.mark.parametrize("group", [128, 256, 512, 1024])
def test_d1000_shapes(group):
    res = pu.run_case_on_gpu_and_oracle("nuts_illcond_d1000", knobs=dict(group=group))
    pu.assert_parity(res, rtol=RTOL)
