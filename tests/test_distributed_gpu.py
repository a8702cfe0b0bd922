"""The multi-GPU driver on a real NCCL process group (world of one rank: the GPU test box has one GPU; N = 2..8 is
exercised by bench.py --gpus N and, for the host logic, by tests/test_distributed_cpu.py on gloo)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_distributed_sample_on_one_rank_equals_sample():
    """The N>1 driver on a world of one NCCL rank: same draws as the single-GPU driver, gathered shape intact."""
    import os
    import socket
    import torch
    import torch.distributed as dist
    import gc
    import littlemcmc_b200 as lmc
    gc.collect()
    torch.cuda.empty_cache()   # communicator creation is slow when the caching allocator holds many large segments
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()  # noqa: E702
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda:0"))
    try:
        D = 20
        tgt = lmc.targets.DiagGaussian(sigma=np.linspace(0.5, 2, D))
        tr_d, st_d = lmc.distributed.sample(tgt, D, draws=15, tune=15, chains=64, random_seed=9)
        tr_s, st_s = lmc.sample(tgt, D, draws=15, tune=15, chains=64, random_seed=9, return_device=True)
        assert torch.equal(tr_d, tr_s)
        for k in st_s:
            assert torch.equal(st_d[k], st_s[k]), k
    finally:
        dist.destroy_process_group()
