"""GPU (-m gpu, needs >= 2 GPUs; skipped on a one-GPU box): the single collective of the path -- the end-of-run gather of the decoded
frames (north_star; reference side scripts/neuroclips_video_enhance.py:39-40,324) -- over NCCL with CUDA tensors, uneven shards included.
One process per GPU, rendezvous on 127.0.0.1."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neurons_b200 import sharding

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, num_clips, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        idx = sharding.shard_indices(num_clips, rank, world)
        shape = (3, 16, 64, 64)                                   # a (small) decoded clip: [3, F, H, W]
        local = (torch.stack([torch.full(shape, float(i), device=dev, dtype=torch.float16) for i in idx]) if idx
                 else torch.zeros((0,) + shape, device=dev, dtype=torch.float16))
        out = sharding.gather_clips(local, num_clips)
        torch.cuda.synchronize()
        if rank == 0:
            ok = out is not None and out.is_cuda and out.shape == (num_clips,) + shape and all(
                bool((out[i] == float(i)).all()) for i in range(num_clips))
            ret.put(bool(ok))
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("num_clips", [8, 5])                     # even and uneven shards
def test_gather_clips_nccl(num_clips):
    world = min(torch.cuda.device_count(), 4)
    ctx = mp.get_context("spawn")
    ret = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, num_clips, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert ret.get() is True
