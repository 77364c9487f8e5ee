"""CPU: the N>1 path -- round-robin clip sharding identical to the reference's and the single end-of-run gather,
exercised with world_size 2 over gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neurons_b200 import sharding


def test_round_robin_matches_reference_formula():
    # reference: get_original_index(machine_id, local_index, interval=num_devices) = machine_id + local_index*interval
    for world in (1, 2, 4, 8):
        seen = []
        for rank in range(world):
            idx = sharding.shard_indices(1200, rank, world)
            assert idx == [sharding.original_index(rank, k, world) for k in range(len(idx))]
            seen += idx
        assert sorted(seen) == list(range(1200))          # every clip exactly once
    assert sharding.shard_indices(5, 1, 2) == [1, 3]
    with pytest.raises(ValueError):
        sharding.shard_indices(5, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, num_clips, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        idx = sharding.shard_indices(num_clips, rank, world)
        # each "decoded clip" is filled with its global index so order mistakes are visible
        local = torch.stack([torch.full((3, 2, 4, 4), float(i)) for i in idx]) if idx else torch.zeros((0, 3, 2, 4, 4))
        out = sharding.gather_clips(local, num_clips)
        if rank == 0:
            ok = out is not None and out.shape == (num_clips, 3, 2, 4, 4) and all(
                bool((out[i] == float(i)).all()) for i in range(num_clips))
            ret.put(bool(ok))
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("num_clips", [8, 5])        # even and uneven shards
def test_gather_world2_gloo(num_clips):
    ctx = mp.get_context("spawn")
    ret = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, num_clips, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get() is True
