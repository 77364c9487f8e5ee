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


def _enhance_worker(rank, world, port, num_clips, ret):
    """The whole N2 clip loop on 2 ranks: stage-3 inputs sharded round-robin, every clip through enhance_clip with a toy denoiser,
    final latents gathered to rank 0 in the global clip order; compared with the single-process run of the same loop."""
    from neurons_b200 import pipeline
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        inputs = pipeline.Stage3Inputs(torch.rand(num_clips, 3, 8, 8, generator=g), torch.rand(num_clips, 6, 3, 8, 8, generator=g),
                                       [f"clip {i}" for i in range(num_clips)])
        encode = lambda frames: frames.mean(dim=1, keepdim=True).repeat(1, 4, 1, 1)[..., ::2, ::2]       # [n,3,8,8] -> [n,4,4,4] "latents"
        text = lambda caption: torch.full((2, 3, 4), float(len(caption)))
        make = lambda key_lat: (lambda x2, t, ctx: 0.1 * x2 + 0.01 * key_lat.mean() + 1e-3 * ctx.mean())
        kw = dict(num_inference_steps=3, guidance_scale=8.5, low_strength=0.3, seed=0)
        mine = pipeline.enhance_shard(inputs, rank, world, encode, text, make, **kw)
        assert [i for i, _ in mine] == sharding.shard_indices(num_clips, rank, world)
        local = torch.stack([lat[0] for _, lat in mine]) if mine else torch.zeros((0, 4, 16, 4, 4))
        out = sharding.gather_clips(local, num_clips)
        if rank == 0:
            alone = pipeline.enhance_shard(inputs, 0, 1, encode, text, make, **kw)
            ref = torch.stack([lat[0] for _, lat in alone])
            ret.put(bool(out is not None and out.shape == ref.shape and torch.equal(out, ref)))
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


def test_enhance_shard_world2_gloo():
    ctx = mp.get_context("spawn")
    ret = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_enhance_worker, args=(r, 2, port, 5, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert ret.get() is True
