"""CPU suite: the N > 1 host logic -- (sample, view) sharding and the opacity-map all-gather -- over gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ocrfdet_b200.sharding import gather_opacity_maps, shard_samples, shard_views


def test_shards_partition_the_samples():
    for n in (1, 2, 7, 8, 13):
        for world in (1, 2, 3, 8):
            blocks = [shard_samples(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [e - b for b, e in blocks]
            assert max(sizes) - min(sizes) <= 1
    pairs = sum((shard_views(8, 6, 4, r) for r in range(4)), [])
    assert pairs == [(s, v) for s in range(8) for v in range(6)]
    assert {s for s, _ in shard_views(8, 6, 4, 1)} == {2, 3}  # a sample's views stay on one rank
    with pytest.raises(ValueError):
        shard_samples(4, 2, 2)


def _worker(rank, world, port, num_samples, vps, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, e = shard_samples(num_samples, world, rank)
    H, W = 4, 6
    local = torch.stack([torch.full((1, H, W), float(s * vps + v)) for s in range(b, e) for v in range(vps)]) \
        if e > b else torch.zeros((0, 1, H, W))
    local.requires_grad_(True)
    out = gather_opacity_maps(local, num_samples, vps)
    # the gather is differentiable like the torch.cat it replaces (view_transformer_ocrf.py:1196)
    weight = torch.arange(1, out.shape[0] + 1, dtype=torch.float32).view(-1, 1, 1, 1) * (rank + 1)
    (out * weight).sum().backward()
    g_local = local.grad[:, 0, 0, 0].tolist() if e > b else []
    local.grad = None
    out_sum = gather_opacity_maps(local, num_samples, vps, grad="sum")
    (out_sum * weight).sum().backward()
    g_sum = local.grad[:, 0, 0, 0].tolist() if e > b else []
    q.put((rank, out[:, 0, 0, 0].tolist(), tuple(out.shape), (b * vps, e * vps), g_local, g_sum))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("num_samples,vps", [(2, 6), (3, 2)])
def test_opacity_map_all_gather_world2_gloo(num_samples, vps):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, num_samples, vps, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in procs]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    for rank, vals, shape, (vb, ve), g_local, g_sum in res:
        assert shape == (num_samples * vps, 1, 4, 6)
        assert vals == [float(i) for i in range(num_samples * vps)]  # global sample-major order on every rank
        # grad="local": this rank's own loss only (weight (i+1)*(rank+1)); grad="sum": over both ranks' losses
        assert g_local == [float((i + 1) * (rank + 1)) for i in range(vb, ve)]
        assert g_sum == [float((i + 1) * 3) for i in range(vb, ve)]


def test_single_process_gather_is_identity():
    x = torch.arange(12.0).reshape(3, 1, 2, 2)
    assert gather_opacity_maps(x, 1, 3) is x
