"""The path's single collective on real GPUs (needs >= 2): content of the gathered opacity maps for both transports
(copy-engine peer pushes over symmetric memory, NCCL all-gather), on a side stream, repeated (buffer reuse), uneven
shards, and the gradient of the gather.  Run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_nccl_gather.py`."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        from ocrfdet_b200.sharding import _PeerBuffers, gather_opacity_maps, shard_samples
        H, W, vps = 64, 176, 6
        res = {}
        for transport in ("peer", "nccl"):
            for num_samples in (world, world + 1):  # even and uneven shards
                b, e = shard_samples(num_samples, world, rank)
                side = torch.cuda.Stream()
                for rep in range(3):  # the peer transport reuses one symmetric buffer: later rounds must not race
                    local = torch.stack([torch.full((1, H, W), float(1000 * rep + s * vps + v), device=dev)
                                         for s in range(b, e) for v in range(vps)])
                    local.requires_grad_(True)
                    out = gather_opacity_maps(local, num_samples, vps, stream=side, transport=transport)
                    torch.cuda.current_stream().wait_stream(side)
                    want = torch.arange(num_samples * vps, device=dev, dtype=torch.float32) + 1000 * rep
                    ok = bool((out == want.view(-1, 1, 1, 1)).all()) and tuple(out.shape) == (num_samples * vps, 1, H, W)
                    w = torch.arange(1, out.shape[0] + 1, device=dev, dtype=torch.float32).view(-1, 1, 1, 1)
                    (out * w).sum().backward()
                    gw = w[b * vps:e * vps].expand(-1, 1, H, W)
                    ok = ok and bool(torch.equal(local.grad, gw))
                    res[(transport, num_samples, rep)] = ok
        res["peer_used"] = _PeerBuffers._failed is None and len(_PeerBuffers._cache) > 0
        res["peer_error"] = str(_PeerBuffers._failed)
        res["graphed"] = all(pb._graph is not None for pb in _PeerBuffers._cache.values())
        q.put((rank, res))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as ex:  # noqa: BLE001
        q.put((rank, {"exception": repr(ex)}))
        raise


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_opacity_map_gather_content_on_gpus():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=300) for _ in procs]
    [p.join(timeout=120) for p in procs]
    for rank, r in res:
        assert "exception" not in r, r
        bad = [k for k, v in r.items() if isinstance(k, tuple) and not v]
        assert not bad, "rank %d: wrong gather content / gradient for %s" % (rank, bad)
        assert r["peer_used"], "the copy-engine transport was not used: %s" % r["peer_error"]
        assert r["graphed"], "the peer exchange was not captured as a CUDA graph"
    assert all(p.exitcode == 0 for p in procs)
