"""(sample, view) sharding of the render path over the GPUs of one box, and its single collective.

The reference scales by plain data parallelism over samples (MMDistributedDataParallel,
/root/reference/mmdet3d/apis/train.py:227-231) and its render path issues no collective; the only
cross-sample assembly point is `torch.cat(opacity_alpha_list, 0)` feeding the HOA lift
(/root/reference/mmdet3d/models/necks/view_transformer_ocrf.py:1196).  Here every rank renders a
contiguous block of (sample, view) pairs -- a sample's views stay on one rank so its Gaussian
gradients need no reduction -- and the per-view opacity maps are assembled with ONE all-gather.

On the GPU box the gather does not run on the SMs.  The blend backward it overlaps with is bound by FP32 issue on
all 148 SMs, and an NCCL all-gather takes SMs away from it (measured in round 1: the backward went from 0.268 ms
at 1 GPU to 0.2875 ms at 8).  `gather_opacity_maps` therefore PUSHES the local maps into a symmetric buffer of
every peer with plain device-to-device copies over the NVLink peer mapping -- copy-engine traffic, issued on a side
stream -- bracketed by two single-CTA signal-pad barriers (torch symmetric memory: cuMem allocations exchanged over
the process group's store, no NVSHMEM).  `transport="nccl"` (and every non-CUDA tensor: gloo in the CPU tests)
uses `all_gather_into_tensor`.

The gather is differentiable like the `torch.cat` it replaces: the backward hands every rank the gradient slice of
its own maps (`grad="local"`, the consumer is replicated on every rank as in the reference's data parallelism, so
the local loss already carries the whole gradient), or the sum of that slice over all ranks (`grad="sum"`, one
reduce-scatter, for consumers that compute rank-specific losses from the gathered maps).
"""
import os
import warnings
from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_samples(num_samples: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [begin, end) of samples for `rank`; blocks differ in size by at most one."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    base, extra = divmod(num_samples, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_views(num_samples: int, views_per_sample: int, world_size: int, rank: int) -> List[Tuple[int, int]]:
    """The (sample, view) pairs of `rank`, sample-major; all views of a sample land on the same rank."""
    b, e = shard_samples(num_samples, world_size, rank)
    return [(s, v) for s in range(b, e) for v in range(views_per_sample)]


class _PeerBuffers:
    """Symmetric gather buffers [world, vmax, 1, H, W], one allocation per (group, shape), mapped into every rank."""
    _cache = {}
    _failed = None

    def __init__(self, group, vmax, H, W, dtype, device):
        import torch.distributed._symmetric_memory as symm
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        n = self.world * vmax * H * W
        self.buf = symm.empty(n, dtype=dtype, device=device)
        pg = group if group is not None else dist.group.WORLD
        self.hdl = symm.rendezvous(self.buf, pg)
        shape = (self.world, vmax, 1, H, W)
        self.views = [self.hdl.get_buffer(r, shape, dtype) for r in range(self.world)]

    @classmethod
    def get(cls, group, vmax, H, W, dtype, device):
        key = (id(group), vmax, H, W, dtype, str(device))
        if key not in cls._cache:
            cls._cache[key] = cls(group, vmax, H, W, dtype, device)
        return cls._cache[key]

    def all_gather(self, send):
        """Runs on the CURRENT stream.  Returns this rank's buffer [world, vmax, 1, H, W]; it is overwritten by the
        next gather of the same shape (the opening barrier of that call waits until every rank got there, i.e. is
        done with the previous contents in stream order).

        Ten stream operations per step (two barriers, `world` copies) cost the launching thread ~0.1 ms at 8 ranks --
        with eight processes on one host that showed up as jitter that every rank then waits for.  The exchange
        [barrier, pushes of my slot to the peers, barrier] has fixed addresses, so it is captured once as a CUDA graph:
        a gather is the local copy into my slot plus ONE graph launch.  OCRF_GATHER_GRAPH=0 keeps the eager sequence."""
        mine = self.views[self.rank][self.rank]
        if self._graph is None and self._graph_ok:
            try:
                self._capture()
            except Exception as e:  # noqa: BLE001  (capture of the signal-pad barrier not supported: stay eager)
                self._graph_ok = False
                warnings.warn("opacity-map gather: CUDA-graph capture of the peer exchange failed (%s); issuing it "
                              "eagerly" % (e,))
        if self._graph is not None:
            mine.copy_(send)
            self._graph.replay()
            return self.views[self.rank]
        self.hdl.barrier(channel=0)
        for i in range(self.world):  # start with myself, then ring order: no two ranks push to the same peer first
            peer = (self.rank + i) % self.world
            self.views[peer][self.rank].copy_(send)  # contiguous same-dtype copy: cudaMemcpyAsync -> copy engine
        self.hdl.barrier(channel=1)  # every rank's pushes are complete and visible
        return self.views[self.rank]

    _graph = None
    _graph_ok = os.environ.get("OCRF_GATHER_GRAPH", "1") != "0"

    def _exchange(self):
        mine = self.views[self.rank][self.rank]
        self.hdl.barrier(channel=0)  # every peer is done with the previous contents of ITS buffer
        for i in range(1, self.world):
            peer = (self.rank + i) % self.world
            self.views[peer][self.rank].copy_(mine)
        self.hdl.barrier(channel=1)

    def _capture(self):
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            self._exchange()  # warm-up outside the capture (collective: every rank does the same)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._exchange()
        self._graph = g


def _nccl_all_gather(send, world, vmax, group):
    out = send.new_empty((world * vmax,) + tuple(send.shape[1:]))
    dist.all_gather_into_tensor(out, send, group=group)
    return out.view((world, vmax) + tuple(send.shape[1:]))


def _gather_raw(send, world, vmax, group, transport):
    """send [vmax,1,H,W] contiguous -> [world, vmax, 1, H, W] on the current stream."""
    if send.is_cuda and transport == "peer" and _PeerBuffers._failed is None:
        try:
            pb = _PeerBuffers.get(group, vmax, send.shape[-2], send.shape[-1], send.dtype, send.device)
        except Exception as e:  # noqa: BLE001  (no peer access / symmetric memory unavailable on this system)
            _PeerBuffers._failed = e
            warnings.warn("opacity-map gather: symmetric-memory peer copies unavailable (%s); using the NCCL "
                          "all-gather, which shares the SMs with the blend backward" % (e,))
        else:
            return pb.all_gather(send)
    return _nccl_all_gather(send, world, vmax, group)


def _gather_forward(local_maps, counts, group, stream, transport, after):
    world = dist.get_world_size(group)
    vmax = max(counts)
    send = local_maps.detach()
    if send.shape[0] != vmax:
        pad = send.new_zeros((vmax,) + tuple(send.shape[1:]))
        pad[:send.shape[0]] = send
        send = pad
    send = send.contiguous()
    side = stream is not None and send.is_cuda
    if side:
        if after is not None:
            stream.wait_event(after)
        else:
            stream.wait_stream(torch.cuda.current_stream())
        send.record_stream(stream)
    with torch.cuda.stream(stream) if side else _null():
        out = _gather_raw(send, world, vmax, group, transport)
        if all(c == vmax for c in counts):
            out = out.view((world * vmax,) + tuple(send.shape[1:]))
        else:
            out = torch.cat([out[r, :counts[r]] for r in range(world)], 0)
    if side:
        out.record_stream(torch.cuda.current_stream())  # produced on `stream`, consumed on the caller's stream
    return out


class _GatherMaps(torch.autograd.Function):
    @staticmethod
    def forward(ctx, local_maps, counts, group, stream, transport, grad_mode, after):
        out = _gather_forward(local_maps, counts, group, stream, transport, after)
        ctx.meta = (counts, dist.get_rank(group), dist.get_world_size(group), group, grad_mode, max(counts))
        return out

    @staticmethod
    def backward(ctx, g):
        counts, rank, world, group, grad_mode, vmax = ctx.meta
        begin = sum(counts[:rank])
        if grad_mode == "local":
            return g[begin:begin + counts[rank]], None, None, None, None, None, None
        # grad == "sum": every rank's consumer produced a gradient for my maps
        padded = g.new_zeros((world, vmax) + tuple(g.shape[1:]))
        off = 0
        for r in range(world):
            padded[r, :counts[r]] = g[off:off + counts[r]]
            off += counts[r]
        mine = g.new_empty((vmax,) + tuple(g.shape[1:]))
        dist.reduce_scatter_tensor(mine, padded.view((world * vmax,) + tuple(g.shape[1:])), group=group)
        return mine[:counts[rank]], None, None, None, None, None, None


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


_COUNTS = {}


def gather_opacity_maps(local_maps: torch.Tensor, num_samples: int, views_per_sample: int, group=None,
                        stream: "torch.cuda.Stream" = None, transport: str = None, grad: str = "local",
                        after: "torch.cuda.Event" = None) -> torch.Tensor:
    """All-gather per-view opacity maps [V_local,1,H,W] into [num_samples*views_per_sample,1,H,W]
    (global sample-major order).  Works for uneven shards (pads to the largest shard).  Differentiable (see the
    module docstring for `grad`).
    With `stream`, the transfer is enqueued there after the producer stream's current work -- or, with `after`, after
    that event only, so a caller may queue the backward first and the gather behind it on the host while the device
    still runs them side by side -- and the caller must `torch.cuda.current_stream().wait_stream(stream)` before
    consuming the result.
    `transport`: "peer" (default on CUDA: copy-engine pushes over NVLink peer mappings; the result aliases a
    persistent symmetric buffer that the next gather of the same shape overwrites) or "nccl"; the environment
    variable OCRF_GATHER overrides the default."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_maps
    if grad not in ("local", "sum"):
        raise ValueError("grad must be 'local' or 'sum'")
    if transport is None:
        transport = os.environ.get("OCRF_GATHER", "peer")
    world = dist.get_world_size(group)
    key = (num_samples, views_per_sample, world)
    counts = _COUNTS.get(key)
    if counts is None:
        counts = _COUNTS[key] = [(shard_samples(num_samples, world, r)[1] - shard_samples(num_samples, world, r)[0])
                                 * views_per_sample for r in range(world)]
    if not (local_maps.requires_grad and torch.is_grad_enabled()):
        return _gather_forward(local_maps, counts, group, stream, transport, after)  # nothing to differentiate
    return _GatherMaps.apply(local_maps, counts, group, stream, transport, grad, after)
