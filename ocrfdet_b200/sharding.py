"""(sample, view) sharding of the render path over the GPUs of one box, and its single collective.

The reference scales by plain data parallelism over samples (MMDistributedDataParallel,
/root/reference/mmdet3d/apis/train.py:227-231) and its render path issues no collective; the only
cross-sample assembly point is `torch.cat(opacity_alpha_list, 0)` feeding the HOA lift
(/root/reference/mmdet3d/models/necks/view_transformer_ocrf.py:1196).  Here every rank renders a
contiguous block of (sample, view) pairs -- a sample's views stay on one rank so its Gaussian
gradients need no reduction -- and the per-view opacity maps are assembled with ONE all-gather
(NCCL over NVLink on the GPU box, gloo in the CPU tests), issued on a side stream so it overlaps the
backward pass.
"""
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


def gather_opacity_maps(local_maps: torch.Tensor, num_samples: int, views_per_sample: int, group=None,
                        stream: "torch.cuda.Stream" = None) -> torch.Tensor:
    """All-gather per-view opacity maps [V_local,1,H,W] into [num_samples*views_per_sample,1,H,W]
    (global sample-major order).  Works for uneven shards (pads to the largest shard).
    With `stream`, the collective is enqueued there after the producer stream's current work and the
    caller must `torch.cuda.current_stream().wait_stream(stream)` before consuming the result."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_maps
    world = dist.get_world_size(group)
    counts = [(shard_samples(num_samples, world, r)[1] - shard_samples(num_samples, world, r)[0]) * views_per_sample
              for r in range(world)]
    vmax = max(counts)
    H, W = local_maps.shape[-2:]
    send = local_maps
    if local_maps.shape[0] != vmax:
        send = local_maps.new_zeros((vmax, 1, H, W))
        send[:local_maps.shape[0]] = local_maps
    out = local_maps.new_empty((world * vmax, 1, H, W))

    def run():
        dist.all_gather_into_tensor(out, send.contiguous(), group=group)

    if stream is not None and local_maps.is_cuda:
        stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(stream):
            run()
        send.record_stream(stream)
    else:
        run()
    if all(c == vmax for c in counts):
        return out
    parts = [out[r * vmax:r * vmax + counts[r]] for r in range(world)]
    if stream is not None and local_maps.is_cuda:
        with torch.cuda.stream(stream):
            return torch.cat(parts, 0)
    return torch.cat(parts, 0)
