"""Batch sharding of has_conv_fwd_t::run_fwd over the GPUs of one box (the reference has no multi-GPU path).

Every op on the rtc_fwd path is per-image (`img` is the outermost dim of every activation node, src/conv_util.cc:482-503), so the
ranks split the batch and never exchange activations:
  * `shard_range`       -- rank r owns the contiguous images [r*ceil(B/G), min(B, (r+1)*ceil(B/G)))
  * `broadcast_params`  -- ONE collective broadcast of all filts / biases from rank 0 at init (flattened into a single buffer)
  * `gather_logits`     -- ONE all-gather of the per-rank output node per forward
  * `GatherPipeline`    -- the same gather, issued asynchronously one step late so that it overlaps the next batch's forward
Only `torch.distributed` calls; the tensors live wherever the process group's backend wants them (CUDA for NCCL over NVLink on the
GPU box, CPU for the gloo tests), so the same code is exercised by world_size-2 CPU tests.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Tuple

import numpy as np


def shard_range(global_batch: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous split of the `img` dim: [begin, end) for `rank`; trailing ranks may get fewer (or zero) images.
    The gathers below (`gather_logits`, `GatherPipeline`, `b200_shard_gather_*`) need EQUAL shards: use `equal_shard_size` to check a global
    batch before building the per-rank forwards (bench.py is weak-scaling: every rank runs the same per-GPU batch)."""
    if world < 1 or not (0 <= rank < world) or global_batch < 0:
        raise ValueError("shard_range: bad world/rank/batch %r/%r/%r" % (world, rank, global_batch))
    per = -(-global_batch // world)
    b = min(global_batch, rank * per)
    return b, min(global_batch, b + per)


def equal_shard_size(global_batch: int, world: int) -> int:
    """Images per rank when `global_batch` splits evenly over `world` ranks; raises otherwise (an uneven split would make the fixed-size
    all-gather hang or mis-order images -- pad the batch to a multiple of `world` and slice the gathered result with `shard_range`)."""
    if world < 1 or global_batch < 0 or global_batch % world:
        raise ValueError("global batch %d does not split evenly over %d ranks: pad it to %d" % (global_batch, world, -(-global_batch // max(world, 1)) * max(world, 1)))
    return global_batch // world


def param_layout(shapes: Dict[str, Tuple[int, ...]]) -> List[Tuple[str, int, int, Tuple[int, ...]]]:
    """Deterministic (sorted-by-name) packing of the parameter nodes into one flat buffer: [(name, offset, numel, shape)]."""
    out, off = [], 0
    for n in sorted(shapes):
        sz = int(np.prod(shapes[n]))
        out.append((n, off, sz, tuple(shapes[n])))
        off += sz
    return out


def broadcast_params(dist, shapes: Dict[str, Tuple[int, ...]], params_rank0, device="cpu", src: int = 0) -> Dict[str, np.ndarray]:
    """All filts / biases in ONE broadcast from `src` (north_star: "a single NCCL broadcast of weights"). `params_rank0` is a dict of
    numpy arrays on the source rank (ignored elsewhere; may be None). Returns the full dict as host arrays on every rank."""
    import torch
    layout = param_layout(shapes)
    total = layout[-1][1] + layout[-1][2] if layout else 0
    flat = torch.empty(total, dtype=torch.float32, device=device)
    if dist.get_rank() == src:
        if params_rank0 is None:
            raise ValueError("broadcast_params: the source rank needs the parameters")
        host = np.empty(total, np.float32)
        for n, off, sz, shape in layout:
            a = np.asarray(params_rank0[n], np.float32)
            if tuple(a.shape) != shape:
                raise ValueError("broadcast_params: '%s' has shape %r, the net wants %r" % (n, a.shape, shape))
            host[off:off + sz] = a.ravel()
        flat.copy_(torch.from_numpy(host))
    dist.broadcast(flat, src=src)
    host = flat.cpu().numpy()
    return {n: host[off:off + sz].reshape(shape) for n, off, sz, shape in layout}


def gather_logits(dist, local, out=None):
    """All-gather of each rank's output node ([B_local, ...] with equal B_local on every rank) into [world*B_local, ...], rank-major --
    i.e. image order of the global batch under `shard_range`."""
    import torch
    world = dist.get_world_size()
    if out is None:
        out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    if out.shape[0] != world * local.shape[0]:
        raise ValueError("gather_logits: output holds %d rows, %d ranks x %d local rows expected (unequal shards? see equal_shard_size)" % (out.shape[0], world, local.shape[0]))
    dist.all_gather_into_tensor(out, local.contiguous())
    return out


class GatherPipeline:
    """Double-buffered, one-step-late logits gather for serving loops: step i stages its logits (a device copy on the caller's current
    stream), step i+1 opens by issuing the asynchronous all-gather of that staged copy and closes by joining it, so the collective runs
    beside forward i+1 instead of after forward i. Per step:  begin_step(i); <enqueue forward i>; stage(i, logits); out = end_step().
    `end_step` returns the gathered logits of step i-1 (None at i == 0); `drain(i)` after the last step returns the final one.
    With NCCL the joins are stream-level (no host block); with gloo they block the host -- same results either way."""

    def __init__(self, dist, local_like):
        import torch
        self.dist = dist
        world = dist.get_world_size()
        self.staging = [torch.empty_like(local_like) for _ in range(2)]
        self.out = torch.empty((world * local_like.shape[0],) + tuple(local_like.shape[1:]), dtype=local_like.dtype, device=local_like.device)
        self.work = None

    def begin_step(self, i: int):
        if i > 0:
            self.work = self.dist.all_gather_into_tensor(self.out, self.staging[(i - 1) & 1], async_op=True)

    def stage(self, i: int, local):
        self.staging[i & 1].copy_(local, non_blocking=True)

    def end_step(self):
        if self.work is None:
            return None
        self.work.wait()
        self.work = None
        return self.out

    def drain(self, n_steps: int):
        self.begin_step(n_steps)
        return self.end_step()


def max_over_ranks(dist, value: float, device="cpu") -> float:
    """Timing rule for every multi-GPU number: the slowest rank's device time."""
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def split_inputs(x: np.ndarray, world: int) -> Iterable[np.ndarray]:
    """Host-side scatter of a global NCHW batch into per-rank shards (each rank H2D-copies only its own)."""
    return [x[slice(*shard_range(x.shape[0], world, r))] for r in range(world)]
