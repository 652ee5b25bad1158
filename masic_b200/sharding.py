"""Multi-GPU host logic: one process per GPU, `torch.distributed` for the plumbing.

Inference (BASELINE.json configs 2-4): the independent unit is the stereo pair.  Rank r
owns pairs r, r+W, r+2W, ... with a full weight replica; there is NO collective on the data path — only
the per-pair scalars (bpp, PSNR, bitstream sizes) are gathered at the end.  A pair's result must not depend
on the rank or batch it ran in (tests/test_hsic_gpu.py asserts bit-identity on the device;
tests/test_sharding_cpu.py covers this module with world_size-2 gloo groups).

Training (config 5): data parallel, the one collective is an all-reduce (sum, then / world) of the fp32
gradients — `GradBuckets` flattens them into a few large buffers (NVSwitch: size buckets for launch latency,
not link count) and writes the averaged values back into `.grad`.  The reference has no multi-GPU training
(single process, `torch.cuda.set_device(args.cuda)`, newtrain_codec_real.py:369-372); the loss normalises by
the LOCAL batch (`:76`), so averaging gradients reproduces a single-GPU step on the concatenated batch.
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist


def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_pairs(n_pairs: int, rank: int, world: int) -> List[int]:
    """Pair indices owned by `rank`: r, r+W, ... (round-robin keeps ragged tails within one pair)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    if n_pairs < 0:
        raise ValueError("n_pairs must be >= 0")
    return list(range(rank, n_pairs, world))


def gather_pair_results(local: Dict[int, Sequence[float]], n_pairs: int, width: int, group=None,
                        device: Optional[torch.device] = None) -> Optional[torch.Tensor]:
    """Collect per-pair result rows (`width` floats each, keyed by global pair index) on rank 0.

    Returns a (n_pairs, width) float64 tensor on rank 0 (None elsewhere).  Implemented as one all_gather of a
    fixed-size padded block per rank (works on gloo and nccl; `device` is where the exchange buffers live).
    Raises if a pair is missing or reported by two ranks.
    """
    rank, world = _world(group)
    per_rank = -(-n_pairs // world) if n_pairs else 0
    block = torch.full((max(per_rank, 1), width + 1), -1.0, dtype=torch.float64, device=device)
    if len(local) > per_rank:
        raise ValueError(f"rank {rank} reports {len(local)} pairs, more than its share of {per_rank}")
    for i, (idx, row) in enumerate(sorted(local.items())):
        if len(row) != width:
            raise ValueError(f"pair {idx}: expected {width} values, got {len(row)}")
        block[i, 0] = float(idx)
        block[i, 1:] = torch.as_tensor(list(row), dtype=torch.float64)
    if world == 1:
        blocks = [block]
    else:
        blocks = [torch.empty_like(block) for _ in range(world)]
        dist.all_gather(blocks, block, group=group)
    if rank != 0:
        return None
    out = torch.full((n_pairs, width), float("nan"), dtype=torch.float64)
    seen = torch.zeros(n_pairs, dtype=torch.int32)
    for b in blocks:
        b = b.cpu()
        for row in b:
            idx = int(row[0])
            if idx < 0:
                continue
            if idx >= n_pairs:
                raise ValueError(f"pair index {idx} out of range")
            seen[idx] += 1
            out[idx] = row[1:]
    if n_pairs and not bool((seen == 1).all()):
        bad = [i for i in range(n_pairs) if int(seen[i]) != 1]
        raise ValueError(f"pairs missing or duplicated across ranks: {bad[:8]}")
    return out


def evaluate_sharded(n_pairs: int, load_pair: Callable[[int], tuple], run_pair: Callable[..., Sequence[float]],
                     width: int, group=None, device: Optional[torch.device] = None) -> Optional[torch.Tensor]:
    """The sharded eval loop: every rank runs `run_pair(*load_pair(i))` for its own pairs; rank 0 gets the table."""
    rank, world = _world(group)
    local = {i: run_pair(*load_pair(i)) for i in shard_pairs(n_pairs, rank, world)}
    return gather_pair_results(local, n_pairs, width, group=group, device=device)


class _MeanWork:
    """Handle of all_reduce_mean_: wait() blocks (stream-side on NCCL) until the buffer holds the mean."""

    def __init__(self, work, tensor: Optional[torch.Tensor], scale: float):
        self._work, self._t, self._scale = work, tensor, scale

    def wait(self):
        if self._work is not None:
            self._work.wait()
            self._work = None
        if self._t is not None:                      # backends without ReduceOp.AVG: sum, then scale
            self._t.mul_(self._scale)
            self._t = None


def all_reduce_mean_(t: torch.Tensor, group=None, async_op: bool = False) -> _MeanWork:
    """In-place mean of `t` over the ranks of `group` — the ONE collective of the data-parallel training step
    (HSICTrainer.train_step: the flat fp32 gradient buffer, or a contiguous slice of it).  NCCL takes the mean inside the
    collective (ReduceOp.AVG); other backends sum and scale.  Always returns a handle; with async_op=False it has
    already completed."""
    _, world = _world(group)
    if world == 1:
        return _MeanWork(None, None, 1.0)
    if dist.get_backend(group) == "nccl":
        w = dist.all_reduce(t, op=dist.ReduceOp.AVG, group=group, async_op=async_op)
        return _MeanWork(w if async_op else None, None, 1.0)
    w = dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
    h = _MeanWork(w if async_op else None, t, 1.0 / world)
    if not async_op:
        h.wait()
    return h


class GradBuckets:
    """Flat fp32 gradient buckets for the data-parallel all-reduce.

    `params` is the ordered parameter list (main + aux); buckets are filled in REVERSE order (the order
    backward produces gradients) up to `bucket_bytes`.  `allreduce_()` launches one all_reduce(SUM) per bucket
    (asynchronously when `async_op`), scales by 1/world and scatters the values back into `.grad`.
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 64 << 20, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.buckets: List[List[torch.nn.Parameter]] = []
        cur, size = [], 0
        for p in reversed(self.params):
            nb = p.numel() * 4
            if cur and size + nb > bucket_bytes:
                self.buckets.append(cur)
                cur, size = [], 0
            cur.append(p)
            size += nb
        if cur:
            self.buckets.append(cur)
        self._flat: List[Optional[torch.Tensor]] = [None] * len(self.buckets)

    def _flatten(self, i: int) -> torch.Tensor:
        ps = self.buckets[i]
        n = sum(p.numel() for p in ps)
        flat = self._flat[i]
        if flat is None or flat.device != ps[0].device:
            flat = self._flat[i] = torch.empty(n, dtype=torch.float32, device=ps[0].device)
        off = 0
        for p in ps:
            k = p.numel()
            if p.grad is None:
                flat[off:off + k].zero_()
            else:
                flat[off:off + k].copy_(p.grad.reshape(-1))
            off += k
        return flat

    def _scatter(self, i: int, scale: float):
        flat, off = self._flat[i], 0
        for p in self.buckets[i]:
            k = p.numel()
            g = flat[off:off + k].view_as(p)
            if p.grad is None:
                p.grad = (g * scale).clone()
            else:
                torch.mul(g, scale, out=p.grad)
            off += k

    def allreduce_(self):
        rank, world = _world(self.group)
        if world == 1:
            return
        works = []
        for i in range(len(self.buckets)):
            flat = self._flatten(i)
            works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        for i, w in enumerate(works):
            w.wait()
            self._scatter(i, 1.0 / world)
