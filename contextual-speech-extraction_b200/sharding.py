"""Multi-GPU host logic: mixtures are independent (GroupNorm(1,.) is per sample, attention never
crosses samples — SURVEY.md §8e), so the batch is partitioned across ranks with NO data-path
collective; only the optional result gather uses torch.distributed (NCCL on GPUs, gloo in tests).
"""
from typing import List, Sequence

import torch
import torch.distributed as dist


def contiguous_shard(n_items: int, rank: int, world: int) -> range:
    """Contiguous slice of a fixed-length batch for `rank` (first ranks take the remainder)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def balanced_shards(lengths: Sequence[int], world: int) -> List[List[int]]:
    """Ragged batches: sort by length (longest first) and deal round-robin in a snake order so
    every rank gets a similar amount of audio; idle ranks get an empty list when B < world."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    shards: List[List[int]] = [[] for _ in range(world)]
    for pos, idx in enumerate(order):
        lap, slot = divmod(pos, world)
        r = slot if lap % 2 == 0 else world - 1 - slot
        shards[r].append(idx)
    return shards


def separate_sharded(model_fn, mix: torch.Tensor, ctx, group=None, gather: bool = True):
    """Run `model_fn(mix_shard, ctx_shard) -> est [b,T,spk]` on this rank's contiguous slice of the
    batch and (optionally) all-gather the separated waveforms so every rank holds [B,T,spk].

    `model_fn` is the CUDA model on GPUs; the collective is result-only (4*T*spk bytes per item).
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = mix.shape[0]
    idx = contiguous_shard(B, rank, world)
    if len(idx) > 0:
        sl = slice(idx.start, idx.stop)
        est_local = model_fn(mix[sl], None if ctx is None else ctx[sl])
    else:
        est_local = None
    if not gather or world == 1:
        return est_local
    # every rank must know the trailing shape even when its own shard is empty
    shape = torch.zeros(2, dtype=torch.int64, device=mix.device)
    if est_local is not None:
        shape[0], shape[1] = est_local.shape[1], est_local.shape[2]
    dist.all_reduce(shape, op=dist.ReduceOp.MAX, group=group)
    T, spk = int(shape[0]), int(shape[1])
    per_rank = max(len(contiguous_shard(B, r, world)) for r in range(world))
    buf = torch.zeros(per_rank, T, spk, dtype=torch.float32, device=mix.device)
    if est_local is not None:
        buf[: est_local.shape[0]] = est_local.float()
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    parts = [out[r][: len(contiguous_shard(B, r, world))] for r in range(world)]
    return torch.cat(parts, dim=0)
