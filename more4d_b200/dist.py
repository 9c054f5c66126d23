"""Multi-GPU execution of the denoising path: one process per GPU, `torch.distributed` for the
plumbing (NCCL on the NVSwitch domain; gloo in the CPU tests).

The 4D-STraG denoiser shards on the BATCH axis: samples — and the two CFG
branches of one sample — are independent DiT forwards, so ranks never exchange activations
inside a step.  Two modes:

  * ``sample`` sharding (default): rank r owns samples r, r+P, r+2P, ...; weights replicated;
    ONE all-gather of the final latents after the loop (6 MB/sample at 720p) — the collective
    the north-star names.  Weak scaling, linear by construction.
  * ``cfg`` sharding for a single sample on 2 ranks: rank 0 evaluates the unconditional branch,
    rank 1 the text branch of every step; one all-gather of the two noise predictions
    (2 x 6 MB) per step feeds the CFG combine (pipeline_wan_fun_control.py:820-822) on both.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist

Tensor = torch.Tensor


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_indices(num_samples: int, rank: Optional[int] = None, world_size: Optional[int] = None) -> List[int]:
    """Sample indices owned by `rank`: r, r+P, r+2P, ... (round-robin keeps shards balanced)."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    return list(range(rank, num_samples, world_size))


def gather_latents(local: Sequence[Tensor], num_samples: int, group=None) -> List[Tensor]:
    """All-gather per-sample latents (same shape everywhere) produced under `shard_indices`;
    returns the full list in sample order on every rank."""
    rank, w = world()
    if w == 1:
        return list(local)
    per_rank = (num_samples + w - 1) // w
    ref = local[0] if len(local) else None
    shape = torch.tensor(list(ref.shape) if ref is not None else [0] * 8, dtype=torch.int64)
    # every rank needs the sample shape even if it owns nothing
    shapes = [torch.zeros_like(shape) for _ in range(w)]
    if ref is not None and ref.is_cuda:
        shape = shape.to(ref.device)
        shapes = [s.to(ref.device) for s in shapes]
    dist.all_gather(shapes, shape, group=group)
    full = next(s for s in shapes if int(s.sum()) > 0).tolist()
    like = ref if ref is not None else None
    if like is None:
        raise RuntimeError("gather_latents: a rank without samples needs dtype/device; pass >= world_size samples")
    buf = torch.zeros(per_rank, *full[:like.dim()], dtype=like.dtype, device=like.device)
    for i, t in enumerate(local):
        buf[i].copy_(t)
    out = [torch.empty_like(buf) for _ in range(w)]
    dist.all_gather(out, buf, group=group)
    result: List[Optional[Tensor]] = [None] * num_samples
    for r in range(w):
        for i, idx in enumerate(shard_indices(num_samples, r, w)):
            result[idx] = out[r][i]
    return result  # type: ignore[return-value]


def denoise_sharded(step_fn: Callable[[Tensor, int, object], Tensor], latents: Sequence[Tensor],
                    conds: Sequence[object], steps: Sequence[int], group=None) -> List[Tensor]:
    """Run `step_fn(latent, step, cond)` (e.g. StraGDenoiser.step) for every step on the samples
    this rank owns, then all-gather the final latents.  Bitwise identical to the 1-rank run."""
    mine = shard_indices(len(latents))
    local = []
    for idx in mine:
        lat = latents[idx]
        for s in steps:
            lat = step_fn(lat, s, conds[idx])
        local.append(lat)
    return gather_latents(local, len(latents), group)


def cfg_split_noise(noise_fn: Callable[[int], Tensor], group=None):
    """CFG sharding on 2 ranks: `noise_fn(branch)` returns the noise prediction of branch 0
    (unconditional) or 1 (text); every rank evaluates ONE branch and receives both."""
    rank, w = world()
    if w == 1:
        return noise_fn(0), noise_fn(1)
    if w != 2:
        raise ValueError("cfg sharding pairs exactly two ranks")
    mine = noise_fn(rank)
    both = [torch.empty_like(mine) for _ in range(2)]
    dist.all_gather(both, mine.contiguous(), group=group)
    return both[0], both[1]
