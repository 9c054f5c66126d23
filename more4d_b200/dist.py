"""Multi-GPU execution of the denoising path: one process per GPU, `torch.distributed` for the
plumbing (NCCL on the NVSwitch domain; gloo in the CPU tests).

The 4D-STraG denoiser shards on the BATCH axis: samples — and the two CFG
branches of one sample — are independent DiT forwards, so ranks never exchange activations
inside a step.  Two modes:

  * ``sample`` sharding (default): rank r owns samples r, r+P, r+2P, ...; weights replicated;
    ONE all-gather of the final latents after the loop (6 MB/sample at 720p) — the collective
    the north-star names.  Weak scaling, linear by construction.
  * ``sequence`` (Ulysses) sharding for ONE sample's latency (`SequenceParallel`, the slot of the
    reference's `enable_multi_gpus_inference` / `usp_attn_forward`, wan_transformer4d.py:1038-1044,
    1187-1198,1320-1321): every rank holds L/P tokens through all token-local ops (LayerNorm,
    AdaLN, Linear, cross-attention, FFN); around self-attention one all-to-all trades the token
    shard of all heads for all tokens of heads/P heads, and one trades it back.
  * ``cfg`` sharding for a single sample on 2 ranks: rank 0 evaluates the unconditional branch,
    rank 1 the text branch of every step; one all-gather of the two noise predictions
    (2 x 6 MB) per step feeds the CFG combine (pipeline_wan_fun_control.py:820-822) on both.
"""
from __future__ import annotations

import os
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
    returns the full list in sample order on every rank.  Every rank must own at least one sample
    (num_samples >= world size) — checked on every rank BEFORE any collective, so a mis-sized call
    raises everywhere instead of hanging in NCCL."""
    rank, w = world()
    if w == 1:
        return list(local)
    if num_samples < w:
        raise RuntimeError(f"gather_latents: {num_samples} samples cannot be sharded over {w} ranks "
                           "(every rank needs at least one sample)")
    if len(local) != len(shard_indices(num_samples, rank, w)):
        raise RuntimeError("gather_latents: `local` does not match this rank's shard")
    per_rank = (num_samples + w - 1) // w
    like = local[0]
    buf = torch.zeros(per_rank, *like.shape, dtype=like.dtype, device=like.device)
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


class SequenceParallel:
    """Ulysses sequence parallelism over `group` (NCCL on the NVSwitch domain; gloo in tests).

    Layout contract: token shards are contiguous chunks of the (padded) sequence — rank r holds
    tokens [r*L/P, (r+1)*L/P), like `torch.chunk(x, P, dim=1)[r]` at wan_transformer4d.py:1188 —
    and head shards are contiguous groups of heads/P heads."""

    def __init__(self, group=None, peer_memory: bool = True):
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("SequenceParallel needs an initialised torch.distributed process group")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        # fused exchange over symmetric (peer-mapped) memory; False = NCCL all-to-alls + permutes
        self.peer_memory = peer_memory
        self._peer = {}

    def peer_buffers(self, B: int, L: int, C: int, device) -> "PeerBuffers":
        key = (B, L, C, str(device))
        if key not in self._peer:
            self._peer[key] = PeerBuffers(self, B, L, C, device)       # collective: same order on all ranks
        return self._peer[key]

    def shard_tokens(self, x: Tensor) -> Tensor:
        """[B, L, ...] -> this rank's [B, L/P, ...] (L % P == 0)."""
        L = x.shape[1]
        if L % self.world:
            raise ValueError(f"sequence length {L} is not a multiple of the SP world size {self.world}")
        n = L // self.world
        return x[:, self.rank * n:(self.rank + 1) * n].contiguous()

    def seq_to_heads(self, x: Tensor) -> Tensor:
        """[S, B, L/P, H, D] (token shard, all heads; S stacked tensors, e.g. q/k/v) ->
        [S, B, L, H/P, D] (all tokens, this rank's heads).  One all-to-all."""
        S, B, n, H, D = x.shape
        P = self.world
        if H % P:
            raise ValueError(f"{H} heads cannot be split over {P} ranks")
        send = x.view(S, B, n, P, H // P, D).permute(3, 0, 1, 2, 4, 5).contiguous()    # [P(dst), S, B, n, H/P, D]
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, group=self.group)                           # [P(src), ...]
        return recv.permute(1, 2, 0, 3, 4, 5).reshape(S, B, P * n, H // P, D)

    def heads_to_seq(self, x: Tensor) -> Tensor:
        """[B, L, H/P, D] (all tokens, this rank's heads) -> [B, L/P, H, D].  One all-to-all."""
        B, L, h, D = x.shape
        P = self.world
        n = L // P
        send = x.view(B, P, n, h, D).permute(1, 0, 2, 3, 4).contiguous()               # [P(dst), B, n, h, D]
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, group=self.group)                           # [P(src = head group), ...]
        return recv.permute(1, 2, 0, 3, 4).reshape(B, n, P * h, D)

    def gather_tokens(self, x: Tensor) -> Tensor:
        """[B, L/P, ...] -> [B, L, ...] on every rank (wan_transformer4d.py:1320-1321)."""
        parts = [torch.empty_like(x) for _ in range(self.world)]
        dist.all_gather(parts, x.contiguous(), group=self.group)
        return torch.cat(parts, dim=1)


class PeerBuffers:
    """Receive buffers of the FUSED sequence-parallel exchange, allocated in torch symmetric memory so
    that every rank holds P2P (NVLink) mappings of all peers' buffers: producers — the v-projection
    GEMM epilogue, the q/k RMSNorm kernel, the attention epilogue — store straight into the
    consumer's buffer in its final layout; what remains of the all-to-all is a barrier.

        qkv [3, B, L, C/P]   all tokens, this rank's heads   (written by every rank's projections)
        o   [B, L/P, C]      this rank's tokens, all heads   (written by every rank's attention)
    """

    def __init__(self, sp: "SequenceParallel", B: int, L: int, C: int, device):
        import torch.distributed._symmetric_memory as symm_mem
        P = sp.world
        self.shape = (B, L, C)
        group = sp.group if sp.group is not None else dist.group.WORLD
        self.qkv = symm_mem.empty(3, B, L, C // P, dtype=torch.bfloat16, device=device)
        self.o = symm_mem.empty(B, L // P, C, dtype=torch.bfloat16, device=device)
        self._h_qkv = symm_mem.rendezvous(self.qkv, group)
        self._h_o = symm_mem.rendezvous(self.o, group)
        self.qkv_peers = [self._h_qkv.get_buffer(g, (3, B, L, C // P), torch.bfloat16) for g in range(P)]
        self.o_peers = [self._h_o.get_buffer(g, (B, L // P, C), torch.bfloat16) for g in range(P)]

    def barrier(self) -> None:
        """Stream-ordered barrier over the group's signal pads (a ~7 us kernel)."""
        self._h_qkv.barrier()

