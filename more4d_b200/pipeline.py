"""Host mirror of the 4D-STraG denoising loop — the body of
``WanFunControlPipeline.__call__`` (MoRe4D/pipeline/pipeline_wan_fun_control.py:741-840,
"pctl") with everything around it (T5 / CLIP / VAE / video IO) left to the reference.

One *latent-step* = CFG batch-doubling (pctl:751), conditioning concat control | start-image |
depth (pctl:762-777), one DiT forward on the doubled batch (pctl:796), the CFG combine
(pctl:820-822) and the scheduler step (pctl:825) — the unit BASELINE.json's metric counts.

Scheduler: the reference's default is diffusers' FlowMatchEulerDiscreteScheduler, whose sigma
grid lives in diffusers (absent here).  The only in-tree formula is used instead:
``get_sampling_sigmas`` (MoRe4D/utils/fm_solvers.py:22-26), t = 1000*sigma and the Euler update
x += (sigma_next - sigma) * v in fp32 with a final sigma of 0.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import ops

Tensor = torch.Tensor
BF16 = torch.bfloat16


def get_sampling_sigmas(sampling_steps: int, shift: float) -> np.ndarray:
    sigma = np.linspace(1, 0, sampling_steps + 1)[:sampling_steps]
    return shift * sigma / (1 + (shift - 1) * sigma)


@dataclass
class StraGConditioning:
    """Step-invariant inputs of the loop, for ONE sample (batch 1), as the pipeline holds them."""
    control_latents: Tensor            # [1, 16, T, h, w]  vae.encode(control_video)      pctl:631-643
    depth_latents: Optional[Tensor]    # [1, 16, T, h, w]  vae.encode(depth_image)        pctl:645-657
    ref_latents: Optional[Tensor]      # [1, 16, h, w]     vae.encode(ref_image)[:,:,0]   pctl:703-722
    clip_context: Tensor               # [1, 257, 1280]    CLIP image tokens               pctl:696
    prompt_embeds: Tensor              # [Lp, 4096]
    negative_prompt_embeds: Tensor     # [Ln, 4096]

    def to(self, device) -> "StraGConditioning":
        mv = lambda t: None if t is None else t.to(device=device, dtype=BF16, non_blocking=True)
        return StraGConditioning(mv(self.control_latents), mv(self.depth_latents),
                                 mv(self.ref_latents), mv(self.clip_context),
                                 mv(self.prompt_embeds), mv(self.negative_prompt_embeds))

    def nbytes(self) -> int:
        return sum(t.numel() * 2 for t in (self.control_latents, self.depth_latents, self.ref_latents,
                                           self.clip_context, self.prompt_embeds,
                                           self.negative_prompt_embeds) if t is not None)


class StraGDenoiser:
    """Runs latent-steps of the 4D-STraG loop on the B200 kernels."""

    def __init__(self, transformer, guidance_scale: float = 6.0, shift: float = 5.0,
                 num_inference_steps: int = 50, hoist_conditioning: bool = False):
        """hoist_conditioning: compute the step-invariant context embedding and cross-attention
        K/V once per conditioning object instead of on every step (bit-identical results; the
        reference recomputes them, so the default keeps the reference's work per step)."""
        self.transformer = transformer
        self.hoist_conditioning = hoist_conditioning
        self._hoisted = None                  # (id(cond), Conditioning)
        self.guidance_scale = float(guidance_scale)
        self.num_inference_steps = num_inference_steps
        sig = get_sampling_sigmas(num_inference_steps, shift)
        self.sigmas = np.append(sig, 0.0)
        self.timesteps = sig * 1000.0
        transformer.num_inference_steps = num_inference_steps

    def seq_len(self, latents: Tensor) -> int:
        ps = self.transformer.config.patch_size                         # pctl:737
        _, _, T, h, w = latents.shape
        return int(np.ceil((h * w) / (ps[1] * ps[2]) * T))

    def _model_inputs(self, cond) -> dict:
        """Conditioning tensors of one CFG-doubled forward (pctl:751-794)."""
        start = torch.zeros_like(cond.control_latents)                               # start_image_latentes_conv_in
        parts = [cond.control_latents, start]
        if cond.depth_latents is not None:
            parts.append(cond.depth_latents)
        y1 = torch.cat(parts, dim=1)                                                 # pctl:762-777
        ref = None if cond.ref_latents is None else torch.cat([cond.ref_latents] * 2)
        return dict(y=torch.cat([y1] * 2), full_ref=ref, clip_fea=torch.cat([cond.clip_context] * 2))

    @torch.no_grad()
    def step(self, latents: Tensor, i: int, cond) -> Tensor:
        """Advance `latents` ([1, 16, T, h, w] bf16, CUDA, updated in place) by scheduler step i."""
        tr = self.transformer
        tr.current_steps = i
        x = torch.cat([latents] * 2)                                                 # pctl:751
        kw = self._model_inputs(cond)
        t = torch.full((2,), float(self.timesteps[i]), device=latents.device, dtype=torch.float32)
        context = [cond.negative_prompt_embeds, cond.prompt_embeds]
        pre = None
        if self.hoist_conditioning:
            if self._hoisted is None or self._hoisted[0] is not cond:
                self._hoisted = (cond, tr.precompute_conditioning(context, kw["clip_fea"]))
            pre = self._hoisted[1]
        noise = tr(x=x, context=context, t=t, seq_len=self.seq_len(latents), conditioning=pre, **kw)  # pctl:796
        dt = float(self.sigmas[i + 1] - self.sigmas[i])
        ops.cfg_euler_step_(latents, noise[0:1], noise[1:2], self.guidance_scale, dt)  # pctl:820-825
        return latents

    @torch.no_grad()
    def __call__(self, latents: Tensor, cond: StraGConditioning, steps: Optional[Sequence[int]] = None,
                 device="cuda") -> Tensor:
        """Public entry: host (or device) latents + conditioning in, denoised latents out on the
        host.  `steps` defaults to the whole schedule."""
        lat = latents.to(device=device, dtype=BF16, non_blocking=True).contiguous()
        c = cond.to(device)
        for i in (range(self.num_inference_steps) if steps is None else steps):
            self.step(lat, i, c)
        return lat.to("cpu")


@dataclass
class ViSMConditioning:
    """Step-invariant inputs of the 4D-ViSM (Wan-InP) loop for one sample
    (pipeline_wan_fun_inpaint.py:612-690)."""
    mask_latents: Tensor               # [1, 4, T, h, w]   resize_mask(1 - mask_condition)      pinp:642-646
    masked_video_latents: Tensor       # [1, 16, T, h, w]  vae.encode(masked video)             pinp:653-664
    clip_context: Tensor               # [1, 257, 1280]
    prompt_embeds: Tensor              # [Lp, 4096]
    negative_prompt_embeds: Tensor     # [Ln, 4096]

    def to(self, device) -> "ViSMConditioning":
        mv = lambda t: t.to(device=device, dtype=BF16, non_blocking=True)
        return ViSMConditioning(mv(self.mask_latents), mv(self.masked_video_latents), mv(self.clip_context),
                                mv(self.prompt_embeds), mv(self.negative_prompt_embeds))

    def nbytes(self) -> int:
        return sum(t.numel() * 2 for t in (self.mask_latents, self.masked_video_latents, self.clip_context,
                                           self.prompt_embeds, self.negative_prompt_embeds))


class ViSMDenoiser(StraGDenoiser):
    """The 4D-ViSM denoise loop (pipeline_wan_fun_inpaint.py:693-743) on the same kernels: the
    DiT call differs from 4D-STraG only in its conditioning — y = cat(mask latents [4], masked-
    video latents [16]) (pinp:707-713), no reference frame — and runs WanTransformer3DModel."""

    def _model_inputs(self, cond: ViSMConditioning) -> dict:
        y1 = torch.cat([cond.mask_latents, cond.masked_video_latents], dim=1)
        return dict(y=torch.cat([y1] * 2), full_ref=None, clip_fea=torch.cat([cond.clip_context] * 2))


def synthetic_visim_conditioning(latent_shape, seed: int = 0, device="cpu", prompt_tokens: int = 32,
                                 negative_tokens: int = 1, text_dim: int = 4096, clip_dim: int = 1280):
    """Synthetic latents + ViSM conditioning (denoise part)."""
    g = torch.Generator().manual_seed(seed)
    _, c, T, h, w = latent_shape
    rn = lambda *shape: torch.randn(*shape, generator=g).to(BF16).to(device)
    mask = (torch.rand(1, 4, T, h, w, generator=g) < 0.5).to(BF16).to(device)
    return rn(1, c, T, h, w), ViSMConditioning(mask, rn(1, 16, T, h, w), rn(1, 257, clip_dim),
                                               rn(prompt_tokens, text_dim), rn(negative_tokens, text_dim))


def synthetic_conditioning(latent_shape, seed: int = 0, device="cpu", pin: bool = False,
                           prompt_tokens: int = 32, negative_tokens: int = 1, text_dim: int = 4096,
                           clip_dim: int = 1280):
    """Synthetic latents + conditioning of the BASELINE shapes."""
    g = torch.Generator().manual_seed(seed)
    _, c, T, h, w = latent_shape

    def rn(*shape):
        t = torch.randn(*shape, generator=g).to(BF16)
        if pin and device == "cpu":
            t = t.pin_memory()
        return t.to(device)

    latents = rn(1, c, T, h, w)
    cond = StraGConditioning(rn(1, 16, T, h, w), rn(1, 16, T, h, w), rn(1, 16, h, w),
                             rn(1, 257, clip_dim), rn(prompt_tokens, text_dim), rn(negative_tokens, text_dim))
    return latents, cond
