"""Drop-in installation into live reference objects (SURVEY.md §8b seams 5-7).

``install(transformer=..., vae=..., encoder_prompt=..., decoder_prompt=...)`` takes the
reference's own module instances (as built by scripts/inference/infer.py:537-629), builds the
B200 host mirrors around THE SAME parameter tensors (no copies: a later ``load_state_dict`` /
LoRA merge on the reference module is seen by the kernels) and re-points the reference
objects' entry methods at the mirrors — the same per-instance patching the reference itself
uses at wan_transformer4d.py:1042-1044 and scripts/inference/infer.py:608-610.  After this
``WanFunControlPipeline.__call__`` and ``scripts/inference/infer_vae.py`` run unchanged.
"""
from __future__ import annotations

import types

import torch
import torch.nn as nn


def adopt_parameters(ours: nn.Module, theirs: nn.Module) -> int:
    """Make every parameter of `ours` BE the same-named parameter of `theirs` (shared storage)."""
    src = dict(theirs.named_parameters())
    n = 0
    for name, _ in list(ours.named_parameters()):
        if name not in src:
            raise KeyError(f"reference module has no parameter `{name}`")
        p = src[name]
        if tuple(p.shape) != tuple(dict(ours.named_parameters())[name].shape):
            raise ValueError(f"shape mismatch for `{name}`")
        mod = ours
        *path, leaf = name.split(".")
        for part in path:
            mod = mod._modules[part]
        mod._parameters[leaf] = p
        n += 1
    return n


def _check(p: torch.Tensor, what: str) -> None:
    if p.dtype != torch.bfloat16:
        raise RuntimeError(f"more4d_b200.install: {what} must be bf16 (got {p.dtype}); the reference runs "
                           "this path with weight_dtype = torch.bfloat16")


def install_transformer(ref):
    """`ref`: a reference WanTransformer4DModel (or 3D model without MPM).  Returns `ref`."""
    from .dit import WanTransformer4DModel
    w = ref.patch_embedding.weight
    _check(w, "transformer")
    blk = ref.blocks[0]
    ours = WanTransformer4DModel(
        model_type=ref.model_type, patch_size=tuple(ref.patch_size), text_len=ref.text_len,
        in_dim=w.shape[1], dim=ref.dim, ffn_dim=ref.ffn_dim, freq_dim=ref.freq_dim, text_dim=ref.text_dim,
        out_dim=ref.out_dim, num_heads=ref.num_heads, num_layers=ref.num_layers, qk_norm=ref.qk_norm,
        cross_attn_norm=ref.cross_attn_norm, eps=ref.eps, add_ref_conv=getattr(ref, "ref_conv", None) is not None,
        in_dim_ref_conv=(ref.ref_conv.weight.shape[1] if getattr(ref, "ref_conv", None) is not None else 16),
        use_spatial_guidance=getattr(blk, "spatial_guidance_self", None) is not None,
        use_cls_token=getattr(ref, "use_cls_token", False), device="meta")
    adopt_parameters(ours, ref)
    ours.freqs = ref.freqs
    ref._m4d = ours

    def forward(self, x, t, context, seq_len, clip_fea=None, y=None, y_camera=None, full_ref=None,
                subject_ref=None, cond_flag=True, first_frame=None):
        m = self._m4d
        # state the pipeline toggles on the reference object (pctl:743-746, t4d:961-1008)
        m.teacache, m.cfg_skip_ratio = self.teacache, self.cfg_skip_ratio
        m.current_steps, m.num_inference_steps = self.current_steps, self.num_inference_steps
        m.freqs = self.freqs
        return m.forward(x, t, context, seq_len, clip_fea=clip_fea, y=y, y_camera=y_camera,
                         full_ref=full_ref, subject_ref=subject_ref, cond_flag=cond_flag,
                         first_frame=first_frame)

    ref.forward = types.MethodType(forward, ref)
    return ref


def install_vae(ref):
    """`ref`: a reference AutoencoderKLWan.  Patches encode / decode (and _encode / _decode)."""
    from .vae import AutoencoderKLWan
    _check(ref.model.conv1.weight, "vae")
    ours = AutoencoderKLWan(device="meta")
    adopt_parameters(ours, ref)
    ref._m4d = ours
    ref.encode = lambda x, return_dict=True: ours.encode(x, return_dict)
    ref.decode = lambda z, return_dict=True: ours.decode(z, return_dict)
    ref._encode = ours._encode
    ref._decode = lambda zs: ours.decode(zs)
    # the checkpointed "memory saver" twins (vae:783-821,842-847) are the same forward
    ref.encode_memory_saver, ref.decode_memory_saver = ref.encode, ref.decode
    ref._encode_memory_saver, ref._decode_memory_saver = ref._encode, ref._decode
    return ref


def install_adaptor(ref):
    """`ref`: a reference VAEEncoderadaptor / VAEDecoderadaptor."""
    from .vae import VAEDecoderadaptor, VAEEncoderadaptor
    _check(ref.conv_in.weight, "adaptor")
    cls = VAEEncoderadaptor if hasattr(ref, "down") else VAEDecoderadaptor
    ours = cls(device="meta")
    adopt_parameters(ours, ref)
    ref._m4d = ours
    ref.forward = lambda x: ours.forward(x)
    return ref


def install(transformer=None, vae=None, encoder_prompt=None, decoder_prompt=None):
    if transformer is not None:
        install_transformer(transformer)
    if vae is not None:
        install_vae(vae)
    for a in (encoder_prompt, decoder_prompt):
        if a is not None:
            install_adaptor(a)
