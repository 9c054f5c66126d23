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


def _kernel_path_ok(p: torch.Tensor, what: str = "module") -> bool:
    """True -> run the B200 kernels; False -> run the reference's own method.

    The ONLY fall-through is `torch.is_grad_enabled()`: training / backward is out of scope
    (SURVEY.md §8b "Threading / streams": "under torch.is_grad_enabled() the shim must fall through
    to the reference modules"), mirroring the reference's own rule for its optional fast paths
    (wan_transformer4d.py:192-193, MoRe4D/models/__init__.py:51-55).  Inference calls never fall
    back: weights that are not bf16 raise here, tensors that are not on an sm_100 device raise in
    more4d_b200.ops (no CPU path exists)."""
    if torch.is_grad_enabled():
        return False
    if p.dtype != torch.bfloat16:
        raise RuntimeError(f"more4d_b200.install: {what} must be bf16 (got {p.dtype}); the reference runs "
                           "this path with weight_dtype = torch.bfloat16 and there is no other kernel path")
    return True


def install_transformer(ref):
    """`ref`: a reference WanTransformer4DModel (or 3D model without MPM).  Returns `ref`.

    The original bound `forward` (the reference's own, `@cfg_skip`-decorated) is kept as
    `ref._m4d_reference_forward` and is what runs whenever the kernel path does not apply:
    grad enabled (scripts/4D_STraG_training/train_wan.py:1938-1950 keeps working after install()),
    or the `y_camera` / `subject_ref` / DINO inputs the 4D-STraG inference path never passes."""
    from .dit import WanTransformer4DModel
    w = ref.patch_embedding.weight
    blk = ref.blocks[0]
    omni = bool(getattr(ref, "use_omnimae_guidance", False)) and getattr(ref, "feature_adapter", None) is not None
    ours = WanTransformer4DModel(
        model_type=ref.model_type, patch_size=tuple(ref.patch_size), text_len=ref.text_len,
        in_dim=w.shape[1], dim=ref.dim, ffn_dim=ref.ffn_dim, freq_dim=ref.freq_dim, text_dim=ref.text_dim,
        out_dim=ref.out_dim, num_heads=ref.num_heads, num_layers=ref.num_layers, qk_norm=ref.qk_norm,
        cross_attn_norm=ref.cross_attn_norm, eps=ref.eps, add_ref_conv=getattr(ref, "ref_conv", None) is not None,
        in_dim_ref_conv=(ref.ref_conv.weight.shape[1] if getattr(ref, "ref_conv", None) is not None else 16),
        use_omnimae_guidance=omni,
        use_spatial_guidance=getattr(blk, "spatial_guidance_self", None) is not None,
        use_cls_token=getattr(ref, "use_cls_token", False), device="meta")
    adopt_parameters(ours, ref)
    ours.freqs = ref.freqs
    ours.omnimae_extractor = getattr(ref, "omnimae_extractor", None)   # frozen trunk: the reference's own module
    ref._m4d = ours
    if not hasattr(ref, "_m4d_reference_forward"):
        ref._m4d_reference_forward = ref.forward

    def forward(self, x, t, context, seq_len, clip_fea=None, y=None, y_camera=None, full_ref=None,
                subject_ref=None, cond_flag=True, first_frame=None):
        m = self._m4d
        if (not _kernel_path_ok(self.patch_embedding.weight, "transformer") or y_camera is not None or subject_ref is not None
                or (first_frame is not None and not m.use_omnimae_guidance)):
            return self._m4d_reference_forward(x, t, context, seq_len, clip_fea=clip_fea, y=y, y_camera=y_camera,
                                               full_ref=full_ref, subject_ref=subject_ref, cond_flag=cond_flag,
                                               first_frame=first_frame)
        # state the pipeline toggles on the reference object (pctl:743-746, t4d:961-1008); the
        # TeaCache object stays the reference's own (cache_utils.teacache_decide works on its fields)
        m.teacache, m.cfg_skip_ratio = self.teacache, self.cfg_skip_ratio
        m.current_steps, m.num_inference_steps = self.current_steps, self.num_inference_steps
        m.freqs = self.freqs
        return m.forward(x, t, context, seq_len, clip_fea=clip_fea, y=y, y_camera=y_camera,
                         full_ref=full_ref, subject_ref=subject_ref, cond_flag=cond_flag,
                         first_frame=first_frame)

    ref.forward = types.MethodType(forward, ref)
    return ref


def _fallthrough(ours_fn, ref_fn, probe: torch.Tensor):
    def call(*args, **kwargs):
        if _kernel_path_ok(probe):
            return ours_fn(*args, **kwargs)
        return ref_fn(*args, **kwargs)
    return call


def install_vae(ref):
    """`ref`: a reference AutoencoderKLWan.  Patches encode / decode (and _encode / _decode); with
    grad enabled (scripts/4D_STraG_training/train_vae.py trains the decoder) the reference's own
    methods run."""
    from .vae import AutoencoderKLWan
    ours = AutoencoderKLWan(device="meta")
    adopt_parameters(ours, ref)
    ref._m4d = ours
    probe = ref.model.conv1.weight
    orig = {n: getattr(ref, n) for n in ("encode", "decode", "_encode", "_decode") if hasattr(ref, n)}
    ref._m4d_reference = orig
    ref.encode = _fallthrough(lambda x, return_dict=True: ours.encode(x, return_dict), orig["encode"], probe)
    ref.decode = _fallthrough(lambda z, return_dict=True: ours.decode(z, return_dict), orig["decode"], probe)
    ref._encode = _fallthrough(ours._encode, orig["_encode"], probe)
    ref._decode = _fallthrough(lambda zs: ours.decode(zs), orig["_decode"], probe)
    # the checkpointed "memory saver" twins (vae:783-821,842-847) are the same forward
    for n in ("encode_memory_saver", "decode_memory_saver", "_encode_memory_saver", "_decode_memory_saver"):
        if hasattr(ref, n):
            setattr(ref, n, _fallthrough(getattr(ref, n.replace("_memory_saver", "")), getattr(ref, n), probe))
    return ref


def install_adaptor(ref):
    """`ref`: a reference VAEEncoderadaptor / VAEDecoderadaptor (trained with grad by train_vae.py:
    then the reference forward runs)."""
    from .vae import VAEDecoderadaptor, VAEEncoderadaptor
    cls = VAEEncoderadaptor if hasattr(ref, "down") else VAEDecoderadaptor
    ours = cls(device="meta")
    adopt_parameters(ours, ref)
    ref._m4d = ours
    ref._m4d_reference_forward = ref.forward
    ref.forward = _fallthrough(lambda x: ours.forward(x), ref._m4d_reference_forward, ref.conv_in.weight)
    return ref


def install(transformer=None, vae=None, encoder_prompt=None, decoder_prompt=None):
    if transformer is not None:
        install_transformer(transformer)
    if vae is not None:
        install_vae(vae)
    for a in (encoder_prompt, decoder_prompt):
        if a is not None:
            install_adaptor(a)
