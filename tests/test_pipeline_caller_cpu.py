"""Caller smoke: the REAL `WanFunControlPipeline.__call__`
(MoRe4D/pipeline/pipeline_wan_fun_control.py:477-858) imported in place and driven with
`prompt_embeds` given and `output_type="latent"`, once against the plain reference transformer and
once against the same transformer after `more4d_b200.install()`.

There is no GPU here, so the installed mirror's innermost `_forward` (the part that launches
kernels) is replaced by a RECORDER that checks the call surface it receives and answers with the CPU
oracle; everything the caller touches — the patched `forward`, state sync (current_steps,
num_inference_steps, teacache, cfg_skip_ratio, freqs), the `@cfg_skip` batch halving, TeaCache
decisions on the REFERENCE's own TeaCache object, the grad-enabled fall-through — runs for real.
Authoring container only (needs /root/reference)."""
import contextlib

import pytest
import torch

from oracle import ref_import

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference tree not present")

BF16 = torch.bfloat16


class _TextEncoder(torch.nn.Module):
    dtype = BF16


class _Clip(torch.nn.Module):
    def forward(self, images):
        g = torch.Generator().manual_seed(7)
        return torch.randn(1, 257, 1280, generator=g).to(BF16)


class _Vae(torch.nn.Module):
    """Only the attributes the loop reads (pctl:185,328-330,595,736); control_video=None, so no
    encode happens in this smoke."""
    latent_channels, temporal_compression_ratio, spatial_compression_ratio = 16, 4, 8
    dtype = BF16

    class config:
        latent_channels = 16


class _Cfg(dict):
    __getattr__ = dict.get


class _Embeds(list):
    """`prompt_embeds` as the pipeline's own encode_prompt returns them — a LIST of [L_i, 4096]
    tensors, concatenated with `+` for CFG (pctl:233,571) — carrying the `.shape` that check_inputs
    and the batch-size logic read when embeddings are passed in directly (pctl:449-455,545)."""
    shape = (1,)


def _make(t4d, cfg, dtype, seed=0):
    from more4d_b200 import synth
    m = t4d.WanTransformer4DModel(model_type="i2v", in_dim=cfg.in_dim, dim=cfg.dim, ffn_dim=cfg.ffn_dim,
                                  num_heads=cfg.num_heads, num_layers=cfg.num_layers, text_dim=cfg.text_dim,
                                  text_len=cfg.text_len, add_ref_conv=True, use_dino_guidance=False,
                                  use_omnimae_guidance=False).to(dtype)
    m.load_state_dict({k: v.to(dtype) for k, v in synth.dit_state_dict(cfg, seed).items()}, strict=True)
    m.eval()
    # diffusers' register_to_config would create this (pctl:703,737 read it)
    m.config = _Cfg(add_ref_conv=True, patch_size=(1, 2, 2))
    return m


@contextlib.contextmanager
def _cpu_pipeline_env(pipe_cls):
    """pctl:472-474 hard-codes cuda:0 and wraps the DiT call in torch.cuda.device(...) (pctl:794)."""
    old_dev, old_ctx = pipe_cls._execution_device, torch.cuda.device
    pipe_cls._execution_device = property(lambda self: torch.device("cpu"))
    torch.cuda.device = lambda device=None: contextlib.nullcontext()
    try:
        yield
    finally:
        pipe_cls._execution_device, torch.cuda.device = old_dev, old_ctx


def _run_pipeline(pctl, Sched, transformer, steps=3, **kw):
    pipe = pctl.WanFunControlPipeline(tokenizer=None, text_encoder=_TextEncoder(), vae=_Vae(),
                                      transformer=transformer, clip_image_encoder=_Clip(), scheduler=Sched(5.0))
    g = torch.Generator().manual_seed(3)
    prompt = _Embeds([torch.randn(9, 128, generator=g).to(BF16)])
    negative = _Embeds([torch.randn(4, 128, generator=g).to(BF16)])
    lat = torch.randn(1, 16, 2, 8, 12, generator=g).to(BF16)
    with _cpu_pipeline_env(pctl.WanFunControlPipeline), contextlib.redirect_stdout(None):
        out = pipe(prompt=None, prompt_embeds=prompt, negative_prompt_embeds=negative, height=64, width=96,
                   num_frames=5, num_inference_steps=steps, guidance_scale=6.0, latents=lat.clone(),
                   output_type="latent", return_dict=True, **kw)
    return out.videos


def _autocast_like(ref_fp32):
    """CPU stand-in for the CUDA autocast the pipeline relies on (pctl:794): bf16 tensors in, the
    reference's fp32 arithmetic on bf16-valued weights, bf16 out."""
    orig = ref_fp32.forward

    def fwd(x, t, context, seq_len, clip_fea=None, y=None, y_camera=None, full_ref=None, **kw):
        up = lambda v: None if v is None else v.float()
        return orig(x=up(x), t=t, context=[c.float() for c in context], seq_len=seq_len, clip_fea=up(clip_fea),
                    y=up(y), y_camera=y_camera, full_ref=up(full_ref), **kw).to(BF16)
    ref_fp32.forward = fwd
    return ref_fp32


def _install_with_recorder(ref_bf16, cfg, calls):
    from more4d_b200 import install, synth
    from oracle import dit_oracle as O
    install.install(transformer=ref_bf16)
    m = ref_bf16._m4d
    sd = synth.dit_state_dict(cfg, 0)

    def recorder(x, t, context, seq_len, clip_fea, y, full_ref, guidance_features, cond_flag=True,
                 conditioning=None, first_frame=None):
        calls.append(dict(x=tuple(x.shape), dtype=x.dtype, y=tuple(y.shape), t=t.clone(), seq_len=seq_len,
                          n_ctx=len(context), full_ref=tuple(full_ref.shape), clip=tuple(clip_fea.shape),
                          step=m.current_steps, n_steps=m.num_inference_steps))
        from more4d_b200.cache_utils import teacache_decide, teacache_step_done
        B = x.shape[0]
        tc = m.teacache
        if tc is not None:                       # same hook order as dit._forward / t4d:1200-1270,1336-1339
            _, e0 = O.time_embed(t.float(), sd, cfg.freq_dim, cfg.dim)
            calls[-1]["should_calc"] = teacache_decide(tc, e0, cond_flag)
        out = O.dit_forward(sd, cfg, x.float(), t.float(), [c.float() for c in context], seq_len,
                            clip_fea=clip_fea.float(), y=y.float(), full_ref=full_ref.float())
        if tc is not None:
            teacache_step_done(tc, cond_flag)
        assert out.shape[0] == B
        return out.to(BF16)
    m._forward = recorder
    return ref_bf16


def test_pipeline_call_through_installed_transformer_matches_reference_run():
    from more4d_b200.config import WAN_TINY
    cfg = WAN_TINY.with_(in_dim=48)      # control_video=None: y = control(16) | start image(16), pctl:762-777
    pctl, Sched = ref_import.load_pipeline()
    t4d, _, _ = ref_import.load()
    want = _run_pipeline(pctl, Sched, _autocast_like(_make(t4d, cfg, torch.float32)))
    calls = []
    got = _run_pipeline(pctl, Sched, _install_with_recorder(_make(t4d, cfg, BF16), cfg, calls))
    assert got.shape == want.shape == (1, 16, 2, 8, 12) and got.dtype == BF16
    assert len(calls) == 3
    for i, c in enumerate(calls):                                  # the call surface of pctl:796-806
        assert c["x"] == (2, 16, 2, 8, 12) and c["y"] == (2, 32, 2, 8, 12) and c["dtype"] == BF16
        assert c["full_ref"] == (2, 16, 8, 12) and c["clip"] == (2, 257, 1280) and c["n_ctx"] == 2
        assert c["seq_len"] == 2 * 4 * 6 and c["step"] == i and c["n_steps"] == 3
        assert c["t"].shape == (2,)
    err = float((got.float() - want.float()).norm() / want.float().norm())
    assert err < 2e-3, err                                         # bf16 output rounding flips only


def test_pipeline_cfg_skip_and_reference_teacache_object_after_install():
    """cfg_skip state set on the REFERENCE object (t4d:986-1008) and a TeaCache created by the
    reference's own enable_teacache (t4d:961-970, MoRe4D/models/cache_utils.py) must drive the
    installed path (the reference class has no decide()/step_done())."""
    from more4d_b200.config import WAN_TINY
    cfg = WAN_TINY.with_(in_dim=48)      # control_video=None: y = control(16) | start image(16), pctl:762-777
    pctl, Sched = ref_import.load_pipeline()
    t4d, _, _ = ref_import.load()
    ref = _make(t4d, cfg, BF16)
    ref.enable_cfg_skip(0.34, 3)                                   # last step drops the uncond half
    ref.enable_teacache([0.0, 0.0, 0.0, 1.0, 0.0], 3, rel_l1_thresh=0.0, num_skip_start_steps=0, offload=False)
    assert type(ref.teacache).__module__.startswith("MoRe4D.")     # the reference's class, not ours
    calls = []
    out = _run_pipeline(pctl, Sched, _install_with_recorder(ref, cfg, calls))
    assert out.shape == (1, 16, 2, 8, 12) and bool(torch.isfinite(out.float()).all())
    assert [c["x"][0] for c in calls] == [2, 2, 1]                 # cfg_skip halved the last call's batch
    assert all(c["should_calc"] for c in calls)                    # threshold 0: never skip
    assert ref.teacache.cnt == 0                                   # 3 steps of 3: the reference object was reset
    assert ref._m4d.teacache is ref.teacache


def test_grad_enabled_call_falls_through_to_the_reference_forward():
    """under torch.is_grad_enabled() the shim runs the reference's own modules
    (train_wan.py:1938-1950 keeps working after install())."""
    from more4d_b200 import install, synth
    from more4d_b200.config import WAN_TINY as cfg
    t4d, _, _ = ref_import.load()
    ref = _make(t4d, cfg, torch.float32)
    inp = synth.dit_inputs(cfg, (3, 4, 6), 2, 0)
    kw = dict(x=inp["x"].float(), t=inp["t"], context=[c.float() for c in inp["context"]], seq_len=inp["seq_len"],
              clip_fea=inp["clip_fea"].float(), y=inp["y"].float(), full_ref=inp["full_ref"].float())
    with torch.no_grad():
        want = ref(**kw)
    install.install(transformer=ref)
    ref._m4d._forward = None                                       # the kernel path must not be entered
    for p in ref.parameters():
        p.requires_grad_(True)
    with torch.enable_grad():
        got = ref(first_frame=None, **kw)
        got.float().pow(2).mean().backward()                       # the training step's backward works
    assert torch.equal(got.detach(), want)
    assert ref.blocks[0].ffn[0].weight.grad is not None
    # inference on non-bf16 weights does NOT silently take another path
    with torch.no_grad(), pytest.raises(RuntimeError, match="must be bf16"):
        ref(**kw)


def test_vae_and_adaptor_fall_through_under_grad():
    import io
    from more4d_b200 import install, synth
    _, vae_mod, traj = ref_import.load()
    with contextlib.redirect_stdout(io.StringIO()):
        ea = traj.VAEEncoderadaptor()
    ea.load_state_dict({k: v.float() for k, v in synth.adaptor_state_dict("encoder", 0).items()}, strict=True)
    x = synth.trajectory_video(2, 16, 16, 0).float()
    with torch.no_grad():
        want = ea(x)
    install.install(encoder_prompt=ea)
    with torch.enable_grad():
        got = ea(x)
        got.mean().backward()
    assert torch.equal(got.detach(), want) and ea.conv_in.weight.grad is not None
