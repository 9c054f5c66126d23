"""Generate the committed golden vectors by running the REAL reference modules.

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports the reference in place through oracle/ref_import.py (stubs for the missing
MoRe4D.dist and diffusers — SURVEY.md Appendix A), loads the synthetic state dicts of
more4d_b200.synth (bf16-valued, upcast to fp32 = oracle mode A), runs the reference's own
forward on CPU through its SDPA attention branch, and stores the outputs as safetensors.
Weights and inputs are regenerated from seeds at test time; each file also stores input /
weight checksums so RNG drift is detected rather than silently mis-compared.
"""
from __future__ import annotations

import os
import sys
import warnings

import torch
from safetensors.torch import save_file

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from more4d_b200 import synth                      # noqa: E402
from more4d_b200.config import DiTConfig, WAN_1_3B, WAN_TINY   # noqa: E402
from oracle import ref_import                      # noqa: E402
from oracle.dit_oracle import time_embed           # noqa: E402  (only to build e0 from the seeded MLP)

OUT = os.path.dirname(os.path.abspath(__file__))
torch.set_grad_enabled(False)


def checksum(t: torch.Tensor) -> torch.Tensor:
    t = t.double()
    return torch.stack([t.sum(), t.abs().sum(), (t * t).sum()]).float()


def f32(sd):
    return {k: v.float() for k, v in sd.items()}


def block_case(t4d, name, cfg: DiTConfig, grid, seq_len, seed, guidance=False):
    C = cfg.dim
    sd = synth.block_state_dict(cfg, 0, seed)
    blk = t4d.WanAttentionBlock("i2v_cross_attn", C, cfg.ffn_dim, cfg.num_heads, (-1, -1), True,
                                True, cfg.eps, use_spatial_guidance=cfg.use_spatial_guidance)
    missing = blk.load_state_dict(f32(sd), strict=True)
    x = synth._randn(seed, "blk.x", (1, seq_len, C), 1.0, "cpu", torch.bfloat16)
    ctx = synth._randn(seed, "blk.ctx", (1, 257 + cfg.text_len, C), 1.0, "cpu", torch.bfloat16)
    tsd = synth.dit_state_dict(cfg, seed, prefix_filter="time_")
    _, e0 = time_embed(torch.tensor([500.0]), tsd, cfg.freq_dim, C)
    d = cfg.head_dim
    freqs = torch.cat([t4d.rope_params(1024, d - 4 * (d // 6)), t4d.rope_params(1024, 2 * (d // 6)),
                       t4d.rope_params(1024, 2 * (d // 6))], dim=1)
    n_tok = grid[0] * grid[1] * grid[2]
    feats = None
    if guidance:
        feats = (synth._randn(seed, "blk.dino", (1, n_tok, cfg.guidance_dim), 1.0, "cpu", torch.float32),
                 synth._randn(seed, "blk.cls", (1, 1, cfg.guidance_dim), 1.0, "cpu", torch.float32))
    y = blk(x.float(), e0, torch.tensor([n_tok]), torch.tensor([list(grid)]), freqs, ctx.float(),
            None, dtype=torch.float32, t=torch.tensor([500.0]), dino_features=feats)
    out = {"y": y.contiguous(), "x_sum": checksum(x), "ctx_sum": checksum(ctx), "e0": e0.contiguous(),
           "w_sum": checksum(torch.cat([v.flatten().float() for v in sd.values()]))}
    save_file(out, os.path.join(OUT, name + ".safetensors"))
    print(name, tuple(y.shape), float(y.abs().mean()))


def block14b_case(t4d, name="block_14b", seed=6):
    """ONE block at the headline dims (Wan2.1-14B: C 5120, F 13824, 40 heads) and L = 1152 tokens
    (grid 2x24x24: 9 key tiles, a 2-CTA GEMM M tail, K = 13 824 accumulation in ffn.2) through the
    real WanAttentionBlock (t4d:585-688).  The full output is 23 MB in fp32, so 96 sampled rows are
    stored, plus the increment y - x of the same rows (the residual pass-through masks errors in the
    output-level number)."""
    from more4d_b200.config import WAN_14B
    cfg = WAN_14B.with_(num_layers=1)
    C, grid, L = cfg.dim, (2, 24, 24), 1152
    sd = synth.block_state_dict(cfg, 0, seed)
    blk = t4d.WanAttentionBlock("i2v_cross_attn", C, cfg.ffn_dim, cfg.num_heads, (-1, -1), True, True, cfg.eps,
                                use_spatial_guidance=False)
    blk.load_state_dict(f32(sd), strict=True)
    x = synth._randn(seed, "blk.x", (1, L, C), 1.0, "cpu", torch.bfloat16)
    ctx = synth._randn(seed, "blk.ctx", (1, 257 + cfg.text_len, C), 1.0, "cpu", torch.bfloat16)
    e0 = synth._randn(seed, "blk.e0", (1, 6, C), 0.3, "cpu", torch.float32)
    d = cfg.head_dim
    freqs = torch.cat([t4d.rope_params(1024, d - 4 * (d // 6)), t4d.rope_params(1024, 2 * (d // 6)),
                       t4d.rope_params(1024, 2 * (d // 6))], dim=1)
    y = blk(x.float(), e0, torch.tensor([L]), torch.tensor([list(grid)]), freqs, ctx.float(), None,
            dtype=torch.float32, t=torch.tensor([500.0]))
    rows = torch.arange(0, L, 12)                                   # 96 rows incl. both frames
    out = {"rows": rows.to(torch.int32), "y_rows": y[0, rows].contiguous(),
           "inc_rows": (y[0, rows] - x[0, rows].float()).contiguous(),
           "y_sum": checksum(y), "x_sum": checksum(x), "ctx_sum": checksum(ctx), "e0_sum": checksum(e0),
           "w_sum": checksum(torch.cat([v.flatten().float() for v in sd.values()]))}
    save_file(out, os.path.join(OUT, name + ".safetensors"))
    print(name, tuple(y.shape), float(y.abs().mean()), float((y - x.float()).abs().mean()))


def ops_case(t4d):
    """Leaf ops: rope_apply, WanRMSNorm, attention (SDPA branch)."""
    seed = 3
    B, N, D = 2, 2, 128
    grid = (2, 3, 4)
    L = 30                                           # 24 rope'd tokens + 6 pass-through
    q = synth._randn(seed, "op.q", (B, L, N, D), 1.0, "cpu", torch.float32)
    k = synth._randn(seed, "op.k", (B, L, N, D), 1.0, "cpu", torch.float32)
    v = synth._randn(seed, "op.v", (B, L, N, D), 1.0, "cpu", torch.float32)
    freqs = torch.cat([t4d.rope_params(1024, D - 4 * (D // 6)), t4d.rope_params(1024, 2 * (D // 6)),
                       t4d.rope_params(1024, 2 * (D // 6))], dim=1)
    gs = torch.tensor([list(grid)] * B)
    qr = t4d.rope_apply(q, gs, freqs)
    nrm = t4d.WanRMSNorm(N * D, eps=1e-6)
    w = synth._randn(seed, "op.w", (N * D,), 0.1, "cpu", torch.float32, mean=1.0)
    nrm.weight.data.copy_(w)
    qn = nrm(q.flatten(2))
    att = t4d.attention(q, k, v, k_lens=torch.tensor([L, L]))
    sin = t4d.sinusoidal_embedding_1d(256, torch.tensor([500.0, 999.0, 0.0]))
    save_file({"rope": qr.contiguous(), "rms": qn.contiguous(), "attn": att.contiguous(),
               "sinus": sin.float().contiguous(), "q_sum": checksum(q)},
              os.path.join(OUT, "dit_ops.safetensors"))
    print("dit_ops", tuple(att.shape))


def model_case(t4d, name, cfg: DiTConfig, grid, batch, seed):
    sd = synth.dit_state_dict(cfg, seed)
    m = t4d.WanTransformer4DModel(
        model_type="i2v", in_dim=cfg.in_dim, dim=cfg.dim, ffn_dim=cfg.ffn_dim,
        num_heads=cfg.num_heads, num_layers=cfg.num_layers, text_dim=cfg.text_dim,
        text_len=cfg.text_len, add_ref_conv=True, use_dino_guidance=False,
        use_omnimae_guidance=False)
    m.load_state_dict(f32(sd), strict=True)
    m.eval()
    inp = synth.dit_inputs(cfg, grid, batch, seed)
    y = m(x=inp["x"].float(), t=inp["t"], context=[c.float() for c in inp["context"]],
          seq_len=inp["seq_len"], clip_fea=inp["clip_fea"].float(), y=inp["y"].float(),
          full_ref=inp["full_ref"].float())
    save_file({"y": y.contiguous(), "x_sum": checksum(inp["x"]),
               "w_sum": checksum(torch.cat([v.flatten().float() for v in sd.values()]))},
              os.path.join(OUT, name + ".safetensors"))
    print(name, tuple(y.shape), float(y.abs().mean()))


def mpm_model_case(name="dit_tiny_mpm", seed=8):
    """Full tiny model WITH the Motion-Perception front end: the real forward's first_frame branch
    (t4d:1127-1156: ImageNet normalise -> OmniMAE trunk [stub, out of scope] -> feature_adapter ->
    bilinear -> repeat over latent_T) feeding the real SpatialGuidanceModule in every block."""
    t4d, _, _ = ref_import.load_with_stub_omnimae()
    cfg = WAN_TINY.with_(use_spatial_guidance=True, use_omnimae_guidance=True)
    grid, batch = (3, 4, 6), 2
    sd = synth.dit_state_dict(cfg, seed)
    m = t4d.WanTransformer4DModel(
        model_type="i2v", in_dim=cfg.in_dim, dim=cfg.dim, ffn_dim=cfg.ffn_dim, num_heads=cfg.num_heads,
        num_layers=cfg.num_layers, text_dim=cfg.text_dim, text_len=cfg.text_len, add_ref_conv=True,
        use_dino_guidance=False, use_omnimae_guidance=True)
    m.load_state_dict(f32(sd), strict=True)
    m.eval()
    inp = synth.dit_inputs(cfg, grid, batch, seed)
    ff = synth._randn(seed, "in.first_frame", (batch, 3, 40, 56), 0.25, "cpu", torch.float32, mean=0.5).clamp(0, 1)
    captured = {}
    orig_adapter = m.feature_adapter.forward
    m.feature_adapter.forward = lambda t: captured.setdefault("adapter_out", orig_adapter(t))
    y = m(x=inp["x"].float(), t=inp["t"], context=[c.float() for c in inp["context"]], seq_len=inp["seq_len"],
          clip_fea=inp["clip_fea"].float(), y=inp["y"].float(), full_ref=inp["full_ref"].float(), first_frame=ff)
    save_file({"y": y.contiguous(), "adapter_out": captured["adapter_out"].contiguous(), "x_sum": checksum(inp["x"]),
               "ff_sum": checksum(ff),
               "w_sum": checksum(torch.cat([v.flatten().float() for v in sd.values()]))},
              os.path.join(OUT, name + ".safetensors"))
    print(name, tuple(y.shape), float(y.abs().mean()), tuple(captured["adapter_out"].shape))


TEACACHE_TS = [999.0, 990.0, 975.0, 940.0, 900.0, 700.0, 650.0, 400.0]
TEACACHE_KW = dict(coefficients=[0.0, 0.0, 0.0, 1.0, 0.0], rel_l1_thresh=0.5, num_skip_start_steps=1)


def teacache_case(t4d, name="dit_tiny_teacache", seed=4):
    """The REAL model with the reference's own TeaCache (MoRe4D/models/cache_utils.py, hooks
    t4d:1200-1270,1336-1339) over 8 steps of a toy Euler loop: the per-step skip decisions and outputs
    (skipped steps re-use the previous residual of the block stack)."""
    cfg, grid, batch = WAN_TINY, (3, 4, 6), 2
    sd = synth.dit_state_dict(cfg, seed)
    m = t4d.WanTransformer4DModel(
        model_type="i2v", in_dim=cfg.in_dim, dim=cfg.dim, ffn_dim=cfg.ffn_dim, num_heads=cfg.num_heads,
        num_layers=cfg.num_layers, text_dim=cfg.text_dim, text_len=cfg.text_len, add_ref_conv=True,
        use_dino_guidance=False, use_omnimae_guidance=False)
    m.load_state_dict(f32(sd), strict=True)
    m.eval()
    inp = synth.dit_inputs(cfg, grid, batch, seed)
    m.enable_teacache(TEACACHE_KW["coefficients"], len(TEACACHE_TS), rel_l1_thresh=TEACACHE_KW["rel_l1_thresh"],
                      num_skip_start_steps=TEACACHE_KW["num_skip_start_steps"], offload=False)
    x = inp["x"].float()
    ys, dec = [], []
    for t in TEACACHE_TS:
        y = m(x=x, t=torch.tensor([t, t]), context=[c.float() for c in inp["context"]], seq_len=inp["seq_len"],
              clip_fea=inp["clip_fea"].float(), y=inp["y"].float(), full_ref=inp["full_ref"].float())
        ys.append(y)
        dec.append(int(bool(m.should_calc)))
        x = (x - 0.05 * y).to(torch.bfloat16).float()              # toy Euler update, bf16 latents like the loop
    save_file({"y": torch.stack(ys).contiguous(), "should_calc": torch.tensor(dec, dtype=torch.int32),
               "x_sum": checksum(inp["x"])}, os.path.join(OUT, name + ".safetensors"))
    print(name, dec)


def model3d_case(t3d, name, cfg: DiTConfig, grid, batch, seed):
    """Real WanTransformer3DModel (the 4D-ViSM / Wan-InP backbone): in_dim 36, no reference conv."""
    sd = synth.dit_state_dict(cfg, seed)
    m = t3d.WanTransformer3DModel(
        model_type="i2v", in_dim=cfg.in_dim, dim=cfg.dim, ffn_dim=cfg.ffn_dim,
        num_heads=cfg.num_heads, num_layers=cfg.num_layers, text_dim=cfg.text_dim,
        text_len=cfg.text_len)
    m.load_state_dict(f32(sd), strict=True)
    m.eval()
    inp = synth.dit_inputs(cfg, grid, batch, seed, with_ref=False)
    y = m(x=inp["x"].float(), t=inp["t"], context=[c.float() for c in inp["context"]],
          seq_len=inp["seq_len"], clip_fea=inp["clip_fea"].float(), y=inp["y"].float())
    save_file({"y": y.contiguous(), "x_sum": checksum(inp["x"]),
               "w_sum": checksum(torch.cat([v.flatten().float() for v in sd.values()]))},
              os.path.join(OUT, name + ".safetensors"))
    print(name, tuple(y.shape), float(y.abs().mean()))


def project_case():
    """Outputs of the REAL `render_with_project` (scripts/inference/infer.py:222-258): its source
    and MoRe4D/utils/project_utils.py are exec'd in place (infer.py itself imports packages that
    are absent here); `scatter` (torch_scatter, absent) is stubbed with an index_add mean."""
    import ast
    import importlib.util
    import numpy as np
    ref = os.path.dirname(os.path.dirname(ref_import._PKG)) if False else os.path.dirname(ref_import._PKG)
    spec = importlib.util.spec_from_file_location("m4d_ref_project_utils",
                                                  os.path.join(ref_import._PKG, "utils", "project_utils.py"))
    pu = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(pu)
    src = open(os.path.join(ref, "scripts", "inference", "infer.py")).read()
    tree = ast.parse(src)
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "render_with_project"][0]

    def scatter(srcv, index, dim=0, reduce="mean"):
        assert dim == 0 and reduce == "mean"
        n = int(index.max()) + 1
        out = torch.zeros(n, srcv.shape[1], dtype=srcv.dtype)
        cnt = torch.zeros(n, dtype=srcv.dtype)
        out.index_add_(0, index, srcv)
        cnt.index_add_(0, index, torch.ones_like(index, dtype=srcv.dtype))
        return out / cnt.clamp_min(1)[:, None]

    from typing import Tuple
    ns = {"torch": torch, "np": np, "project": pu.project, "scatter": scatter, "Tuple": Tuple}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "infer.py:render_with_project", "exec"), ns)
    render = ns["render_with_project"]
    out = {}
    for name, (H, W, seed, tilt) in {"a": (48, 64, 0, 0.0), "b": (37, 53, 1, 0.15)}.items():
        pts, col, ext, K = synth.point_cloud(H, W, seed, tilt)
        img, mask = render(pts, ext, K, col, H, W, torch.device("cpu"))
        out[f"{name}.image"] = torch.from_numpy(np.ascontiguousarray(img))
        out[f"{name}.mask"] = torch.from_numpy(np.ascontiguousarray(mask.astype(np.uint8)))
        out[f"{name}.pts_sum"] = checksum(pts)
        print("project", name, img.shape, float(mask.mean()))
    save_file(out, os.path.join(OUT, "project.safetensors"))


def vae_cases(vae_mod, traj_mod):
    """Real AutoencoderKLWan (chunked encode/decode with the feature cache) and the two
    trajectory adaptors on small clips: 13 frames = 1 + three 4-frame chunks on the encoder
    side; 4 latent frames = 'Rep' chunk + three cached chunks on the decoder side."""
    import contextlib, io
    seed = 11
    sd = synth.vae_state_dict(seed=seed)
    m = vae_mod.AutoencoderKLWan()
    m.load_state_dict(f32(sd), strict=True)
    m.eval()
    x = synth._randn(seed, "vae.x", (1, 3, 13, 32, 48), 0.5, "cpu", torch.bfloat16).float()
    dist = m.encode(x)[0]
    params = dist.parameters
    z = synth._randn(seed, "vae.z", (1, 16, 4, 4, 6), 1.0, "cpu", torch.bfloat16).float()
    rec = m.decode(z).sample
    x1 = x[:, :, :1]
    p1 = m.encode(x1)[0].parameters                       # single-frame clip (T = 1 edge case)
    r1 = m.decode(z[:, :, :1]).sample
    out = {"enc_params": params.contiguous(), "dec": rec.contiguous(), "enc_params_T1": p1.contiguous(),
           "dec_T1": r1.contiguous(), "x_sum": checksum(x), "z_sum": checksum(z),
           "w_sum": checksum(torch.cat([v.flatten().float() for v in sd.values()]))}
    with contextlib.redirect_stdout(io.StringIO()):
        ea, da = traj_mod.VAEEncoderadaptor(), traj_mod.VAEDecoderadaptor()
    esd, dsd = synth.adaptor_state_dict("encoder", seed), synth.adaptor_state_dict("decoder", seed)
    ea.load_state_dict(f32(esd), strict=True)
    da.load_state_dict(f32(dsd), strict=True)
    tv = synth.trajectory_video(5, 32, 48, seed).float()
    out["adaptor_enc"] = ea(tv).contiguous()
    out["adaptor_dec"] = da(tv).contiguous()
    out["tv_sum"] = checksum(tv)
    save_file(out, os.path.join(OUT, "vae.safetensors"))
    print("vae", tuple(params.shape), tuple(rec.shape), float(rec.abs().mean()))


def vae_bf16_case(vae_mod, traj_mod):
    """The REAL reference VAE and adaptors run IN BF16 (weights, activations) on CPU — the dtype the
    reference's inference path uses (pipeline_wan_fun_control.py:356-386 calls the VAE outside
    autocast with weight_dtype bf16).  Every op of the path accepts bf16 on CPU (probed: Conv3d,
    F.normalize, SiLU, nearest-exact Upsample via its own .float() round trip vae:67, SDPA,
    GroupNorm).  Same inputs as vae_cases(); replaces the bf16-EMULATION argument for the
    end-to-end tolerances."""
    import contextlib, io
    seed = 11
    sd = synth.vae_state_dict(seed=seed)
    m = vae_mod.AutoencoderKLWan().to(torch.bfloat16)
    m.load_state_dict(sd, strict=True)
    m.eval()
    x = synth._randn(seed, "vae.x", (1, 3, 13, 32, 48), 0.5, "cpu", torch.bfloat16)
    z = synth._randn(seed, "vae.z", (1, 16, 4, 4, 6), 1.0, "cpu", torch.bfloat16)
    out = {"enc_params": m.encode(x)[0].parameters.float().contiguous(),
           "dec": m.decode(z).sample.float().contiguous(), "x_sum": checksum(x), "z_sum": checksum(z)}
    with contextlib.redirect_stdout(io.StringIO()):
        ea, da = traj_mod.VAEEncoderadaptor().to(torch.bfloat16), traj_mod.VAEDecoderadaptor().to(torch.bfloat16)
    ea.load_state_dict(synth.adaptor_state_dict("encoder", seed), strict=True)
    da.load_state_dict(synth.adaptor_state_dict("decoder", seed), strict=True)
    tv = synth.trajectory_video(5, 32, 48, seed)
    out["adaptor_enc"] = ea(tv).float().contiguous()
    out["adaptor_dec"] = da(tv).float().contiguous()
    # the whole Motion-Sensitive round trip (infer_vae.py:276-281 with .mode())
    pseudo = ea(tv) * 2 - 1
    lat = m.encode(pseudo)[0].mode()
    rec = m.decode(lat).sample
    out["rt_latent"], out["rt_video"], out["rt_out"] = (t.float().contiguous() for t in (lat, rec, da(rec)))
    save_file(out, os.path.join(OUT, "vae_bf16.safetensors"))
    print("vae_bf16", tuple(out["enc_params"].shape), tuple(out["rt_out"].shape))


def main():
    t4d, _vae, _traj = ref_import.load()
    if "--only-vae-bf16" in sys.argv:
        vae_bf16_case(_vae, _traj)
        return
    if "--only-14b" in sys.argv:
        block14b_case(t4d)
        return
    if "--only-teacache" in sys.argv:
        teacache_case(t4d)
        return
    if "--only-mpm" in sys.argv:
        mpm_model_case()
        return
    from more4d_b200.config import WAN_TINY_INP
    model3d_case(ref_import.load3d(), "dit3d_tiny", WAN_TINY_INP, (3, 4, 6), 2, seed=5)
    if "--only-3d" in sys.argv:
        return
    project_case()
    if "--only-new" in sys.argv:
        return
    vae_cases(_vae, _traj)
    ops_case(t4d)
    tiny = WAN_TINY
    block_case(t4d, "block_tiny", tiny, (2, 3, 4), 30, seed=1)
    block_case(t4d, "block_tiny_mpm", tiny.with_(use_spatial_guidance=True), (2, 3, 4), 30, seed=2,
               guidance=True)
    # BASELINE.json configs[0]: single DiT block, [1, 2x9x16, 1536], t=500
    block_case(t4d, "block_config1", WAN_1_3B.with_(num_layers=1), (2, 9, 16), 288, seed=0)
    model_case(t4d, "dit_tiny", tiny, (3, 4, 6), 2, seed=4)


if __name__ == "__main__":
    main()
