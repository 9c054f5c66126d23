"""Seam test against the REAL reference modules (authoring container only — the GPU box has no
/root/reference): install() must share parameter storage with the reference objects, re-point
their entry methods, and — on a box without CUDA — fail loudly instead of computing anywhere
else."""
import contextlib
import io

import pytest
import torch

from oracle import ref_import

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference tree not present")


def test_install_transformer_shares_parameters_and_has_no_fallback():
    from more4d_b200 import install, synth
    from more4d_b200.config import WAN_TINY as cfg
    t4d, _, _ = ref_import.load()
    ref = t4d.WanTransformer4DModel(model_type="i2v", in_dim=cfg.in_dim, dim=cfg.dim, ffn_dim=cfg.ffn_dim,
                                    num_heads=cfg.num_heads, num_layers=cfg.num_layers, text_dim=cfg.text_dim,
                                    text_len=cfg.text_len, add_ref_conv=True, use_dino_guidance=False,
                                    use_omnimae_guidance=False).to(torch.bfloat16)
    ref.load_state_dict(synth.dit_state_dict(cfg, 0), strict=True)
    orig_forward = ref.forward
    install.install(transformer=ref)
    ours = ref._m4d
    assert ref.forward != orig_forward
    for name, p in ref.named_parameters():
        assert dict(ours.named_parameters())[name] is p                  # same Parameter object
    # a later in-place weight update on the reference is visible to the mirror
    with torch.no_grad():
        ref.blocks[0].self_attn.q.weight.zero_()
    assert float(ours.blocks[0].self_attn.q.weight.abs().sum()) == 0.0
    inp = synth.dit_inputs(cfg, (3, 4, 6), 2, 0)
    if not torch.cuda.is_available():
        with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
            ref(x=inp["x"], t=inp["t"], context=inp["context"], seq_len=inp["seq_len"],
                clip_fea=inp["clip_fea"], y=inp["y"], full_ref=inp["full_ref"])


def test_install_vae_and_adaptors():
    from more4d_b200 import install, synth
    _, vae_mod, traj = ref_import.load()
    ref = vae_mod.AutoencoderKLWan().to(torch.bfloat16)
    ref.load_state_dict(synth.vae_state_dict(seed=0), strict=True)
    with contextlib.redirect_stdout(io.StringIO()):
        ea, da = traj.VAEEncoderadaptor().to(torch.bfloat16), traj.VAEDecoderadaptor().to(torch.bfloat16)
    install.install(vae=ref, encoder_prompt=ea, decoder_prompt=da)
    for mod in (ref, ea, da):
        mine = dict(mod._m4d.named_parameters())
        for name, p in mod.named_parameters():
            assert mine[name] is p
    assert type(ea._m4d).__name__ == "VAEEncoderadaptor" and type(da._m4d).__name__ == "VAEDecoderadaptor"
    if not torch.cuda.is_available():
        with torch.no_grad(), pytest.raises(RuntimeError):
            ref.decode(torch.zeros(1, 16, 1, 4, 4, dtype=torch.bfloat16))


def test_install_3d_transformer_visim_backbone():
    """install() also adopts the 4D-ViSM backbone (wan_transformer3d.WanTransformer3DModel)."""
    from more4d_b200 import install, synth
    from more4d_b200.config import WAN_TINY_INP as cfg
    t3d = ref_import.load3d()
    ref = t3d.WanTransformer3DModel(model_type="i2v", in_dim=cfg.in_dim, dim=cfg.dim, ffn_dim=cfg.ffn_dim,
                                    num_heads=cfg.num_heads, num_layers=cfg.num_layers, text_dim=cfg.text_dim,
                                    text_len=cfg.text_len).to(torch.bfloat16)
    ref.load_state_dict(synth.dit_state_dict(cfg, 0), strict=True)
    install.install(transformer=ref)
    mine = dict(ref._m4d.named_parameters())
    assert set(mine) == set(dict(ref.named_parameters()))
    for name, p in ref.named_parameters():
        assert mine[name] is p
    assert ref._m4d.ref_conv is None and ref._m4d.blocks[0].spatial_guidance_self is None
    inp = synth.dit_inputs(cfg, (3, 4, 6), 2, 0, with_ref=False)
    if not torch.cuda.is_available():
        with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
            ref(x=inp["x"], t=inp["t"], context=inp["context"], seq_len=inp["seq_len"],
                clip_fea=inp["clip_fea"], y=inp["y"])


def test_riflex_tables_match_reference():
    """enable_riflex / disable_riflex (t4d:1011-1036): identical `freqs` tables to the reference's."""
    from more4d_b200.config import WAN_TINY as cfg
    from more4d_b200.dit import WanTransformer4DModel
    t4d, _, _ = ref_import.load()
    ref = t4d.WanTransformer4DModel(model_type="i2v", in_dim=cfg.in_dim, dim=cfg.dim, ffn_dim=cfg.ffn_dim,
                                    num_heads=cfg.num_heads, num_layers=1, text_dim=cfg.text_dim,
                                    text_len=cfg.text_len, add_ref_conv=True, use_dino_guidance=False,
                                    use_omnimae_guidance=False)
    ours = WanTransformer4DModel.from_config(cfg.with_(num_layers=1), device="meta")
    assert torch.equal(ours.freqs, ref.freqs)
    for kw in (dict(), dict(k=4, L_test=33, L_test_scale=2.0)):
        ref.enable_riflex(**kw)
        ours.enable_riflex(**kw)
        assert ours.freqs.dtype == ref.freqs.dtype and torch.equal(ours.freqs, ref.freqs)
    ref.disable_riflex()
    ours.disable_riflex()
    assert torch.equal(ours.freqs, ref.freqs)
