"""from_pretrained of the mirrors follows the reference's loading rules (wan_transformer4d.py:
1393-1534, wan_vae.py:849-871): config.json -> constructor, sharded safetensors, zero-padded
patch embedding for the 48 -> 64 channel widening, size-mismatched tensors skipped, `model.` prefix
for the VAE."""
import json
import os

import torch
from safetensors.torch import save_file

from more4d_b200 import synth
from more4d_b200.config import WAN_TINY


def test_transformer_from_pretrained_widens_patch_embedding(tmp_path):
    from more4d_b200.dit import WanTransformer4DModel
    cfg = WAN_TINY
    sd = synth.dit_state_dict(cfg, 3)
    narrow = sd["patch_embedding.weight"][:, :48].contiguous()              # released Control checkpoint: 48 ch
    sd_ckpt = dict(sd, **{"patch_embedding.weight": narrow})
    sd_ckpt["blocks.0.ffn.0.bias"] = torch.zeros(7, dtype=torch.bfloat16)   # wrong size -> skipped
    sd_ckpt["not.a.parameter"] = torch.zeros(3, dtype=torch.bfloat16)       # unknown -> ignored
    keys = sorted(sd_ckpt)
    save_file({k: sd_ckpt[k] for k in keys[::2]}, str(tmp_path / "model-00001-of-00002.safetensors"))
    save_file({k: sd_ckpt[k] for k in keys[1::2]}, str(tmp_path / "model-00002-of-00002.safetensors"))
    config = dict(model_type="i2v", dim=cfg.dim, ffn_dim=cfg.ffn_dim, num_heads=cfg.num_heads,
                  num_layers=cfg.num_layers, text_dim=cfg.text_dim, text_len=cfg.text_len, in_dim=48,
                  add_ref_conv=True, _class_name="ignored", some_unknown_field=1)
    with open(tmp_path / "config.json", "w") as f:
        json.dump(config, f)
    m = WanTransformer4DModel.from_pretrained(str(tmp_path), transformer_additional_kwargs={"in_dim": 64},
                                              device="cpu")
    own = m.state_dict()
    assert own["patch_embedding.weight"].shape[1] == 64
    assert torch.equal(own["patch_embedding.weight"][:, :48], narrow)
    assert float(own["patch_embedding.weight"][:, 48:].abs().sum()) == 0.0
    assert torch.equal(own["blocks.1.self_attn.q.weight"], sd["blocks.1.self_attn.q.weight"])
    assert m.load_report["missing"] == ["blocks.0.ffn.0.bias"] and m.load_report["unexpected"] == []
    m2 = WanTransformer4DModel.from_pretrained(str(tmp_path.parent), subfolder=tmp_path.name, device="meta",
                                               transformer_additional_kwargs={"dict_mapping": {"in_dim": "in_dim"}})
    assert m2.in_dim == 48


def test_vae_from_pretrained_adds_model_prefix(tmp_path):
    from more4d_b200.vae import AutoencoderKLWan
    sd = synth.vae_state_dict(seed=2)
    inner = {k[len("model."):]: v for k, v in sd.items()}
    path = str(tmp_path / "Wan2.1_VAE.safetensors")
    save_file(inner, path)
    m = AutoencoderKLWan.from_pretrained(path, additional_kwargs={"latent_channels": 16, "bogus": 1}, device="cpu")
    assert m.load_report == {"missing": [], "unexpected": []}
    for k, v in m.state_dict().items():
        assert torch.equal(v, sd[k])
    torch.save(inner, str(tmp_path / "vae.pth"))
    m2 = AutoencoderKLWan.from_pretrained(str(tmp_path / "vae.pth"), device="cpu")
    assert torch.equal(m2.state_dict()["model.conv1.weight"], sd["model.conv1.weight"])
