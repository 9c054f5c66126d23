"""Pin oracle/vae_oracle.py (full-sequence formulation) against the real reference VAE run
chunk-by-chunk with its feature cache, and the real trajectory adaptors (goldens from
tests/golden/make_golden.py).  fp32 vs fp32."""
import torch

from more4d_b200 import synth
from oracle import vae_oracle as V
from tests.helpers import checksum, rel_err

SEED = 11
TOL = 3e-5


def test_vae_encode_decode(golden):
    g = golden("vae")
    sd = synth.vae_state_dict(seed=SEED)
    x = synth._randn(SEED, "vae.x", (1, 3, 13, 32, 48), 0.5, "cpu", torch.bfloat16).float()
    z = synth._randn(SEED, "vae.z", (1, 16, 4, 4, 6), 1.0, "cpu", torch.bfloat16).float()
    assert torch.allclose(checksum(x), g["x_sum"], rtol=1e-6), "RNG drift: regenerate goldens"
    assert torch.allclose(checksum(z), g["z_sum"], rtol=1e-6)
    params = V.encode(x, sd)
    assert params.shape == g["enc_params"].shape == (1, 32, 4, 4, 6)
    assert rel_err(params, g["enc_params"]) < TOL
    rec = V.decode(z, sd)
    assert rec.shape == g["dec"].shape == (1, 3, 13, 32, 48)
    assert rel_err(rec, g["dec"]) < TOL
    # single-frame clips: only the 'first chunk' rules are exercised
    assert rel_err(V.encode(x[:, :, :1], sd), g["enc_params_T1"]) < TOL
    assert rel_err(V.decode(z[:, :, :1], sd), g["dec_T1"]) < TOL


def test_causality_chunk_equivalence():
    """Encoding a prefix gives the prefix of the encoding (what makes the full-sequence
    formulation equal to the reference's chunk loop)."""
    sd = synth.vae_state_dict(seed=SEED)
    x = synth._randn(SEED, "vae.x", (1, 3, 13, 32, 48), 0.5, "cpu", torch.bfloat16).float()
    full = V.encode(x, sd)
    part = V.encode(x[:, :, :9], sd)
    assert rel_err(part, full[:, :, :3]) < 1e-5


def test_adaptors(golden):
    g = golden("vae")
    tv = synth.trajectory_video(5, 32, 48, SEED).float()
    assert torch.allclose(checksum(tv), g["tv_sum"], rtol=1e-6), "RNG drift: regenerate goldens"
    esd, dsd = synth.adaptor_state_dict("encoder", SEED), synth.adaptor_state_dict("decoder", SEED)
    assert rel_err(V.encoder_adaptor(tv, esd), g["adaptor_enc"]) < TOL
    assert rel_err(V.decoder_adaptor(tv, dsd), g["adaptor_dec"]) < TOL
    # bf16 emulation stays close to the fp32 gold
    assert rel_err(V.encoder_adaptor(tv, esd, True), g["adaptor_enc"]) < 5e-3
    assert rel_err(V.decoder_adaptor(tv, dsd, True), g["adaptor_dec"]) < 3e-2
