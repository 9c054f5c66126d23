"""Import the real reference modules *in place* from /root/reference (oracle-side only).

The reference cannot be imported as shipped: ``MoRe4D/dist`` is missing from the release
(imported at MoRe4D/models/wan_transformer4d.py:23-25) and ``diffusers`` is not installed.
This module pre-seeds ``sys.modules`` with the minimum stand-ins so that the three hot-path
files import unmodified:

  MoRe4D/models/wan_transformer4d.py   (needs diffusers.{configuration_utils,loaders,models,utils}, MoRe4D.dist, MoRe4D.utils.cfg_skip)
  MoRe4D/models/wan_vae.py             (needs diffusers.models.autoencoders.vae, modeling_outputs, accelerate_utils)
  MoRe4D/models/trajectory_module.py

Nothing is copied out of, or written into, the reference tree.  ``available()`` is False on
the GPU box, where /root/reference does not exist; callers must gate on it.
"""
from __future__ import annotations

import importlib
import logging
import os
import sys
import types

import torch
import torch.nn as nn

REF_ROOT = os.environ.get("MORE4D_REFERENCE_ROOT", "/root/reference")
_PKG = os.path.join(REF_ROOT, "MoRe4D")


def available() -> bool:
    return os.path.isfile(os.path.join(_PKG, "models", "wan_transformer4d.py"))


def _mod(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _GaussianPosterior:
    """Stand-in with the semantics of diffusers' DiagonalGaussianDistribution that
    wan_vae.py:797 relies on (mean/logvar split on dim 1, logvar clamp, mode, sample)."""

    def __init__(self, parameters):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)

    def mode(self):
        return self.mean

    def sample(self, generator=None):
        eps = torch.randn(self.mean.shape, generator=generator, dtype=self.mean.dtype,
                          device=self.mean.device)
        return self.mean + self.std * eps


class _DecoderOutput:
    def __init__(self, sample):
        self.sample = sample


class _EncoderOutput:
    def __init__(self, latent_dist):
        self.latent_dist = latent_dist

    def __getitem__(self, i):
        return (self.latent_dist,)[i]


def _install_stubs() -> None:
    if "MoRe4D.models.wan_transformer4d" in sys.modules:
        return

    class ConfigMixin:
        config_name = "config.json"

    class ModelMixin(nn.Module):
        _supports_gradient_checkpointing = False

    class FromOriginalModelMixin:
        pass

    class _Logging:
        @staticmethod
        def get_logger(name):
            return logging.getLogger(name)

    _mod("diffusers")
    _mod("diffusers.loaders")
    _mod("diffusers.models")
    _mod("diffusers.models.autoencoders")
    _mod("diffusers.configuration_utils", ConfigMixin=ConfigMixin,
         register_to_config=lambda f: f)
    _mod("diffusers.loaders.single_file_model", FromOriginalModelMixin=FromOriginalModelMixin)
    _mod("diffusers.models.modeling_utils", ModelMixin=ModelMixin)
    _mod("diffusers.utils", is_torch_version=lambda op, v: True, logging=_Logging())
    _mod("diffusers.utils.accelerate_utils", apply_forward_hook=lambda f: f)
    _mod("diffusers.models.autoencoders.vae", DecoderOutput=_DecoderOutput,
         DiagonalGaussianDistribution=_GaussianPosterior)
    _mod("diffusers.models.modeling_outputs", AutoencoderKLOutput=_EncoderOutput)

    # MoRe4D package skeleton: bypass the heavy __init__ files and the missing MoRe4D/dist.
    _mod("MoRe4D").__path__ = [_PKG]
    _mod("MoRe4D.models").__path__ = [os.path.join(_PKG, "models")]
    _mod("MoRe4D.dist",
         get_sequence_parallel_rank=lambda: 0,
         get_sequence_parallel_world_size=lambda: 1,
         get_sp_group=lambda: None,
         usp_attn_forward=None,
         xFuserLongContextAttention=None)
    utils = _mod("MoRe4D.utils")
    utils.__path__ = [os.path.join(_PKG, "utils")]
    utils.cfg_skip = importlib.import_module("MoRe4D.utils.cfg_optimization").cfg_skip


def load():
    """Return (wan_transformer4d, wan_vae, trajectory_module) reference modules."""
    if not available():
        raise RuntimeError(f"reference tree not found under {REF_ROOT}")
    # Anything but FLASH_ATTENTION / SAGE_ATTENTION selects the reference's own SDPA branch
    # (wan_transformer4d.py:221-235); the flash branch asserts CUDA (line 96).
    os.environ["VIDEOX_ATTENTION_TYPE"] = "SDPA"
    _install_stubs()
    t4d = importlib.import_module("MoRe4D.models.wan_transformer4d")
    vae = importlib.import_module("MoRe4D.models.wan_vae")
    traj = importlib.import_module("MoRe4D.models.trajectory_module")
    return t4d, vae, traj


def load3d():
    """The reference's 4D-ViSM backbone module (MoRe4D/models/wan_transformer3d.py)."""
    load()
    return importlib.import_module("MoRe4D.models.wan_transformer3d")

