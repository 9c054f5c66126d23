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



# --------------------------------------------------------------------------------------
# The CALLER of the hot path: MoRe4D/pipeline/pipeline_wan_fun_control.py
# --------------------------------------------------------------------------------------
class EulerFlowScheduler:
    """Stand-in for diffusers' FlowMatchEulerDiscreteScheduler (absent here; its sigma grid lives in
    diffusers and is NOT guessed): the only in-tree schedule, get_sampling_sigmas
    (MoRe4D/utils/fm_solvers.py:22-26), with the Euler update x += (sigma_next - sigma) * v in
    fp32.  Interface = what pctl:576-577,741,825 uses: set_timesteps(n, device=, mu=), .timesteps,
    .order, .step(model_output, t, sample, return_dict=False)."""
    order = 1

    def __init__(self, shift: float = 5.0):
        self.shift = shift
        self.timesteps = None
        self.sigmas = None
        self._i = 0

    def set_timesteps(self, num_inference_steps, device=None, mu=None):
        import numpy as np
        sig = np.linspace(1, 0, num_inference_steps + 1)[:num_inference_steps]
        sig = self.shift * sig / (1 + (self.shift - 1) * sig)
        self.sigmas = np.append(sig, 0.0)
        self.timesteps = torch.tensor(sig * 1000.0, dtype=torch.float32, device=device)
        self._i = 0

    def step(self, model_output, timestep, sample, return_dict=False):
        dt = float(self.sigmas[self._i + 1] - self.sigmas[self._i])
        self._i += 1
        prev = (sample.float() + dt * model_output.float()).to(sample.dtype)
        return (prev,)


def _install_pipeline_stubs() -> None:
    if "MoRe4D.pipeline.pipeline_wan_fun_control" in sys.modules:
        return
    import contextlib
    import enum
    from dataclasses import dataclass  # noqa: F401

    class DiffusionPipeline:
        """register_modules / components / progress_bar / maybe_free_model_hooks of diffusers'
        base class, as pipeline_wan_fun_control.py:170-189,741,853 use them."""

        def __init__(self):
            self._components = {}

        def register_modules(self, **kw):
            for k, v in kw.items():
                self._components[k] = v
                setattr(self, k, v)

        @property
        def components(self):
            return dict(self._components)

        @contextlib.contextmanager
        def progress_bar(self, total=None):
            class _Bar:
                def update(self, n=1):
                    pass
            yield _Bar()

        def maybe_free_model_hooks(self):
            pass

    class _Proc:
        def __init__(self, *a, **kw):
            pass

        def preprocess(self, x, height=None, width=None):
            return x

        def postprocess_video(self, video, output_type="np"):
            return video

    class BaseOutput:
        pass

    class KarrasDiffusionSchedulers(enum.Enum):
        EulerDiscreteScheduler = 1

    class SchedulerMixin:
        pass

    class SchedulerOutput:
        def __init__(self, prev_sample):
            self.prev_sample = prev_sample

    def randn_tensor(shape, generator=None, device=None, dtype=None, layout=None):
        return torch.randn(shape, generator=generator, dtype=dtype).to(device)

    d = sys.modules["diffusers"]
    d.FlowMatchEulerDiscreteScheduler = EulerFlowScheduler
    _mod("diffusers.callbacks", MultiPipelineCallbacks=type("MultiPipelineCallbacks", (), {}),
         PipelineCallback=type("PipelineCallback", (), {}))
    _mod("diffusers.image_processor", VaeImageProcessor=_Proc)
    _mod("diffusers.models.embeddings", get_1d_rotary_pos_embed=None)
    _mod("diffusers.pipelines")
    _mod("diffusers.pipelines.pipeline_utils", DiffusionPipeline=DiffusionPipeline)
    _mod("diffusers.schedulers", FlowMatchEulerDiscreteScheduler=EulerFlowScheduler)
    _mod("diffusers.schedulers.scheduling_utils", KarrasDiffusionSchedulers=KarrasDiffusionSchedulers,
         SchedulerMixin=SchedulerMixin, SchedulerOutput=SchedulerOutput)
    du = sys.modules["diffusers.utils"]
    du.BaseOutput = BaseOutput
    du.replace_example_docstring = lambda doc: (lambda f: f)
    du.deprecate = lambda *a, **kw: None
    du.is_scipy_available = lambda: True
    _mod("diffusers.utils.torch_utils", randn_tensor=randn_tensor)
    _mod("diffusers.video_processor", VideoProcessor=_Proc)
    models = sys.modules["MoRe4D.models"]
    models.AutoencoderKLWan = importlib.import_module("MoRe4D.models.wan_vae").AutoencoderKLWan
    models.WanTransformer3DModel = importlib.import_module("MoRe4D.models.wan_transformer3d").WanTransformer3DModel
    for name in ("AutoTokenizer", "CLIPModel", "WanT5EncoderModel"):
        setattr(models, name, type(name, (), {}))
    _mod("MoRe4D.pipeline").__path__ = [os.path.join(_PKG, "pipeline")]


def load_pipeline():
    """The reference's 4D-STraG pipeline module, imported in place with the stand-ins above; returns
    (module, EulerFlowScheduler)."""
    load()
    _install_pipeline_stubs()
    return importlib.import_module("MoRe4D.pipeline.pipeline_wan_fun_control"), EulerFlowScheduler


# --------------------------------------------------------------------------------------
# Motion-Perception front end (t4d:1127-1156): stand-in for the frozen OmniMAE ViT-B trunk
# --------------------------------------------------------------------------------------
class StubOmniMAE(nn.Module):
    """Parameter-free stand-in for `vit_base_mae_pretraining()` (MoRe4D/models/omnimae.py:77-143;
    the real one loads a checkpoint from a hard-coded path, omnimae.py:64-74).  The trunk is OUT OF
    SCOPE — only its interface matters: `.trunk.forward_patch_features(img [1,3,H,W], None)` ->
    (tokens [1, 196, 768], cls [1, 768]) (omnivision/models/vision_transformer.py:688-703).  Tokens
    are a fixed deterministic function of the image so that the reference forward, the oracle and
    the CUDA mirror see the same trunk output."""

    class _Trunk(nn.Module):
        def forward_patch_features(self, x, use_checkpoint=False):
            import torch.nn.functional as F
            g = torch.Generator().manual_seed(1234)
            proj = torch.randn(3, 768, generator=g)
            table = torch.randn(196, 768, generator=g) * 0.5
            pooled = F.adaptive_avg_pool2d(x.float().cpu(), 14).flatten(2).transpose(1, 2)    # [1, 196, 3]
            tok = (pooled @ proj + table).to(device=x.device)
            return tok, tok.mean(dim=1)

    def __init__(self):
        super().__init__()
        self.trunk = StubOmniMAE._Trunk()


def load_with_stub_omnimae():
    """load() with `MoRe4D.models.omnimae` pre-seeded by the stub, so that the REAL
    WanTransformer4DModel(use_omnimae_guidance=True) constructs (t4d:883-892) and its REAL forward
    runs the Motion-Perception branch (t4d:1127-1156) on CPU."""
    mods = load()
    _mod("MoRe4D.models.omnimae", vit_base_mae_pretraining=lambda pretrained=True: StubOmniMAE())
    return mods
