"""more4d_b200 — B200-native (sm_100a) implementation of MoRe4D's 4D-STraG denoising hot path.

Host side is Python/PyTorch (device memory, streams, torch.distributed) mirroring the
reference's call surface; all arithmetic on the path runs in hand-written CUDA kernels
reached through the C ABI declared in ``include/more4d_b200.h``.  There is no CPU fallback:
every op raises if ``libmore4d_sm100.so`` is missing or the device is not sm_100.
"""
from .config import DiTConfig, WAN_14B, WAN_1_3B, WAN_TINY, token_grid  # noqa: F401

__version__ = "0.1.0"
