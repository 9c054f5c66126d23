"""Shared test utilities."""
import torch


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """Relative Frobenius error ||a-b|| / ||b|| (the parity metric of BASELINE.md §4)."""
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def checksum(t: torch.Tensor) -> torch.Tensor:
    t = t.double()
    return torch.stack([t.sum(), t.abs().sum(), (t * t).sum()]).float()
