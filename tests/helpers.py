"""Shared helpers for the GPU parity tests (oracle = checker, CUDA path = thing under test)."""
import numpy as np
import torch


def rel_err(got: torch.Tensor, want: torch.Tensor) -> float:
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    return float((got - want).abs().max() / want.abs().max().clamp(min=1e-30))


def random_coords(n: int, seed: int = 0, scale: float = 0.3) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return scale * torch.randn(n, 3, generator=g)


def wish_from_map(adj: torch.Tensor, factor: float = 1.0) -> torch.Tensor:
    from oracle.wish import cont2dist

    return cont2dist(adj.clone().double(), factor)


def small_map(n: int, density: float, seed: int = 0) -> torch.Tensor:
    from hic_gnn_b200 import synth

    return synth.synthetic_map(n, density, seed=seed)
