"""Oracle (test infrastructure): contact -> "wish distance" conversion.

Follows reference ``utils.py:75-80`` (``cont2dist``) line by line.
"""
from __future__ import annotations

import torch


def cont2dist(adj: torch.Tensor, factor: float) -> torch.Tensor:
    """``utils.py:75-80``: ``(1/adj)**factor``, diagonal 0, ``inf -> max finite``, ``/ max``."""
    dist = (1 / adj) ** factor  # utils.py:76
    dist.fill_diagonal_(0)  # utils.py:77
    mx = torch.max(torch.nan_to_num(dist, posinf=0))  # utils.py:78
    dist = torch.nan_to_num(dist, posinf=mx)  # utils.py:79
    return dist / mx  # utils.py:80
