#!/usr/bin/env python
"""ncu target: the GATConv projection GEMM of the C5 step (49 850 x 512 x 512, 3xTF32) a few times."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from hic_gnn_b200 import ops

g = torch.Generator(device="cuda").manual_seed(0)
x = 0.5 * torch.randn(49850, 512, generator=g, device="cuda")
w = torch.randn(512, 512, generator=g, device="cuda") / 22.6
xs, ws = ops.split_tf32(x, 0b100), ops.split_tf32(w, 0b010)
for _ in range(4):
    y = ops.gemm_tf32_tn(xs, ws)
torch.cuda.synchronize()
print("ok", float(y.abs().max()))
