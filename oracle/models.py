"""Oracle (test infrastructure): the three live networks of the reference.

* ``Net``                                    -- ``models.py:14-55``
* ``GATNetSelectiveResidualsUpdated``        -- ``models.py:614-691``
* ``GATNetHeadsChanged3LayersLeakyReLUv2``   -- ``models.py:1010-1047``

Same attribute names, construction order and ``state_dict`` keys as the reference, so a
reference ``*_weights.pt`` loads unchanged.  ``forward`` returns the N x N ``cdist`` matrix,
``get_model`` the N x 3 coordinates.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import cdist, nn
from torch.nn import LayerNorm, Linear

from .conv import GATConv, SAGEConv


class Net(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv = SAGEConv(512, 512)  # models.py:17
        self.densea = Linear(512, 256)
        self.dense1 = Linear(256, 128)
        self.dense2 = Linear(128, 64)
        self.dense3 = Linear(64, 3)

    def get_model(self, x, edge_index):  # models.py:44-55
        x = self.conv(x, edge_index).relu()
        x = self.densea(x).relu()
        x = self.dense1(x).relu()
        x = self.dense2(x).relu()
        return self.dense3(x)

    def forward(self, x, edge_index):  # models.py:23-42
        x = self.get_model(x, edge_index)
        return cdist(x, x, p=2)


class GATNetSelectiveResidualsUpdated(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv = GATConv(512, 256, heads=2)  # models.py:619
        self.densea = Linear(512, 256)
        self.norm_a = LayerNorm(256)
        self.align_densea = Linear(512, 256)
        self.dense1 = Linear(256, 128)
        self.norm1 = LayerNorm(128)
        self.align_dense1 = Linear(256, 128)
        self.dense2 = Linear(128, 64)
        self.norm2 = LayerNorm(64)
        self.dense3 = Linear(64, 3)
        self.dense_graph = False

    def get_model(self, x, edge_index):  # models.py:664-691
        x = F.relu(self.conv(x, edge_index, dense=self.dense_graph))
        x_initial = self.align_densea(x)
        x = F.relu(self.norm_a(self.densea(x))) + x_initial
        x_initial = self.align_dense1(x)
        x = F.relu(self.norm1(self.dense1(x))) + x_initial
        x = F.relu(self.norm2(self.dense2(x)))
        return self.dense3(x)

    def forward(self, x, edge_index):  # models.py:634-662
        x = self.get_model(x, edge_index)
        return cdist(x, x, p=2)


class GATNetHeadsChanged3LayersLeakyReLUv2(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv = GATConv(512, 256, heads=2)  # models.py:1013
        self.densea = Linear(512, 256)
        self.dense1 = Linear(256, 64)
        self.dense2 = Linear(64, 3)
        self.dense_graph = False

    def get_model(self, x, edge_index):  # models.py:1036-1047
        x = F.leaky_relu(self.conv(x, edge_index, dense=self.dense_graph))
        x = F.leaky_relu(self.densea(x))
        x = F.leaky_relu(self.dense1(x))
        return self.dense2(x)

    def forward(self, x, edge_index):  # models.py:1020-1034
        x = self.get_model(x, edge_index)
        return cdist(x, x, p=2)
