"""hic_gnn_b200 -- B200-native (sm_100a) training hot path of HiC-GNN / GAT-HiC.

Python surface mirrors the reference (``models.Net`` / the GAT nets with
``forward(x, edge_index[, edge_weight])`` + ``get_model``, ``utils.load_input``,
``utils.cont2dist``, the training loops); the arithmetic runs in hand-written CUDA kernels
behind the C ABI of ``include/hicgat.h`` (``libhicgat_sm100.so``).  No Triton, no
torch_geometric, no CPU fallback.
"""
from . import _native, ops  # noqa: F401
from .ops import SparseWishTarget, WishTarget, pairwise_loss, pair_moments, pairdist  # noqa: F401

__version__ = "0.1.0"
