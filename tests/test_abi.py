"""The C-ABI library loads and exports every symbol include/hicgat.h declares (CPU box: no
compute calls)."""
import ctypes
import os

import pytest

from hic_gnn_b200 import _native as N


@pytest.fixture(scope="module")
def built():
    from hic_gnn_b200 import build

    return build.build()


def test_library_exports_every_declared_symbol(built):
    handle = ctypes.CDLL(built)
    names = N.declared_symbols()
    assert len(names) >= 15
    for name in names:
        assert hasattr(handle, name), f"{name} declared in hicgat.h but not exported"
    assert sorted(N.SIGNATURES) == names, "ctypes table and header disagree"


def test_version_and_error_text(built):
    lib = N.lib()
    assert lib.hicgat_version() >= 100
    # argument validation happens before any CUDA call, so this is safe without a GPU
    rc = lib.hicgat_pairloss_fwd_bwd(None, None, 8, 8, 0, 8, 0, 0.0, 0.0, None, None, None, 0, None)
    assert rc == -1
    assert b"null pointer" in lib.hicgat_last_error()
    with pytest.raises(RuntimeError, match="null pointer"):
        N.check(rc, "pairloss")
    assert lib.hicgat_pairloss_set_tuning(7, 0) == -1


def test_missing_library_fails_loudly(monkeypatch, built):
    monkeypatch.setattr(N, "_lib", None)
    monkeypatch.setattr(N, "LIB_PATH", os.path.join(os.path.dirname(built), "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU / PyTorch fallback"):
        N.lib()


def test_ops_reject_cpu_tensors(built):
    import torch

    from hic_gnn_b200 import ops

    with pytest.raises(RuntimeError, match="CUDA"):
        ops.cont2dist(torch.ones(4, 4, dtype=torch.float64), 1.0)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.csr_from_dense(torch.ones(4, 4, dtype=torch.float64))
