"""Host-side logic of the row-sharded loss on CPU: world_size-2 (and 3) gloo process groups.
The CUDA kernel is replaced by a torch stand-in that fills the same packed buffer for the
rank's row block; what is tested is the partition, the packing and the single all-reduce."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hic_gnn_b200 import sharding


def _local_block_cpu(coords, truth, r0, r1, packed):
    """Reference arithmetic of one row block (moments layout of include/hicgat.h)."""
    n = coords.shape[0]
    c = coords.double()
    d = torch.cdist(c[r0:r1], c)
    t = truth[r0:r1].double()
    e = d - t
    rows = torch.arange(r0, r1).unsqueeze(1)
    cols = torch.arange(n).unsqueeze(0)
    up = rows < cols
    m = torch.zeros(8, dtype=torch.float64)
    m[0] = (e * e).sum()
    m[1] = e[up].abs().sum()
    m[2], m[3], m[4], m[5], m[6], m[7] = d[up].sum(), (d[up] ** 2).sum(), t[up].sum(), (t[up] ** 2).sum(), (d[up] * t[up]).sum(), (e[up] ** 2).sum()
    w = torch.where(d > 0, e / d.clamp(min=1e-30), torch.zeros_like(d))  # [rows, n]
    diff = c.unsqueeze(0) - c[r0:r1].unsqueeze(1)  # x_j - x_i
    g = (4.0 / n**2) * (w.unsqueeze(-1) * diff).sum(0)  # column-side sum over this block's rows
    packed[:8] = m
    packed[8:] = g.float().double().reshape(-1)


def _worker(rank, world, port, n, tmpdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        truth = torch.rand(n, n, generator=g, dtype=torch.float64)
        truth = (truth + truth.t()) / 2
        truth.fill_diagonal_(0)
        coords = (0.3 * torch.randn(n, 3, generator=g)).requires_grad_(True)
        r0, r1 = sharding.row_block(n, rank, world)
        loss_fn = sharding.ShardedPairLoss(n, lambda c, p: _local_block_cpu(c.detach(), truth, r0, r1, p), "cpu")
        moments, grad = loss_fn(coords)
        # single-process answer
        full = torch.zeros(8 + 3 * n, dtype=torch.float64)
        _local_block_cpu(coords.detach(), truth, 0, n, full)
        m_full, g_full = sharding.unpack(full, n)
        assert torch.allclose(moments, m_full, rtol=1e-12), (rank, moments, m_full)
        assert torch.allclose(grad, g_full, rtol=1e-5, atol=1e-9)
        # and the autograd answer of the reference formulation
        l = ((torch.cdist(coords, coords) - truth.float()) ** 2).mean()
        (ga,) = torch.autograd.grad(l, coords)
        assert abs(float(moments[0]) / n**2 - float(l)) / float(l) < 1e-5
        assert float((grad - ga).abs().max() / ga.abs().max()) < 1e-4
        # sharded wish-distance max
        mx = torch.tensor([float(rank + 1)], dtype=torch.float64)
        sharding.allreduce_max_(mx)
        assert float(mx) == world
        open(os.path.join(tmpdir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 37), (3, 10)])
def test_sharded_loss_gloo(tmp_path, world, n):
    port = 29500 + os.getpid() % 2000 + world
    mp.spawn(_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_row_blocks_cover_exactly():
    for n in (1, 7, 58, 1000, 49850):
        for world in (1, 2, 3, 4, 8):
            blocks = sharding.all_blocks(n, world)
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            for (a0, a1), (b0, b1) in zip(blocks, blocks[1:]):
                assert a1 == b0 and a0 <= a1
            assert sum(b - a for a, b in blocks) == n


def test_upper_triangle_blocks_and_measured_rebalance():
    """Area-balanced blocks of the upper-triangle kernel cover the rows exactly and carry equal triangle area; the measured rebalance
    step moves boundaries towards the slow ranks and is a fixed point for equal times."""
    for n in (9970, 49850):
        for world in (2, 4, 8):
            blocks = sharding.all_blocks(n, world, "upper")
            assert blocks[0][0] == 0 and blocks[-1][1] == n and all(a1 == b0 for (_, a1), (b0, _) in zip(blocks, blocks[1:]))
            area = [sum(n - i for i in range(a, b)) for a, b in blocks]
            assert max(area) / min(area) < 1.08
            cuts = [b[0] for b in blocks] + [n]
            same = sharding.rebalance_cuts(n, cuts, [1.0] * world)
            assert all(abs(a - b) <= 64 for a, b in zip(same, cuts))
            times = [1.0] * world
            times[0] = 1.2  # rank 0 slow: its block must shrink, everything stays ordered and covered
            new = sharding.rebalance_cuts(n, cuts, times)
            assert new[0] == 0 and new[-1] == n and all(a <= b for a, b in zip(new, new[1:])) and new[1] < cuts[1]
    assert sharding.rebalance_cuts(100, [0, 100], [1.0]) == [0, 100]
