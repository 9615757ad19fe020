"""N>1 logic of the batch-sharded path on CPU with gloo (world_size 2): shard bounds, parameter sharding, the
single all-gather of the final reconstructions, and the two whole-batch reductions (early-stop mean, CG all-converged).
The per-rank 'reconstruction' is the oracle's PnP-PGD so that the test needs no GPU; the sharded result must equal the
single-process result exactly (every sample is independent)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world_size, port, n_total, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    import sys
    from pathlib import Path

    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    from deepinv_b200 import sharding as S
    from oracle import ref_ops as R

    torch.manual_seed(0)
    H = W = 16
    x = torch.randn(n_total, 2, H, W)
    mask = (torch.rand(n_total, 1, 1, W) > 0.5).float().expand(n_total, 2, H, W).contiguous()
    shared = (torch.rand(1, 2, H, W) > 0.5).float()
    y = R.mri_A(x, mask)
    den = lambda v, s: v * 0.9

    class Algo:
        def __call__(self, ys, phys):
            m = phys
            return R.pgd(ys, lambda v: R.mri_A(v, m), lambda v: R.mri_At(v, m), den, 1.0, 0.05, 3)

    full = Algo()(y, mask)
    got = S.reconstruct_sharded(Algo(), y, lambda lo, hi: mask[lo:hi], n_total=n_total)
    assert torch.equal(got, full), "sharded reconstruction differs from the single-process one"
    assert S.shard_batch(shared).shape[0] == 1 and S.shard_batch(mask).shape[0] == S.shard_bounds(n_total, rank, world_size)[1] - S.shard_bounds(n_total, rank, world_size)[0]
    lo, hi = S.shard_bounds(n_total, rank, world_size)
    local_mean = full[lo:hi].flatten(1).norm(dim=1).mean()
    gm = S.allreduce_mean(local_mean.clone(), hi - lo, n_total)
    assert torch.allclose(gm, full.flatten(1).norm(dim=1).mean(), atol=1e-6)
    flag = torch.tensor([1 if rank == 0 else 0])
    assert int(S.allreduce_all(flag)) == 0
    assert int(S.allreduce_all(torch.tensor([1]))) == 1
    if rank == 0:
        torch.save(got, out_path)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [6, 5])
def test_sharded_pgd_matches_single_process(tmp_path, n_total):
    port = _free_port()
    out = tmp_path / "x.pt"
    mp.spawn(_worker, args=(2, port, n_total, str(out)), nprocs=2, join=True)
    assert out.exists() and torch.load(out).shape[0] == n_total


def test_shard_bounds_cover_everything():
    from deepinv_b200 import sharding as S

    for n in (1, 7, 64, 257):
        for w in (1, 2, 3, 8):
            pieces = [S.shard_bounds(n, r, w) for r in range(w)]
            assert pieces[0][0] == 0 and pieces[-1][1] == n
            assert all(pieces[i][1] == pieces[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in pieces) - min(h - l for l, h in pieces) <= 1
