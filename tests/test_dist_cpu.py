"""world_size-2 gloo tests of the multi-rank plumbing (runs on CPU).

The CUDA kernels cannot run here, so each rank stands in for its shard's kernel output with exact
integer arithmetic (numpy) for the cross-product and the oracle for the scan; what is under test is
the sharding, the integer all-reduce, the epilogue from all-reduced sums and the rank-ordered gather.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gapit_b200 import shard, synth
    from oracle import gapit_oracle as go
    n, m = 40, 333
    G = synth.make_genotypes(n, m, seed=17, threads=1)      # every rank generates the same matrix ...
    G[5, :] = 2
    lo, hi = shard.marker_shard(m, rank, world)              # ... and owns one column block of it
    x = G[lo:hi].astype(np.int64)
    gram = torch.from_numpy(x.T @ x)                         # partial cross-product of the shard
    c = x.sum(1)
    sums = torch.tensor([int(c.sum()), int((c * c).sum()), int((c == 2 * n).sum())], dtype=torch.int64)
    dist.all_reduce(gram)
    dist.all_reduce(sums)
    K = shard.vanraden_from_sums(gram.numpy(), sums.numpy(), n)
    K_ref = go.kinship_vanraden(G.T)
    assert np.abs(K - K_ref).max() < 1e-12
    # scan: every rank scans its shard with the same eigen-system, rank 0 gathers in rank order
    Y = synth.make_phenotypes(G, seed=17)[:, 0]
    X0 = np.ones((n, 1))
    eL = go.emma_eigen_L_wo_Z(K_ref)
    reml = go.emma_REMLE(Y, X0, K_ref)
    part = go.emmax_p3d(Y, G[lo:hi].T, K_ref, X0, eig_L=eL, reml=reml)
    ps = shard.gather_markers(np.nan_to_num(part.ps, nan=-1.0), m, rank, world, dst=0)
    if rank == 0:
        assert ps.shape == (m,)
        full = go.emmax_p3d(Y, G.T, K_ref, X0, eig_L=eL, reml=reml)
        ref = np.nan_to_num(full.ps, nan=-1.0)
        # a duplicated marker that straddles a shard boundary is recomputed, not copied: same value
        assert np.allclose(ps, ref, rtol=1e-12, atol=0)
        open(os.path.join(tmp, "ok"), "w").write("ok")
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_shard_allreduce_gather(tmp_path):
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()
