"""world_size-2 gloo test (CPU): the Q-sharded decomposition used at N>1 GPUs.  Each rank builds the
partial J/K of its Q shard (with the oracle standing in for the GPU kernels), the partials are summed
with an all-reduce, and the result must equal the unsharded build."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, lr, out):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import dfjk_oracle as oracle
    from psi4_b200 import DFHelper
    from psi4_b200.sharding import q_range

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(42)  # same inputs on every rank
    n, a = 30, 45
    r = rng.random((n, n))
    keep = (r + r.T) < 1.2
    np.fill_diagonal(keep, True)
    B = rng.standard_normal((a, n, n))
    B = B + B.transpose(0, 2, 1)
    Cl = [rng.standard_normal((n, 6)), rng.standard_normal((n, 3))]
    Cr = None if lr else [rng.standard_normal((n, 6)), rng.standard_normal((n, 3))]
    q0, q1 = q_range(a, rank, world)
    d = DFHelper(n, q1 - q0)
    d.prepare_sparsity(keep=keep)
    sp = oracle.Sparsity(keep, q1 - q0)
    J, K, _, _ = oracle.build_JK(sp, d.pack(B[q0:q1]), Cl, Cr, nthreads=1)
    buf = torch.from_numpy(np.stack(J + K))
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    if rank == 0:
        spf = oracle.Sparsity(keep, a)
        df = DFHelper(n, a)
        df.prepare_sparsity(keep=keep)
        Jf, Kf, _, _ = oracle.build_JK(spf, df.pack(B), Cl, Cr, nthreads=1)
        err = float(np.abs(buf.numpy() - np.stack(Jf + Kf)).max())
        with open(out, "w") as f:
            f.write(repr(err))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("lr", [True, False])
def test_q_sharded_partials_sum_to_full_build(tmp_path, lr):
    out = str(tmp_path / "err.txt")
    port = 29500 + (os.getpid() % 2000) + (1 if lr else 0)
    mp.spawn(_worker, args=(2, port, lr, out), nprocs=2, join=True)
    assert float(open(out).read()) < 1e-10


def test_q_ranges_tile_the_aux_index():
    from psi4_b200.sharding import q_range

    for naux in (1, 7, 116, 4740, 5560):
        for w in (1, 2, 4, 8):
            edges = [q_range(naux, r, w) for r in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == naux
            assert all(edges[i][1] == edges[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1


def _fit_worker(rank, world, port, out):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import dfjk_oracle as oracle
    from psi4_b200 import DFHelper
    from psi4_b200.sharding import q_range

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(11)
    n, a = 26, 31
    r = rng.random((n, n))
    keep = (r + r.T) < 1.3
    np.fill_diagonal(keep, True)
    U = rng.standard_normal((a, n, n))
    U = U + U.transpose(0, 2, 1)
    g = rng.standard_normal((a, a))
    met = g @ g.T / a + np.eye(a)
    d = DFHelper(n, a)
    d.prepare_sparsity(keep=keep)
    # what b200jk_fit_rows does on shard `rank`: the WHOLE unfitted block, this shard's rows of the metric only
    q0, q1 = q_range(a, rank, world)
    mine = np.einsum("QR,Rmn->Qmn", met[q0:q1], U * keep[None])
    rows = [None] * world  # shards are uneven (31 rows over 2 ranks): gather as objects
    dist.all_gather_object(rows, mine)
    if rank == 0:
        sp = oracle.Sparsity(keep.astype(np.uint8), a)
        ref = oracle.contract_metric_AO_core_symm(sp, d.pack_symm(U), met, nthreads=1)
        got = d.pack(np.concatenate(rows))
        with open(out, "w") as f:
            f.write(repr(float(np.abs(got - ref).max())))
    dist.barrier()
    dist.destroy_process_group()


def test_q_sharded_fitting_rows_concatenate_to_the_full_fit(tmp_path):
    """On-device fitting at N > 1 (f2): every shard receives the whole unfitted block and contracts it with ITS rows of the
    metric; the fitted tensor is the concatenation over shards -- no collective on the data path."""
    out = str(tmp_path / "fit.txt")
    mp.spawn(_fit_worker, args=(2, 29500 + (os.getpid() % 2000) + 7, out), nprocs=2, join=True)
    assert float(open(out).read()) < 1e-11
