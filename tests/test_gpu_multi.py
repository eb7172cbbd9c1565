"""Multi-GPU parity (-m gpu, skipped below 2 devices): both ways of driving N GPUs -- one process
(psi4's situation: b200jk_create(ngpu)) and one process per GPU (torchrun: b200jk_create_rank)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch

    return torch.cuda.device_count()


def _case(rng, oracle, n=90):
    from psi4_b200 import DFHelper

    a = 131
    r = rng.random((n, n))
    keep = (r + r.T) < 1.4
    np.fill_diagonal(keep, True)
    d = DFHelper(n, a)
    d.prepare_sparsity(keep=keep)
    B = rng.standard_normal((a, n, n)) * 0.1
    B = B + B.transpose(0, 2, 1)
    P = d.pack(B)
    Cl = [rng.standard_normal((n, 13)), rng.standard_normal((n, 4))]
    Cr = [rng.standard_normal((n, 13)), rng.standard_normal((n, 4))]
    return d, oracle.Sparsity(keep, a), P, Cl, Cr


@pytest.mark.parametrize("lr", [True, False])
@pytest.mark.parametrize("n", [90, 91])  # 91: odd nbf^2 -> the cross-GPU sum takes its unaligned (8-byte) path
def test_single_process_multi_gpu(oracle, lr, n):
    if _ngpu() < 2:
        pytest.skip("needs >= 2 GPUs")
    from psi4_b200 import Engine

    rng = np.random.default_rng(8)
    d, sp, P, Cl, Cr = _case(rng, oracle, n)
    Cr = None if lr else Cr
    D = [x @ (x if lr else y).T for x, y in zip(Cl, Cl if lr else Cr)]
    Jo, Ko, _, _ = oracle.build_JK(sp, P, Cl, Cr, D=D)
    for ng in [g for g in (2, 4, 8) if g <= _ngpu()]:
        e = Engine(ng)
        e.set_layout(d.nbf_, d.naux_, d.small_skips_, d.big_skips_, d.schwarz_fun_index_)
        e.upload(0, P)
        J, K, _ = e.compute(Cl, Cr, D)
        for i in range(2):
            assert np.abs(J[i] - Jo[i]).max() < 1e-10
            assert np.abs(K[i] - Ko[i]).max() < 1e-10
        assert e.stats()["n_shards"] == ng
        assert e.stats()["reduce_kind"] == 1  # the fixed-order peer-memory sum, not NCCL
        # deterministic: a second build gives the same bits
        J2, K2, _ = e.compute(Cl, Cr, D)
        assert all(np.array_equal(a, b) for a, b in zip(J + K, J2 + K2))
        e.close()


RANK_SCRIPT = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["B2_ROOT"]); sys.path.insert(0, os.path.join(os.environ["B2_ROOT"], "oracle"))
sys.path.insert(0, os.path.join(os.environ["B2_ROOT"], "tests"))
import dfjk_oracle as oracle
from psi4_b200 import Engine
from test_gpu_multi import _case
rank, world, lrk = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lrk)
dist.init_process_group("nccl", device_id=torch.device("cuda", lrk))
idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    idt = torch.tensor(list(Engine.nccl_unique_id()), dtype=torch.uint8, device="cuda")
dist.broadcast(idt, 0)
d, sp, P, Cl, Cr = _case(np.random.default_rng(8), oracle, int(os.environ.get("B2_NBF", "90")))
e = Engine(rank=rank, world=world, device=lrk, nccl_id=bytes(idt.cpu().tolist()))
e.set_layout(d.nbf_, d.naux_, d.small_skips_, d.big_skips_, d.schwarz_fun_index_)
e.upload(0, P)   # every rank holds the host tensor here; the engine copies only its Q shard
J, K, _ = e.compute(Cl, Cr, [x @ y.T for x, y in zip(Cl, Cr)])
Jo, Ko, _, _ = oracle.build_JK(sp, P, Cl, Cr)
err = max(max(np.abs(J[i] - Jo[i]).max(), np.abs(K[i] - Ko[i]).max()) for i in range(2))
st = e.stats()
print(f"RANK{rank} err={err:.3e} q=[{st['q_begin']},{st['q_end']}) allreduce_ms={st['ms_allreduce']:.3f}", flush=True)
assert err < 1e-10
# worker ranks may bring nothing home (b200jk_compute with NULL outputs on rank != 0): rank 0's result is unchanged
J2, K2, _ = e.compute(Cl, Cr, [x @ y.T for x, y in zip(Cl, Cr)], fetch=(rank == 0))
if rank == 0:
    assert all(np.array_equal(a, b) for a, b in zip(J + K, J2 + K2))
else:
    assert J2 is None and K2 is None
print(f"RANK{rank} fetch_ok", flush=True)
want_kind = 2 if os.environ.get("B200JK_REDUCE") == "nccl" else 1
assert st["reduce_kind"] == want_kind, st["reduce_kind"]
# determinism (SURVEY.md 8e): run to run, and the device-operand entry point against the host-operand one
n = d.nbf_
dC = [e.dev_put(x) for x in Cl]; dCr = [e.dev_put(x) for x in Cr]; dD = [e.dev_put(x @ y.T) for x, y in zip(Cl, Cr)]
dJ = [e.dev_alloc(n * n * 8) for _ in Cl]; dK = [e.dev_alloc(n * n * 8) for _ in Cl]
for rep in range(3):
    e.compute_device(dC, dCr, [x.shape[1] for x in Cl], dD, dJ, dK, None)
    got = [e.dev_get(p, (n, n)) for p in dJ + dK]
    t = torch.tensor([float(np.frombuffer(g.tobytes(), dtype=np.uint64).sum() % (1 << 52)) for g in got], dtype=torch.float64, device="cuda")
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert torch.equal(lo, hi), "ranks hold different bits"
    if rank == 0:
        assert all(np.array_equal(a, b) for a, b in zip(J + K, got)), f"device arm differs from host arm (rep {rep})"
print(f"RANK{rank} determinism_ok kind={st['reduce_kind']}", flush=True)
e.close()
dist.barrier(); dist.destroy_process_group()
"""


@pytest.mark.parametrize("nproc", [2, 4, 8])
@pytest.mark.parametrize("mode", ["p2p", "p2p_odd", "nccl"])
def test_one_process_per_gpu_torchrun(tmp_path, nproc, mode):
    """Parity against the oracle, worker ranks that bring nothing home, and bitwise determinism (run to run, rank to
    rank, device-operand arm against host-operand arm) of the cross-GPU sum at 2 / 4 / 8 ranks."""
    if _ngpu() < nproc:
        pytest.skip(f"needs >= {nproc} GPUs")
    script = tmp_path / "rank.py"
    script.write_text(RANK_SCRIPT)
    env = dict(os.environ, B2_ROOT=ROOT, B2_NBF="91" if mode == "p2p_odd" else "90")
    if mode == "nccl":
        env["B200JK_REDUCE"] = "nccl"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(29611 + nproc), str(script)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    for k in range(nproc):
        assert f"RANK{k} err=" in r.stdout and f"RANK{k} fetch_ok" in r.stdout and f"RANK{k} determinism_ok" in r.stdout


LEGACY_SCRIPT = r"""
import os, sys
import numpy as np
sys.path.insert(0, os.environ["B2_ROOT"]); sys.path.insert(0, os.path.join(os.environ["B2_ROOT"], "oracle"))
import dfjk_oracle as oracle
from psi4_b200 import DFHelper, Engine
rng = np.random.default_rng(3)
n, a, o = 140, 150, 21
r = rng.random((n, n)); keep = (r + r.T) < 1.3; np.fill_diagonal(keep, True)
d = DFHelper(n, a); d.prepare_sparsity(keep=keep)
B = rng.standard_normal((a, n, n)) * 0.1; B = B + B.transpose(0, 2, 1)
P = d.pack(B); C = rng.standard_normal((n, o))
e = Engine(1); e.set_layout(n, a, d.small_skips_, d.big_skips_, d.schwarz_fun_index_); e.upload(0, P)
J, K, _ = e.compute([C], None, [C @ C.T])
Jo, Ko, _, _ = oracle.build_JK(oracle.Sparsity(keep, a), P, [C])
err = max(np.abs(J[0] - Jo[0]).max(), np.abs(K[0] - Ko[0]).max())
print(f"LEGACY={os.environ.get('B200JK_LEGACY')} NOFUSE={os.environ.get('B200JK_NO_JFUSE')} "
      f"CGATHER={os.environ.get('B200JK_CGATHER')} err={err:.3e}")
assert err < 1e-10
"""


@pytest.mark.parametrize("env", [{"B200JK_LEGACY": "1"}, {"B200JK_NO_JFUSE": "1"}, {"B200JK_CGATHER": "1"},
                                 {"B200JK_CGATHER": "1", "B200JK_NO_JFUSE": "1"},
                                 {"B200JK_KGEMM": "i8"}, {"B200JK_KGEMM": "dmma", "B200JK_HALF": "i8"},
                                 {"B200JK_KGEMM": "i8", "B200JK_HALF": "i8", "B200JK_I8_CLUSTER": "4"},
                                 {"B200JK_HALF": "i8", "B200JK_I8_CLUSTER": "1", "B200JK_NO_JFUSE": "1",
                                  "B200JK_I8_HALF_MODULI": "13", "B200JK_I8_MODULI": "12"},
                                 {"B200JK_HALF": "i8", "B200JK_I8_JCOL": "0"}])
def test_ab_switches_still_correct(tmp_path, env):
    """The A/B switches documented in DESIGN.md (first-generation kernels, unfused J, pre-gathered C^T forced on for a
    small screened system, with and without the density row riding in it, the INT8-tensor-core arms of the two GEMMs forced
    from the environment with 1 / 2 / 4 CTAs per cluster, the first J sweep as its own kernel instead of a column of the
    residue GEMM) stay parity-green."""
    script = tmp_path / "ab.py"
    script.write_text(LEGACY_SCRIPT)
    r = subprocess.run([sys.executable, str(script)], env=dict(os.environ, B2_ROOT=ROOT, **env), capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]


def test_multi_shard_fitting_matches_oracle(oracle):
    """b200jk_fit_rows on a 2-shard handle: each Q shard contracts with its own metric rows."""
    if _ngpu() < 2:
        pytest.skip("needs >= 2 GPUs")
    from psi4_b200 import DFHelper, Engine

    rng = np.random.default_rng(12)
    n, a = 70, 91
    r = rng.random((n, n))
    keep = (r + r.T) < 1.5
    np.fill_diagonal(keep, True)
    d = DFHelper(n, a)
    d.prepare_sparsity(keep=keep)
    U = rng.standard_normal((a, n, n))
    U = U + U.transpose(0, 2, 1)
    g = rng.standard_normal((a, a))
    met = g @ g.T / a + np.eye(a)
    sp = oracle.Sparsity(keep.astype(np.uint8), a)
    ref = oracle.contract_metric_AO_core_symm(sp, d.pack_symm(U), met)
    e = Engine(2)
    e.set_layout(n, a, d.small_skips_, d.big_skips_, d.schwarz_fun_index_)
    e.set_metric(met)
    for m0 in range(0, n, 16):
        m1 = min(n, m0 + 16)
        e.fit_rows(0, m0, m1, d.pack_symm(U, m0, m1))
    for m in (0, 33, n - 1):
        got = e.download_rows(0, m, 0, a).ravel()
        want = ref[int(d.big_skips_[m]):int(d.big_skips_[m + 1])]
        assert np.abs(got - want).max() < 1e-12 * max(1.0, np.abs(ref).max())
    C = rng.standard_normal((n, 5))
    J, K, _ = e.compute([C], None, [C @ C.T])
    Jo, Ko, _, _ = oracle.build_JK(sp, ref, [C])
    assert np.abs(J[0] - Jo[0]).max() < 1e-10 * max(1.0, np.abs(Jo[0]).max())
    assert np.abs(K[0] - Ko[0]).max() < 1e-10 * max(1.0, np.abs(Ko[0]).max())
    e.close()
