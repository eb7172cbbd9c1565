"""Full-shape checks of the other BASELINE.json configurations (cfg3 n-C20H42 cc-pVTZ, cfg5 (H2O)40 cc-pVTZ UHF): a
Q slice of the synthetic tensor element-wise against the oracle at the full nbf / nocc / nmat, and the whole tensor on
one GPU against the on-the-fly spot oracle plus the size-independent properties (own module: the C60 fixture of
test_gpu_fullsize.py must have released its 113 GB first)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_hbm_gb():
    import torch

    free, _ = torch.cuda.mem_get_info(0)
    return free / 1e9


# ---- the other BASELINE.json configurations at their full nbf / nocc / nmat -----------------
# cfg3 n-C20H42 cc-pVTZ: nocc 81 -> one 88-column orbital tile whose last column carries the density row
# cfg5 (H2O)40 cc-pVTZ UHF: two densities with different occupied blocks, nbf 2320 (19 tile rows, 16 live in the last)
def _workload_inputs(name):
    from psi4_b200 import workloads

    cfg = workloads.CONFIGS[name]
    n, o, nmat = cfg["nbf"], cfg["nocc"], cfg["nmat"]
    keep = workloads.pair_mask(n, cfg["mask"])
    amp = workloads.amplitude(n)
    Cl = [workloads.orbitals(n, o, workloads.SEED + 10 * i) for i in range(nmat)]
    D = [c @ c.T for c in Cl]
    return cfg, keep, amp, Cl, D


@pytest.mark.parametrize("name", ["c20h42_tz", "h2o40_tz"])
def test_workload_q_slice_engine_matches_oracle_on_slice(name, oracle):
    """Q rows [0,40) of the workload's synthetic tensor at the full nbf / nocc / nmat, element-wise against the oracle."""
    from psi4_b200 import DFHelper, Engine, workloads

    cfg, keep, amp, Cl, D = _workload_inputs(name)
    n, nq = cfg["nbf"], 40
    sp = oracle.Sparsity(keep.astype(np.uint8), nq)
    P = oracle.synth_fill(sp, 0, nq, workloads.SEED, amp)
    d = DFHelper(n, nq)
    d.prepare_sparsity(keep=keep)
    e = Engine(1)
    e.set_layout(n, nq, d.small_skips_, d.big_skips_, d.schwarz_fun_index_)
    e.fill_synthetic(0, workloads.SEED, amp)
    J, K, _ = e.compute(Cl, None, D)
    Jo, Ko, _, _ = oracle.build_JK(sp, P, Cl, D=D)
    for i in range(cfg["nmat"]):
        assert np.abs(J[i] - Jo[i]).max() < 1e-10, f"J[{i}]"
        assert np.abs(K[i] - Ko[i]).max() < 1e-10, f"K[{i}]"
    # general (C_right != C_left) path at the same shape: full K square, two transforms, full-row J
    Cr = [workloads.orbitals(n, cfg["nocc"], workloads.SEED + 77 + i) for i in range(cfg["nmat"])]
    Dg = [a @ b.T for a, b in zip(Cl, Cr)]
    J, K, _ = e.compute(Cl, Cr, Dg)
    Jo, Ko, _, _ = oracle.build_JK(sp, P, Cl, Cr, D=Dg)
    for i in range(cfg["nmat"]):
        assert np.abs(J[i] - Jo[i]).max() < 1e-10, f"general J[{i}]"
        assert np.abs(K[i] - Ko[i]).max() < 1e-10, f"general K[{i}]"
    e.close()


@pytest.mark.parametrize("name", ["c20h42_tz", "h2o40_tz"])
def test_workload_full_size_spot_parity_and_properties(name, oracle):
    """The whole tensor of the workload on one GPU ((H2O)40: 76 GB + 22 GB of work buffers): rows of J and a sample of
    K against the on-the-fly oracle, exact symmetry, bit-exact power-of-two scaling, repeatability."""
    from psi4_b200 import DFHelper, Engine, workloads

    cfg, keep, amp, Cl, D = _workload_inputs(name)
    n, a, nmat = cfg["nbf"], cfg["naux"], cfg["nmat"]
    need_gb = 8.0 * keep.sum() * a / 1e9 * 1.35 + 4
    if _free_hbm_gb() < need_gb:
        pytest.skip(f"needs ~{need_gb:.0f} GB of free HBM")
    d = DFHelper(n, a)
    d.prepare_sparsity(keep=keep)
    e = Engine(1)
    e.set_layout(n, a, d.small_skips_, d.big_skips_, d.schwarz_fun_index_)
    e.fill_synthetic(0, workloads.SEED, amp)
    J, K, _ = e.compute(Cl, None, D)
    keep8 = keep.astype(np.uint8)
    rows = [0, 7, n // 2 + 1, n - 1]
    Bm = {m: oracle.synth_rowblock(keep8, a, workloads.SEED, amp, m) for m in rows}
    for i in range(nmat):
        Dsym = np.triu(D[i]) + np.triu(D[i], 1).T
        dq = oracle.synth_dq(keep8, a, workloads.SEED, amp, Dsym)
        T = {m: Bm[m] @ Cl[i] for m in rows}
        for m in rows:
            jref = Bm[m].T @ dq
            assert np.abs(J[i][m] - jref).max() < 1e-10 * max(1.0, np.abs(jref).max()), f"J[{i}] row {m}"
            for nn in rows:
                kref = float(np.vdot(T[m], T[nn]))
                assert abs(K[i][m, nn] - kref) < 1e-10 * max(1.0, abs(kref)), f"K[{i}][{m},{nn}]"
        assert np.array_equal(K[i], K[i].T) and np.array_equal(J[i], J[i].T)
    J2, K2, _ = e.compute([2.0 * c for c in Cl], None, [4.0 * x for x in D])
    J3, K3, _ = e.compute(Cl, None, D)
    for i in range(nmat):
        assert np.array_equal(J2[i], 4.0 * J[i]) and np.array_equal(K2[i], 4.0 * K[i])  # exact: powers of two
        assert np.array_equal(J3[i], J[i]) and np.array_equal(K3[i], K[i])  # deterministic reductions
    e.close()
