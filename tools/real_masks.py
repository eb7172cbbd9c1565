"""Real Schwarz pair masks (DFHelper::prepare_sparsity, cutoff 1e-12) for the geometries behind the BASELINE.json
configurations, from the host integral front end -- calibrates the synthetic masks of psi4_b200/workloads.py.

    python tools/real_masks.py c60 | c20h42 | h2o40 [--threads 8]

Geometries: C60 truncated icosahedron with 1.43 A edges; all-trans n-C20H42 (CC 1.53 A, CH 1.09 A,
109.5 deg); (H2O)40 on a jittered 4x5x2 lattice with O-O 2.8 A (seed 20251017)."""
import argparse
import itertools
import json
import math
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as ct  # noqa: E402

import numpy as np  # noqa: E402

from psi4_b200.integrals import BasisSet, Molecule, _ints  # noqa: E402


def c60():
    phi = (1 + math.sqrt(5)) / 2
    pts = set()
    for base in [(0, 1, 3 * phi), (1, 2 + phi, 2 * phi), (phi, 2, 2 * phi + 1)]:
        for signs in itertools.product([1, -1], repeat=3):
            v = [s * b for s, b in zip(signs, base)]
            for k in range(3):  # even (cyclic) permutations
                pts.add(tuple(round(x, 9) for x in (v[k % 3], v[(k + 1) % 3], v[(k + 2) % 3])))
    xyz = np.array(sorted(pts)) * (1.43 / 2.0)  # edge length 2 in these coordinates
    assert len(xyz) == 60
    return Molecule.from_angstrom(["C"] * 60, xyz)


def alkane(nc=20):
    cc, ch, th = 1.53, 1.09, math.radians(109.5)
    dx, dz = cc * math.sin(th / 2), cc * math.cos(th / 2)
    sym, xyz = [], []
    for i in range(nc):
        c = np.array([i * dx, 0.0, (i % 2) * dz])
        sym.append("C")
        xyz.append(c)
        up = 1.0 if i % 2 else -1.0
        hy, hz = ch * math.sin(th / 2), ch * math.cos(th / 2) * up
        for sgn in (1, -1):
            sym.append("H")
            xyz.append(c + np.array([0.0, sgn * hy, hz]))
    for end, i in ((-1, 0), (1, nc - 1)):  # terminal hydrogens along the chain
        c = xyz[[k for k, s in enumerate(sym) if s == "C"][i]]
        sym.append("H")
        xyz.append(c + np.array([end * ch * math.sin(th / 2), 0.0, -ch * math.cos(th / 2) * (1.0 if i % 2 else -1.0)]))
    return Molecule.from_angstrom(sym, np.array(xyz))


def water_cluster(n=40, seed=20251017):
    rng = np.random.default_rng(seed)
    sym, xyz = [], []
    r, a = 0.9572, math.radians(104.52)
    for ix, iy, iz in itertools.product(range(4), range(5), range(2)):
        o = np.array([ix, iy, iz]) * 2.8 + rng.normal(0, 0.15, 3)
        q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
        h1 = o + q @ np.array([r, 0, 0])
        h2 = o + q @ np.array([r * math.cos(a), r * math.sin(a), 0])
        sym += ["O", "H", "H"]
        xyz += [o, h1, h2]
    return Molecule.from_angstrom(sym[: 3 * n], np.array(xyz)[: 3 * n])


def schwarz(P: BasisSet, threads: int) -> np.ndarray:
    n = P.nbf()
    out = np.zeros((n, n))
    dp = ct.POINTER(ct.c_double)
    coff = np.concatenate([[0], np.cumsum([(l + 1) * (l + 2) // 2 for l in P.l])]).astype(int)
    args = P._args()
    lib = _ints()

    def work(MU):
        fm, nm = P.shell_first_function[MU], P.shell_nfunction(MU)
        Um = P.U[fm:fm + nm, coff[MU]:coff[MU + 1]]
        for NU in range(MU + 1):
            fn, nn = P.shell_first_function[NU], P.shell_nfunction(NU)
            Un = P.U[fn:fn + nn, coff[NU]:coff[NU + 1]]
            blk = np.zeros((Um.shape[1], Un.shape[1], Um.shape[1], Un.shape[1]))
            lib.ints_pair_diagonal(*args, MU, NU, blk.ctypes.data_as(dp))
            v = np.abs(np.einsum("ia,jb,abcd,ic,jd->ij", Um, Un, blk, Um, Un, optimize=True))
            out[fm:fm + nm, fn:fn + nn] = v
            out[fn:fn + nn, fm:fm + nm] = v.T

    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(work, range(P.nshell())))
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("system", choices=["c60", "c20h42", "h2o40"])
    ap.add_argument("--basis", default="cc-pvtz")
    ap.add_argument("--threads", type=int, default=os.cpu_count())
    ap.add_argument("--save", default=None)
    a = ap.parse_args()
    mol = {"c60": c60, "c20h42": alkane, "h2o40": water_cluster}[a.system]()
    P = BasisSet.build(mol, a.basis)
    t0 = time.time()
    f = schwarz(P, a.threads)
    tol = 1e-12 ** 2 / f.max()
    keep = f >= tol
    res = {"system": a.system, "basis": a.basis, "natom": len(mol.symbols), "nbf": P.nbf(), "cutoff": 1e-12,
           "kept_pairs": int(keep.sum()), "mask_sparsity_percent": 100.0 * (1.0 - keep.sum() / keep.size),
           "min_kept_per_row": int(keep.sum(1).min()), "max_kept_per_row": int(keep.sum(1).max()), "seconds": time.time() - t0}
    print(json.dumps(res))
    if a.save:
        np.save(a.save, np.packbits(keep))
