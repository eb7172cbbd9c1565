"""Host-side molecule / basis-set / integral layer feeding the DF-JK engine.

psi4 takes these from libmints + Libint2; here: a Gaussian94 (.gbs) parser (psi4/share/psi4/basis format,
libmints/basisset.cc:866-884 conventions: coefficients refer to normalised primitives, contracted
functions normalised to unity), cartesian integrals from libb200ints.so (csrc/ints.c) and the
cartesian -> real-solid-harmonic transformation.  SCF energies are invariant to the ordering / phase /
normalisation of functions inside a shell, so only the spanned space has to match the reference.

Units: bohr = 0.52917721067 Angstrom (psi4/include/psi4/physconst.h:402).
"""
from __future__ import annotations

import ctypes as ct
import math
import os
from dataclasses import dataclass, field

import numpy as np

BOHR_TO_ANGSTROM = 0.52917721067
_HERE = os.path.dirname(os.path.abspath(__file__))
BASIS_DIR = os.path.join(_HERE, "share", "basis")
INTS_LIB_PATH = os.path.join(_HERE, "libb200ints.so")
Z_OF = {"H": 1, "HE": 2, "LI": 3, "BE": 4, "B": 5, "C": 6, "N": 7, "O": 8, "F": 9, "NE": 10, "AR": 18}
L_OF = {"S": 0, "P": 1, "D": 2, "F": 3, "G": 4, "H": 5, "I": 6}

_lib = None


def _ints():
    global _lib
    if _lib is None:
        if not os.path.exists(INTS_LIB_PATH):
            raise ImportError(f"{INTS_LIB_PATH} not found: run `python -m psi4_b200.build`")
        _lib = ct.CDLL(INTS_LIB_PATH)
    return _lib


# ------------------------------------------------------------------------------------------------
@dataclass
class Molecule:
    symbols: list
    xyz: np.ndarray  # bohr, (natom, 3)
    charge: int = 0

    @property
    def Z(self):
        return np.array([Z_OF[s.upper()] for s in self.symbols], dtype=np.float64)

    def nelectron(self):
        return int(self.Z.sum()) - self.charge

    def nuclear_repulsion(self) -> float:
        e = 0.0
        for i in range(len(self.symbols)):
            for j in range(i):
                e += self.Z[i] * self.Z[j] / np.linalg.norm(self.xyz[i] - self.xyz[j])
        return float(e)

    @staticmethod
    def from_angstrom(symbols, xyz_angstrom, charge=0):
        return Molecule(list(symbols), np.asarray(xyz_angstrom, dtype=np.float64) / BOHR_TO_ANGSTROM, charge)

    @staticmethod
    def from_zmat_h2o(r_oh=0.96, angle_deg=104.5):
        """The Z-matrix of tests/tu1-h2o-energy/input.dat:  O / H 1 r / H 1 r 2 angle."""
        a = math.radians(angle_deg)
        xyz = [[0.0, 0.0, 0.0], [0.0, 0.0, r_oh], [r_oh * math.sin(a), 0.0, r_oh * math.cos(a)]]
        return Molecule.from_angstrom(["O", "H", "H"], xyz)


@dataclass
class Shell:
    l: int
    exps: np.ndarray
    coefs: np.ndarray  # as in the file (normalised primitives)
    center: np.ndarray
    atom: int


def parse_gbs(name: str) -> tuple[bool, dict]:
    """-> (spherical?, {ELEMENT: [(l, exps, coefs), ...]}).  SP shells are split."""
    path = os.path.join(BASIS_DIR, name.lower() + ".gbs")
    if not os.path.exists(path):
        raise FileNotFoundError(f"basis {name} not in {BASIS_DIR} (tools/extract_basis.py adds elements)")
    lines = [l.split("!")[0].rstrip() for l in open(path)]
    lines = [l for l in lines if l.strip()]
    spherical = lines[0].strip().lower() == "spherical"
    out, i = {}, 1
    while i < len(lines):
        parts = lines[i].split()
        if parts[0] == "****":
            i += 1
            continue
        elem = parts[0].upper()
        i += 1
        shells = []
        while i < len(lines) and lines[i].strip() != "****":
            typ, nprim = lines[i].split()[0].upper(), int(lines[i].split()[1])
            rows = [[float(x.replace("D", "E").replace("d", "e")) for x in lines[i + 1 + k].split()] for k in range(nprim)]
            rows = np.array(rows)
            if typ == "SP":
                shells.append((0, rows[:, 0], rows[:, 1]))
                shells.append((1, rows[:, 0], rows[:, 2]))
            else:
                shells.append((L_OF[typ], rows[:, 0], rows[:, 1]))
            i += 1 + nprim
        out[elem] = shells
    return spherical, out


def basis_shape(mol: "Molecule", name: str, puream: bool | None = None) -> tuple[int, int]:
    """(number of functions, largest shell) of a basis on a molecule without building any integral: all that
    DFHelper::get_core_size (dfhelper.cc:216-236) needs from the auxiliary basis (naux_, Qshell_max_)."""
    spherical, table = parse_gbs(name)
    if puream is not None:
        spherical = puream
    nf = lambda l: 2 * l + 1 if spherical else (l + 1) * (l + 2) // 2  # noqa: E731
    tot, big = 0, 0
    for sym in mol.symbols:
        for l, _, _ in table[sym.upper()]:
            tot += nf(l)
            big = max(big, nf(l))
    return tot, big


def _dfact(n):
    return 1.0 if n <= 0 else n * _dfact(n - 2)


def _cart_list(l):
    return [(lx, ly, l - lx - ly) for lx in range(l, -1, -1) for ly in range(l - lx, -1, -1)]


def solid_harmonic_matrix(l: int) -> np.ndarray:
    """(2l+1, ncart) coefficients of the real solid harmonics S_lm in the cartesian monomials of _cart_list(l)
    (Helgaker, Jorgensen, Olsen eq. 6.4.47-50), order m = 0, +1, -1, ..., +l, -l as psi4 ("gaussian" order)."""
    cart = {c: i for i, c in enumerate(_cart_list(l))}
    rows = []
    for m in [0] + [s * k for k in range(1, l + 1) for s in (1, -1)]:
        am = abs(m)
        row = np.zeros(len(cart))
        vm = 0.0 if m >= 0 else 0.5
        nlm = (1.0 / (2.0 ** am * math.factorial(l))) * math.sqrt(
            2.0 * math.factorial(l + am) * math.factorial(l - am) / (2.0 if m == 0 else 1.0))
        for t in range((l - am) // 2 + 1):
            for u in range(t + 1):
                v = vm
                vmax = math.floor(am / 2.0 - vm) + vm
                while v <= vmax + 1e-9:
                    c = ((-1.0) ** int(round(t + v - vm))) * (0.25 ** t) * math.comb(l, t) * math.comb(l - t, am + t) * \
                        math.comb(t, u) * math.comb(am, int(round(2 * v)))
                    lx = int(round(2 * t + am - 2 * (u + v)))
                    ly = int(round(2 * (u + v)))
                    lz = l - 2 * t - am
                    row[cart[(lx, ly, lz)]] += nlm * c
                    v += 1.0
        rows.append(row)
    return np.array(rows)


@dataclass
class BasisSet:
    """Counterpart of psi4's BasisSet for the C1 AO basis: shells on centres, cartesian bookkeeping for the C
    integral library and the cartesian->function transformation (pure or cartesian, per the .gbs header)."""
    name: str
    shells: list = field(default_factory=list)
    spherical: bool = True

    @staticmethod
    def build(mol: Molecule, name: str, puream: bool | None = None) -> "BasisSet":
        spherical, table = parse_gbs(name)
        if puream is not None:
            spherical = puream
        bs = BasisSet(name, [], spherical)
        bs.mol = mol
        for ia, sym in enumerate(mol.symbols):
            for l, e, c in table[sym.upper()]:
                bs.shells.append(Shell(l, np.array(e), np.array(c), mol.xyz[ia].copy(), ia))
        bs._finalize()
        return bs

    def _finalize(self):
        ns = len(self.shells)
        self.l = np.array([s.l for s in self.shells], dtype=np.int32)
        self.nprim = np.array([len(s.exps) for s in self.shells], dtype=np.int32)
        self.poff = np.concatenate([[0], np.cumsum(self.nprim)[:-1]]).astype(np.int32)
        self.xyz = np.ascontiguousarray(np.array([s.center for s in self.shells], dtype=np.float64).reshape(ns, 3))
        self.exps = np.concatenate([s.exps for s in self.shells]).astype(np.float64)
        coefs = []
        for s in self.shells:
            # coefficient x normalisation of the primitive x^l exp(-a r^2)  (radial convention shared by the shell)
            nrm = (2.0 * s.exps / math.pi) ** 0.75 * (4.0 * s.exps) ** (s.l / 2.0) / math.sqrt(_dfact(2 * s.l - 1))
            coefs.append(s.coefs * nrm)
        self.coefs = np.concatenate(coefs).astype(np.float64)
        self.ncart = int(sum((l + 1) * (l + 2) // 2 for l in self.l))
        # block-diagonal cartesian -> basis-function transformation U (nbf x ncart), rows normalised later
        blocks, self.shell_first_function, nf = [], [], 0
        for s in self.shells:
            self.shell_first_function.append(nf)
            b = solid_harmonic_matrix(s.l) if (self.spherical and s.l >= 2) else np.eye((s.l + 1) * (s.l + 2) // 2)
            blocks.append(b)
            nf += b.shape[0]
        self._nbf = nf
        U = np.zeros((nf, self.ncart))
        r = c = 0
        for b in blocks:
            U[r:r + b.shape[0], c:c + b.shape[1]] = b
            r += b.shape[0]
            c += b.shape[1]
        # normalise every function to unit self-overlap (basisset.cc: normalised contracted functions)
        S = self._one_electron_cart(np.zeros(0), np.zeros((0, 3)))[0]
        d = np.einsum("ic,cd,id->i", U, S, U)
        self.U = U / np.sqrt(d)[:, None]

    def nbf(self) -> int:
        return self._nbf

    def molecule(self) -> "Molecule":
        return self.mol

    def nshell(self) -> int:
        return len(self.shells)

    def shell_nfunction(self, s: int) -> int:
        l = self.shells[s].l
        return 2 * l + 1 if (self.spherical and l >= 2) else (l + 1) * (l + 2) // 2

    # ---- raw C calls -------------------------------------------------------------------------
    def _args(self):
        p = lambda a, t: a.ctypes.data_as(ct.POINTER(t))  # noqa: E731
        return [len(self.shells), p(self.xyz, ct.c_double), p(self.l, ct.c_int), p(self.nprim, ct.c_int),
                p(self.poff, ct.c_int), p(self.exps, ct.c_double), p(self.coefs, ct.c_double)]

    def _one_electron_cart(self, Z, axyz):
        n = self.ncart
        S, T, V = np.zeros((n, n)), np.zeros((n, n)), np.zeros((n, n))
        Z = np.ascontiguousarray(Z, dtype=np.float64)
        axyz = np.ascontiguousarray(axyz, dtype=np.float64)
        dp = ct.POINTER(ct.c_double)
        rc = _ints().ints_one_electron(*self._args(), len(Z), Z.ctypes.data_as(dp), axyz.ctypes.data_as(dp),
                                       S.ctypes.data_as(dp), T.ctypes.data_as(dp), V.ctypes.data_as(dp))
        if rc:
            raise RuntimeError("ints_one_electron: angular momentum above the library limit")
        return S, T, V


class MintsHelper:
    """The slice of psi4's MintsHelper the DF-SCF path uses (ao_overlap / ao_kinetic / ao_potential) plus the
    DF integral producers of DFHelper."""

    def __init__(self, mol: Molecule, primary: BasisSet):
        self.mol, self.primary = mol, primary

    def one_electron(self):
        U = self.primary.U
        S, T, V = self.primary._one_electron_cart(self.mol.Z, self.mol.xyz)
        return U @ S @ U.T, U @ T @ U.T, U @ V @ U.T

    def metric(self, aux: BasisSet) -> np.ndarray:
        """(A|B), fittingmetric.cc:72-155."""
        n = aux.ncart
        out = np.zeros((n, n))
        rc = _ints().ints_two_center(*aux._args(), out.ctypes.data_as(ct.POINTER(ct.c_double)))
        if rc:
            raise RuntimeError("ints_two_center failed")
        return aux.U @ out @ aux.U.T

    def three_center(self, aux: BasisSet, omega: float = 0.0) -> np.ndarray:
        """(A|mn) as a dense (naux, nbf, nbf) array, dfhelper.cc:1284-1347; omega > 0: (A|erf(omega r)/r|mn), the
        integrals of IntegralFactory::erf_eri that fill wPpq_ (dfhelper.cc:596, :683-692)."""
        P = self.primary
        out = np.zeros((aux.ncart, P.ncart, P.ncart))
        if omega > 0.0:
            rc = _ints().ints_three_center_erf(*aux._args(), *P._args(), ct.c_double(omega),
                                               out.ctypes.data_as(ct.POINTER(ct.c_double)))
        else:
            rc = _ints().ints_three_center(*aux._args(), *P._args(), out.ctypes.data_as(ct.POINTER(ct.c_double)))
        if rc:
            raise RuntimeError("ints_three_center failed")
        t = np.einsum("Aa,amn->Amn", aux.U, out, optimize=True)
        t = np.einsum("Mm,Amn->AMn", P.U, t, optimize=True)
        return np.ascontiguousarray(np.einsum("Nn,AMn->AMN", P.U, t, optimize=True))

    def schwarz_function_maxima(self) -> np.ndarray:
        """fun_max_vals[m,n] = |(mn|mn)| per function pair, dfhelper.cc:330-369."""
        P = self.primary
        n = P.nbf()
        out = np.zeros((n, n))
        dp = ct.POINTER(ct.c_double)
        coff = np.concatenate([[0], np.cumsum([(l + 1) * (l + 2) // 2 for l in P.l])]).astype(int)
        for MU in range(P.nshell()):
            fm, nm = P.shell_first_function[MU], P.shell_nfunction(MU)
            Um = P.U[fm:fm + nm, coff[MU]:coff[MU + 1]]
            for NU in range(MU + 1):
                fn, nn = P.shell_first_function[NU], P.shell_nfunction(NU)
                Un = P.U[fn:fn + nn, coff[NU]:coff[NU + 1]]
                blk = np.zeros((Um.shape[1], Un.shape[1], Um.shape[1], Un.shape[1]))
                rc = _ints().ints_pair_diagonal(*P._args(), MU, NU, blk.ctypes.data_as(dp))
                if rc:
                    raise RuntimeError("ints_pair_diagonal failed")
                v = np.einsum("ia,jb,abcd,ic,jd->ij", Um, Un, blk, Um, Un, optimize=True)
                out[fm:fm + nm, fn:fn + nn] = np.abs(v)
                out[fn:fn + nn, fm:fm + nm] = np.abs(v).T
        return out
