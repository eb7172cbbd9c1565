"""DF-RHF on the B200 JK engine, printed like a psi4 SCF iteration table.

    python tools/run_scf.py            # water / cc-pVDZ (tests/tu1-h2o-energy of the reference)
    python tools/run_scf.py decane     # n-decane / def2-SVP (tests/dlpnocc-4), 29.6 % screened
    python tools/run_scf.py bz2 --fit-on-device --gpus 2
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from psi4_b200 import scf  # noqa: E402
from psi4_b200.integrals import BasisSet, Molecule  # noqa: E402

ANCH = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_anchors.json")))
ap = argparse.ArgumentParser()
ap.add_argument("system", nargs="?", default="h2o", choices=["h2o", "decane", "bz2"])
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--fit-on-device", action="store_true")
args = ap.parse_args()

if args.system == "h2o":
    a = ANCH["tu1_h2o_ccpvdz"]
    mol = Molecule.from_zmat_h2o(a["zmat"]["r_oh_angstrom"], a["zmat"]["angle_deg"])
    ref = a["scf_total_energy"]
elif args.system == "decane":
    a = ANCH["dlpnocc4_decane_def2svp"]
    geo = a["geometry_angstrom_input"]
    mol = Molecule.from_angstrom([g[0] for g in geo], [g[1:] for g in geo])
    ref = a["scf_total_energy"]
else:
    a = ANCH["dfscf_bz2_ccpvdz"]
    geo = a["geometry_angstrom_output_ref"]
    mol = Molecule.from_angstrom([g[0] for g in geo], [g[1:] for g in geo])
    ref = a["scf_total_energy"]

t0 = time.perf_counter()
P, A = BasisSet.build(mol, a["basis"]), BasisSet.build(mol, a["aux"])
jk = scf.build_jk(mol, P, A, ngpu=args.gpus, fit_on_device=args.fit_on_device)
jk.initialize()
print(jk.print_header())
print(f"  basis functions {P.nbf()}, auxiliary {A.nbf()}, nuclear repulsion {mol.nuclear_repulsion():.12f}")
print(f"  integrals + tensor on device: {time.perf_counter() - t0:.2f} s\n")
rhf = scf.RHF(mol, P, jk)
t0 = time.perf_counter()
E = rhf.compute_energy()
dt = time.perf_counter() - t0
print("                           Total Energy        Delta E     RMS |[F,P]|")
for i, (e, de, dr) in enumerate(rhf.iterations):
    print(f"   @DF-RHF iter {i:3d}:  {e:20.14f}   {de:12.5e}   {dr:11.5e}")
st = jk.stats()
print(f"\n  SCF energy {E:.12f}   reference {ref:.12f}   diff {E - ref:+.2e}")
print(f"  {len(rhf.iterations)} iterations in {dt:.2f} s; last JK build {st['ms_total']:.3f} ms on {st['n_shards']} GPU(s)")
