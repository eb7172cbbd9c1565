"""Copy the element blocks the tests need out of the reference's basis library
(/root/reference/psi4/share/psi4/basis/*.gbs, Gaussian94 format, published EMSL/BSE data)
into psi4_b200/share/basis/ so nothing reads /root/reference at run time."""
import os
import sys

SRC = "/root/reference/psi4/share/psi4/basis"
DST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "psi4_b200", "share", "basis")
WANT = {
    "cc-pvdz": ["H", "C", "N", "O", "Ar"],
    "cc-pvdz-jkfit": ["H", "C", "N", "O", "Ar"],
    "cc-pv5z": ["Ar"],
    "cc-pv5z-jkfit": ["Ar"],
    "cc-pvtz": ["H", "C", "O"],
    "cc-pvtz-jkfit": ["H", "C", "O"],
    "aug-cc-pvdz": ["H", "C", "O"],
    "aug-cc-pvdz-jkfit": ["H", "C", "O"],
    "def2-svp": ["H", "C"],
    "def2-universal-jkfit": ["H", "C", "O"],
    "sto-3g": ["H", "O"],
}


def extract(name, elements):
    lines = open(os.path.join(SRC, name + ".gbs")).read().splitlines()
    head = next(l for l in lines if l.strip() in ("spherical", "cartesian"))
    out = [head, "", f"! subset of {name}.gbs ({', '.join(elements)}) extracted by tools/extract_basis.py", "****"]
    i = 0
    while i < len(lines):
        parts = lines[i].split()
        if len(parts) == 2 and parts[1] == "0" and parts[0].capitalize() in elements:
            j = i
            while lines[j].strip() != "****":
                j += 1
            out += lines[i:j + 1]
            i = j
        i += 1
    open(os.path.join(DST, name + ".gbs"), "w").write("\n".join(out) + "\n")


if __name__ == "__main__":
    os.makedirs(DST, exist_ok=True)
    for k, v in WANT.items():
        extract(k, v)
        print(k, os.path.getsize(os.path.join(DST, k + ".gbs")))
