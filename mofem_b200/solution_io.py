"""``solution.txt`` -- the reference's result file (outputFunctions/writeSolution.py:8-20) and the wire format of its
shipped golden vector.  Writer, reader and a comparison helper."""
from __future__ import annotations

import re

import numpy as np

_HEADER = "                      DISPLACEMENT\n\n  Node No.           x-displacement                   y-displacement\n"


def format_solution(displacement, energy=None, input_file_name=None) -> str:
    """Exactly the text ``writeDisplacement(displacement, energy, inputFileName)`` writes."""
    d = np.asarray(displacement, dtype=np.float64).reshape(-1)
    out = []
    if input_file_name is not None:
        out.append("The input file name is %s.txt.\n\n" % input_file_name)
    out.append(_HEADER)
    for i in range(d.size // 2):
        out.append("%6d           %22.15e           %22.15e\n" % (i, d[2 * i], d[2 * i + 1]))
    out.append("\n")
    if energy is not None:
        out.append("The strain ENERGY is: %.15e." % float(np.asarray(energy).reshape(-1)[0]))
    return "".join(out)


def write_solution(path, displacement, energy=None, input_file_name=None):
    with open(path, "w") as f:
        f.write(format_solution(displacement, energy, input_file_name))


def parse_solution(text: str):
    """-> (displacement (2N,1) float64, energy (1,1) or None, input file stem or None)."""
    name = None
    m = re.search(r"The input file name is (.*)\.txt\.", text)
    if m:
        name = m.group(1)
    rows = re.findall(r"^\s*(\d+)\s+([-+0-9.eE]+)\s+([-+0-9.eE]+)\s*$", text, flags=re.M)
    if not rows:
        raise ValueError("no displacement rows found")
    idx = np.array([int(r[0]) for r in rows])
    if not np.array_equal(idx, np.arange(len(rows))):
        raise ValueError("node numbers are not 0..N-1 in order")
    d = np.array([[float(r[1]), float(r[2])] for r in rows]).reshape(-1, 1)
    m = re.search(r"The strain ENERGY is: ([-+0-9.eE]+?)\.?\s*$", text.strip())
    energy = np.array([[float(m.group(1))]]) if m else None
    return d, energy, name


def read_solution(path):
    with open(path) as f:
        return parse_solution(f.read())


def compare(displacement, energy, ref_displacement, ref_energy=None):
    """Relative inf-norm displacement error and relative energy error against a reference solution."""
    a = np.asarray(displacement).reshape(-1)
    b = np.asarray(ref_displacement).reshape(-1)
    if a.shape != b.shape:
        raise ValueError("different number of DOFs")
    du = float(np.abs(a - b).max() / np.abs(b).max())
    de = None
    if energy is not None and ref_energy is not None:
        e, er = float(np.asarray(energy).reshape(-1)[0]), float(np.asarray(ref_energy).reshape(-1)[0])
        de = abs(e - er) / abs(er)
    return du, de
