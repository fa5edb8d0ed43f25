"""femMain.py of the UNMODIFIED reference, end to end, with the drop-in: the reference's parser, both
plane sweeps, triangulation, writeDisplacement and writeBoundary run as shipped (baseline/_ref/MOFEM, staged by
baseline/stage_ref.py); the two solver modules femMain.py imports are mofem_b200.solvers and writeSampleData2D is
mofem_b200.sampler's.  Output files are compared with the files the reference ships (femMain.py:77-112)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import REPO
from mofem_b200 import solution_io

pytestmark = pytest.mark.gpu
TREE = os.path.join(REPO, "baseline", "_ref", "MOFEM")
RUNNER = os.path.join(REPO, "baseline", "ref_runner.py")


def _run(*argv):
    r = subprocess.run([sys.executable, RUNNER] + list(argv), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


@pytest.mark.skipif(not os.path.isdir(TREE), reason="baseline/_ref/MOFEM is staged in the build container only")
def test_femmain_bracket_with_the_drop_in(tmp_path):
    out = _run("femmain", "AMOREBracket", "b200", str(tmp_path))
    assert {"solution.txt", "SampleData.txt", "nSampling.txt", "boundaryCoord.txt", "boundaryNumber.txt"} <= set(out["files"])
    assert set(out["phases"]) == {"assembly_s", "force_s", "constraints_s", "solve_s"}        # the drop-in prints the reference's phase lines
    # solution.txt against the shipped golden file (1111 nodes to 16 digits + strain energy 1.061718786130466)
    u, e, stem = solution_io.read_solution(str(tmp_path / "solution.txt"))
    u_ref, e_ref, stem_ref = solution_io.read_solution(os.path.join(TREE, "solution.txt"))
    assert u.shape == u_ref.shape == (2222, 1) and stem == stem_ref == "AMOREBracket"
    du, de = solution_io.compare(u, e, u_ref, e_ref)
    assert du <= 1e-9 and de <= 1e-9 and abs(float(e_ref[0, 0]) - 1.061718786130466) < 1e-15
    # SampleData.txt (mofem_b200.sampler on the GPU) against the shipped file: same layout, values to 1e-8 of each component's range
    got, ref = np.loadtxt(tmp_path / "SampleData.txt"), np.loadtxt(os.path.join(TREE, "SampleData.txt"))
    assert got.shape == ref.shape == (57834, 3)
    bg, br = got.reshape(-1, 6, 3, 3, 3), ref.reshape(-1, 6, 3, 3, 3)
    assert np.abs(bg[:, :, :, :2] - br[:, :, :, :2]).max() <= 1e-12 * np.abs(br[:, :, :, :2]).max()
    for comp in range(6):
        assert np.abs(bg[:, comp, :, 2] - br[:, comp, :, 2]).max() <= 1e-8 * np.abs(br[:, comp, :, 2]).max(), comp
    assert (tmp_path / "nSampling.txt").read_text() == open(os.path.join(TREE, "nSampling.txt")).read()
    # the boundary polylines come from the reference's own code: identical to the shipped files up to the sweep's round-off
    assert np.abs(np.loadtxt(tmp_path / "boundaryCoord.txt") - np.loadtxt(os.path.join(TREE, "boundaryCoord.txt"))).max() < 1e-12


@pytest.mark.skipif(not os.path.isdir(TREE), reason="baseline/_ref/MOFEM is staged in the build container only")
def test_entry_point_on_live_reference_objects(tmp_path):
    # the drop-in and the reference entry point on the SAME live objects (parser + sweep of the staged reference), simpleOFE3
    g = _run("entry", "simpleOFE3", "b200", str(tmp_path / "g.npz"))
    r = _run("entry", "simpleOFE3", "ref", str(tmp_path / "r.npz"))
    ug, ur = np.load(tmp_path / "g.npz"), np.load(tmp_path / "r.npz")
    assert np.abs(ug["displacement"] - ur["displacement"]).max() <= 1e-9 * np.abs(ur["displacement"]).max()
    assert abs(ug["energy"][0] - ur["energy"][0]) <= 1e-9 * ur["energy"][0] and abs(ug["energy"][0] - 19.2) < 19.2e-8
    assert g["cells"] == r["cells"] == 21 and g["runs"][0]["gpu_launches"] > 0
