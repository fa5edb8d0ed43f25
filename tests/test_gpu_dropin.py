"""The drop-in entry points (mofem_b200/solvers.py) called exactly like the reference's linearSolvers/AMORE.py and
traditionalElement.py -- reference argument structure in, reference result structure out -- against the golden
results of the unmodified reference (P3/P4 of SURVEY.md section 8)."""
import contextlib
import io
import pickle

import numpy as np
import pytest

from conftest import load_golden
from graph_fixture import install_bracket_load, load_graph
from mofem_b200 import solvers

pytestmark = pytest.mark.gpu

TOL = 1e-9


def _check(out, d, n_out):
    assert len(out) == n_out
    displacement, energy = out[0], out[1]
    n = len(d["displacement"])
    assert isinstance(displacement, np.ndarray) and displacement.shape == (n, 1) and displacement.dtype == np.float64
    assert isinstance(energy, np.ndarray) and energy.shape == (1, 1)
    ur = d["displacement"]
    assert np.abs(displacement[:, 0] - ur).max() <= TOL * np.abs(ur).max()
    assert abs(energy[0][0] - d["energy"][0]) <= TOL * abs(d["energy"][0])
    _ = displacement[2 * 3][0], energy[0][0]          # the writer's access pattern (writeSolution.py:16,20)


@pytest.mark.parametrize("name,fn,n_out", [("simpleOFE3", "quadraticAMORE", 3), ("AMOREBracket", "quadraticICMAMORE", 3),
                                           ("patch16", "lowerOrderAMORE", 2)])
def test_amore_entry_points(name, fn, n_out):
    _, d = load_golden(name)
    input_data, overlapping_mesh, polygons, node_elements_partial = load_graph(d)
    if name == "AMOREBracket":
        install_bracket_load()
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        out = getattr(solvers, fn)(input_data, overlapping_mesh, polygons, node_elements_partial)
    _check(out, d, n_out)
    log = buf.getvalue()
    for phrase in ("Assembling stiffness matrix costs", "Calculating force term costs", "Imposing constraints costs",
                   "Solving the linear system costs"):
        assert phrase in log
    if n_out == 3:
        assert all(m.shape == (3, 3) for m in out[2])
        assert "Number of equations = %d." % len(d["red_rhs"]) in log
        assert "Number of entries = %d." % len(d["red_data"]) in log
    if name == "AMOREBracket":
        # the reference's shipped golden output (solution.txt: 1111 x 2 displacements + strain energy)
        import os
        from conftest import GOLDEN
        from mofem_b200 import solution_io
        g = dict(np.load(os.path.join(GOLDEN, "shipped_solution_AMOREBracket.npz")))
        du, de = solution_io.compare(out[0], out[1], g["displacement"], g["energy"])
        assert du <= TOL and de <= TOL
        assert abs(out[1][0][0] - 1.061718786130466e+00) <= TOL * 1.061718786130466e+00
    if name == "simpleOFE3":
        assert abs(out[1][0][0] - 19.2) <= 1e-8 * 19.2          # analytic patch-test energy


@pytest.mark.parametrize("kind,fn", [("lower", "lowerOrderFE"), ("icm", "ICMFE"), ("quad", "quadraticFE")])
def test_single_mesh_entry_points(kind, fn):
    _, d = load_golden("single_" + kind)
    deck = pickle.loads(d["deck"].tobytes())
    with contextlib.redirect_stdout(io.StringIO()):
        out = getattr(solvers, fn)(deck)
    _check(out, d, 2)


def test_wrong_element_type_raises_like_the_reference():
    _, d = load_golden("single_quad")
    deck = pickle.loads(d["deck"].tobytes())
    with pytest.raises(ValueError, match="Wrong element numbering"):
        solvers.lowerOrderFE(deck)
