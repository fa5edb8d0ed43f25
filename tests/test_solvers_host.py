"""Host glue of the drop-in entry points (mofem_b200/solvers.py) against the reference: the force vector and the
constraint handling, on the reference's own argument structure rebuilt from the golden fixtures.  CPU only."""
import pickle

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import load_golden
from graph_fixture import install_bracket_load, load_graph
from mofem_b200 import flatten as fl
from mofem_b200 import solvers


def _reduced_rhs(d, f, fixed, value):
    n = len(f)
    K = sp.csr_matrix((d["csr_data"], d["csr_indices"], d["csr_indptr"]), shape=(n, n))
    b = f.reshape(-1, 1) - K.dot(value.reshape(-1, 1))          # AMORE.py:227
    return b.reshape(-1)[fixed == 0]


@pytest.mark.parametrize("name,quadratic", [("simpleOFE3", True), ("AMOREBracket", True), ("patch16", False)])
def test_force_vector_and_constraints_match_the_reference(name, quadratic):
    _, d = load_golden(name)
    input_data, om, polygons, nep = load_graph(d)
    if name == "AMOREBracket":
        install_bracket_load()
    f = solvers.force_vector(input_data, om, nep, quadratic=quadratic)
    fixed, value = solvers.constraints(input_data)
    b = _reduced_rhs(d, f.reshape(-1), fixed, value)
    ref = d["red_rhs"]
    assert b.shape == ref.shape
    # the fixture stores overlay vertices rounded to fp64 (the live sweep keeps np.longdouble), hence not bit-exact
    assert np.abs(b - ref).max() <= 1e-13 * np.abs(ref).max()
    assert input_data[-3] == sorted(input_data[-3], key=lambda c: c[0])     # sorted in place, like the reference


def test_flatten_of_the_rebuilt_graph_reproduces_the_fixture():
    for name in ("simpleOFE3", "AMOREBracket", "patch16"):
        fp0, d = load_golden(name)
        input_data, om, polygons, nep = load_graph(d)
        fp = fl.flatten(input_data, om, polygons)
        for k in fl.FlatProblem.ARRAYS:
            a, b = getattr(fp0, k), getattr(fp, k)
            assert (a is None) == (b is None), k
            if a is not None:
                assert np.array_equal(a, b, equal_nan=True), (name, k)


@pytest.mark.parametrize("kind,quadratic", [("lower", False), ("icm", False), ("quad", True)])
def test_single_mesh_force_vector(kind, quadratic):
    _, d = load_golden("single_" + kind)
    deck = pickle.loads(d["deck"].tobytes())
    f = solvers.force_vector(deck, quadratic=quadratic)
    fixed, value = solvers.constraints(deck)
    b = _reduced_rhs(d, f.reshape(-1), fixed, value)
    assert np.array_equal(b, d["red_rhs"])          # same expressions in the same order: bit-exact


def test_element_type_checks():
    _, d = load_golden("single_quad")
    deck = pickle.loads(d["deck"].tobytes())
    with pytest.raises(ValueError, match="Wrong element numbering"):
        solvers._check_single(deck[3], (3, 4))          # lowerOrderFE rejects 6/9-node elements before any GPU work
    with pytest.raises(ValueError, match="Wrong element numbering"):
        solvers._check_single(deck[3], (4,))
