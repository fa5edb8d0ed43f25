"""The oracle (oracle/mofem_oracle.c) against the golden vectors generated from the unmodified
reference (tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import FIXTURES, GOLDEN, coo_to_csr, load_golden, row_scaled_error
from oracle import oracle


@pytest.mark.parametrize("name", FIXTURES)
def test_assembly_matches_reference_triplets(name):
    fp, d = load_golden(name)
    I, J, V, off, newton = oracle.assemble_coo(fp)
    assert np.array_equal(I, d["coo_I"]) and np.array_equal(J, d["coo_J"])      # emission order, bit-exact
    n = fp.n_dof
    err = row_scaled_error(coo_to_csr(n, I, J, V), coo_to_csr(n, I, J, d["coo_V"]))
    assert err <= 1e-13, err                                                      # bar is 1e-12; oracle sits at ~1e-15
    # per-triplet check scaled by the largest value of its block row
    assert np.abs(V - d["coo_V"]).max() <= 1e-13 * np.abs(d["coo_V"]).max()


@pytest.mark.parametrize("name", FIXTURES)
def test_coo_tocsr_matches_scipy(name):
    fp, d = load_golden(name)
    indptr, indices, data = oracle.coo_tocsr(fp.n_dof, d["coo_I"], d["coo_J"], d["coo_V"])
    assert np.array_equal(indptr, d["csr_indptr"]) and np.array_equal(indices, d["csr_indices"])   # P1 incl. explicit zeros
    K = sp.csr_matrix((data, indices, indptr), shape=(fp.n_dof,) * 2)
    Kr = sp.csr_matrix((d["csr_data"], d["csr_indices"], d["csr_indptr"]), shape=(fp.n_dof,) * 2)
    assert row_scaled_error(K, Kr) <= 1e-14


def test_threaded_assembly_is_identical():
    fp, d = load_golden("AMOREBracket")
    a = oracle.assemble_coo(fp, 1)
    b = oracle.assemble_coo(fp, 4)
    assert np.array_equal(a[2], b[2]) and a[4] == b[4]


def test_newton_known_answers():
    d = dict(np.load(os.path.join(GOLDEN, "newton.npz")))
    for n in (4, 9, 6):
        for c, xy, iso in zip(d[f"coord{n}"], d[f"xy{n}"], d[f"iso{n}"]):
            got, it = oracle.inv_newton(c, xy)
            assert np.allclose(got, iso, rtol=0, atol=1e-14)
            assert 0 <= it < 20
    for c, xy, iso in zip(d["coord3"], d["xy3"], d["iso3"]):
        assert np.allclose(oracle.inv_tri(c, xy)[:2], iso, rtol=0, atol=1e-14)


def test_newton_failure_is_reported():
    # a self-intersecting quad and a far-away point: Newton never meets the 1e-15 tolerance and the
    # reference asserts "Too many iterations!" (invShapeFunction.py:68)
    c = np.array([[0.37, 0.3], [0.38, -0.22], [-0.73, 0.44], [0.05, -0.38]])
    with pytest.raises(AssertionError):
        oracle.inv_newton(c, np.array([-0.08, 2.34]))


def test_gauss_tables():
    for n in (1, 3, 4, 6, 7, 9, 12, 13, 16):
        pos, wei = oracle.gauss_tri(n)
        assert abs(wei.sum() - 1.0) < 1e-13 and np.allclose(pos.sum(1), 1.0, atol=1e-14)
    with pytest.raises(ValueError):
        oracle.gauss_tri(5)
    for n in range(1, 6):
        pos, wei = oracle.gauss_quad(n)
        assert abs(wei.sum() - 2.0) < 1e-13
    with pytest.raises(ValueError):
        oracle.gauss_quad(6)


def test_patch_test_energy_from_golden():
    # simpleOFE3 is the reference's patch test: energy 19.2 (README.md:12-13)
    _, d = load_golden("simpleOFE3")
    assert abs(d["energy"][0] - 19.2) / 19.2 < 1e-8


def test_wrong_element_numbering():
    fp, _ = load_golden("zoo_s2")
    fp.solver = 1            # lowerOrderAMORE rejects 6/9-node elements (AMORE.py:51,74)
    with pytest.raises(ValueError):
        oracle.assemble_coo(fp)
