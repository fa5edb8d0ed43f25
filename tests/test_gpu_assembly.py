"""Parity of the CUDA assembly path (through the C-ABI) against the golden vectors of the reference and
against the oracle.  Metrics P1/P2 of SURVEY.md section 8."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import FIXTURES, coo_to_csr, load_golden, row_scaled_error

pytestmark = pytest.mark.gpu

TOL_VALUES = 1e-12   # north_star: stiffness entries within 1e-12 relative (row-scaled)


@pytest.mark.parametrize("name", FIXTURES)
def test_csr_pattern_bit_exact_and_values(gpu_ctx, name):
    fp, d = load_golden(name)
    gpu_ctx.set_problem(fp).symbolic().assemble()
    indptr, indices, data = gpu_ctx.csr()
    assert indptr.dtype == np.int32 and indices.dtype == np.int32
    assert np.array_equal(indptr, d["csr_indptr"])            # P1: bit-exact with scipy coo->csr of the reference
    assert np.array_equal(indices, d["csr_indices"])
    n = fp.n_dof
    K = sp.csr_matrix((data, indices, indptr), shape=(n, n))
    Kr = sp.csr_matrix((d["csr_data"], d["csr_indices"], d["csr_indptr"]), shape=(n, n))
    err = row_scaled_error(K, Kr)
    assert err <= TOL_VALUES, err
    info = gpu_ctx.info()
    assert info["n_coo"] == len(d["coo_V"]) and info["nnz"] == len(d["csr_data"])


@pytest.mark.parametrize("name", FIXTURES)
def test_block_values_in_emission_order(gpu_ctx, name):
    fp, d = load_golden(name)
    gpu_ctx.set_problem(fp).symbolic().assemble()
    V = gpu_ctx.coo_values()
    I, J, Vr = d["coo_I"], d["coo_J"], d["coo_V"]
    assert V.shape == Vr.shape
    n = fp.n_dof
    # per-triplet error scaled by the row maximum of the assembled reference matrix
    Kr = coo_to_csr(n, I, J, Vr)
    rm = abs(Kr).max(axis=1).toarray().ravel(); rm[rm == 0] = 1.0
    assert (np.abs(V - Vr) / rm[I]).max() <= TOL_VALUES


@pytest.mark.parametrize("name", ["AMOREBracket", "zoo_s2"])
def test_against_oracle_same_inputs(gpu_ctx, name):
    from oracle import oracle
    fp, _ = load_golden(name)
    I, J, V, _, newton = oracle.assemble_coo(fp)
    gpu_ctx.set_problem(fp).symbolic().assemble()
    Vg = gpu_ctx.coo_values()
    rm = abs(coo_to_csr(fp.n_dof, I, J, V)).max(axis=1).toarray().ravel(); rm[rm == 0] = 1.0
    assert (np.abs(Vg - V) / rm[I]).max() <= TOL_VALUES
    assert gpu_ctx.info()["newton_updates"] == newton      # same Newton trajectory (integer count)


def test_assembly_is_bit_reproducible(gpu_ctx):
    fp, _ = load_golden("AMOREBracket")
    gpu_ctx.set_problem(fp).symbolic().assemble()
    a = gpu_ctx.csr()[2].copy()
    gpu_ctx.set_problem(fp).symbolic().assemble()
    b = gpu_ctx.csr()[2]
    assert np.array_equal(a, b)       # deterministic reduce, no float atomics


def test_symmetry_of_assembled_matrix(gpu_ctx):
    fp, _ = load_golden("simpleOFE3")
    gpu_ctx.set_problem(fp).symbolic().assemble()
    K = gpu_ctx.csr_matrix()
    assert abs(K - K.T).max() <= 1e-12 * abs(K).max()


def test_empty_problem(gpu_ctx):
    fp, _ = load_golden("zoo_s1")
    fp.cell_elem = fp.cell_elem[:0]; fp.cell_kind = fp.cell_kind[:0]; fp.cell_tri_ptr = fp.cell_tri_ptr[:1]
    fp.tri_xy = fp.tri_xy[:0]; fp.tri_rho = fp.tri_rho[:0]; fp.tri_mid = None; fp.tri_curved = None
    gpu_ctx.set_problem(fp).symbolic().assemble()
    indptr, indices, data = gpu_ctx.csr()
    assert indices.size == 0 and not indptr.any()


def test_wrong_element_numbering(gpu_ctx):
    fp, _ = load_golden("zoo_s2")
    fp.solver = 1
    with pytest.raises(ValueError, match="Wrong element numbering"):
        gpu_ctx.set_problem(fp).symbolic()


def test_newton_failure_raises_assertion(gpu_ctx):
    fp, _ = load_golden("zoo_s1")
    # self-intersecting quad + far-away integration triangle (same case as tests/test_oracle.py)
    from mofem_b200.flatten import FlatProblem
    bad = FlatProblem(solver=1, n_mesh=1,
                      node_xy=np.array([[0.37, 0.3], [0.38, -0.22], [-0.73, 0.44], [0.05, -0.38]]),
                      elem_nodes=np.array([[0, 1, 2, 3, -1, -1, -1, -1, -1]], np.int32), elem_nn=np.array([4], np.int32),
                      elem_mat=np.array([0], np.int32), elem_mesh=np.array([0], np.int32), mat_D=fp.mat_D[:1].copy(),
                      cell_elem=np.array([[0]], np.int32), cell_kind=np.array([1], np.int32),
                      cell_tri_ptr=np.array([0, 1], np.int64),
                      tri_xy=np.array([[[-0.08, 2.34], [-0.07, 2.34], [-0.08, 2.35]]]), tri_rho=np.ones((1, 3, 1)))
    with pytest.raises(AssertionError, match="Too many iterations"):
        gpu_ctx.set_problem(bad).symbolic().assemble()


def test_spmv_matches_scipy(gpu_ctx):
    fp, _ = load_golden("AMOREBracket")
    gpu_ctx.set_problem(fp).symbolic().assemble()
    K = gpu_ctx.csr_matrix()
    x = np.random.default_rng(5).standard_normal(fp.n_dof)
    y = gpu_ctx.spmv(x)
    yr = K @ x
    assert np.abs(y - yr).max() <= 1e-13 * np.abs(yr).max()


@pytest.mark.parametrize("name", ["patch16", "patch16_m3", "patch16_m3s", "patch32_m3s", "patch16_tri", "simpleOFE3", "zoo_s1"])
def test_cell_kernel_is_bit_identical_to_the_block_kernels(gpu_ctx, name):
    """The cell kernel (all-quad4 overlay cells with 1..3 incident elements: one evaluation per point and element, strain pairs
    through shared memory, the library default) keeps the per-block kernels' fma order: identical bits -- block values and
    assembled CSR values -- and identical Newton counts.  patch*_m3s have cells with three incident elements; patch16_tri and
    the shipped decks have cells the cell kernel must leave to the block kernels."""
    fp, d = load_golden(name)
    try:
        gpu_ctx.set_option("cell_kernel", 0)
        gpu_ctx.set_problem(fp).symbolic().assemble()
        a, na, da = gpu_ctx.coo_values(), gpu_ctx.info()["newton_updates"], gpu_ctx.csr()[2]
    finally:
        gpu_ctx.set_option("cell_kernel", 1)       # the default: the session context goes on with it
    gpu_ctx.set_problem(fp).symbolic().assemble()
    b, nb, db = gpu_ctx.coo_values(), gpu_ctx.info()["newton_updates"], gpu_ctx.csr()[2]
    assert np.array_equal(a, b) and np.array_equal(da, db) and na == nb
