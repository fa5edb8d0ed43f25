"""Jacobi-PCG (replacing scipy spsolve) against the reference's own reduced system and displacements.
Metrics P3/P4 of SURVEY.md section 8."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from conftest import fixed_from_constraints, load_golden

pytestmark = pytest.mark.gpu

TOL_DISP = 1e-9    # north_star: displacements within 1e-9 relative of scipy's solution
TOL_ENERGY = 1e-9


def _full_rhs(d, fixed, val, n):
    """A global force vector f whose reduced right-hand side (f - K u_fix on free DOFs, AMORE.py:227) is the
    reference's own captured ``red_rhs``."""
    K = sp.csr_matrix((d["csr_data"], d["csr_indices"], d["csr_indptr"]), shape=(n, n))
    f = np.zeros(n)
    free = fixed == 0
    f[free] = d["red_rhs"] + (K @ val)[free]
    return f


@pytest.mark.parametrize("name", ["simpleOFE3", "AMOREBracket", "single_lower", "single_icm", "single_quad", "patch16", "patch16_tri",
                                  "patch16_m3s"])
def test_displacement_and_energy_match_reference(gpu_ctx, name):
    fp, d = load_golden(name)
    n = fp.n_dof
    if "constraints" in d:
        cons = d["constraints"]
    else:
        import pickle
        cons = pickle.loads(d["deck"].tobytes())[-3]
    fixed, val = fixed_from_constraints(cons, n)
    # DOFs no element touches have an empty row in K: the reference's spsolve would be singular, so the
    # fixtures never contain free empty rows
    f = _full_rhs(d, fixed, val, n)
    gpu_ctx.set_problem(fp).symbolic().assemble()
    gpu_ctx.set_system(f, fixed, val)
    u, info = gpu_ctx.solve(rtol=1e-12, maxit=100000)
    assert info["converged"] in (1, 2) and info["relres_rec"] <= 1e-12 and info["relres_true"] <= 1e-8
    if name != "AMOREBracket":   # kappa 1e9: fp64 attainable true residual is 5e-10 (spsolve itself leaves 6e-11)
        assert info["converged"] == 1 and info["relres_true"] <= 1e-12
    ur = d["displacement"]
    assert np.abs(u - ur).max() <= TOL_DISP * np.abs(ur).max()
    e = gpu_ctx.energy()
    assert abs(e - d["energy"][0]) <= TOL_ENERGY * abs(d["energy"][0])
    if name == "simpleOFE3":
        assert abs(e - 19.2) / 19.2 <= 1e-8        # analytic patch-test energy


def test_fixed_values_are_returned_exactly(gpu_ctx):
    fp, d = load_golden("single_lower")
    import pickle
    cons = pickle.loads(d["deck"].tobytes())[-3]
    fixed, val = fixed_from_constraints(cons, fp.n_dof)
    gpu_ctx.set_problem(fp).symbolic().assemble()
    gpu_ctx.set_system(_full_rhs(d, fixed, val, fp.n_dof), fixed, val)
    u, _ = gpu_ctx.solve()
    assert np.array_equal(u[fixed == 1], val[fixed == 1])


def test_maxit_reports_no_convergence(gpu_ctx):
    from mofem_b200 import _lib
    fp, d = load_golden("AMOREBracket")
    fixed, val = fixed_from_constraints(d["constraints"], fp.n_dof)
    gpu_ctx.set_problem(fp).symbolic().assemble()
    gpu_ctx.set_system(_full_rhs(d, fixed, val, fp.n_dof), fixed, val)
    with pytest.raises(_lib.MofemError):
        gpu_ctx.solve(rtol=1e-12, maxit=5)
    u, info = gpu_ctx.solve(rtol=1e-12, maxit=5, check=False)
    assert not info["converged"] and info["iterations"] == 5


def test_zero_rhs(gpu_ctx):
    fp, d = load_golden("simpleOFE3")
    fixed, val = fixed_from_constraints(d["constraints"], fp.n_dof)
    gpu_ctx.set_problem(fp).symbolic().assemble()
    gpu_ctx.set_system(np.zeros(fp.n_dof), fixed, np.zeros(fp.n_dof))
    u, info = gpu_ctx.solve()
    assert info["converged"] and not u.any()
