"""ctypes wrapper around oracle/libmofem_oracle.so.

TEST INFRASTRUCTURE ONLY (parity oracle; see mofem_oracle.c).  Imported by
tests/, ``__graft_entry__.smoke()`` and bench.py's CPU-baseline legs -- never by
the product package ``mofem_b200``.  Takes any object exposing the flattened
problem's array attributes (duck-typed, so nothing is imported from the product).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_d_p = C.POINTER(C.c_double)
c_i32_p = C.POINTER(C.c_int32)
c_i64_p = C.POINTER(C.c_int64)
c_u8_p = C.POINTER(C.c_uint8)


class OrcProblem(C.Structure):
    _fields_ = [("n_node", C.c_int64), ("n_elem", C.c_int64), ("n_cell", C.c_int64), ("n_tri", C.c_int64),
                ("n_mesh", C.c_int32), ("n_mat", C.c_int32), ("solver", C.c_int32), ("_pad", C.c_int32),
                ("node_xy", c_d_p), ("elem_nodes", c_i32_p), ("elem_nn", c_i32_p), ("elem_mat", c_i32_p),
                ("mat_D", c_d_p), ("cell_elem", c_i32_p), ("cell_kind", c_i32_p), ("cell_tri_ptr", c_i64_p),
                ("tri_xy", c_d_p), ("tri_rho", c_d_p), ("tri_mid", c_d_p), ("tri_curved", c_u8_p)]


def build(force=False):
    so = os.path.join(_HERE, "libmofem_oracle.so")
    src = os.path.join(_HERE, "mofem_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_coo_count.restype = C.c_int64
        L.orc_coo_count.argtypes = [C.POINTER(OrcProblem)]
        L.orc_assemble_coo.restype = C.c_int
        L.orc_assemble_coo.argtypes = [C.POINTER(OrcProblem), c_i64_p, c_i64_p, c_d_p, c_i64_p, c_i64_p, C.c_int]
        L.orc_coo_tocsr.restype = C.c_int64
        L.orc_coo_tocsr.argtypes = [C.c_int64, C.c_int64, c_i64_p, c_i64_p, c_d_p, c_i64_p, c_i64_p, c_d_p]
        L.orc_element_stiffness.restype = C.c_int
        L.orc_element_stiffness.argtypes = [c_d_p, C.c_int, c_d_p, C.c_int, c_d_p]
        L.orc_amore_block.restype = C.c_int
        L.orc_amore_block.argtypes = [C.c_int, c_d_p, C.c_int, c_d_p, c_d_p, C.c_int, c_d_p,
                                      c_d_p, C.c_int, c_d_p, c_d_p, C.POINTER(C.c_int)]
        L.orc_inv_newton.restype = C.c_int
        L.orc_inv_newton.argtypes = [c_d_p, C.c_int, c_d_p, c_d_p]
        L.orc_inv_tri.restype = None
        L.orc_inv_tri.argtypes = [c_d_p, c_d_p, c_d_p]
        L.orc_shape.restype = None
        L.orc_shape.argtypes = [c_d_p, C.c_int, C.c_double, C.c_double, c_d_p, c_d_p, c_d_p]
        L.orc_gauss_tri.argtypes = [C.c_int, c_d_p, c_d_p]
        L.orc_gauss_quad.argtypes = [C.c_int, c_d_p, c_d_p]
        L.orc_material_matrix.argtypes = [C.c_double, C.c_double, C.c_int, c_d_p]
        L.orc_amore_nint.argtypes = [C.c_int, C.c_int, C.c_int]
        L.orc_rho_eval.restype = C.c_int
        L.orc_rho_eval.argtypes = [C.c_int64, C.c_int32, c_d_p, c_i32_p, c_i32_p, c_d_p, c_d_p, c_d_p, c_i64_p]
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class _Held:
    """Keeps the arrays behind an OrcProblem alive."""

    def __init__(self, fp):
        k = self.keep = dict(
            node_xy=_d(fp.node_xy), elem_nodes=np.ascontiguousarray(fp.elem_nodes, np.int32),
            elem_nn=np.ascontiguousarray(fp.elem_nn, np.int32), elem_mat=np.ascontiguousarray(fp.elem_mat, np.int32),
            mat_D=_d(fp.mat_D), cell_elem=np.ascontiguousarray(fp.cell_elem, np.int32),
            cell_kind=np.ascontiguousarray(fp.cell_kind, np.int32),
            cell_tri_ptr=np.ascontiguousarray(fp.cell_tri_ptr, np.int64),
            tri_xy=_d(fp.tri_xy), tri_rho=_d(fp.tri_rho),
            tri_mid=_d(fp.tri_mid) if getattr(fp, "tri_mid", None) is not None else None,
            tri_curved=(np.ascontiguousarray(fp.tri_curved, np.uint8)
                        if getattr(fp, "tri_curved", None) is not None else None))
        self.c = OrcProblem(
            n_node=k["node_xy"].shape[0], n_elem=k["elem_nn"].shape[0], n_cell=k["cell_kind"].shape[0],
            n_tri=k["tri_xy"].shape[0], n_mesh=int(fp.n_mesh), n_mat=k["mat_D"].shape[0], solver=int(fp.solver),
            node_xy=_p(k["node_xy"], c_d_p), elem_nodes=_p(k["elem_nodes"], c_i32_p),
            elem_nn=_p(k["elem_nn"], c_i32_p), elem_mat=_p(k["elem_mat"], c_i32_p), mat_D=_p(k["mat_D"], c_d_p),
            cell_elem=_p(k["cell_elem"], c_i32_p), cell_kind=_p(k["cell_kind"], c_i32_p),
            cell_tri_ptr=_p(k["cell_tri_ptr"], c_i64_p), tri_xy=_p(k["tri_xy"], c_d_p),
            tri_rho=_p(k["tri_rho"], c_d_p), tri_mid=_p(k["tri_mid"], c_d_p), tri_curved=_p(k["tri_curved"], c_u8_p))


def assemble_coo(fp, n_threads=1):
    """(I, J, V, cell_off, newton_updates) in the reference's emission order."""
    L = lib()
    h = _Held(fp)
    n = L.orc_coo_count(C.byref(h.c))
    I = np.empty(n, np.int64); J = np.empty(n, np.int64); V = np.empty(n, np.float64)
    off = np.empty(h.c.n_cell + 1, np.int64)
    nw = C.c_int64(0)
    rc = L.orc_assemble_coo(C.byref(h.c), _p(I, c_i64_p), _p(J, c_i64_p), _p(V, c_d_p), _p(off, c_i64_p),
                            C.byref(nw), int(n_threads))
    if rc == -1:
        raise AssertionError("Too many iterations!")       # invShapeFunction.py:68
    if rc == -2:
        raise ValueError("Wrong element numbering!")       # AMORE.py:51
    return I, J, V, off, nw.value


def coo_tocsr(n_row, I, J, V):
    """Oracle restatement of scipy coo->csr; returns (indptr, indices, data)."""
    L = lib()
    I = np.ascontiguousarray(I, np.int64); J = np.ascontiguousarray(J, np.int64); V = _d(V)
    indptr = np.empty(n_row + 1, np.int64); indices = np.empty(len(I), np.int64); data = np.empty(len(I))
    nnz = L.orc_coo_tocsr(n_row, len(I), _p(I, c_i64_p), _p(J, c_i64_p), _p(V, c_d_p),
                          _p(indptr, c_i64_p), _p(indices, c_i64_p), _p(data, c_d_p))
    return indptr, indices[:nnz].copy(), data[:nnz].copy()


def element_stiffness(coord, D, quad4_rule=0):
    coord = _d(coord); D = _d(D); n = coord.shape[0]
    K = np.empty((2 * n, 2 * n))
    if lib().orc_element_stiffness(_p(coord, c_d_p), n, _p(D, c_d_p), int(quad4_rule), _p(K, c_d_p)):
        raise ValueError
    return K


def amore_block(quadratic, coord_tri, D, coord1, rho1, coord2=None, rho2=None):
    """One triangle's contribution: stiffnessMatrix.lowerOrderAMORE / quadraticAMORE."""
    coord_tri = _d(coord_tri); D = _d(D); coord1 = _d(coord1); rho1 = _d(rho1)
    n1 = coord1.shape[0]
    if coord2 is None:
        n2 = n1; c2 = r2 = None
    else:
        c2 = _d(coord2); r2 = _d(rho2); n2 = c2.shape[0]
    K = np.empty((2 * n1, 2 * n2))
    nw = C.c_int(0)
    rc = lib().orc_amore_block(int(quadratic), _p(coord_tri, c_d_p), coord_tri.shape[0], _p(D, c_d_p),
                               _p(coord1, c_d_p), n1, _p(rho1, c_d_p), _p(c2, c_d_p), n2, _p(r2, c_d_p),
                               _p(K, c_d_p), C.byref(nw))
    if rc:
        raise AssertionError("Too many iterations!")
    return K


def inv_newton(coord, xy):
    coord = _d(coord); xy = _d(xy); iso = np.empty(2)
    it = lib().orc_inv_newton(_p(coord, c_d_p), coord.shape[0], _p(xy, c_d_p), _p(iso, c_d_p))
    if it < 0:
        raise AssertionError("Too many iterations!")
    return iso, it


def inv_tri(coord, xy):
    coord = _d(coord); xy = _d(xy); iso = np.empty(3)
    lib().orc_inv_tri(_p(coord, c_d_p), _p(xy, c_d_p), _p(iso, c_d_p))
    return iso


def shape(coord, r, s):
    coord = _d(coord); n = coord.shape[0]
    Nmat = np.empty((2, 2 * n)); Bmat = np.empty((3, 2 * n)); jac = C.c_double(0)
    lib().orc_shape(_p(coord, c_d_p), n, float(r), float(s), _p(Nmat, c_d_p), _p(Bmat, c_d_p), C.byref(jac))
    return Nmat, Bmat, jac.value


def gauss_tri(n):
    pos = np.empty((n, 3)); wei = np.empty(n)
    if lib().orc_gauss_tri(n, _p(pos, c_d_p), _p(wei, c_d_p)):
        raise ValueError("Wrong number of quadrature points!")
    return pos, wei


def gauss_quad(n):
    pos = np.empty(n); wei = np.empty(n)
    if lib().orc_gauss_quad(n, _p(pos, c_d_p), _p(wei, c_d_p)):
        raise ValueError("Wrong number of quadrature points!")
    return pos, wei


def material_matrix(E, nu, problem_type):
    D = np.empty((3, 3))
    if lib().orc_material_matrix(float(E), float(nu), int(problem_type), _p(D, c_d_p)):
        raise ValueError("No such a problem type!")
    return D


def rho_eval(vert_xy, vert_elem, elem_nv, elem_xy, elem_w):
    """planeSweepMeshes.rhoCalculation:991-1030 on flat arrays -> (rho [V,M], newton updates)."""
    vert_xy = _d(vert_xy); vert_elem = np.ascontiguousarray(vert_elem, np.int32)
    elem_nv = np.ascontiguousarray(elem_nv, np.int32); elem_xy = _d(elem_xy); elem_w = _d(elem_w)
    V, M = vert_elem.shape
    rho = np.empty((V, M))
    nw = C.c_int64(0)
    with np.errstate(all="ignore"):
        rc = lib().orc_rho_eval(V, M, _p(vert_xy, c_d_p), _p(vert_elem, c_i32_p), _p(elem_nv, c_i32_p),
                                _p(elem_xy, c_d_p), _p(elem_w, c_d_p), _p(rho, c_d_p), C.byref(nw))
    if rc == -1:
        raise AssertionError("Too many iterations!")
    if rc == -2:
        raise ValueError
    return rho, nw.value
