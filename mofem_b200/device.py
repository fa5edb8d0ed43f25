"""Python face of the C-ABI (include/mofem_b200.h): one :class:`Context` per GPU.

Every method is a thin call into ``libmofem_b200.so``; there is no arithmetic here and no CPU
fallback -- without the library or a GPU the constructor raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .flatten import FlatProblem


_PROBLEM_DTYPES = dict(node_xy=np.float64, elem_nodes=np.int32, elem_nn=np.int32, elem_mat=np.int32, mat_D=np.float64,
                       cell_elem=np.int32, cell_kind=np.int32, cell_tri_ptr=np.int64, tri_xy=np.float64, tri_rho=np.float64,
                       tri_mid=np.float64, tri_curved=np.uint8)


def _c(a, dtype):
    return None if a is None else np.ascontiguousarray(a, dtype=dtype)


class Context:
    """Owns one ``mofem_ctx`` (one CUDA stream on one device)."""

    def __init__(self, device: int = 0):
        self._L = _lib.load()
        h = C.c_void_p()
        rc = self._L.mofem_create(C.byref(h), int(device))
        if rc != _lib.MOFEM_OK:
            raise _lib.MofemError(f"mofem_create(device={device}) failed with {rc}: no usable CUDA device "
                                  "(mofem_b200 has no CPU fallback)")
        self._h = h
        self.device = int(device)
        self._keep = None
        self.n_dof = 0

    # -- lifecycle -------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._L.mofem_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _p(self, a, dtype):
        """Validated raw pointer: exact dtype, contiguous, and (CUDA tensors) on this context's GPU."""
        return _lib.ptr(a, dtype, self.device)

    def _ck(self, rc):
        _lib.raise_for(rc, self._h)

    @property
    def stream(self) -> int:
        """``cudaStream_t`` of the context as an integer (for CUDA-event timing on the launching stream)."""
        return int(self._L.mofem_stream(self._h) or 0)

    # -- assembly --------------------------------------------------------------------------
    @staticmethod
    def problem_struct(fp: FlatProblem, device=None):
        """(struct, keep-alive dict) for a flattened problem held in host numpy arrays or torch tensors."""
        def arr(a, dt):
            if a is None or hasattr(a, "data_ptr"):
                return a
            return _c(a, dt)
        k = dict(node_xy=arr(fp.node_xy, np.float64), elem_nodes=arr(fp.elem_nodes, np.int32),
                 elem_nn=arr(fp.elem_nn, np.int32), elem_mat=arr(fp.elem_mat, np.int32),
                 mat_D=arr(fp.mat_D, np.float64), cell_elem=arr(fp.cell_elem, np.int32),
                 cell_kind=arr(fp.cell_kind, np.int32), cell_tri_ptr=arr(fp.cell_tri_ptr, np.int64),
                 tri_xy=arr(fp.tri_xy, np.float64), tri_rho=arr(fp.tri_rho, np.float64),
                 tri_mid=arr(fp.tri_mid, np.float64), tri_curved=arr(fp.tri_curved, np.uint8))
        s = _lib.MofemProblem(
            n_node=int(k["node_xy"].shape[0]), n_elem=int(k["elem_nn"].shape[0]), n_cell=int(k["cell_kind"].shape[0]),
            n_tri=int(k["tri_xy"].shape[0]), n_mesh=int(fp.n_mesh), n_mat=int(k["mat_D"].shape[0]),
            solver=int(fp.solver), **{name: _lib.ptr(v, _PROBLEM_DTYPES[name], device, name) for name, v in k.items()})
        return s, k

    def set_problem(self, fp: FlatProblem):
        s, keep = self.problem_struct(fp, self.device)
        self._ck(self._L.mofem_set_problem(self._h, C.byref(s)))
        self._keep = keep
        self.n_dof = 2 * int(s.n_node)
        return self

    def symbolic(self):
        self._ck(self._L.mofem_symbolic(self._h))
        return self

    def assemble(self):
        self._ck(self._L.mofem_assemble(self._h))
        return self

    def info(self) -> dict:
        i = _lib.MofemInfo()
        self._ck(self._L.mofem_get_info(self._h, C.byref(i)))
        return {f: int(getattr(i, f)) for f, _ in _lib.MofemInfo._fields_}

    def csr(self, values=True):
        """(indptr int32, indices int32, data float64 | None) as host arrays -- scipy's canonical CSR."""
        i = self.info()
        indptr = np.empty(i["n_dof"] + 1, np.int32)
        indices = np.empty(i["nnz"], np.int32)
        data = np.empty(i["nnz"], np.float64) if values else None
        self._ck(self._L.mofem_get_csr(self._h, self._p(indptr, np.int32), self._p(indices, np.int32), self._p(data, np.float64)))
        return indptr, indices, data

    def csr_matrix(self):
        import scipy.sparse
        indptr, indices, data = self.csr()
        n = len(indptr) - 1
        return scipy.sparse.csr_matrix((data, indices, indptr), shape=(n, n))

    def coo_values(self):
        """``Vglo`` in the reference's emission order (parity / debug)."""
        V = np.empty(self.info()["n_coo"], np.float64)
        self._ck(self._L.mofem_get_coo_values(self._h, self._p(V, np.float64)))
        return V

    # -- solve -----------------------------------------------------------------------------
    def set_system(self, rhs, fixed, fixed_value):
        rhs = rhs if hasattr(rhs, "data_ptr") else _c(np.asarray(rhs).reshape(-1), np.float64)
        fixed = fixed if hasattr(fixed, "data_ptr") else _c(np.asarray(fixed).reshape(-1), np.uint8)
        fixed_value = fixed_value if hasattr(fixed_value, "data_ptr") else _c(np.asarray(fixed_value).reshape(-1), np.float64)
        for a in (rhs, fixed, fixed_value):
            if a.shape[0] != self.n_dof:
                raise ValueError(f"system arrays must have {self.n_dof} entries")
        self._ck(self._L.mofem_set_system(self._h, self._p(rhs, np.float64), self._p(fixed, np.uint8), self._p(fixed_value, np.float64)))
        return self

    def solve(self, rtol=1e-12, maxit=200000, out=None, check=True):
        """Jacobi-PCG; returns (u, info dict).  ``out`` may be a numpy array or a CUDA tensor."""
        u = np.empty(self.n_dof, np.float64) if out is None else out
        si = _lib.MofemSolveInfo()
        rc = self._L.mofem_solve(self._h, float(rtol), int(maxit), self._p(u, np.float64), C.byref(si))
        info = dict(iterations=si.iterations, converged=int(si.converged), relres_true=si.relres_true,
                    relres_rec=si.relres_rec, restarts=si.restarts)
        if rc == _lib.E_NOCONV and not check:
            return u, info
        self._ck(rc)
        return u, info

    def energy(self) -> float:
        e = C.c_double(0.0)
        self._ck(self._L.mofem_energy(self._h, C.byref(e)))
        return e.value

    def spmv(self, x, out=None):
        x = x if hasattr(x, "data_ptr") else _c(np.asarray(x).reshape(-1), np.float64)
        y = np.empty(self.n_dof, np.float64) if out is None else out
        self._ck(self._L.mofem_spmv(self._h, self._p(x, np.float64), self._p(y, np.float64)))
        return y

    def assemble_solve(self, fp: FlatProblem, rhs, fixed, fixed_value, rtol=1e-12, maxit=200000):
        """One call, host buffers in / host displacement out (the end-to-end timing path)."""
        s, keep = self.problem_struct(fp, self.device)
        rhs = _c(np.asarray(rhs).reshape(-1), np.float64)
        fixed = _c(np.asarray(fixed).reshape(-1), np.uint8)
        fixed_value = _c(np.asarray(fixed_value).reshape(-1), np.float64)
        u = np.empty(2 * int(s.n_node), np.float64)
        e = C.c_double(0.0)
        si = _lib.MofemSolveInfo()
        rc = self._L.mofem_assemble_solve(self._h, C.byref(s), self._p(rhs, np.float64), self._p(fixed, np.uint8),
                                          self._p(fixed_value, np.float64), float(rtol), int(maxit), self._p(u, np.float64),
                                          C.byref(e), C.byref(si))
        self._keep = keep
        self.n_dof = 2 * int(s.n_node)
        self._ck(rc)
        return u, e.value, dict(iterations=si.iterations, converged=int(si.converged), relres_true=si.relres_true,
                                relres_rec=si.relres_rec, restarts=si.restarts)

    # -- row-partitioned solve (one process per GPU) -------------------------------------------------
    def dist_unique_id(self) -> bytes:
        buf = (C.c_uint8 * 128)()
        self._ck(self._L.mofem_dist_unique_id(buf))
        return bytes(buf)

    def dist_init(self, unique_id: bytes, rank: int, world: int):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self._ck(self._L.mofem_dist_init(self._h, buf, int(rank), int(world)))
        return self

    def set_halo(self, part):
        """Install the halo plan of a :class:`mofem_b200.partition.LocalPart` (after ``exchange_send_lists``)."""
        nb = _c(part.neighbors, np.int32)
        sp = _c(part.send_ptr, np.int64); si = _c(part.send_idx, np.int32); rp = _c(part.recv_ptr, np.int64)
        self._ck(self._L.mofem_set_halo(self._h, int(part.n_owned_node), int(len(nb)), _lib.ptr(nb), _lib.ptr(sp),
                                        _lib.ptr(si) if len(si) else None, _lib.ptr(rp)))
        return self

    def p2p_export(self):
        """-> (64-byte CUDA IPC handle of this rank's exchange buffer, its ghost-node capacity)."""
        buf = (C.c_uint8 * 64)()
        g = C.c_int64(0)
        self._ck(self._L.mofem_p2p_export(self._h, buf, C.byref(g)))
        return bytes(buf), int(g.value)

    def p2p_import(self, handles, peer_n_owned, send_remote_off):
        hb = b"".join(handles)
        hbuf = (C.c_uint8 * len(hb)).from_buffer_copy(hb)
        ng = _c(np.asarray(peer_n_owned), np.int64)
        so = _c(np.asarray(send_remote_off), np.int64)
        self._ck(self._L.mofem_p2p_import(self._h, hbuf, _lib.ptr(ng), _lib.ptr(so) if len(so) else None))
        return self

    def p2p_close(self):
        self._ck(self._L.mofem_p2p_close(self._h))
        return self

    def clear_halo(self):
        self._ck(self._L.mofem_clear_halo(self._h))
        return self

    def sample(self, n_sampling: int, solver: int, displacement=None):
        """Sampler on the current problem: array [n_unit, 7, nS, nS] = X, Y, U, V, Sxx, Syy, Sxy per unit (regular element or
        integration triangle, in cell order) at the parent-grid points ``np.linspace(-1, 1, nS)``."""
        grid = np.ascontiguousarray(np.linspace(-1, 1, n_sampling))
        nu = C.c_int64(0)
        self._ck(self._L.mofem_sample(self._h, None, int(n_sampling), int(solver), None, None, 0, C.byref(nu)))
        out = np.empty((nu.value, 7, n_sampling, n_sampling))
        u = None if displacement is None else _c(np.asarray(displacement).reshape(-1), np.float64)
        self._ck(self._L.mofem_sample(self._h, _lib.ptr(u), int(n_sampling), int(solver), _lib.ptr(grid), _lib.ptr(out),
                                      int(nu.value), C.byref(nu)))
        return out

    def rho_eval(self, vert_xy, vert_elem, elem_nv, elem_xy, elem_w, out=None):
        """Weight functions of overlay vertices (``mofem_rho_eval``): rho [V, M].  Inputs may be numpy arrays or device
        tensors; ``out`` (optional) a preallocated [V, M] array / device tensor.  Returns (rho, newton updates, kernel ms)."""
        if isinstance(vert_elem, np.ndarray):
            vert_xy = _c(vert_xy, np.float64); vert_elem = _c(vert_elem, np.int32); elem_nv = _c(elem_nv, np.int32)
            elem_xy = _c(elem_xy, np.float64); elem_w = _c(elem_w, np.float64)
        V, M = int(vert_elem.shape[0]), int(vert_elem.shape[1])
        if out is None:
            out = np.empty((V, M))
        nw = C.c_int64(0)
        ms = C.c_double(0)
        self._ck(self._L.mofem_rho_eval(self._h, V, M, self._p(vert_xy, np.float64), self._p(vert_elem, np.int32), int(elem_nv.shape[0]),
                                        self._p(elem_nv, np.int32), self._p(elem_xy, np.float64), self._p(elem_w, np.float64),
                                        self._p(out, np.float64),
                                        C.byref(nw), C.byref(ms)))
        return out, int(nw.value), float(ms.value)

    def set_option(self, name: str, value: int):
        self._ck(self._L.mofem_set_option(self._h, name.encode(), int(value)))
        return self

    def set_profile(self, on=True):
        self._ck(self._L.mofem_set_profile(self._h, int(on)))
        return self

    def pcg_profile(self) -> dict:
        """Summed device time (ms) of the two PCG kernels (SpMV incl. exchange / vector kernel) over the sampled iterations of
        the last solve (needs ``set_profile(N)``)."""
        ms = (C.c_double * 3)()
        it = C.c_int64(0)
        self._ck(self._L.mofem_get_pcg_profile(self._h, ms, C.byref(it)))
        return dict(spmv_ms=ms[0], vec_ms=ms[1], iterations=int(it.value))

    def timing(self) -> dict:
        ms = (C.c_double * 6)()
        ln = (C.c_int64 * 4)()
        self._ck(self._L.mofem_get_timing(self._h, ms, ln))
        return dict(symbolic_ms=ms[0], integrate_ms=ms[1], numeric_ms=ms[2], solve_ms=ms[3], h2d_ms=ms[4], d2h_ms=ms[5],
                    launches=dict(symbolic=int(ln[0]), integrate=int(ln[1]), numeric=int(ln[2]), solve=int(ln[3])))
