"""ctypes binding of libmofem_b200.so (the C-ABI in include/mofem_b200.h).

There is no CPU fallback: if the library is missing or no GPU is present the compute calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# MOFEM_B200_LIB: another build of the same library (kernel experiments); the default is the in-tree one
LIB_PATH = os.environ.get("MOFEM_B200_LIB") or os.path.join(_HERE, "libmofem_b200.so")

c_d_p = C.POINTER(C.c_double)
c_i32_p = C.POINTER(C.c_int32)
c_i64_p = C.POINTER(C.c_int64)
c_u8_p = C.POINTER(C.c_uint8)

MOFEM_OK = 0
E_CUDA, E_ARG, E_NEWTON, E_ELEMENT, E_NOMEM, E_LIMIT, E_NOCONV = -1, -2, -3, -4, -5, -6, -7


class MofemProblem(C.Structure):
    _fields_ = [("n_node", C.c_int64), ("n_elem", C.c_int64), ("n_cell", C.c_int64), ("n_tri", C.c_int64),
                ("n_mesh", C.c_int32), ("n_mat", C.c_int32), ("solver", C.c_int32), ("_pad", C.c_int32),
                ("node_xy", C.c_void_p), ("elem_nodes", C.c_void_p), ("elem_nn", C.c_void_p), ("elem_mat", C.c_void_p),
                ("mat_D", C.c_void_p), ("cell_elem", C.c_void_p), ("cell_kind", C.c_void_p), ("cell_tri_ptr", C.c_void_p),
                ("tri_xy", C.c_void_p), ("tri_rho", C.c_void_p), ("tri_mid", C.c_void_p), ("tri_curved", C.c_void_p)]


class MofemInfo(C.Structure):
    _fields_ = [("n_dof", C.c_int64), ("nnz", C.c_int64), ("n_block", C.c_int64), ("n_coo", C.c_int64),
                ("n_pair", C.c_int64), ("newton_updates", C.c_int64), ("n_eval", C.c_int64)]


class MofemSolveInfo(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("converged", C.c_int32), ("relres_true", C.c_double),
                ("relres_rec", C.c_double), ("restarts", C.c_int32), ("_pad", C.c_int32)]


# every symbol include/mofem_b200.h declares (tests check the library exports all of them)
EXPORTS = ("mofem_create", "mofem_destroy", "mofem_last_error", "mofem_version", "mofem_stream", "mofem_set_problem",
           "mofem_symbolic", "mofem_assemble", "mofem_get_info", "mofem_get_csr", "mofem_get_coo_values",
           "mofem_set_system", "mofem_solve", "mofem_energy", "mofem_spmv", "mofem_assemble_solve", "mofem_get_timing", "mofem_set_profile",
           "mofem_get_pcg_profile", "mofem_dist_unique_id", "mofem_dist_init", "mofem_set_halo", "mofem_clear_halo", "mofem_set_option", "mofem_p2p_export", "mofem_p2p_import",
           "mofem_p2p_close", "mofem_sample", "mofem_rho_eval", "mofem_overlay_patch_sizes", "mofem_overlay_patch_family")

_lib = None


def load():
    """Load the CUDA library; raises RuntimeError if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not found: build it with `python -m mofem_b200.build` "
                           "(mofem_b200 has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.mofem_create.argtypes = [C.POINTER(vp), C.c_int]; L.mofem_create.restype = C.c_int
    L.mofem_destroy.argtypes = [vp]; L.mofem_destroy.restype = None
    L.mofem_last_error.argtypes = [vp]; L.mofem_last_error.restype = C.c_char_p
    L.mofem_version.argtypes = []; L.mofem_version.restype = C.c_char_p
    L.mofem_stream.argtypes = [vp]; L.mofem_stream.restype = vp
    L.mofem_set_problem.argtypes = [vp, C.POINTER(MofemProblem)]; L.mofem_set_problem.restype = C.c_int
    L.mofem_symbolic.argtypes = [vp]; L.mofem_symbolic.restype = C.c_int
    L.mofem_assemble.argtypes = [vp]; L.mofem_assemble.restype = C.c_int
    L.mofem_get_info.argtypes = [vp, C.POINTER(MofemInfo)]; L.mofem_get_info.restype = C.c_int
    L.mofem_get_csr.argtypes = [vp, vp, vp, vp]; L.mofem_get_csr.restype = C.c_int
    L.mofem_get_coo_values.argtypes = [vp, vp]; L.mofem_get_coo_values.restype = C.c_int
    L.mofem_set_system.argtypes = [vp, vp, vp, vp]; L.mofem_set_system.restype = C.c_int
    L.mofem_solve.argtypes = [vp, C.c_double, C.c_int32, vp, C.POINTER(MofemSolveInfo)]; L.mofem_solve.restype = C.c_int
    L.mofem_energy.argtypes = [vp, c_d_p]; L.mofem_energy.restype = C.c_int
    L.mofem_spmv.argtypes = [vp, vp, vp]; L.mofem_spmv.restype = C.c_int
    L.mofem_assemble_solve.argtypes = [vp, C.POINTER(MofemProblem), vp, vp, vp, C.c_double, C.c_int32, vp, c_d_p,
                                       C.POINTER(MofemSolveInfo)]
    L.mofem_assemble_solve.restype = C.c_int
    L.mofem_get_timing.argtypes = [vp, c_d_p, c_i64_p]; L.mofem_get_timing.restype = C.c_int
    L.mofem_dist_unique_id.argtypes = [vp]; L.mofem_dist_unique_id.restype = C.c_int
    L.mofem_dist_init.argtypes = [vp, vp, C.c_int32, C.c_int32]; L.mofem_dist_init.restype = C.c_int
    L.mofem_set_halo.argtypes = [vp, C.c_int64, C.c_int32, vp, vp, vp, vp]; L.mofem_set_halo.restype = C.c_int
    L.mofem_clear_halo.argtypes = [vp]; L.mofem_clear_halo.restype = C.c_int
    L.mofem_p2p_export.argtypes = [vp, vp, c_i64_p]; L.mofem_p2p_export.restype = C.c_int
    L.mofem_p2p_import.argtypes = [vp, vp, vp, vp]; L.mofem_p2p_import.restype = C.c_int
    L.mofem_p2p_close.argtypes = [vp]; L.mofem_p2p_close.restype = C.c_int
    L.mofem_sample.argtypes = [vp, vp, C.c_int32, C.c_int32, vp, vp, C.c_int64, c_i64_p]; L.mofem_sample.restype = C.c_int
    L.mofem_rho_eval.argtypes = [vp, C.c_int64, C.c_int32, vp, vp, C.c_int64, vp, vp, vp, vp, c_i64_p, c_d_p]; L.mofem_rho_eval.restype = C.c_int
    L.mofem_overlay_patch_sizes.argtypes = [C.c_int32] * 5 + [c_i64_p]; L.mofem_overlay_patch_sizes.restype = C.c_int
    L.mofem_overlay_patch_family.argtypes = [vp] + [C.c_int32] * 5 + [vp] * 7 + [c_d_p]; L.mofem_overlay_patch_family.restype = C.c_int
    L.mofem_set_profile.argtypes = [vp, C.c_int]; L.mofem_set_profile.restype = C.c_int
    L.mofem_set_option.argtypes = [vp, C.c_char_p, C.c_int64]; L.mofem_set_option.restype = C.c_int
    L.mofem_get_pcg_profile.argtypes = [vp, c_d_p, c_i64_p]; L.mofem_get_pcg_profile.restype = C.c_int
    _lib = L
    return L


class MofemError(RuntimeError):
    pass


def raise_for(code, ctx_handle):
    """Map a C-ABI status to the exception the reference would raise."""
    if code == MOFEM_OK:
        return
    msg = load().mofem_last_error(ctx_handle)
    msg = msg.decode() if msg else ""
    if code == E_NEWTON:
        raise AssertionError(msg or "Too many iterations!")        # invShapeFunction.py:68,102,136
    if code == E_ELEMENT:
        raise ValueError(msg or "Wrong element numbering!")        # AMORE.py:51,74
    if code == E_ARG:
        raise ValueError(msg)
    if code == E_NOMEM:
        raise MemoryError(msg)
    raise MofemError(f"mofem_b200 error {code}: {msg}")


_TORCH_NAMES = {"float64": "torch.float64", "int32": "torch.int32", "int64": "torch.int64", "uint8": "torch.uint8"}


def ptr(a, dtype=None, device=None, what="array"):
    """Raw address of a numpy array or a torch tensor (host or device), or None.

    The C-ABI takes plain pointers, so a wrong element type, a strided view or a tensor on another GPU would be reinterpreted
    silently: with ``dtype`` (a numpy dtype) the array must have exactly that type and be C-contiguous, and a CUDA tensor must
    live on GPU ``device`` (the context's device) when that is given -- otherwise TypeError / ValueError."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        if dtype is not None and a.dtype != np.dtype(dtype):
            raise TypeError(f"{what}: expected {np.dtype(dtype).name}, got {a.dtype.name}")
        if not a.flags["C_CONTIGUOUS"]:
            raise ValueError(f"{what}: array must be C-contiguous")
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        if dtype is not None and str(a.dtype) != _TORCH_NAMES.get(np.dtype(dtype).name):
            raise TypeError(f"{what}: expected {_TORCH_NAMES.get(np.dtype(dtype).name)}, got {a.dtype}")
        if not a.is_contiguous():
            raise ValueError(f"{what}: tensor must be contiguous")
        if device is not None and getattr(a, "is_cuda", False) and a.device.index != device:
            raise ValueError(f"{what}: tensor is on cuda:{a.device.index}, the context on cuda:{device}")
        return a.data_ptr()
    raise TypeError(type(a))
