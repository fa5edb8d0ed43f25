"""Overlay of the structured patch family by pairwise convex clipping (csrc/overlay_core.h + overlay.cu).

CPU: the header the CUDA kernels compile is built with the host compiler (csrc/overlay_host.cpp -> liboverlay_host.so,
-ffp-contract=off) and compared BIT FOR BIT with the NumPy generator mofem_b200/overlay_gen.py -- which tests/test_overlay_gen.py pins against the
reference's own plane sweep / rhoCalculation / triangulation (fixtures patch16, patch16_m3).
GPU: the kernels, through the C-ABI, against the same generator."""
import ctypes as C

import numpy as np
import pytest

from mofem_b200 import overlay_gen

ARRAYS = ("cell_elem", "cell_kind", "cell_tri_ptr", "tri_xy", "tri_rho")
CASES = [dict(n0=16), dict(n0=16, n_mesh=3), dict(n0=24, ny=16), dict(n0=16, ny=32, patch_rows=(1, 3)),
         dict(n0=64, ny=32, patch_rows=(2, 4), n_mesh=3), dict(n0=96, jitter=0.2), dict(n0=32, jitter=0.0)]


def _bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.int64) if a.dtype == np.float64 else a


def _host_lib():
    from mofem_b200 import build
    return C.CDLL(build.build_host())


@pytest.mark.parametrize("case", CASES, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items()))
def test_host_build_of_the_kernel_source_is_bit_identical_to_the_generator(case):
    L = _host_lib()
    fp, _, _, _, pm = overlay_gen.patch_family(return_meshes=True, **case)
    n_mesh = case.get("n_mesh", 2)
    b_lo, b_hi = case.get("patch_rows", (0, pm.Py))
    sz = (C.c_int64 * 4)()
    L.ovl_host_sizes(pm.nx, pm.ny, b_lo, b_hi, n_mesh, sz)
    n_reg, n_ring, n_piece, n_tri = list(sz)
    n_cell = n_reg + n_ring + n_piece
    assert (n_cell, n_tri) == (fp.n_cell, fp.n_tri)
    out = dict(cell_elem=np.empty((n_cell, n_mesh), np.int32), cell_kind=np.empty(n_cell, np.int32),
               cell_tri_ptr=np.empty(n_cell + 1, np.int64), tri_xy=np.empty((n_tri, 3, 2)), tri_rho=np.empty((n_tri, 3, n_mesh)))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    X0 = np.ascontiguousarray(pm.X0); X1 = np.ascontiguousarray(pm.X1)
    L.ovl_host_generate(pm.nx, pm.ny, b_lo, b_hi, n_mesh, p(X0), p(X1), *[p(out[k]) for k in ARRAYS])
    for k in ARRAYS:
        assert np.array_equal(_bits(out[k]), _bits(getattr(fp, k))), k


@pytest.mark.parametrize("case", CASES, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items()))
def test_native_engine_of_patch_family_is_bit_identical(case):
    """``patch_family(engine="native")`` (what the multi-GPU driver uses for its strips) against the NumPy engine: every
    array of the problem and the linear system."""
    a = overlay_gen.patch_family(**case)
    b = overlay_gen.patch_family(engine="native", **case)
    for k in a[0].ARRAYS:
        x, y = getattr(a[0], k), getattr(b[0], k)
        assert (x is None and y is None) or (x.dtype == y.dtype and x.shape == y.shape and np.array_equal(_bits(x), _bits(y))), k
    assert (a[0].solver, a[0].n_mesh, a[0].mesh_offset) == (b[0].solver, b[0].n_mesh, b[0].mesh_offset)
    assert all(np.array_equal(x, y) for x, y in zip(a[1:], b[1:]))


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES + [dict(n0=256)], ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items()))
def test_gpu_overlay_is_bit_identical_to_the_generator(gpu_ctx, case):
    fp, rhs, fixed, val = overlay_gen.patch_family(**case)
    g, rhs_g, fixed_g, val_g, ms = overlay_gen.patch_family_gpu(gpu_ctx, to_host=True, **case)
    for k in ARRAYS:
        assert np.array_equal(_bits(getattr(g, k)), _bits(getattr(fp, k))), k
    for k in ("node_xy", "elem_nodes", "elem_nn", "elem_mat", "elem_mesh", "mat_D"):
        assert np.array_equal(getattr(g, k), getattr(fp, k)), k
    assert np.array_equal(rhs, rhs_g) and np.array_equal(fixed, fixed_g) and np.array_equal(val, val_g)


@pytest.mark.gpu
def test_gpu_overlay_feeds_the_solver_from_device_memory(gpu_ctx):
    """Device-resident overlay arrays go straight into mofem_set_problem (zero-copy); same matrix as from the host arrays."""
    fp, rhs, fixed, val = overlay_gen.patch_family(64)
    gpu_ctx.set_problem(fp).symbolic().assemble()
    ref = gpu_ctx.csr()
    g, rhs_g, fixed_g, val_g, _ = overlay_gen.patch_family_gpu(gpu_ctx, 64)
    gpu_ctx.set_problem(g).symbolic().assemble()
    got = gpu_ctx.csr()
    assert all(np.array_equal(a, b) for a, b in zip(ref, got))
    gpu_ctx.set_system(rhs_g, fixed_g, val_g)
    u, info = gpu_ctx.solve(rtol=1e-10)
    assert info["converged"] == 1


@pytest.mark.gpu
def test_gpu_overlay_argument_errors(gpu_ctx):
    from mofem_b200 import _lib
    L = _lib.load()
    sz = (C.c_int64 * 5)()
    assert L.mofem_overlay_patch_sizes(12, 16, 0, 2, 2, sz) == _lib.E_ARG        # nx not a multiple of 8
    assert L.mofem_overlay_patch_sizes(16, 16, 1, 1, 2, sz) == _lib.E_ARG        # empty row range
    assert L.mofem_overlay_patch_sizes(16, 16, 0, 2, 4, sz) == _lib.E_ARG        # n_mesh
    assert L.mofem_overlay_patch_sizes(16, 16, 0, 2, 2, sz) == 0 and sz[4] == 592


def test_ear_rule_matches_the_python_restatement_on_random_polygons():
    """The C++ ear rule (explicit stack) against overlay_gen.ear_triangulate (the restatement of the reference's recursive
    simpleTriangulation that tests/test_overlay_gen.py checks) on random star-shaped polygons with reflex vertices."""
    L = _host_lib()
    rng = np.random.default_rng(7)
    n_reflex_cases = 0
    for trial in range(3000):
        n = int(rng.integers(3, 7))
        ang = np.sort(rng.uniform(0, 2 * np.pi, n))
        if np.min(np.diff(np.concatenate([ang, [ang[0] + 2 * np.pi]]))) < 0.15:
            continue
        rad = rng.uniform(0.25, 1.0, n)                           # star-shaped around the origin: simple, often non-convex
        pts = np.stack([rad * np.cos(ang), rad * np.sin(ang)], 1) + rng.uniform(-3, 3, 2)
        if trial % 5 == 0:                                        # some exactly horizontal edges / equal heights
            pts[1, 1] = pts[0, 1]
        want = overlay_gen.ear_triangulate([tuple(p) for p in pts.tolist()])
        out = np.zeros(12, np.int32)
        nt = L.ovl_host_ear(np.ascontiguousarray(pts).ctypes.data_as(C.c_void_p), n, out.ctypes.data_as(C.c_void_p))
        assert nt == len(want) == n - 2
        assert [tuple(t) for t in out[:3 * nt].reshape(-1, 3).tolist()] == [tuple(t) for t in want], (trial, pts)
        e1, e2 = np.roll(pts, -1, 0) - pts, np.roll(pts, -2, 0) - np.roll(pts, -1, 0)
        cross = e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0]
        n_reflex_cases += int((cross < 0).any())
    assert n_reflex_cases > 500
