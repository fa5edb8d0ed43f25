"""The synthetic patch-family generator against the reference's own plane sweep (fixture patch16.npz, produced by
tests/golden/make_golden.py from the same deck) and against the reference's triangulation rule.  CPU only."""
import numpy as np
import scipy.sparse as sp

from conftest import coo_to_csr, load_golden, row_scaled_error
from mofem_b200 import overlay_gen
from oracle import oracle


def _canon_tris(t):
    t = np.round(t.reshape(-1, 3, 2) / 1e-13).astype(np.int64)
    return sorted(tuple(sorted(map(tuple, x))) for x in t)


import pytest


@pytest.mark.parametrize("engine", ["numpy", "native"])
@pytest.mark.parametrize("name,n_mesh", [("patch16", 2), ("patch16_m3", 3)])
def test_generator_reproduces_the_reference_overlay(name, n_mesh, engine):
    """Both engines (NumPy; the overlay core of the CUDA kernels built for the host) against the reference's own plane sweep."""
    fpr, d = load_golden(name)
    fp, rhs, fixed, val = overlay_gen.patch_family(16, n_mesh=n_mesh, engine=engine)
    assert fp.n_mesh == fpr.n_mesh == n_mesh
    assert np.array_equal(fp.node_xy, fpr.node_xy)
    assert fp.n_cell == fpr.n_cell and fp.n_tri == fpr.n_tri
    assert _canon_tris(fp.tri_xy) == _canon_tris(fpr.tri_xy)          # same triangle set as the sweep + ear rule
    I, J, V, _, _ = oracle.assemble_coo(fp)
    assert len(V) == len(d["coo_V"])
    n = fp.n_dof
    K = coo_to_csr(n, I, J, V)
    Kr = sp.csr_matrix((d["csr_data"], d["csr_indices"], d["csr_indptr"]), shape=(n, n))
    assert np.array_equal(K.indptr, Kr.indptr) and np.array_equal(K.indices, Kr.indices)
    assert row_scaled_error(K, Kr) <= 1e-12
    # the linear system: same clamped edge and load as the deck
    cons = d["constraints"]
    f2 = np.zeros(n, np.uint8)
    for c in cons:
        f2[2 * int(c[0])] = c[1]; f2[2 * int(c[0]) + 1] = c[2]
    assert np.array_equal(fixed, f2)
    free = fixed == 0
    assert np.allclose(rhs[free], d["red_rhs"], rtol=0, atol=0)


def test_counts_per_period():
    # per period 100 two-element cells, 20 ring cells, 264 triangles, 28672 COO triplets
    fp, *_ = overlay_gen.patch_family(24, jitter=0.0)
    P = 3
    two = (fp.cell_elem >= 0).sum(1) == 2
    assert two.sum() == 100 * P * P
    ring = (fp.cell_kind == 1) & ~two
    assert ring.sum() == 20 * P * P
    assert fp.n_tri == 264 * P * P
    assert oracle.assemble_coo(fp)[2].size == 28672 * P * P + 64 * (24 * 24 - 64 * P * P)
    # triangles tile the patches + rings: area = 36 mesh-0 elements per period
    a = fp.tri_xy
    area = 0.5 * ((a[:, 1, 0] - a[:, 0, 0]) * (a[:, 2, 1] - a[:, 0, 1]) - (a[:, 2, 0] - a[:, 0, 0]) * (a[:, 1, 1] - a[:, 0, 1]))
    assert (area > 0).all()
    assert abs(area.sum() - 36 * P * P / 24 ** 2) < 1e-12
    # partition of unity
    assert np.allclose(fp.tri_rho.sum(-1), 1.0, atol=1e-14)


def test_ear_rule_on_simple_polygons():
    sq = [(0.0, 0.0), (1.0, 0.0), (1.0, 1.0), (0.0, 1.0)]
    # up-left-most vertex is 3: triangles (U.next, opposite, U.prev) then (U.prev, U, U.next)
    assert overlay_gen.ear_triangulate(sq) == [(0, 1, 2), (2, 3, 0)]
    # L-shape: the reflex vertex lies inside the first candidate ear -> split through it
    L = [(0.0, 0.0), (2.0, 0.0), (2.0, 1.0), (1.0, 1.0), (1.0, 2.0), (0.0, 2.0)]
    tris = overlay_gen.ear_triangulate(L)
    assert len(tris) == 4
    area = 0.0
    for t in tris:
        (x0, y0), (x1, y1), (x2, y2) = (L[k] for k in t)
        a = 0.5 * ((x1 - x0) * (y2 - y0) - (x2 - x0) * (y1 - y0))
        assert a > 0
        area += a
    assert abs(area - 3.0) < 1e-15


def test_jitter_is_seeded():
    a = overlay_gen.patch_meshes(16).X0
    b = overlay_gen.patch_meshes(16).X0
    c = overlay_gen.patch_meshes(16, seed=1).X0
    assert np.array_equal(a, b) and not np.array_equal(a, c)


@pytest.mark.parametrize("name,family,n0", [("patch16", "quad", 16), ("patch16_tri", "tri", 16), ("patch16_m3s", "m3s", 16),
                                            ("patch32_m3s", "m3s", 32)])
def test_general_overlay_reproduces_the_reference_sweep(name, family, n0):
    """csrc/overlay_general.h (clipping per mesh-0 element) against the reference's plane sweep + ear rule + rhoCalculation +
    AMORE assembly on the same deck: BASELINE config 3 variant b (tri3 patches) and config 4 (staggered third mesh: cells with
    1, 2 and 3 incident elements)."""
    fpr, d = load_golden(name)
    fp, rhs, fixed, val = overlay_gen.general_family(n0, family)
    assert np.array_equal(fp.node_xy, fpr.node_xy) and np.array_equal(fp.elem_nodes, fpr.elem_nodes)
    assert fp.n_cell == fpr.n_cell and fp.n_tri == fpr.n_tri
    assert np.array_equal(np.bincount((fp.cell_elem >= 0).sum(1)), np.bincount((fpr.cell_elem >= 0).sum(1)))
    assert _canon_tris(fp.tri_xy) == _canon_tris(fpr.tri_xy)
    I, J, V, _, _ = oracle.assemble_coo(fp)
    n = fp.n_dof
    K = coo_to_csr(n, I, J, V)
    Kr = sp.csr_matrix((d["csr_data"], d["csr_indices"], d["csr_indptr"]), shape=(n, n))
    assert np.array_equal(K.indptr, Kr.indptr) and np.array_equal(K.indices, Kr.indices)
    assert row_scaled_error(K, Kr) <= 1e-12
    free = fixed == 0
    assert np.allclose(rhs[free], d["red_rhs"], rtol=0, atol=0)
    assert np.allclose(fp.tri_rho.sum(-1), 1.0, atol=1e-14)


def test_general_overlay_three_element_cells():
    """The staggered family really has cells with 1, 2 and 3 incident elements (356 / 2284 / 378 overlay cells of each kind in
    SURVEY.md section 8d's probe at n0 = 32 -- ours: the fixture's counts) and its triangles tile the touched area exactly."""
    fp, *_ = overlay_gen.general_family(32, "m3s")
    m = (fp.cell_elem >= 0).sum(1)[fp.cell_kind == 1]
    assert (np.bincount(m)[1:] > 0).all() and m.max() == 3
    a = fp.tri_xy
    area = 0.5 * ((a[:, 1, 0] - a[:, 0, 0]) * (a[:, 2, 1] - a[:, 0, 1]) - (a[:, 2, 0] - a[:, 0, 0]) * (a[:, 1, 1] - a[:, 0, 1]))
    assert (area > 0).all()
    # overlay cells + regular cells tile the unit square: touched area = 1 - regular elements' area
    reg = fp.cell_elem[fp.cell_kind == 0, 0]
    xy = fp.node_xy[fp.elem_nodes[reg, :4]]
    quad_area = 0.5 * np.abs(((xy[:, 2, 0] - xy[:, 0, 0]) * (xy[:, 3, 1] - xy[:, 1, 1]) - (xy[:, 3, 0] - xy[:, 1, 0]) * (xy[:, 2, 1] - xy[:, 0, 1])))
    assert abs(area.sum() + quad_area.sum() - 1.0) < 1e-12


def test_general_overlay_strips_and_threads():
    """Element-row strips concatenate to the whole problem, independent of the thread count."""
    whole, *_ = overlay_gen.general_family(16, "m3s", threads=1)
    a, *_ = overlay_gen.general_family(16, "m3s", elem_rows=(0, 7), threads=3)
    b, *_ = overlay_gen.general_family(16, "m3s", elem_rows=(7, 16), threads=2)
    assert np.array_equal(np.concatenate([a.cell_elem, b.cell_elem]), whole.cell_elem)
    assert np.array_equal(np.concatenate([a.tri_xy, b.tri_xy]), whole.tri_xy)
    assert np.array_equal(np.concatenate([a.tri_rho, b.tri_rho]), whole.tri_rho)
