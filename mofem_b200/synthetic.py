"""Synthetic flattened problems (the reference ships no generator).

* :func:`zoo_problem` -- a small randomised problem that exercises every
  element type pair, every quadrature rule of stiffnessMatrix.py:104-106 /
  137-144, straight and curved integration triangles and all regular-element
  rules.  Used for parity tests against the oracle / reference.
* :func:`patch_family` lives in :mod:`mofem_b200.overlay_gen`.
"""
from __future__ import annotations

import numpy as np

from .flatten import CELL_OVERLAY, CELL_REGULAR, FlatProblem, material_matrix

_QUAD = np.array([[-1.0, -1.0], [1.0, -1.0], [1.0, 1.0], [-1.0, 1.0]])
_TRI = np.array([[-1.2, -1.0], [1.4, -0.9], [-0.1, 1.5]])


def _element_coords(n, rng, centre, size, distort):
    """Random, valid (positive Jacobian) element of ``n`` nodes around ``centre``."""
    if n in (4, 9):
        c = _QUAD * size + rng.uniform(-distort, distort, (4, 2)) * size
        if n == 9:
            mids = 0.5 * (c + c[[1, 2, 3, 0]]) + rng.uniform(-0.3, 0.3, (4, 2)) * distort * size
            cen = c.mean(0) + rng.uniform(-0.3, 0.3, 2) * distort * size
            c = np.vstack([c, mids, cen])
    else:
        c = _TRI * size + rng.uniform(-distort, distort, (3, 2)) * size
        if n == 6:
            mids = 0.5 * (c + c[[1, 2, 0]]) + rng.uniform(-0.3, 0.3, (3, 2)) * distort * size
            c = np.vstack([c, mids])
    return c + centre


def zoo_problem(seed=0, solver=2, n_mesh=3, cells_per_combo=2, distort=0.15, n_regular=3) -> FlatProblem:
    """Randomised all-types problem.  ``solver`` 1 restricts to tri3/quad4 and
    straight triangles (AMORE.lowerOrderAMORE); 2/3 add tri6/quad9 and curved
    triangles (quadraticAMORE / quadraticICMAMORE)."""
    rng = np.random.default_rng(seed)
    types = (3, 4) if solver == 1 else (3, 4, 6, 9)
    M = n_mesh
    nodes, elems, emat, emesh = [], [], [], []
    cell_elem, cell_kind, tri_ptr = [], [], [0]
    tri_xy, tri_rho, tri_mid, tri_curved = [], [], [], []
    materials = [[1000.0, 0.3], [2.5e4, 0.2], [7.0, 0.45]]

    def add_element(n, mesh, centre, size):
        c = _element_coords(n, rng, centre, size, distort)
        base = len(nodes)
        nodes.extend(c.tolist())
        num = list(range(base, base + n))
        perm = rng.permutation(len(num))  # scramble global numbering so DOF order is not trivially sorted
        _ = perm
        elems.append(num + [-1] * (9 - n))
        emat.append(int(rng.integers(0, len(materials))))
        emesh.append(mesh)
        return len(elems) - 1

    # regular cells of every type (AMORE.py:42-56 and the untriangulated branch :60-81)
    for n in types:
        for _ in range(n_regular):
            centre = rng.uniform(-50, 50, 2)
            mesh = int(rng.integers(0, M))
            e = add_element(n, mesh, centre, rng.uniform(0.3, 2.0))
            row = [-1] * M
            row[mesh] = e
            cell_elem.append(row); cell_kind.append(CELL_REGULAR); tri_ptr.append(tri_ptr[-1])
    n_mesh_regular = len(cell_kind) // 2  # first half play "mesh regular", rest "untriangulated cells"

    combos = [(a,) for a in types] + [(a, b) for a in types for b in types]
    if M >= 3:
        combos += [(types[i % len(types)], types[(i // 2) % len(types)], types[(i + 1) % len(types)])
                   for i in range(len(types) * 2)]
    for combo in combos:
        for _ in range(cells_per_combo):
            centre = rng.uniform(-50, 50, 2)
            size = rng.uniform(0.5, 3.0)
            slots = sorted(rng.choice(M, size=len(combo), replace=False).tolist())
            row = [-1] * M
            for n, mesh in zip(combo, slots):
                row[mesh] = add_element(n, mesh, centre, size)
            cell_elem.append(row); cell_kind.append(CELL_OVERLAY)
            T = int(rng.integers(1, 4))
            for _t in range(T):
                # a small counter-clockwise triangle well inside all incident elements
                ang = np.sort(rng.uniform(0, 2 * np.pi, 3))
                rad = rng.uniform(0.08, 0.3, 3) * size
                xy = centre + np.stack([rad * np.cos(ang), rad * np.sin(ang)], 1)
                a2 = (xy[1, 0] - xy[0, 0]) * (xy[2, 1] - xy[0, 1]) - (xy[2, 0] - xy[0, 0]) * (xy[1, 1] - xy[0, 1])
                if a2 < 0:
                    xy = xy[[0, 2, 1]]
                rho = np.zeros((3, M))
                raw = rng.uniform(0.05, 1.0, (3, len(slots)))
                rho[:, slots] = raw / raw.sum(1, keepdims=True)
                tri_xy.append(xy); tri_rho.append(rho)
                curved = solver != 1 and rng.random() < 0.4
                mid = 0.5 * (xy + xy[[1, 2, 0]])
                if curved:
                    mid = mid + rng.uniform(-0.02, 0.02, (3, 2)) * size
                tri_mid.append(mid); tri_curved.append(1 if curved else 0)
            tri_ptr.append(tri_ptr[-1] + T)

    # duplicate a few cells (same elements, new triangles) so CSR entries receive several contributions
    C0 = len(cell_kind)
    for c in range(n_mesh_regular * 2, C0, 3):
        cell_elem.append(list(cell_elem[c])); cell_kind.append(CELL_OVERLAY)
        t0 = tri_ptr[c]
        tri_xy.append(tri_xy[t0][[1, 2, 0]] if isinstance(tri_xy[t0], np.ndarray) else tri_xy[t0])
        tri_rho.append(tri_rho[t0][[1, 2, 0]])
        tri_mid.append(tri_mid[t0][[1, 2, 0]]); tri_curved.append(tri_curved[t0])
        tri_ptr.append(tri_ptr[-1] + 1)

    T = len(tri_xy)
    problem_type = 1 if seed % 2 == 0 else 2
    mat_D = np.array([material_matrix(m, problem_type) for m in materials])
    any_curved = any(tri_curved)
    return FlatProblem(
        solver=solver, n_mesh=M, node_xy=np.array(nodes, dtype=np.float64).reshape(-1, 2),
        elem_nodes=np.array(elems, dtype=np.int32), elem_nn=np.array([sum(1 for v in e if v >= 0) for e in elems], dtype=np.int32),
        elem_mat=np.array(emat, dtype=np.int32), elem_mesh=np.array(emesh, dtype=np.int32), mat_D=mat_D,
        cell_elem=np.array(cell_elem, dtype=np.int32).reshape(-1, M), cell_kind=np.array(cell_kind, dtype=np.int32),
        cell_tri_ptr=np.array(tri_ptr, dtype=np.int64),
        tri_xy=np.array(tri_xy, dtype=np.float64).reshape(T, 3, 2), tri_rho=np.array(tri_rho, dtype=np.float64).reshape(T, 3, M),
        tri_mid=np.array(tri_mid, dtype=np.float64).reshape(T, 3, 2) if any_curved else None,
        tri_curved=np.array(tri_curved, dtype=np.uint8) if any_curved else None,
        n_mesh_regular=n_mesh_regular, mesh_offset=[0])
