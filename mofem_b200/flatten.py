"""Host-side flattening boundary: the reference's object graph -> SoA arrays.

The reference hands its solvers Python lists and a half-edge object graph
(``inputData`` from inputFunctions/inputF.py:165,204; ``polygons`` /
``overlappingMesh`` from computGeometry/planeSweepMeshes.py:354 and
meshTools/toDC.py:48-72).  The CUDA path consumes plain arrays.  This module
walks the graph once and produces a :class:`FlatProblem` whose *cell order is
the reference's emission order* (AMORE.py:42-123, 307-398, 729-821): first the
regular (non-overlapped) elements mesh by mesh, then the overlay cells in
``polygons`` order.

Rules replicated here (all host-side, O(cells)):
  * ``getCoord`` (AMORE.py:1208-1238): missing mid-side nodes = edge midpoint,
    missing centre node = mean of the four corners.
  * ``getRho`` (AMORE.py:1250-1252) and ``getCoordFromHalfEdge``
    (stiffnessMatrix.py:224-236): triangle vertices are
    ``edge.origin, edge.next.origin, edge.next.next.origin``.
  * ``coordCurvedTriangle`` (AMORE.py:1148-1199): a triangle edge that coincides
    exactly with a curved edge of a higher-order incident element picks up that
    element's mid-node; the scan stops at the first low-order incident element.
  * untriangulated cells use the first non-null incident element
    (AMORE.py:60-66).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np

SOLVER_LOWER = 1      # AMORE.lowerOrderAMORE / traditionalElement.lowerOrderFE
SOLVER_QUADRATIC = 2  # AMORE.quadraticAMORE  / traditionalElement.quadraticFE
SOLVER_ICM = 3        # AMORE.quadraticICMAMORE / traditionalElement.ICMFE

CELL_REGULAR = 0
CELL_OVERLAY = 1


@dataclass
class FlatProblem:
    """Flattened assembly problem (host numpy arrays, C-contiguous)."""
    solver: int
    n_mesh: int
    node_xy: np.ndarray        # f64 [N,2]  (NaN for nodes no element defines)
    elem_nodes: np.ndarray     # i32 [E,9]  (-1 padded)
    elem_nn: np.ndarray        # i32 [E]
    elem_mat: np.ndarray       # i32 [E]
    elem_mesh: np.ndarray      # i32 [E]
    mat_D: np.ndarray          # f64 [nmat,3,3]
    cell_elem: np.ndarray      # i32 [C,M]
    cell_kind: np.ndarray      # i32 [C]
    cell_tri_ptr: np.ndarray   # i64 [C+1]
    tri_xy: np.ndarray         # f64 [T,3,2]
    tri_rho: np.ndarray        # f64 [T,3,M]
    tri_mid: Optional[np.ndarray] = None     # f64 [T,3,2]
    tri_curved: Optional[np.ndarray] = None  # u8  [T]
    n_mesh_regular: int = 0    # leading pseudo-cells that are regular mesh elements
    mesh_offset: list = field(default_factory=list)  # global element id = mesh_offset[i]+j

    @property
    def n_node(self): return int(self.node_xy.shape[0])
    @property
    def n_elem(self): return int(self.elem_nn.shape[0])
    @property
    def n_cell(self): return int(self.cell_kind.shape[0])
    @property
    def n_tri(self): return int(self.tri_xy.shape[0])
    @property
    def n_dof(self): return 2 * self.n_node

    ARRAYS = ("node_xy", "elem_nodes", "elem_nn", "elem_mat", "elem_mesh", "mat_D", "cell_elem",
              "cell_kind", "cell_tri_ptr", "tri_xy", "tri_rho", "tri_mid", "tri_curved")

    def to_npz_dict(self):
        d = {k: getattr(self, k) for k in self.ARRAYS if getattr(self, k) is not None}
        d["meta"] = np.array([self.solver, self.n_mesh, self.n_mesh_regular], dtype=np.int64)
        d["mesh_offset"] = np.array(self.mesh_offset, dtype=np.int64)
        return d

    @classmethod
    def from_npz_dict(cls, d, prefix=""):
        g = lambda k: (np.ascontiguousarray(d[prefix + k]) if (prefix + k) in d else None)
        meta = d[prefix + "meta"]
        return cls(solver=int(meta[0]), n_mesh=int(meta[1]), n_mesh_regular=int(meta[2]),
                   mesh_offset=[int(x) for x in d[prefix + "mesh_offset"]],
                   **{k: g(k) for k in cls.ARRAYS})

    def validate(self):
        """Input checks with the reference's error behaviour (ValueError)."""
        nn = self.elem_nn
        if not np.isin(nn, (3, 4, 6, 9)).all():
            raise ValueError("Wrong element numbering!")
        used = np.unique(self.cell_elem[self.cell_elem >= 0])
        if self.solver == SOLVER_LOWER and not np.isin(nn[used], (3, 4)).all():
            raise ValueError("Wrong element numbering!")  # AMORE.py:51,74
        if self.cell_tri_ptr[-1] != self.n_tri:
            raise ValueError("cell_tri_ptr does not cover the triangle list")


def material_matrix(material, problem_type):
    """``twoDMaterialMatrix`` (AMORE.py:1128-1146): [E, nu] -> 3x3 D."""
    E, nu = float(material[0]), float(material[1])
    D = np.zeros((3, 3))
    if problem_type == 1:    # plane stress
        D[0, 0] = D[1, 1] = E / (1.0 - nu ** 2)
        D[0, 1] = D[1, 0] = E * nu / (1.0 - nu ** 2)
        D[2, 2] = E / 2.0 / (1.0 + nu)
    elif problem_type == 2:  # plane strain
        cc = E * (1.0 - nu) / (1.0 + nu) / (1.0 - 2.0 * nu)
        D[0, 0] = D[1, 1] = cc
        D[0, 1] = D[1, 0] = cc * nu / (1.0 - nu)
        D[2, 2] = cc * (1.0 - 2.0 * nu) / (2.0 * (1.0 - nu))
    else:
        raise ValueError("No such a problem type!")
    return D


_NEXT3 = (1, 2, 0)
_NEXT4 = (1, 2, 3, 0)


def _fill_element_nodes(node_xy, coordinates, numbering):
    """Apply the ``getCoord`` completion rule for one element (AMORE.py:1208)."""
    n = len(numbering)
    if n not in (3, 4, 6, 9):
        raise ValueError
    nc = 3 if n in (3, 6) else 4
    nxt = _NEXT3 if nc == 3 else _NEXT4
    for a in range(nc):
        node_xy[numbering[a]] = coordinates[numbering[a]]
    if n in (6, 9):
        for a in range(nc, 2 * nc):
            nd = numbering[a]
            if coordinates[nd]:
                node_xy[nd] = coordinates[nd]
            else:
                node_xy[nd] = 0.5 * (np.array(coordinates[numbering[a - nc]])
                                     + np.array(coordinates[numbering[nxt[a - nc]]]))
    if n == 9:
        nd = numbering[8]
        if coordinates[nd]:
            node_xy[nd] = coordinates[nd]
        else:
            node_xy[nd] = 0.25 * (np.array(coordinates[numbering[0]]) + np.array(coordinates[numbering[1]])
                                  + np.array(coordinates[numbering[2]]) + np.array(coordinates[numbering[3]]))


def _curved_triangle(edge, incident_elements, coordinates):
    """``coordCurvedTriangle`` (AMORE.py:1148-1199).  Returns (xy[3][2], mid[3][2] | None)."""
    start = edge
    verts = [start, start.next, start.next.next]
    xy = [[float(h.origin.x), float(h.origin.y)] for h in verts]
    for el in incident_elements:
        if not el:
            continue
        n = len(el.numbering)
        if n <= 4:
            return xy, None          # quirk: the scan ends at the first low-order element
        if n == 9:
            nc, nxt = 4, _NEXT4
        elif n == 6:
            nc, nxt = 3, _NEXT3
        else:
            raise ValueError
        curved_edges = [j for j in range(nc) if coordinates[el.numbering[j + nc]] is not None]
        hit = [None, None, None]
        for t, h in enumerate(verts):
            for k in curved_edges:
                a = coordinates[el.numbering[k]]
                b = coordinates[el.numbering[nxt[k]]]
                if (h.origin.x == a[0] and h.origin.y == a[1]
                        and h.destination.x == b[0] and h.destination.y == b[1]):
                    hit[t] = k
                    break
        if any(k is not None for k in hit):
            mid = []
            for t, h in enumerate(verts):
                if hit[t] is not None:
                    c = coordinates[el.numbering[hit[t] + nc]]
                    mid.append([float(c[0]), float(c[1])])
                else:
                    mid.append([float((h.origin.x + h.destination.x) / 2),
                                float((h.origin.y + h.destination.y) / 2)])
            return xy, mid
    return xy, None


def flatten(input_data, overlapping_mesh=None, polygons=None, solver=None) -> FlatProblem:
    """Flatten ``inputData`` (+ the overlay object graph) into a :class:`FlatProblem`.

    ``overlapping_mesh``/``polygons`` are ``None`` for the single-mesh solvers of
    linearSolvers/traditionalElement.py.
    """
    parameters = input_data[0]
    meshes = input_data[3]
    material_mesh_list = input_data[4]
    coordinates = input_data[5]
    material_list = input_data[6]
    if solver is None:
        solver = int(parameters[0])
    M = len(meshes)
    N = len(coordinates)

    mesh_offset = [0]
    for m in meshes:
        mesh_offset.append(mesh_offset[-1] + len(m))
    E = mesh_offset[-1]

    node_xy = np.full((N, 2), np.nan)
    elem_nodes = np.full((E, 9), -1, dtype=np.int32)
    elem_nn = np.zeros(E, dtype=np.int32)
    elem_mat = np.zeros(E, dtype=np.int32)
    elem_mesh = np.zeros(E, dtype=np.int32)

    def put_element(i, j, numbering):
        e = mesh_offset[i] + j
        n = len(numbering)
        if n not in (3, 4, 6, 9):
            raise ValueError("Wrong element numbering!")
        elem_nodes[e, :n] = numbering
        elem_nn[e] = n
        elem_mat[e] = material_mesh_list[i][j]
        elem_mesh[e] = i
        _fill_element_nodes(node_xy, coordinates, numbering)
        return e

    cell_elem, cell_kind, tri_ptr = [], [], [0]
    tri_xy, tri_rho, tri_mid, tri_curved = [], [], [], []

    # regular elements, mesh by mesh (AMORE.py:42-56; traditionalElement.py:36-48)
    for i in range(M):
        for j in range(len(meshes[i])):
            if meshes[i][j] is not None:
                e = put_element(i, j, meshes[i][j])
                row = [-1] * M
                row[i] = e
                cell_elem.append(row)
                cell_kind.append(CELL_REGULAR)
                tri_ptr.append(tri_ptr[-1])
    n_mesh_regular = len(cell_kind)

    if overlapping_mesh is not None:
        for i in range(M):
            if overlapping_mesh[i] is None:
                continue
            for el in overlapping_mesh[i]:
                put_element(el.position[0], el.position[1], el.numbering)

    for poly in (polygons or []):
        row = [-1] * M
        if poly.triangles is None:
            for el in poly.incidentElements:
                if el:
                    row[el.position[0]] = mesh_offset[el.position[0]] + el.position[1]
                    break
            cell_elem.append(row)
            cell_kind.append(CELL_REGULAR)
            tri_ptr.append(tri_ptr[-1])
            continue
        for j in range(M):
            el = poly.incidentElements[j]
            if el:
                row[j] = mesh_offset[el.position[0]] + el.position[1]
        cell_elem.append(row)
        cell_kind.append(CELL_OVERLAY)
        for k in poly.triangles:
            verts = (k, k.next, k.next.next)
            if solver == SOLVER_LOWER:
                xy, mid = [[float(h.origin.x), float(h.origin.y)] for h in verts], None
            else:
                xy, mid = _curved_triangle(k, poly.incidentElements, coordinates)
            tri_xy.append(xy)
            tri_rho.append([[float(h.origin.rho[j]) for j in range(M)] for h in verts])
            tri_curved.append(0 if mid is None else 1)
            tri_mid.append(mid if mid is not None else [[0.0, 0.0]] * 3)
        tri_ptr.append(tri_ptr[-1] + len(poly.triangles))

    problem_type = int(parameters[1])
    mat_D = np.array([material_matrix(m, problem_type) for m in material_list]).reshape(-1, 3, 3)
    T = len(tri_xy)
    any_curved = bool(np.any(tri_curved)) if T else False
    fp = FlatProblem(
        solver=int(solver), n_mesh=M, node_xy=node_xy, elem_nodes=elem_nodes, elem_nn=elem_nn,
        elem_mat=elem_mat, elem_mesh=elem_mesh, mat_D=np.ascontiguousarray(mat_D),
        cell_elem=np.array(cell_elem, dtype=np.int32).reshape(-1, M),
        cell_kind=np.array(cell_kind, dtype=np.int32),
        cell_tri_ptr=np.array(tri_ptr, dtype=np.int64),
        tri_xy=np.array(tri_xy, dtype=np.float64).reshape(T, 3, 2),
        tri_rho=np.array(tri_rho, dtype=np.float64).reshape(T, 3, M),
        tri_mid=np.array(tri_mid, dtype=np.float64).reshape(T, 3, 2) if any_curved else None,
        tri_curved=np.array(tri_curved, dtype=np.uint8) if any_curved else None,
        n_mesh_regular=n_mesh_regular, mesh_offset=mesh_offset)
    return fp
