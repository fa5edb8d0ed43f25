"""Drop-in entry points: same names, positional arguments, side effects and return values as the reference's
``linearSolvers/AMORE.py`` and ``linearSolvers/traditionalElement.py``.

    AMORE side           lowerOrderAMORE / quadraticAMORE / quadraticICMAMORE (inputData, overlappingMesh, polygons,
                         nodeElementsPartial) -> (displacement (2N,1), energy (1,1)[, materialMatrix])
    single-mesh side     lowerOrderFE / ICMFE / quadraticFE (inputData) -> (displacement, energy)

What runs where
  * GPU (through the C-ABI, :mod:`mofem_b200.device`): the element / overlay-cell loops, COO->CSR, the constraint
    elimination (as a mask), the linear solve (Jacobi-PCG instead of SuperLU) and the energy.
  * host, O(boundary): flattening of the object graph (:mod:`mofem_b200.flatten`), the force vector (edge tractions,
    ``rho``-weighted on overlapped edges, AMORE.py:134-208 / 409-630 / 832-1053; traditionalElement.py:59-89, 209-239,
    362-403) and the fixed-DOF list (AMORE.py:213-226).  The force loop below follows the reference's branch
    structure and floating-point expression order, so ``f`` is reproduced to the last bit.

Mirrored behaviour: the constraint list ``inputData[-3]`` is sorted IN PLACE (AMORE.py:219); the phase timings are
printed with the reference's wording; errors are ``ValueError("Wrong element numbering!")`` / ``AssertionError``.
There is no CPU fallback: without the CUDA library and a GPU these functions raise.
"""
from __future__ import annotations

import math
import time
import warnings

import numpy as np

from . import flatten as _fl
from .device import Context

# solver options (module level, like the reference has none: defaults reproduce its results)
RTOL = 1e-12          # PCG stop: true relative residual (or the fp64 attainable floor, see csrc/pcg.cu)
MAXIT = 500000
DEVICE = 0
_ctx = None
last_solve_info = None   # dict of the last PCG run (iterations, residuals), for callers that want it


def _context():
    global _ctx
    if _ctx is None:
        _ctx = Context(DEVICE)
    return _ctx


# ---------------------------------------------------------------------------------------------------------------
# quadrature and 1-D shape functions used by the force term
# ---------------------------------------------------------------------------------------------------------------
def _gauss_line(n):
    """otherFunctions/numericalIntegration.py:116-134 (gaussQuad), literals verbatim."""
    if n == 2:
        return [-0.577350269189626, 0.577350269189626], [1.0, 1.0]
    if n == 3:
        return [-0.774596669241483, 0.0, 0.774596669241483], [5.0 / 9, 8.0 / 9, 5.0 / 9]
    if n == 5:
        return ([-0.906179845938664, -0.538469310105683, 0.0, 0.538469310105683, 0.906179845938664],
                [0.236926885056189, 0.478628670499366, 128.0 / 225, 0.478628670499366, 0.236926885056189])
    raise ValueError("Wrong number of quadrature points!")


def _len(a, b):
    return ((a[0] - b[0]) ** 2 + (a[1] - b[1]) ** 2) ** 0.5          # lenEdge, AMORE.py:1240


def _lin(r):
    return 0.5 * (1.0 - r), 0.5 * (1.0 + r)                            # shapeFunction.oneDLinear:5


def _quad(r, coord):
    """shapeFunction.oneDQuadratic:15 -> (N1, N2, N3), |dx/dr|."""
    n1 = 0.5 * (1.0 - r) - 0.5 * (1.0 - r ** 2)
    n2 = 0.5 * (1.0 + r) - 0.5 * (1.0 - r ** 2)
    n3 = 1.0 - r ** 2
    v = np.array([[-0.5 + r, 0.5 + r, -2.0 * r]]) @ coord
    return (n1, n2, n3), (v[0, 0] ** 2 + v[0, 1] ** 2) ** 0.5


def _apply(ns, force):
    """``Nmat.T @ force`` for the block-diagonal 1-D shape matrices: (2k,1) with entries N_a*fx, N_a*fy."""
    out = np.empty((2 * len(ns), 1))
    for a, n in enumerate(ns):
        out[2 * a, 0] = n * force[0, 0]
        out[2 * a + 1, 0] = n * force[1, 0]
    return out


def _count_edges(edge):
    n, e = 1, edge.next
    while e is not edge:
        e = e.next
        n += 1
    return n


def _cross(x1, y1, x2, y2):
    return x1 * y2 - x2 * y1


def _triangle_on_boundary(edge, c1, c2):
    """AMORE.isTriangleOnBoundary:1254 -- the triangle edge lying on the straight segment c1-c2, or None."""
    length = math.sqrt((c1[0] - c2[0]) ** 2 + (c1[1] - c2[1]) ** 2)
    l1 = math.sqrt((edge.origin.x - c2[0]) ** 2 + (edge.origin.y - c2[1]) ** 2)
    for _ in range(_count_edges(edge)):
        l2 = math.sqrt((edge.destination.x - c2[0]) ** 2 + (edge.destination.y - c2[1]) ** 2)
        if abs(_cross(edge.origin.x - c2[0], edge.origin.y - c2[1], c1[0] - c2[0], c1[1] - c2[1])) <= 10 ** (-13) * length * l1 \
                and abs(_cross(edge.destination.x - c2[0], edge.destination.y - c2[1], c1[0] - c2[0], c1[1] - c2[1])) \
                <= 10 ** (-13) * length * l2:
            return edge
        l1 = l2
        edge = edge.next
    return None


def _triangle_containing_boundary(edge, c1, c2):
    """AMORE.isTriangleContainingBoundary:1274 -- exact end-point match (possibly reversed -> twin)."""
    for _ in range(_count_edges(edge)):
        if edge.origin.x == c1[0] and edge.origin.y == c1[1] and edge.destination.x == c2[0] and edge.destination.y == c2[1]:
            return edge
        if edge.origin.x == c2[0] and edge.origin.y == c2[1] and edge.destination.x == c1[0] and edge.destination.y == c1[1]:
            return edge.twin
        edge = edge.next
    return None


def _mid_node(numbering, n1, n2):
    """traditionalElement.findMidNode:479."""
    k = len(numbering)
    if k not in (6, 9):
        return None
    nc = 3 if k == 6 else 4
    for i in range(nc):
        a, b = numbering[i], numbering[(i + 1) % nc]
        if (a == n1 and b == n2) or (a == n2 and b == n1):
            return numbering[i + nc]
    return None


def _node_elements(coordinates, meshes):
    """toDC.nodeElementList:5 (corner nodes only)."""
    out = [[] for _ in coordinates]
    for i, mesh in enumerate(meshes):
        for j, el in enumerate(mesh):
            if el is None:
                continue
            if len(el) in (4, 9):
                nc = 4
            elif len(el) in (3, 6):
                nc = 3
            else:
                raise ValueError
            for k in range(nc):
                out[el[k]].append((i, j))
    return out


def _custom_force(x, y):
    import importlib
    try:
        cforce = importlib.import_module("cForce")          # the user's module of the reference (repo root)
    except ImportError as e:
        raise ImportError("customised boundary force requested but no cForce module is importable") from e
    return cforce.customizedForce(x, y)


def _edge3(coordinates, n1, n2, n3):
    coord = np.array([coordinates[n1], coordinates[n2], [0.0, 0.0]])
    if coordinates[n3]:
        coord[2, :] = np.array(coordinates[n3])
    else:
        coord[2, :] = 0.5 * (coord[0, :] + coord[1, :])
    return coord


# ---------------------------------------------------------------------------------------------------------------
# force vectors
# ---------------------------------------------------------------------------------------------------------------
def _straight_piece(floc, edge, pos, wei, coord, mesh_id, shape_at, force_at):
    """A straight triangle edge lying on the loaded element edge: rho-weighted 2-node rule (AMORE.py:176-195)."""
    rho = np.array([edge.origin.rho[mesh_id], edge.destination.rho[mesh_id]])
    ce = np.array([[edge.origin.x, edge.origin.y], [edge.destination.x, edge.destination.y]])
    length = _len(ce[0], ce[1])
    for l in range(len(pos)):
        n2 = np.array([[0.5 * (1.0 - pos[l]), 0.5 * (1.0 + pos[l])]])
        rho_v = (n2 @ rho)[0]
        xy = (n2 @ ce).reshape(-1)
        iso = 2.0 * _len(coord[0], xy) / _len(coord[0], coord[1]) - 1.0        # invLine:7
        ns = shape_at(iso)
        floc += 0.5 * wei[l] * length * rho_v * _apply(ns, force_at(ns, xy))
    return floc


def _force_amore(input_data, overlapping_mesh, node_elements_partial, quadratic):
    coordinates, meshes = input_data[5], input_data[3]
    ind = input_data[-2]
    if not quadratic:
        assert ind[1] == 0, "Customized force is not allowed in this solver!"          # AMORE.py:139
    f = np.zeros((2 * len(coordinates), 1))
    if len(ind) == 2 or (len(ind) == 3 and ind[2] == 0):
        pairs = input_data[-1]
    elif len(ind) == 3 and ind[2] == 1:
        pairs = input_data[-1][0]
        for c in input_data[-1][1]:
            f[2 * c[0], 0] += c[1]
            f[2 * c[0] + 1, 0] += c[2]
    else:
        raise ValueError
    if quadratic:
        n_int = 3 + (2 if ind[1] else 0)
    else:
        n_int = 2 + (3 if ind[1] else 0)
    pos, wei = _gauss_line(n_int)
    custom = bool(ind[1])
    node_elements = _node_elements(coordinates, meshes) if quadratic else None

    for i in range(len(pairs) // 2):
        n1, n2 = pairs[2 * i][0], pairs[2 * i + 1][0]
        f1 = np.array([pairs[2 * i][1:3]]).transpose()
        f2 = np.array([pairs[2 * i + 1][1:3]]).transpose()
        f3 = 0.5 * (f1 + f2)
        where = None
        common = set(node_elements_partial[n1]) & set(node_elements_partial[n2])
        if common:
            where = common.pop()
        el = overlapping_mesh[where[0]][where[1]] if where else None
        triangulated = el is not None and el.polygons[0].triangles is not None
        n3 = None

        def lin_force(ns, xy):
            return _custom_force(xy[0], xy[1]) if custom else ns[0] * f1 + ns[1] * f2

        def quad_force(ns, xy):
            return _custom_force(xy[0], xy[1]) if custom else ns[0] * f1 + ns[1] * f2 + ns[2] * f3

        if not quadratic:
            coord = np.array([coordinates[n1], coordinates[n2]])
            floc = np.zeros((4, 1))
            if triangulated:
                for poly in el.polygons:
                    for tri in poly.triangles:
                        edge = _triangle_on_boundary(tri, coord[0], coord[1])
                        if edge:
                            _straight_piece(floc, edge, pos, wei, coord, where[0], _lin, lin_force)
                            break
            else:
                length = _len(coord[0], coord[1])
                for j in range(n_int):
                    ns = _lin(pos[j])
                    floc += 0.5 * wei[j] * length * _apply(ns, ns[0] * f1 + ns[1] * f2)
        else:
            if el is not None:
                numbering = el.numbering
            else:
                where = (set(node_elements[n1]) & set(node_elements[n2])).pop()
                numbering = meshes[where[0]][where[1]]
            n3 = _mid_node(numbering, n1, n2)
            if triangulated and n3 is not None:
                floc = np.zeros((6, 1))
                coord = _edge3(coordinates, n1, n2, n3)
                for poly in el.polygons:
                    for tri in poly.triangles:
                        edge = _triangle_containing_boundary(tri, coord[0], coord[1])
                        if edge:     # the whole (possibly curved) element edge is one triangle edge
                            rho = np.array([edge.origin.rho[where[0]], edge.destination.rho[where[0]], 0.0])
                            rho[2] = 0.5 * (rho[0] + rho[1])
                            for l in range(n_int):
                                ns, jac = _quad(pos[l], coord)
                                rho_v = ns[0] * rho[0] + ns[1] * rho[1] + ns[2] * rho[2]
                                if custom:
                                    xy = ns[0] * coord[0] + ns[1] * coord[1] + ns[2] * coord[2]
                                    force = _custom_force(xy[0], xy[1])
                                else:
                                    force = ns[0] * f1 + ns[1] * f2 + ns[2] * f3
                                floc += wei[l] * jac * rho_v * _apply(ns, force)
                            break
                        edge = _triangle_on_boundary(tri, coord[0], coord[1])
                        if edge:
                            _straight_piece(floc, edge, pos, wei, coord, where[0], lambda r: _quad(r, coord)[0], quad_force)
                            break
            elif triangulated:
                floc = np.zeros((4, 1))
                coord = np.array([coordinates[n1], coordinates[n2]])
                for poly in el.polygons:
                    for tri in poly.triangles:
                        edge = _triangle_on_boundary(tri, coord[0], coord[1])
                        if edge:
                            _straight_piece(floc, edge, pos, wei, coord, where[0], _lin, lin_force)
                            break
            elif n3 is not None:        # untriangulated overlapping element or regular element, curved edge
                floc = np.zeros((6, 1))
                coord = _edge3(coordinates, n1, n2, n3)
                for j in range(n_int):
                    ns, jac = _quad(pos[j], coord)
                    if custom:
                        xy = ns[0] * coord[0] + ns[1] * coord[1] + ns[2] * coord[2]
                        force = _custom_force(xy[0], xy[1])
                    else:
                        force = ns[0] * f1 + ns[1] * f2 + ns[2] * f3
                    floc += wei[j] * jac * _apply(ns, force)
            else:
                floc = np.zeros((4, 1))
                length = _len(coordinates[n1], coordinates[n2])
                coord = np.array([coordinates[n1], coordinates[n2]])
                for j in range(n_int):
                    ns = _lin(pos[j])
                    if custom:
                        xy = ns[0] * coord[0] + ns[1] * coord[1]
                        force = _custom_force(xy[0], xy[1])
                    else:
                        force = ns[0] * f1 + ns[1] * f2
                    floc += 0.5 * wei[j] * length * _apply(ns, force)
        f[2 * n1:2 * n1 + 2, 0] += floc[0:2, 0]
        f[2 * n2:2 * n2 + 2, 0] += floc[2:4, 0]
        if len(floc) == 6:
            f[2 * n3:2 * n3 + 2, 0] += floc[4:6, 0]
    return f


def _force_single(input_data, quadratic):
    """traditionalElement.py:59-89 (lowerOrderFE / ICMFE) and :362-403 (quadraticFE)."""
    coordinates, meshes = input_data[5], input_data[3]
    ind = input_data[-2]
    pairs = input_data[-1]
    n_int = (3 + (2 if ind[1] else 0)) if quadratic else (2 + (3 if ind[1] else 0))
    pos, wei = _gauss_line(n_int)
    f = np.zeros((2 * len(coordinates), 1))
    node_elements = _node_elements(coordinates, meshes) if quadratic else None
    for i in range(len(pairs) // 2):
        n1, n2 = pairs[2 * i][0], pairs[2 * i + 1][0]
        f1 = np.array([pairs[2 * i][1:3]]).transpose()
        f2 = np.array([pairs[2 * i + 1][1:3]]).transpose()
        if not quadratic:
            length = _len(coordinates[n1], coordinates[n2])
            floc = np.zeros((4, 1))
            for j in range(n_int):
                ns = _lin(pos[j])
                floc += 0.5 * wei[j] * length * _apply(ns, ns[0] * f1 + ns[1] * f2)
        else:
            where = (set(node_elements[n1]) & set(node_elements[n2])).pop()
            n3 = _mid_node(meshes[where[0]][where[1]], n1, n2)
            coord = _edge3(coordinates, n1, n2, n3)
            floc = np.zeros((6, 1))
            for j in range(n_int):
                l1, l2 = _lin(pos[j])
                ns, jac = _quad(pos[j], coord)
                floc += wei[j] * jac * _apply(ns, l1 * f1 + l2 * f2)
            f[2 * n3:2 * n3 + 2, 0] += floc[4:6, 0]
        f[2 * n1:2 * n1 + 2, 0] += floc[0:2, 0]
        f[2 * n2:2 * n2 + 2, 0] += floc[2:4, 0]
    return f


def force_vector(input_data, overlapping_mesh=None, node_elements_partial=None, quadratic=False):
    """Global force vector ``fglo`` (2N,1) exactly as the reference's solvers build it."""
    if overlapping_mesh is None:
        return _force_single(input_data, quadratic)
    return _force_amore(input_data, overlapping_mesh, node_elements_partial, quadratic)


def constraints(input_data):
    """Fixed-DOF mask and values (AMORE.py:213-226); sorts ``inputData[-3]`` in place like the reference."""
    n = 2 * len(input_data[5])
    fixed = np.zeros(n, dtype=np.uint8)
    value = np.zeros(n)
    clist = input_data[-3]
    clist.sort(key=lambda item: item[0])
    for c in clist:
        if c[1]:
            value[2 * c[0]] = c[3]; fixed[2 * c[0]] = 1
        if c[2]:
            value[2 * c[0] + 1] = c[4]; fixed[2 * c[0] + 1] = 1
    return fixed, value


# ---------------------------------------------------------------------------------------------------------------
# the solve skeleton shared by the six entry points
# ---------------------------------------------------------------------------------------------------------------
def _check_single(meshes, allowed):
    for el in meshes[0]:
        if len(el) not in allowed:
            raise ValueError("Wrong element numbering!")


def _run(input_data, solver, overlapping_mesh=None, polygons=None, node_elements_partial=None, report=False):
    global last_solve_info
    t = time.time()
    fp = _fl.flatten(input_data, overlapping_mesh, polygons, solver=solver)
    fp.validate()
    ctx = _context()
    ctx.clear_halo()
    ctx.set_problem(fp).symbolic().assemble()
    print("Assembling stiffness matrix costs %s seconds." % (time.time() - t))
    t = time.time()
    f = force_vector(input_data, overlapping_mesh, node_elements_partial, quadratic=(solver != _fl.SOLVER_LOWER)
                     if overlapping_mesh is not None else (solver == _fl.SOLVER_QUADRATIC))
    print("Calculating force term costs %s seconds." % (time.time() - t))
    t = time.time()
    fixed, value = constraints(input_data)
    ctx.set_system(f.reshape(-1), fixed, value)
    print("Imposing constraints costs %s seconds." % (time.time() - t))
    t = time.time()
    # raises RuntimeError (MOFEM_E_NOCONV) unless the TRUE residual is <= RTOL, or stagnated at the fp64 floor of this
    # matrix below the library's bound (1e-8): the reference's direct solve has no such failure mode, so it is never silent
    u, info = ctx.solve(rtol=RTOL, maxit=MAXIT)
    last_solve_info = info
    if info["converged"] == 2:
        warnings.warn("PCG stopped at the fp64 floor of this system: true relative residual %.2e (requested %.0e)"
                      % (info["relres_true"], RTOL), RuntimeWarning, stacklevel=3)
    if report:      # AMORE.py:685-687 / 1108-1110
        indptr, indices, data = ctx.csr()
        free = fixed == 0
        rows = np.repeat(np.arange(len(indptr) - 1), np.diff(indptr))
        keep = free[rows] & free[indices]
        print("Number of equations = %s." % int(free.sum()))
        print("Number of entries = %s." % int(keep.sum()))
        print("Number of non-zero entries = %s." % int(np.count_nonzero(data[keep])))
    print("Solving the linear system costs %s seconds." % (time.time() - t))
    displacement = u.reshape(-1, 1)
    energy = np.array([[ctx.energy()]])
    return displacement, energy, [fp.mat_D[k].copy() for k in range(fp.mat_D.shape[0])]


def lowerOrderAMORE(inputData, overlappingMesh, polygons, nodeElementsPartial):
    """AMORE.lowerOrderAMORE:17."""
    d, e, _ = _run(inputData, _fl.SOLVER_LOWER, overlappingMesh, polygons, nodeElementsPartial)
    return d, e


def quadraticAMORE(inputData, overlappingMesh, polygons, nodeElementsPartial):
    """AMORE.quadraticAMORE:282."""
    return _run(inputData, _fl.SOLVER_QUADRATIC, overlappingMesh, polygons, nodeElementsPartial, report=True)


def quadraticICMAMORE(inputData, overlappingMesh, polygons, nodeElementsPartial):
    """AMORE.quadraticICMAMORE:705."""
    return _run(inputData, _fl.SOLVER_ICM, overlappingMesh, polygons, nodeElementsPartial, report=True)


def lowerOrderFE(inputData):
    """traditionalElement.lowerOrderFE:13."""
    _check_single(inputData[3], (3, 4))
    d, e, _ = _run(inputData, _fl.SOLVER_LOWER)
    return d, e


def ICMFE(inputData):
    """traditionalElement.ICMFE:164."""
    _check_single(inputData[3], (4,))
    d, e, _ = _run(inputData, _fl.SOLVER_ICM)
    return d, e


def quadraticFE(inputData):
    """traditionalElement.quadraticFE:314."""
    _check_single(inputData[3], (6, 9))
    d, e, _ = _run(inputData, _fl.SOLVER_QUADRATIC)
    return d, e
