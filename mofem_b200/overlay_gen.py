"""Synthetic overlapping-mesh problems emitted directly as flattened SoA (the reference ships no generator).

Patch family ``P(p=5, period=8)`` of SURVEY.md App. C: mesh 0 is an ``n0 x n0`` quad4 grid on the unit square,
mesh 1 is ``(n0/8)^2`` disjoint patches of ``5 x 5`` quad4 elements with the same spacing, lower-left corner of
patch ``(a, b)`` at ``((8a+1.37)h, (8b+1.41)h)``; seeded jitter moves mesh-0 interior nodes and patch-interior
nodes (patch borders stay straight), so the quads are non-affine and the Newton inverse map needs 3-4 updates.

The generator follows the reference's rules for what the plane sweep would hand to the solver:

* overlay cells = pieces with a constant tuple of incident elements (computGeometry/planeSweepMeshes.py:354);
  every mesh-1 element contains exactly one mesh-0 node and is cut into 4 convex 4-gons; the patch border cuts
  the surrounding ring of mesh-0 elements into one-element cells (5-gons on the sides, L-shaped 6-gons at the
  corners); untouched mesh-0 elements are regular quadFE cells;
* triangulation = the reference's recursive "cut at the up-left-most vertex" rule
  (computGeometry/triangulation.py:49-139), triangles in the same order and vertex rotation;
* node weights: 0 on patch-border nodes (a mesh boundary strictly inside the domain), 1 elsewhere
  (planeSweepMeshes.py:434-446); ``rho_j(v) = c_j sum_a N_a(xi_j(v)) w_a / sum_k (...)``, ``c_0 = 1``,
  ``c_{j>=1} = 9`` (planeSweepMeshes.py:1004-1027).

:func:`patch_deck` returns the same problem as a reference ``inputData`` tuple so that small sizes can be pushed
through the reference's own sweep for cross-validation (tests/golden/make_golden.py).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

from .flatten import CELL_OVERLAY, CELL_REGULAR, SOLVER_LOWER, FlatProblem, material_matrix

PERIOD = 8
PATCH = 5
OFF_X, OFF_Y = 1.37, 1.41
OFF2_X, OFF2_Y = 4.71, 4.19        # staggered third mesh
SEED = 20260101


@dataclass
class PatchMeshes:
    nx: int                   # mesh-0 elements per row (x); h = 1/nx
    ny: int                   # mesh-0 element rows (y); the domain is [0,1] x [0,ny/nx]
    Px: int                   # patches per direction
    Py: int
    X0: np.ndarray            # [ny+1, nx+1, 2] mesh-0 node coordinates, indexed [j, i]
    X1: np.ndarray            # [Py, Px, 6, 6, 2] mesh-1 node coordinates, indexed [b, a, t, s]
    X2: np.ndarray = None     # [Py-1, Px-1, 6, 6, 2] staggered mesh-2 patches (family "m3s"), else None

    @property
    def n0(self): return self.nx
    @property
    def P(self): return self.Px
    @property
    def n_node0(self): return (self.nx + 1) * (self.ny + 1)
    @property
    def n_elem0(self): return self.nx * self.ny
    @property
    def n_node(self): return self.n_node0 + 36 * self.Px * self.Py
    @property
    def n_elem(self): return self.n_elem0 + 25 * self.Px * self.Py

    def node_xy(self):
        return np.concatenate([self.X0.reshape(-1, 2), self.X1.reshape(-1, 2)])

    def elem_nodes0(self):
        nx, ny = self.nx, self.ny
        j, i = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
        q = (j * (nx + 1) + i).reshape(-1)
        return np.stack([q, q + 1, q + nx + 2, q + nx + 1], 1)

    def elem_nodes1(self):
        base = self.n_node0 + 36 * np.arange(self.Px * self.Py)
        t, s = np.meshgrid(np.arange(5), np.arange(5), indexing="ij")
        loc = (6 * t + s).reshape(-1)
        e = np.stack([loc, loc + 1, loc + 7, loc + 6], 1)             # [25, 4]
        return (base[:, None, None] + e[None]).reshape(-1, 4)


def patch_meshes(n0: int, jitter: float = 0.12, seed: int = SEED, ny: int = None, stagger: bool = False) -> PatchMeshes:
    nx = n0
    ny = nx if ny is None else ny
    if nx % PERIOD or nx < PERIOD or ny % PERIOD or ny < PERIOD:
        raise ValueError("n0 (and ny) must be positive multiples of 8")
    h = 1.0 / nx
    Px, Py = nx // PERIOD, ny // PERIOD
    rng = np.random.default_rng(seed)
    j, i = np.meshgrid(np.arange(ny + 1), np.arange(nx + 1), indexing="ij")
    X0 = np.stack([i, j], -1).astype(np.float64)
    jit0 = rng.uniform(-jitter, jitter, X0.shape)
    jit0[0, :] = jit0[-1, :] = 0.0
    jit0[:, 0] = jit0[:, -1] = 0.0
    X0 = (X0 + jit0) * h
    b, a, t, s = np.meshgrid(np.arange(Py), np.arange(Px), np.arange(6), np.arange(6), indexing="ij")
    X1 = np.stack([PERIOD * a + OFF_X + s, PERIOD * b + OFF_Y + t], -1).astype(np.float64)
    jit1 = rng.uniform(-jitter, jitter, X1.shape)
    jit1[:, :, 0, :] = jit1[:, :, 5, :] = 0.0
    jit1[:, :, :, 0] = jit1[:, :, :, 5] = 0.0
    X1 = (X1 + jit1) * h
    X2 = None
    if stagger:           # drawn after X0 and X1, so those are the same arrays as without the third mesh
        if Px < 2 or Py < 2:
            raise ValueError("the staggered third mesh needs n0 >= 16")
        b, a, t, s = np.meshgrid(np.arange(Py - 1), np.arange(Px - 1), np.arange(6), np.arange(6), indexing="ij")
        X2 = np.stack([PERIOD * a + OFF2_X + s, PERIOD * b + OFF2_Y + t], -1).astype(np.float64)
        jit2 = rng.uniform(-jitter, jitter, X2.shape)
        jit2[:, :, 0, :] = jit2[:, :, 5, :] = 0.0
        jit2[:, :, :, 0] = jit2[:, :, :, 5] = 0.0
        X2 = (X2 + jit2) * h
    return PatchMeshes(nx=nx, ny=ny, Px=Px, Py=Py, X0=X0, X1=X1, X2=X2)


# ---------------------------------------------------------------------------------------------
# geometry helpers (vectorised over the leading axes)
# ---------------------------------------------------------------------------------------------
def _intersect(p, q, c, d):
    """Intersection of segment p-q with the line through c-d."""
    r = q - p
    e = d - c
    den = r[..., 0] * e[..., 1] - r[..., 1] * e[..., 0]
    t = ((c[..., 0] - p[..., 0]) * e[..., 1] - (c[..., 1] - p[..., 1]) * e[..., 0]) / den
    return p + t[..., None] * r


def _quad_shape(r, s):
    N = 0.25 * np.stack([(1 - r) * (1 - s), (1 + r) * (1 - s), (1 + r) * (1 + s), (1 - r) * (1 + s)], -1)
    dr = 0.25 * np.stack([-(1 - s), (1 - s), (1 + s), -(1 + s)], -1)
    ds = 0.25 * np.stack([-(1 - r), -(1 + r), (1 + r), (1 - r)], -1)
    return N, dr, ds


def _inv_quad_N(coord, xy, iters=6):
    """Shape-function values N_a(xi(xy)) of quad4 elements (Newton inverse map, invShapeFunction.py:41-73), vectorised
    over the leading axis with explicit bilinear formulas (6 updates from (0,0): the jittered quads converge in 3-5)."""
    X = coord[..., 0] - coord[..., :1, 0]
    Y = coord[..., 1] - coord[..., :1, 1]
    x = xy[..., 0] - coord[..., 0, 0]
    y = xy[..., 1] - coord[..., 0, 1]
    X0, X1, X2, X3 = (X[..., k] for k in range(4))
    Y0, Y1, Y2, Y3 = (Y[..., k] for k in range(4))
    r = np.zeros_like(x); s = np.zeros_like(x)
    for _ in range(iters):
        rm, rp, sm, sp = 1.0 - r, 1.0 + r, 1.0 - s, 1.0 + s
        N0, N1, N2, N3 = 0.25 * rm * sm, 0.25 * rp * sm, 0.25 * rp * sp, 0.25 * rm * sp
        fx = x - (N0 * X0 + N1 * X1 + N2 * X2 + N3 * X3)
        fy = y - (N0 * Y0 + N1 * Y1 + N2 * Y2 + N3 * Y3)
        xr = 0.25 * (sm * (X1 - X0) + sp * (X2 - X3)); xs = 0.25 * (rm * (X3 - X0) + rp * (X2 - X1))
        yr = 0.25 * (sm * (Y1 - Y0) + sp * (Y2 - Y3)); ys = 0.25 * (rm * (Y3 - Y0) + rp * (Y2 - Y1))
        det = xr * ys - xs * yr
        r = r + (fx * ys - fy * xs) / det
        s = s + (fy * xr - fx * yr) / det
    rm, rp, sm, sp = 1.0 - r, 1.0 + r, 1.0 - s, 1.0 + s
    return np.stack([0.25 * rm * sm, 0.25 * rp * sm, 0.25 * rp * sp, 0.25 * rm * sp], -1)


def _point_gt(a, b):
    """point_2D.__gt__ (planeSweepMeshes.py:36-40): higher y first, then smaller x."""
    return a[1] > b[1] if a[1] != b[1] else a[0] < b[0]


def _is_left(p1, p2, p3):
    return ((p2[0] - p1[0]) * (p3[1] - p1[1]) - (p2[1] - p1[1]) * (p3[0] - p1[0])) / math.sqrt(
        (p2[1] - p1[1]) ** 2 + (p2[0] - p1[0]) ** 2)


def _in_polygon(pt, poly):
    """triangulation.isInPolygon:151-171 (winding number; boundary counts as inside)."""
    w = 0
    for i in range(len(poly) - 1):
        a, b = poly[i], poly[i + 1]
        if (a[0] == pt[0] and a[1] == pt[1]) or (b[0] == pt[0] and b[1] == pt[1]):
            return 1
        if a[1] == pt[1] == b[1] and (pt[0] - a[0]) * (pt[0] - b[0]) <= 0:
            return 1
        if a[1] <= pt[1]:
            if b[1] > pt[1]:
                v = _is_left(a, b, pt)
                if v > 0: w += 1
                if v == 0: return 1
        elif b[1] <= pt[1]:
            v = _is_left(a, b, pt)
            if v < 0: w -= 1
            if v == 0: return 1
    return w


def ear_triangulate(pts, ids=None):
    """The reference's ``simpleTriangulation`` (triangulation.py:49-139) on a CCW vertex list.

    ``pts``: sequence of (x, y); ``ids``: vertex labels (default 0..n-1), ``ids[0]`` is the origin of the
    half-edge the recursion starts from.  Returns triangles as label triples, in the reference's order and
    vertex rotation."""
    if ids is None:
        ids = list(range(len(pts)))
    n = len(ids)
    if n == 3:
        return [tuple(ids)]
    u = 0
    for k in range(1, n):
        if _point_gt(pts[ids[k]], pts[ids[u]]):
            u = k
    ring = ids[u:] + ids[:u]                       # ring[0] = up-left-most vertex U
    U, nxt, prv = ring[0], ring[1], ring[-1]
    tri = [pts[U], pts[nxt], pts[prv], pts[U]]
    inside, dist = [], []
    for k in range(2, n - 1):
        if _in_polygon(pts[ring[k]], tri):
            inside.append(k)
            dist.append(_is_left(pts[nxt], pts[prv], pts[ring[k]]))
    if not inside:
        t1 = [(prv, U, nxt)]                        # simpleTriangulation(edge1, 3), edge1 = U.prev
        t2 = ear_triangulate(pts, ring[1:])         # start edge2 = U.next
    else:
        best = 0                                    # the reference keeps the FIRST maximum (strict >)
        for m in range(1, len(dist)):
            if dist[m] > dist[best]:
                best = m
        k = inside[best]
        t1 = ear_triangulate(pts, ring[:k + 1])     # U ... X, start edge1 = U
        t2 = ear_triangulate(pts, ring[k:] + [U])   # X ... U, start edge2 = X
    return t2 + t1


# ---------------------------------------------------------------------------------------------
# the flattened problem
# ---------------------------------------------------------------------------------------------
def _quad_triangles(poly):
    """Vectorised ``ear_triangulate`` for convex 4-gons: poly [n,4,2] -> vertex index table [n,2,3]."""
    y = poly[..., 1]; x = poly[..., 0]
    ymax = y.max(1, keepdims=True)
    cand = np.where(y == ymax, x, np.inf)
    u = cand.argmin(1)                                # max y, then min x
    # triangles: (U.next, opposite, U.prev) then (U.prev, U, U.next)
    idx = np.stack([np.stack([u + 1, u + 2, u + 3], 1), np.stack([u + 3, u, u + 1], 1)], 1) % 4
    return idx


_RING_CTX = None


def _gen_workers(n_patches):
    import os
    w = os.environ.get("MOFEM_GEN_WORKERS")
    if w is not None:
        return max(1, int(w))
    if n_patches < 4000:
        return 1
    return max(1, min(8, (os.cpu_count() or 1) // 2))


def _ring_row(c, bb):
    """Ring cells of patch row ``bb``: (bb, element ids, triangle coordinates [T,3,2], triangles per cell)."""
    IW, IE, IS, IN, X0, X1 = c["IW"], c["IE"], c["IS"], c["IN"], c["X0"], c["X1"]
    b_lo, Px, nx = c["b_lo"], c["Px"], c["nx"]
    j1 = PERIOD * bb + 1
    X0l = X0[j1:j1 + 7].tolist()                      # the 7 mesh-0 node rows this patch row touches
    elems, tris_out, cnt = [], [], []
    for aa in range(Px):
        iw = IW[bb - b_lo, aa, :, 0].tolist(); ie = IE[bb - b_lo, aa, :, 4].tolist()
        is_ = IS[bb - b_lo, aa, 0, :].tolist(); in_ = IN[bb - b_lo, aa, 4, :].tolist()
        m = X1[bb, aa].tolist()
        i1 = PERIOD * aa + 1
        polys = []

        def corners(u, v):
            i = i1 + u
            return X0l[v][i], X0l[v][i + 1], X0l[v + 1][i + 1], X0l[v + 1][i]
        for v in range(1, 5):      # left and right sides
            n00, n10, n11, n01 = corners(0, v)
            polys.append((0, v, [n00, iw[v - 1], m[v][0], iw[v], n01]))
            n00, n10, n11, n01 = corners(5, v)
            polys.append((5, v, [ie[v - 1], n10, n11, ie[v], m[v][5]]))
        for u in range(1, 5):      # bottom and top sides
            n00, n10, n11, n01 = corners(u, 0)
            polys.append((u, 0, [n00, n10, is_[u], m[0][u], is_[u - 1]]))
            n00, n10, n11, n01 = corners(u, 5)
            polys.append((u, 5, [in_[u - 1], m[5][u], in_[u], n11, n01]))
        n00, n10, n11, n01 = corners(0, 0)
        polys.append((0, 0, [n00, n10, is_[0], m[0][0], iw[0], n01]))
        n00, n10, n11, n01 = corners(5, 0)
        polys.append((5, 0, [n00, n10, n11, ie[0], m[0][5], is_[4]]))
        n00, n10, n11, n01 = corners(5, 5)
        polys.append((5, 5, [ie[4], n10, n11, n01, in_[4], m[5][5]]))
        n00, n10, n11, n01 = corners(0, 5)
        polys.append((0, 5, [n00, iw[4], m[5][0], in_[0], n11, n01]))
        for u, v, poly in polys:
            tris = ear_triangulate(poly)
            elems.append((j1 + v) * nx + i1 + u)
            for t in tris:
                tris_out.append([poly[t[0]], poly[t[1]], poly[t[2]]])
            cnt.append(len(tris))
    return (bb, np.array(elems, dtype=np.int64), np.array(tris_out, dtype=np.float64).reshape(-1, 3, 2),
            np.array(cnt, dtype=np.int64))


def _ring_rows_worker(rows):
    return [_ring_row(_RING_CTX, bb) for bb in rows]


def patch_family(n0: int, jitter: float = 0.12, seed: int = SEED, E: float = 1000.0, nu: float = 0.3,
                 solver: int = SOLVER_LOWER, return_meshes: bool = False, ny: int = None, patch_rows=None,
                 n_mesh: int = 2, engine: str = "numpy"):
    """Flattened 2-mesh patch-family problem + its linear system.

    Returns ``(FlatProblem, rhs, fixed, fixed_value)``: left edge of mesh 0 clamped, concentrated load
    ``(1, -0.5)`` at the bottom-right corner node of mesh 0.

    ``n_mesh=3`` ("multi-mesh" config) distributes the patches over two overlapping meshes: patch ``(a, b)`` belongs to
    mesh ``1 + (a+b) % 2``.  Cells then carry three mesh slots and three weight functions; the patches of meshes 1 and 2
    do not overlap each other (three-element cells are covered by the simpleOFE3 / zoo fixtures instead).

    ``ny`` makes the domain a strip stack of ``ny`` element rows (weak scaling: one ``n0 x n0`` block per GPU).
    ``patch_rows=(b_lo, b_hi)`` emits only the cells of the patch rows ``b_lo <= b < b_hi`` plus the mesh-0
    element row directly below them (every cell that touches a node of that strip); node and element tables
    stay global, so the result can go straight into :func:`mofem_b200.partition.partition_problem`."""
    if n_mesh not in (2, 3):
        raise ValueError("n_mesh must be 2 or 3")
    if engine not in ("numpy", "native"):
        raise ValueError("engine must be 'numpy' or 'native'")
    pm = patch_meshes(n0, jitter, seed, ny)
    nx, ny = pm.nx, pm.ny
    Px, Py = pm.Px, pm.Py
    b_lo, b_hi = (0, Py) if patch_rows is None else (int(patch_rows[0]), int(patch_rows[1]))
    if not (0 <= b_lo < b_hi <= Py):
        raise ValueError("patch_rows out of range")
    if engine == "native":
        # the overlay core the CUDA kernels compile, built for the host (csrc/overlay_host.cpp): same bits as below, no
        # forked workers, ~10x faster -- what the multi-GPU driver uses before CUDA / NCCL are up
        out = _finish_problem(pm, _native_overlay(pm, b_lo, b_hi, n_mesh), n_mesh, solver, E, nu)
        return out + (pm,) if return_meshes else out
    X0, X1 = pm.X0, pm.X1
    n_elem0 = pm.n_elem0
    node0 = lambda j, i: j * (nx + 1) + i
    elem0 = lambda j, i: j * nx + i

    b, a, q, p = np.meshgrid(np.arange(b_lo, b_hi), np.arange(Px), np.arange(5), np.arange(5), indexing="ij")
    C00 = X1[b, a, q, p]; C10 = X1[b, a, q, p + 1]; C11 = X1[b, a, q + 1, p + 1]; C01 = X1[b, a, q + 1, p]
    ip = PERIOD * a + 2 + p
    jp = PERIOD * b + 2 + q
    Pn = X0[jp, ip]
    IW = _intersect(Pn, X0[jp, ip - 1], C00, C01)
    IE = _intersect(Pn, X0[jp, ip + 1], C10, C11)
    IS = _intersect(Pn, X0[jp - 1, ip], C00, C10)
    IN = _intersect(Pn, X0[jp + 1, ip], C01, C11)
    # straight patch borders: pin the border coordinate exactly
    X1s = X1[b_lo:b_hi]
    IW[:, :, :, 0, 0] = X1s[:, :, 0:1, 0, 0]
    IE[:, :, :, 4, 0] = X1s[:, :, 0:1, 5, 0]
    IS[:, :, 0, :, 1] = X1s[:, :, 0, 0:1, 1]
    IN[:, :, 4, :, 1] = X1s[:, :, 5, 0:1, 1]

    # ---- two-element cells: 4 pieces per mesh-1 element ------------------------------------
    pieces = np.stack([
        np.stack([C00, IS, Pn, IW], -2),     # lower-left  (dx, dy) = (0, 0)
        np.stack([IS, C10, IE, Pn], -2),     # lower-right (1, 0)
        np.stack([Pn, IE, C11, IN], -2),     # upper-right (1, 1)
        np.stack([IW, Pn, IN, C01], -2),     # upper-left  (0, 1)
    ], 4)                                     # [P,P,5,5,4 pieces,4 verts,2]
    dxy = np.array([[0, 0], [1, 0], [1, 1], [0, 1]])
    e0_piece = elem0((jp - 1)[..., None] + dxy[:, 1], (ip - 1)[..., None] + dxy[:, 0])        # [P,P,5,5,4]
    e1_piece = np.broadcast_to((n_elem0 + 25 * (b * Px + a) + 5 * q + p)[..., None], e0_piece.shape)
    pieces = pieces.reshape(-1, 4, 2)
    e0_piece = e0_piece.reshape(-1); e1_piece = e1_piece.reshape(-1)
    n_piece = pieces.shape[0]

    en0 = pm.elem_nodes0(); en1 = pm.elem_nodes1()
    node_xy = pm.node_xy()
    w1 = np.zeros((6, 6)); w1[1:5, 1:5] = 1.0
    node_w = np.concatenate([np.ones(pm.n_node0), np.tile(w1.reshape(-1), Px * Py)])

    c1 = node_xy[en1[e1_piece - n_elem0]]                         # [n_piece,4,2]
    wgt1 = node_w[en1[e1_piece - n_elem0]]                        # [n_piece,4]
    patch_mesh = (1 + ((b + a) % 2 if n_mesh == 3 else 0 * b))            # [Py',Px,5,5] mesh of the patch
    mesh_piece = np.broadcast_to(patch_mesh[..., None], (b_hi - b_lo, Px, 5, 5, 4)).reshape(-1)
    rho_piece = np.zeros((n_piece, 4, n_mesh))
    for v in range(4):
        N1 = _inv_quad_N(c1, pieces[:, v])
        r0 = 1.0          # mesh 0: every node weight is 1 and the shape functions sum to 1 (the reference gets 1 +- 1 ulp)
        r1 = 9.0 * (N1 * wgt1).sum(-1)
        rho_piece[:, v, 0] = r0 / (r0 + r1)
        rho_piece[np.arange(n_piece), v, mesh_piece] = r1 / (r0 + r1)
    tri_idx = _quad_triangles(pieces)                             # [n_piece,2,3]
    ar = np.arange(n_piece)[:, None, None]
    piece_tri_xy = pieces[ar, tri_idx]                            # [n_piece,2,3,2]
    piece_tri_rho = rho_piece[ar, tri_idx]                        # [n_piece,2,3,n_mesh]

    # ---- ring cells (one element, triangulated) -----------------------------------------------
    # scalar work (the reference's recursive ear rule per polygon): spread over forked worker processes when large
    ctx_ring = dict(IW=IW, IE=IE, IS=IS, IN=IN, X0=X0, X1=X1, b_lo=b_lo, Px=Px, nx=nx)
    rows = list(range(b_lo, b_hi))
    n_workers = _gen_workers(len(rows) * Px)
    if n_workers > 1:
        import multiprocessing as mp
        global _RING_CTX
        _RING_CTX = ctx_ring
        chunks = [rows[k::n_workers * 4] for k in range(n_workers * 4)]
        order = np.argsort(np.concatenate([c for c in chunks if c]), kind="stable")
        with mp.get_context("fork").Pool(n_workers) as pool:
            parts = pool.map(_ring_rows_worker, [c for c in chunks if c])
        _RING_CTX = None
        # back to ascending patch-row order (each part holds whole rows)
        by_row = {}
        for part in parts:
            for bb, elems, tris, cnt in part:
                by_row[bb] = (elems, tris, cnt)
        del order
        parts_sorted = [by_row[bb] for bb in rows]
    else:
        parts_sorted = [_ring_row(ctx_ring, bb)[1:] for bb in rows]
    ring_elem = np.concatenate([p[0] for p in parts_sorted]) if parts_sorted else np.zeros(0, np.int64)
    ring_tri_xy = np.concatenate([p[1] for p in parts_sorted]).reshape(-1, 3, 2) if parts_sorted else np.zeros((0, 3, 2))
    ring_cnt = np.concatenate([p[2] for p in parts_sorted]) if parts_sorted else np.zeros(0, np.int64)
    ring_ptr = np.concatenate([[0], np.cumsum(ring_cnt)])
    n_ring = len(ring_elem)

    # ---- untouched mesh-0 elements -> regular quadFE cells ------------------------------------------
    touched = np.zeros((ny, nx), dtype=bool)
    for d in range(6):
        for e in range(6):
            touched[(PERIOD * np.arange(Py) + 1 + d)[:, None], (PERIOD * np.arange(Px) + 1 + e)[None, :]] = True
    in_strip = np.zeros((ny, nx), dtype=bool)
    in_strip[max(PERIOD * b_lo - 1, 0):PERIOD * b_hi, :] = True
    reg_elem = np.flatnonzero((~touched & in_strip).reshape(-1))
    n_reg = len(reg_elem)

    # ---- assemble the SoA ----------------------------------------------------------------------------
    C = n_reg + n_ring + n_piece
    cell_elem = np.full((C, n_mesh), -1, dtype=np.int32)
    cell_kind = np.zeros(C, dtype=np.int32)
    cell_elem[:n_reg, 0] = reg_elem
    cell_elem[n_reg:n_reg + n_ring, 0] = ring_elem
    cell_kind[n_reg:] = CELL_OVERLAY
    cell_elem[n_reg + n_ring:, 0] = e0_piece
    cell_elem[n_reg + n_ring + np.arange(n_piece), mesh_piece] = e1_piece
    ntri = np.zeros(C, dtype=np.int64)
    ntri[n_reg:n_reg + n_ring] = ring_cnt
    ntri[n_reg + n_ring:] = 2
    tri_ptr = np.concatenate([[0], np.cumsum(ntri)])
    ring_rho = np.zeros((len(ring_tri_xy), 3, n_mesh)); ring_rho[:, :, 0] = 1.0
    tri_xy = np.concatenate([ring_tri_xy, piece_tri_xy.reshape(-1, 3, 2)])
    tri_rho = np.concatenate([ring_rho, piece_tri_rho.reshape(-1, 3, n_mesh)])

    elem_nodes = np.full((pm.n_elem, 9), -1, dtype=np.int32)
    elem_nodes[:n_elem0, :4] = en0
    elem_nodes[n_elem0:, :4] = en1
    fp = FlatProblem(
        solver=int(solver), n_mesh=n_mesh, node_xy=np.ascontiguousarray(node_xy), elem_nodes=elem_nodes,
        elem_nn=np.full(pm.n_elem, 4, dtype=np.int32), elem_mat=np.zeros(pm.n_elem, dtype=np.int32),
        elem_mesh=np.concatenate([np.zeros(n_elem0, np.int32), np.repeat(_patch_mesh_ids(Px, Py, n_mesh), 25)]),
        mat_D=material_matrix([E, nu], 1).reshape(1, 3, 3), cell_elem=cell_elem, cell_kind=cell_kind,
        cell_tri_ptr=tri_ptr.astype(np.int64), tri_xy=np.ascontiguousarray(tri_xy),
        tri_rho=np.ascontiguousarray(tri_rho), n_mesh_regular=0, mesh_offset=[0, n_elem0, pm.n_elem])

    n_dof = 2 * pm.n_node
    fixed = np.zeros(n_dof, dtype=np.uint8)
    left = node0(np.arange(ny + 1), 0)
    fixed[2 * left] = 1; fixed[2 * left + 1] = 1
    rhs = np.zeros(n_dof)
    rhs[2 * nx] = 1.0; rhs[2 * nx + 1] = -0.5
    out = (fp, rhs, fixed, np.zeros(n_dof))
    return out + (pm,) if return_meshes else out


def _native_overlay(pm, b_lo, b_hi, n_mesh):
    """Cell / triangle arrays from liboverlay_host.so (built by ``python -m mofem_b200.build``)."""
    import ctypes as C
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "liboverlay_host.so")
    if not os.path.exists(path):
        raise RuntimeError(f"{path} not found: build it with `python -m mofem_b200.build`")
    L = C.CDLL(path)
    sz = (C.c_int64 * 4)()
    L.ovl_host_sizes(C.c_int(pm.nx), C.c_int(pm.ny), C.c_int(b_lo), C.c_int(b_hi), C.c_int(n_mesh), sz)
    n_cell, n_tri = int(sz[0] + sz[1] + sz[2]), int(sz[3])
    out = dict(cell_elem=np.empty((n_cell, n_mesh), np.int32), cell_kind=np.empty(n_cell, np.int32),
               cell_tri_ptr=np.empty(n_cell + 1, np.int64), tri_xy=np.empty((n_tri, 3, 2)), tri_rho=np.empty((n_tri, 3, n_mesh)))
    X0 = np.ascontiguousarray(pm.X0); X1 = np.ascontiguousarray(pm.X1)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    L.ovl_host_generate(C.c_int(pm.nx), C.c_int(pm.ny), C.c_int(b_lo), C.c_int(b_hi), C.c_int(n_mesh), p(X0), p(X1),
                        p(out["cell_elem"]), p(out["cell_kind"]), p(out["cell_tri_ptr"]), p(out["tri_xy"]), p(out["tri_rho"]))
    return out


def _finish_problem(pm, cells, n_mesh, solver, E, nu):
    """(FlatProblem, rhs, fixed, fixed_value) around given cell / triangle arrays: node and element tables, clamped left
    edge, corner load -- identical to the tail of :func:`patch_family`."""
    nx, ny, n_elem0 = pm.nx, pm.ny, pm.n_elem0
    elem_nodes = np.full((pm.n_elem, 9), -1, dtype=np.int32)
    elem_nodes[:n_elem0, :4] = pm.elem_nodes0()
    elem_nodes[n_elem0:, :4] = pm.elem_nodes1()
    fp = FlatProblem(
        solver=int(solver), n_mesh=n_mesh, node_xy=np.ascontiguousarray(pm.node_xy()), elem_nodes=elem_nodes,
        elem_nn=np.full(pm.n_elem, 4, dtype=np.int32), elem_mat=np.zeros(pm.n_elem, dtype=np.int32),
        elem_mesh=np.concatenate([np.zeros(n_elem0, np.int32), np.repeat(_patch_mesh_ids(pm.Px, pm.Py, n_mesh), 25)]),
        mat_D=material_matrix([E, nu], 1).reshape(1, 3, 3), n_mesh_regular=0, mesh_offset=[0, n_elem0, pm.n_elem], **cells)
    n_dof = 2 * pm.n_node
    fixed = np.zeros(n_dof, dtype=np.uint8)
    left = np.arange(ny + 1) * (nx + 1)
    fixed[2 * left] = 1; fixed[2 * left + 1] = 1
    rhs = np.zeros(n_dof)
    rhs[2 * nx] = 1.0; rhs[2 * nx + 1] = -0.5
    return fp, rhs, fixed, np.zeros(n_dof)


def patch_family_gpu(ctx, n0: int, jitter: float = 0.12, seed: int = SEED, E: float = 1000.0, nu: float = 0.3,
                     solver: int = SOLVER_LOWER, ny: int = None, patch_rows=None, n_mesh: int = 2, to_host: bool = False):
    """:func:`patch_family` with the overlay built on the GPU (``mofem_overlay_patch_family``, ``csrc/overlay.cu``): the
    cell and triangle arrays of the returned ``FlatProblem`` are CUDA tensors on the context's device (numpy arrays with
    ``to_host=True``), bit-identical to the host generator's; node / element tables, right-hand side and constraints are
    O(nodes) numpy work as before.  Returns ``(FlatProblem, rhs, fixed, fixed_value, kernel_ms)``."""
    import ctypes as C

    import torch

    from . import _lib
    if n_mesh not in (2, 3):
        raise ValueError("n_mesh must be 2 or 3")
    pm = patch_meshes(n0, jitter, seed, ny)
    nx, ny = pm.nx, pm.ny
    Px, Py = pm.Px, pm.Py
    b_lo, b_hi = (0, Py) if patch_rows is None else (int(patch_rows[0]), int(patch_rows[1]))
    if not (0 <= b_lo < b_hi <= Py):
        raise ValueError("patch_rows out of range")
    L = _lib.load()
    sz = (C.c_int64 * 5)()
    _lib.raise_for(L.mofem_overlay_patch_sizes(nx, ny, b_lo, b_hi, n_mesh, sz), ctx._h)
    n_tri, n_cell = int(sz[3]), int(sz[4])
    dev = torch.device("cuda", ctx.device)
    out = dict(cell_elem=torch.empty((n_cell, n_mesh), dtype=torch.int32, device=dev),
               cell_kind=torch.empty(n_cell, dtype=torch.int32, device=dev),
               cell_tri_ptr=torch.empty(n_cell + 1, dtype=torch.int64, device=dev),
               tri_xy=torch.empty((n_tri, 3, 2), dtype=torch.float64, device=dev),
               tri_rho=torch.empty((n_tri, 3, n_mesh), dtype=torch.float64, device=dev))
    X0 = np.ascontiguousarray(pm.X0); X1 = np.ascontiguousarray(pm.X1)
    ms = C.c_double(0.0)
    _lib.raise_for(L.mofem_overlay_patch_family(ctx._h, nx, ny, b_lo, b_hi, n_mesh, _lib.ptr(X0), _lib.ptr(X1),
                                                _lib.ptr(out["cell_elem"]), _lib.ptr(out["cell_kind"]),
                                                _lib.ptr(out["cell_tri_ptr"]), _lib.ptr(out["tri_xy"]), _lib.ptr(out["tri_rho"]),
                                                C.byref(ms)), ctx._h)
    if to_host:
        out = {k: v.cpu().numpy() for k, v in out.items()}
    return _finish_problem(pm, out, n_mesh, solver, E, nu) + (float(ms.value),)


def _patch_mesh_ids(Px, Py, n_mesh):
    """Mesh id of every patch, patch-major order (b * Px + a)."""
    b, a = np.meshgrid(np.arange(Py), np.arange(Px), indexing="ij")
    return (1 + ((a + b) % 2 if n_mesh == 3 else 0 * a)).reshape(-1).astype(np.int32)


def family_deck(n0: int, family: str, jitter: float = 0.12, seed: int = SEED, E: float = 1000.0, nu: float = 0.3, solver: int = 1):
    """Reference ``inputData`` tuple of the general families (:func:`general_family`): "m3s" = staggered third mesh
    (BASELINE config 4), "tri" = patch quads split into 3-node triangles (config 3 variant b)."""
    gm = general_meshes(n0, family, jitter, seed)
    coords = gm.node_xy.tolist()
    meshes = [gm.elem_nodes[gm.elem_mesh == k][:, :gm.nn_of_mesh[k]].tolist() for k in range(gm.n_mesh)]
    material_mesh = [[0] * len(m) for m in meshes]
    constraints = [[j * (n0 + 1), 1, 1, 0.0, 0.0] for j in range(n0 + 1)]
    forces = [[], [[n0, 1.0, -0.5]]]
    return ([solver, 1], None, None, meshes, material_mesh, coords, [[E, nu]], constraints, [0, 0, 1], forces)


def patch_deck(n0: int, jitter: float = 0.12, seed: int = SEED, E: float = 1000.0, nu: float = 0.3, solver: int = 1,
               n_mesh: int = 2):
    """The same problem as a reference ``inputData`` tuple (inputFunctions/inputF.py:204-205, concentrated-force
    form), for pushing small sizes through the reference's own plane sweep."""
    pm = patch_meshes(n0, jitter, seed)
    coords = pm.node_xy().tolist()
    en1 = pm.elem_nodes1()
    ids = np.repeat(_patch_mesh_ids(pm.Px, pm.Py, n_mesh), 25)
    meshes = [pm.elem_nodes0().tolist()] + [en1[ids == k].tolist() for k in range(1, n_mesh)]
    material_mesh = [[0] * len(m) for m in meshes]
    constraints = [[j * (n0 + 1), 1, 1, 0.0, 0.0] for j in range(n0 + 1)]
    forces = [[], [[n0, 1.0, -0.5]]]
    return ([solver, 1], None, None, meshes, material_mesh, coords, [[E, nu]], constraints, [0, 0, 1], forces)


# ---------------------------------------------------------------------------------------------
# general families: any number of patch meshes, quad4 or tri3 patch elements (overlay by liboverlay_host.so)
# ---------------------------------------------------------------------------------------------
@dataclass
class GeneralMeshes:
    nx: int
    ny: int
    n_mesh: int
    node_xy: np.ndarray        # [n_node, 2]
    elem_nodes: np.ndarray     # [n_elem, 9] padded with -1
    elem_nn: np.ndarray        # [n_elem]
    elem_mesh: np.ndarray      # [n_elem]
    nn_of_mesh: list           # nodes per element of every mesh
    mesh_offset: list          # [n_mesh + 1] first element of every mesh
    patch: list                # per patch mesh: dict(Px, Py, off_x, off_y, tri, node0, elem0, X [Py, Px, 6, 6, 2])


def general_meshes(n0: int, family: str, jitter: float = 0.12, seed: int = SEED, ny: int = None) -> GeneralMeshes:
    """Meshes of a general family: "tri" (2 meshes, every patch quad split into two 3-node triangles), "m3s" (3 meshes: the
    disjoint patches of the 2-mesh family + staggered mesh-2 patches at ((8a+4.71)h, (8b+4.19)h), a, b < n0/8 - 1),
    "quad" (the 2-mesh family itself, for cross-checks against :func:`patch_family`)."""
    if family not in ("quad", "tri", "m3s"):
        raise ValueError("family must be 'quad', 'tri' or 'm3s'")
    pm = patch_meshes(n0, jitter, seed, ny, stagger=(family == "m3s"))
    nx, ny = pm.nx, pm.ny
    tri = family == "tri"
    patch = [dict(Px=pm.Px, Py=pm.Py, off_x=OFF_X, off_y=OFF_Y, tri=tri, X=np.ascontiguousarray(pm.X1))]
    if family == "m3s":
        patch.append(dict(Px=pm.Px - 1, Py=pm.Py - 1, off_x=OFF2_X, off_y=OFF2_Y, tri=False, X=np.ascontiguousarray(pm.X2)))
    node_xy = [pm.X0.reshape(-1, 2)]
    en = [np.pad(pm.elem_nodes0(), ((0, 0), (0, 5)), constant_values=-1)]
    nn = [np.full(pm.n_elem0, 4, np.int32)]
    em = [np.zeros(pm.n_elem0, np.int32)]
    node0, elem0 = pm.n_node0, pm.n_elem0
    offs = [0, pm.n_elem0]
    for k, pt in enumerate(patch):
        npatch = pt["Px"] * pt["Py"]
        pt["node0"], pt["elem0"] = node0, elem0
        node_xy.append(pt["X"].reshape(-1, 2))
        t, s = np.meshgrid(np.arange(5), np.arange(5), indexing="ij")
        loc = (6 * t + s).reshape(-1)
        quad = np.stack([loc, loc + 1, loc + 7, loc + 6], 1)                      # [25, 4] CCW
        if pt["tri"]:
            # (q0, q1, q2), (q0, q2, q3) -- except on the anti-diagonal (p + q == 4), where the cut runs q1-q3: with the plain
            # split the patch corners (5, 0) and (0, 5) belong to one triangle whose three nodes all carry weight 0, rho_1
            # vanishes on it, the corner node gets no stiffness and the reference's own spsolve reports a singular matrix
            anti = ((t + s) == 4).reshape(-1)
            e = np.where(anti[:, None, None], np.stack([quad[:, [0, 1, 3]], quad[:, [1, 2, 3]]], 1),
                         np.stack([quad[:, [0, 1, 2]], quad[:, [0, 2, 3]]], 1)).reshape(-1, 3)   # [50, 3]
        else:
            e = quad
        base = node0 + 36 * np.arange(npatch)
        ee = (base[:, None, None] + e[None]).reshape(-1, e.shape[1])
        en.append(np.pad(ee, ((0, 0), (0, 9 - e.shape[1])), constant_values=-1))
        nn.append(np.full(len(ee), e.shape[1], np.int32))
        em.append(np.full(len(ee), k + 1, np.int32))
        node0 += 36 * npatch
        elem0 += len(ee)
        offs.append(elem0)
    return GeneralMeshes(nx=nx, ny=ny, n_mesh=1 + len(patch), node_xy=np.ascontiguousarray(np.concatenate(node_xy)),
                         elem_nodes=np.concatenate(en).astype(np.int32), elem_nn=np.concatenate(nn), elem_mesh=np.concatenate(em),
                         nn_of_mesh=[4] + [3 if pt["tri"] else 4 for pt in patch], mesh_offset=offs, patch=patch)


def _general_overlay(gm: GeneralMeshes, j_lo: int, j_hi: int, X0: np.ndarray, threads: int = None):
    """Cell / triangle arrays of the mesh-0 element rows ``[j_lo, j_hi)`` from liboverlay_host.so (overlay_general.h)."""
    import ctypes as C
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "liboverlay_host.so")
    if not os.path.exists(path):
        raise RuntimeError(f"{path} not found: build it with `python -m mofem_b200.build`")
    L = C.CDLL(path)
    L.ovl_general_build.restype = C.c_void_p
    L.ovl_general_error.restype = C.c_char_p
    L.ovl_general_error.argtypes = [C.c_void_p]
    L.ovl_general_sizes.argtypes = [C.c_void_p, C.c_void_p]
    L.ovl_general_copy.argtypes = [C.c_void_p] * 6
    L.ovl_general_free.argtypes = [C.c_void_p]
    npm = len(gm.patch)
    px = np.array([pt["Px"] for pt in gm.patch], np.int32); py = np.array([pt["Py"] for pt in gm.patch], np.int32)
    tri = np.array([int(pt["tri"]) for pt in gm.patch], np.int32)
    e0 = np.array([pt["elem0"] for pt in gm.patch], np.int64)
    Xs = (C.c_void_p * npm)(*[pt["X"].ctypes.data for pt in gm.patch])
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    if threads is None:
        threads = int(os.environ.get("MOFEM_GEN_WORKERS", max(1, min(16, (os.cpu_count() or 1)))))
    X0 = np.ascontiguousarray(X0)
    h = L.ovl_general_build(C.c_int(gm.nx), C.c_int(gm.ny), C.c_int(j_lo), C.c_int(j_hi), p(X0), C.c_int(npm), p(px), p(py), p(tri),
                            p(e0), Xs, C.c_int(threads))
    try:
        err = L.ovl_general_error(h)
        if err:
            raise RuntimeError("overlay: " + err.decode())
        sz = np.zeros(2, np.int64)
        L.ovl_general_sizes(h, p(sz))
        n_cell, n_tri = int(sz[0]), int(sz[1])
        out = dict(cell_elem=np.empty((n_cell, gm.n_mesh), np.int32), cell_kind=np.empty(n_cell, np.int32),
                   cell_tri_ptr=np.empty(n_cell + 1, np.int64), tri_xy=np.empty((n_tri, 3, 2)), tri_rho=np.empty((n_tri, 3, gm.n_mesh)))
        L.ovl_general_copy(h, p(out["cell_elem"]), p(out["cell_kind"]), p(out["cell_tri_ptr"]), p(out["tri_xy"]), p(out["tri_rho"]))
    finally:
        L.ovl_general_free(h)
    return out


def general_family(n0: int, family: str, jitter: float = 0.12, seed: int = SEED, E: float = 1000.0, nu: float = 0.3,
                   solver: int = SOLVER_LOWER, ny: int = None, elem_rows=None, threads: int = None):
    """Flattened problem of a general family + its linear system, like :func:`patch_family`:

    * ``"tri"``  -- BASELINE config 3 variant b: the 2-mesh family with every patch quad split into two 3-node triangles
      (quad4 x tri3 cells: closed-form ``invTri``, the 4-point rule of ``n1 + n2 = 7``);
    * ``"m3s"``  -- BASELINE config 4: a third mesh of staggered patches at ``((8a+4.71)h, (8b+4.19)h)`` that overlaps
      mesh 0 everywhere and the mesh-1 patches partially: cells with 1, 2 and 3 incident elements, every ``l > j`` block of
      ``AMORE.py:100-123``;
    * ``"quad"`` -- the 2-mesh family again (cross-check against :func:`patch_family`).

    The overlay comes from ``liboverlay_host.so`` (``csrc/overlay_general.h``: clipping per mesh-0 element), pinned against
    the reference's plane sweep by the ``patch16_tri`` / ``patch16_m3s`` / ``patch32_m3s`` fixtures.
    ``elem_rows=(j_lo, j_hi)`` emits only the cells of those mesh-0 element rows (strips for the row-partitioned solve)."""
    gm = general_meshes(n0, family, jitter, seed, ny)
    nx, ny = gm.nx, gm.ny
    j_lo, j_hi = (0, ny) if elem_rows is None else (int(elem_rows[0]), int(elem_rows[1]))
    if not (0 <= j_lo < j_hi <= ny):
        raise ValueError("elem_rows out of range")
    n_node0 = (nx + 1) * (ny + 1)
    cells = _general_overlay(gm, j_lo, j_hi, gm.node_xy[:n_node0].reshape(ny + 1, nx + 1, 2), threads)
    fp = FlatProblem(
        solver=int(solver), n_mesh=gm.n_mesh, node_xy=gm.node_xy, elem_nodes=gm.elem_nodes, elem_nn=gm.elem_nn,
        elem_mat=np.zeros(len(gm.elem_nn), np.int32), elem_mesh=gm.elem_mesh, mat_D=material_matrix([E, nu], 1).reshape(1, 3, 3),
        n_mesh_regular=0, mesh_offset=list(gm.mesh_offset), **cells)
    n_dof = 2 * len(gm.node_xy)
    fixed = np.zeros(n_dof, dtype=np.uint8)
    left = np.arange(ny + 1) * (nx + 1)
    fixed[2 * left] = 1; fixed[2 * left + 1] = 1
    rhs = np.zeros(n_dof)
    rhs[2 * nx] = 1.0; rhs[2 * nx + 1] = -0.5
    return fp, rhs, fixed, np.zeros(n_dof)


def general_strip(n0: int, family: str, ny: int, rank: int, world: int, jitter: float = 0.12, seed: int = SEED, threads: int = None):
    """What rank ``rank`` of ``world`` needs from a general family on an ``n0 x ny`` domain cut into ``world`` strips of equal
    height (``ny / world`` element rows, a multiple of 8): the cells of its element rows plus a margin of 3 rows (a superset
    of every cell that touches a node it owns -- :func:`mofem_b200.partition.partition_problem` keeps what is needed), the
    node owner array (a node belongs to the strip its nominal row lies in), and the number of cells in exactly its rows.
    Returns ``(FlatProblem, rhs, fixed, fixed_value, owner, own_cells)``; node and element ids are global."""
    if ny % world or (ny // world) % PERIOD:
        raise ValueError("ny / world must be a multiple of 8")
    rows_per_rank = ny // world
    j_lo, j_hi = rank * rows_per_rank, (rank + 1) * rows_per_rank
    strip, rhs, fixed, val = general_family(n0, family, jitter, seed, ny=ny, elem_rows=(max(j_lo - 3, 0), min(j_hi + 3, ny)), threads=threads)
    gm = general_meshes(n0, family, jitter, seed, ny)
    n_node0 = (n0 + 1) * (ny + 1)
    owner = np.empty(strip.n_node, np.int32)
    owner[:n_node0] = np.minimum((np.arange(n_node0) // (n0 + 1)) // rows_per_rank, world - 1)
    for pt in gm.patch:
        b, a, t, s_ = np.meshgrid(np.arange(pt["Py"]), np.arange(pt["Px"]), np.arange(6), np.arange(6), indexing="ij")
        row = np.floor(PERIOD * b + pt["off_y"] + t).astype(np.int64).reshape(-1)
        owner[pt["node0"]:pt["node0"] + row.size] = np.minimum(row // rows_per_rank, world - 1)
    erow = strip.cell_elem[:, 0] // n0
    own_cells = int(((erow >= j_lo) & (erow < j_hi)).sum())
    return strip, rhs, fixed, val, owner, own_cells
