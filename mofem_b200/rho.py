"""Weight functions ("rho") of the overlay vertices.

Drop-in for ``computGeometry/planeSweepMeshes.rhoCalculation:991-1030``: same argument (the ``polygons`` list the plane
sweep returns), same side effect (every vertex on the ring of an overlapping face gets ``vertex.rho`` = array of one
weight per mesh).  The host walks the rings exactly like the reference (a vertex takes the incident elements of the
FIRST face that visits it) and only gathers plain arrays; the per-vertex arithmetic -- inverse map into every incident
element, interpolation of the corner weights, the factor 9 for meshes >= 1, normalisation -- is one CUDA kernel over all
vertices (``csrc/rho.cu``, ``mofem_rho_eval``).
"""
from __future__ import annotations

import numpy as np

from . import device as _device


def not_overlap(face) -> bool:
    """computGeometry/triangulation.notOverlap:24-35."""
    n, first = 0, None
    for e in face.incidentElements:
        if e:
            n += 1
            if first is None:
                first = e
    return n == 1 and len(first.polygons) == 1


class RhoArrays:
    """Flat inputs of ``mofem_rho_eval`` + the vertex objects the result goes back to."""

    __slots__ = ("vert_xy", "vert_elem", "elem_nv", "elem_xy", "elem_w", "vertices", "n_mesh")


def flatten_rho(polygons) -> RhoArrays:
    """Walk ``polygons`` as rhoCalculation does (:993-1001, :1029-1030) and gather the vertices whose ``rho`` is still
    ``None`` together with the incident elements of the face that reaches them first."""
    verts, rows = [], []
    eid, elems = {}, []
    pending = set()
    n_mesh = 0
    for face in polygons:
        if not_overlap(face):
            continue
        inc = face.incidentElements
        n_mesh = max(n_mesh, len(inc))
        row = None
        start = edge = face.outer
        while True:
            p = edge.origin
            if p.rho is None and id(p) not in pending:
                if row is None:
                    # rho has len(incidentElements) entries and is indexed by element.position[0] (:1003,1017)
                    row = [-1] * len(inc)
                    for e in inc:
                        if e:
                            k = eid.get(id(e))
                            if k is None:
                                k = eid[id(e)] = len(elems)
                                elems.append(e)
                            row[e.position[0]] = k
                pending.add(id(p))
                verts.append(p)
                rows.append(row)
            edge = edge.next
            if edge is start:
                break
    out = RhoArrays()
    out.vertices = verts
    out.n_mesh = n_mesh
    V, E = len(verts), len(elems)
    out.vert_xy = np.array([[float(p.x), float(p.y)] for p in verts], dtype=np.float64).reshape(V, 2)
    out.vert_elem = np.full((V, max(n_mesh, 1)), -1, dtype=np.int32)
    for k, row in enumerate(rows):
        out.vert_elem[k, :len(row)] = row
    out.elem_nv = np.zeros(E, dtype=np.int32)
    out.elem_xy = np.zeros((E, 4, 2))
    out.elem_w = np.zeros((E, 4))
    for k, e in enumerate(elems):
        nn = len(e.numbering)
        # the reference branches on len(numbering) (:1014,1018): 4 / 9 -> bilinear corners, 3 / 6 -> area coordinates
        nv = 4 if nn in (4, 9) else 3 if nn in (3, 6) else 0
        out.elem_nv[k] = nv
        for a, v in enumerate(e.vertex[:4]):
            out.elem_xy[k, a] = (float(v.x), float(v.y))
            out.elem_w[k, a] = float(v.weight)
    return out


_ctx = None


def _context():
    global _ctx
    if _ctx is None:
        _ctx = _device.Context(0)
    return _ctx


def rhoCalculation(polygons, ctx=None):
    """planeSweepMeshes.rhoCalculation:991 -- sets ``edge.origin.rho`` on every vertex of every overlapping face."""
    fr = flatten_rho(polygons)
    if not fr.vertices:
        return
    ctx = ctx or _context()
    rho, _, _ = ctx.rho_eval(fr.vert_xy, fr.vert_elem, fr.elem_nv, fr.elem_xy, fr.elem_w)
    for p, row in zip(fr.vertices, rho):
        p.rho = row.copy()
