"""Rebuilds, from a ``tests/golden/rho_*.npz`` fixture (make_golden_rho.py), the part of the reference's overlay object graph
that ``planeSweepMeshes.rhoCalculation`` reads: faces with ``.outer`` half-edge rings and ``.incidentElements``, elements
with ``.vertex`` (x, y, weight), ``.numbering``, ``.position``, ``.polygons``; vertices with ``.rho = None``."""
import os
import types

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RHO_FIXTURES = ("simpleOFE3", "AMOREBracket", "patch16", "patch16_m3", "patch16_tri")


class _Obj(types.SimpleNamespace):
    __hash__ = object.__hash__

    def __eq__(self, other):
        return self is other

    def __bool__(self):
        return True


def load_rho(name):
    return dict(np.load(os.path.join(GOLDEN, f"rho_{name}.npz")))


def build_graph(d):
    """(polygons, vertices) as duck-typed objects."""
    verts = [_Obj(x=float(x), y=float(y), rho=None) for x, y in d["vert_xy"]]
    elems = []
    for k in range(len(d["elem_nv"])):
        nv = int(d["elem_nv"][k])
        vertex = [_Obj(x=float(d["elem_xy"][k, a, 0]), y=float(d["elem_xy"][k, a, 1]), weight=float(d["elem_w"][k, a]))
                  for a in range(nv)]
        elems.append(_Obj(vertex=vertex, numbering=list(range(int(d["elem_nn"][k]))), position=(int(d["elem_mesh"][k]), k),
                          polygons=[None] * int(d["elem_npoly"][k])))
    polygons = []
    ptr, ring = d["face_ring_ptr"], d["face_ring"]
    for f in range(len(ptr) - 1):
        edges = [_Obj(origin=verts[v], next=None) for v in ring[ptr[f]:ptr[f + 1]]]
        for a, e in enumerate(edges):
            e.next = edges[(a + 1) % len(edges)]
        polygons.append(_Obj(outer=edges[0], incidentElements=[None if e < 0 else elems[e] for e in d["face_elem"][f]]))
    return polygons, verts
