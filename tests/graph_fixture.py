"""Rebuilds the reference's overlay object graph (duck-typed) from the plain dump stored in the golden fixtures
(tests/golden/make_golden.py: dump_graph), so that the drop-in entry points can be called with the reference's own
argument structure -- ``(inputData, overlappingMesh, polygons, nodeElementsPartial)`` -- where /root/reference does
not exist (the GPU box)."""
import copy
import math
import pickle
import sys
import types

import numpy as np


class _Obj(types.SimpleNamespace):
    __hash__ = object.__hash__

    def __eq__(self, other):
        return self is other


def load_graph(npz):
    g = pickle.loads(npz["graph"].tobytes())
    pts = [_Obj(x=p[0], y=p[1], rho=None if p[2] is None else np.array(p[2])) for p in g["points"]]
    hes = [_Obj(origin=pts[h[0]], destination=pts[h[1]], next=None, twin=None) for h in g["half_edges"]]
    for h, rec in zip(hes, g["half_edges"]):
        h.next = hes[rec[2]] if rec[2] >= 0 else None
        h.twin = hes[rec[3]] if rec[3] >= 0 else None
    elems = []
    for mesh in g["elements"]:
        elems.append(None if mesh is None else
                     [_Obj(numbering=e[0], position=tuple(e[1]), polygons=[], _faces=e[2]) for e in mesh])
    faces = []
    for inc, tris in g["faces"]:
        faces.append(_Obj(incidentElements=[None if p is None else elems[p[0]][p[1]] for p in inc],
                          triangles=None if tris is None else [hes[t] for t in tris]))
    for mesh in elems:
        for e in (mesh or []):
            e.polygons = [faces[k] for k in e._faces]
    nep = [[tuple(t) for t in lst] for lst in g["node_elements_partial"]]
    return copy.deepcopy(g["input_data"]), elems, faces, nep


def install_bracket_load():
    """The user-supplied traction module the AMOREBracket deck expects (the reference imports ``cForce`` from its
    repository root): pressure p = 2000 on the lower half of the hole centred at x = 27, varying linearly with the
    angle from the downward direction and pointing along the radius."""
    mod = types.ModuleType("cForce")

    def customizedForce(x, y):
        assert (y <= 1.0e-6), "Error in input!"
        p = 2000.0
        x = x - 27.0
        theta = abs(math.atan2(y, x) + math.pi / 2)
        r = math.sqrt(x ** 2 + y ** 2)
        return np.array([[2 * (math.pi / 2 - theta) / math.pi * p * x / r],
                         [2 * (math.pi / 2 - theta) / math.pi * p * y / r]])
    mod.customizedForce = customizedForce
    sys.modules["cForce"] = mod
    return mod
