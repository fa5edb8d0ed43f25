"""Golden fixtures for the weight-function evaluation, from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_rho.py

For each deck the reference's overlay is built (elementFilter -> toDoublyConnected -> meshOverlay, femMain.py:76-83),
the inputs of ``planeSweepMeshes.rhoCalculation`` (planeSweepMeshes.py:991-1030) are dumped as plain arrays, the
reference function is run, and the ``rho`` it stored on every overlay vertex is recorded:

  vert_xy  f64[V][2]      overlay vertices, in order of first visit (polygons order, ring order from face.outer)
  elem_nv  i32[E]         3 or 4 corner vertices (element.vertex)
  elem_nn  i32[E]         len(element.numbering) (3, 4, 6 or 9)
  elem_mesh i32[E]        element.position[0]
  elem_npoly i32[E]       len(element.polygons) (triangulation.notOverlap needs it)
  elem_xy  f64[E][4][2]   corner coordinates, elem_w f64[E][4] corner weights (0 on interior mesh boundaries)
  face_elem i32[F][M]     incident element per mesh slot, -1 = none
  face_ring_ptr i64[F+1], face_ring i32[...]   vertex ids of every face ring
  rho      f64[V][M]      reference output (NaN rows: vertices the reference never visits)
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("MOFEM_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "_mplstub"))
sys.path.insert(0, REF)
sys.path.insert(0, REPO)
os.chdir(REF)

from linearSolvers import traditionalElement, AMORE  # noqa: E402,F401  (import order: circular import)
from inputFunctions import inputF  # noqa: E402
from computGeometry import planeSweepMeshes  # noqa: E402
from meshTools import toDC  # noqa: E402


def overlay_before_rho(input_data):
    meshes, coordinates = input_data[3], input_data[5]
    n_mesh = len(meshes)
    with contextlib.redirect_stdout(io.StringIO()):
        meshes, om = planeSweepMeshes.elementFilter(coordinates, meshes)
        vl, fl, hl, ll, nep = toDC.toDoublyConnected(coordinates, om)
        polygons, _ = planeSweepMeshes.meshOverlay(ll, vl, hl, fl, n_mesh)
    return polygons, n_mesh


def dump(polygons, n_mesh):
    vid, eid = {}, {}
    vert, elems = [], []
    face_elem, ring_ptr, ring = [], [0], []
    for f in polygons:
        row = []
        for e in f.incidentElements:
            if e is None:
                row.append(-1)
                continue
            if id(e) not in eid:
                eid[id(e)] = len(elems)
                elems.append(e)
            row.append(eid[id(e)])
        face_elem.append(row)
        edge = f.outer
        while True:
            p = edge.origin
            if id(p) not in vid:
                vid[id(p)] = len(vert)
                vert.append(p)
            ring.append(vid[id(p)])
            edge = edge.next
            if edge is f.outer:
                break
        ring_ptr.append(len(ring))
    E = len(elems)
    elem_xy = np.zeros((E, 4, 2)); elem_w = np.zeros((E, 4))
    for k, e in enumerate(elems):
        for a, v in enumerate(e.vertex):
            elem_xy[k, a] = [float(v.x), float(v.y)]
            elem_w[k, a] = float(v.weight)
    d = dict(vert_xy=np.array([[float(p.x), float(p.y)] for p in vert]),
             elem_nv=np.array([len(e.vertex) for e in elems], dtype=np.int32),
             elem_nn=np.array([len(e.numbering) for e in elems], dtype=np.int32),
             elem_mesh=np.array([e.position[0] for e in elems], dtype=np.int32),
             elem_npoly=np.array([len(e.polygons) for e in elems], dtype=np.int32),
             elem_xy=elem_xy, elem_w=elem_w,
             face_elem=np.array(face_elem, dtype=np.int32).reshape(len(polygons), n_mesh),
             face_ring_ptr=np.array(ring_ptr, dtype=np.int64), face_ring=np.array(ring, dtype=np.int32))
    return d, vert


def run(name, input_data):
    polygons, n_mesh = overlay_before_rho(input_data)
    d, vert = dump(polygons, n_mesh)
    assert all(p.rho is None for p in vert)
    planeSweepMeshes.rhoCalculation(polygons)
    rho = np.full((len(vert), n_mesh), np.nan)
    for k, p in enumerate(vert):
        if p.rho is not None:
            rho[k] = p.rho
    d["rho"] = rho
    print(name, "faces", len(polygons), "vertices", len(vert), "visited", int(np.isfinite(rho[:, 0]).sum()),
          "elements", len(d["elem_nv"]), "tri elements", int((d["elem_nv"] == 3).sum()))
    np.savez_compressed(os.path.join(HERE, f"rho_{name}.npz"), **d)


def tri_split(deck):
    """Variant b of the patch family: every patch quad split into two 3-node triangles."""
    deck = list(deck)
    meshes = [list(m) for m in deck[3]]
    for k in range(1, len(meshes)):
        out = []
        for q in meshes[k]:
            out += [[q[0], q[1], q[2]], [q[0], q[2], q[3]]]
        meshes[k] = out
    deck[3] = meshes
    deck[4] = [[0] * len(m) for m in meshes]
    return tuple(deck)


def main():
    for stem in ("simpleOFE3", "AMOREBracket"):
        run(stem, inputF.txtInput(stem))
    from mofem_b200 import overlay_gen
    planeSweepMeshes.scale = [1, 1]
    run("patch16", overlay_gen.patch_deck(16))
    run("patch16_m3", overlay_gen.patch_deck(16, n_mesh=3))
    run("patch16_tri", tri_split(overlay_gen.patch_deck(16)))


if __name__ == "__main__":
    main()
