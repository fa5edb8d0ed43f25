"""Generate the golden fixtures in tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference (junbinhuang/MOFEM) is imported from /root/reference with a
matplotlib import stub (tests/golden/_mplstub) and cwd=/root/reference (its
parser opens ./inputFiles/<stem>.txt, inputF.py:44).  Nothing is copied from it.

Fixtures written (npz, float64 / int32):
  simpleOFE3.npz, AMOREBracket.npz  -- entry-point level: flattened inputs (our
      flatten() applied to the reference's live object graph), the reference's
      own COO triplets / CSR / reduced system / displacement / energy captured by
      hooking scipy.sparse.coo_matrix and spsolve inside linearSolvers/AMORE.py.
  single_*.npz  -- traditionalElement.{lowerOrderFE,quadraticFE,ICMFE} on small
      single-mesh decks generated here.
  patch16.npz -- the synthetic patch family (n0=16, jittered) as a reference deck pushed through the
      reference's plane sweep + AMORE.lowerOrderAMORE: pins mofem_b200/overlay_gen.py.
  shipped_solution_AMOREBracket.npz -- the numbers of the reference's shipped solution.txt (its only golden output).
  zoo_s{1,2,3}.npz -- array level: synthetic all-types problems pushed through
      stiffnessMatrix.{lowerOrderAMORE,quadraticAMORE,triFE,...} following the
      reference's cell loop.
  newton.npz -- invQuad / invQuadQuad / invTriQuad / invTri known answers.
"""
import copy
import io
import pickle
import os
import sys
import contextlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("MOFEM_REFERENCE", "/root/reference")

sys.path.insert(0, os.path.join(HERE, "_mplstub"))
sys.path.insert(0, REF)
sys.path.insert(0, REPO)
os.chdir(REF)

import scipy.sparse  # noqa: E402
from linearSolvers import traditionalElement, AMORE  # noqa: E402  (order matters: circular import)
from elementLibrary import stiffnessMatrix, invShapeFunction  # noqa: E402
from inputFunctions import inputF  # noqa: E402
from computGeometry import planeSweepMeshes, triangulation  # noqa: E402
from meshTools import toDC  # noqa: E402

from mofem_b200.flatten import flatten, FlatProblem  # noqa: E402
from mofem_b200 import synthetic  # noqa: E402


class Capture:
    """Hooks scipy.sparse.coo_matrix and the solver modules' spsolve."""

    def __init__(self):
        self.coo = None
        self.reduced = None

    def __enter__(self):
        self._coo = scipy.sparse.coo_matrix
        cap = self

        def coo_hook(arg, *a, **k):
            if isinstance(arg, tuple) and len(arg) == 2 and isinstance(arg[1], tuple):
                V, (I, J) = arg
                cap.coo = (np.array(I), np.array(J), np.array(V))
            return cap._coo(arg, *a, **k)

        scipy.sparse.coo_matrix = coo_hook
        self._sp = (AMORE.spsolve, traditionalElement.spsolve)

        def sp_hook(K, f):
            x = cap._sp[0](K, f)
            cap.reduced = (K.tocsr().copy(), np.array(f).copy(), np.array(x).copy())
            return x

        AMORE.spsolve = sp_hook
        traditionalElement.spsolve = sp_hook
        return self

    def __exit__(self, *a):
        scipy.sparse.coo_matrix = self._coo
        AMORE.spsolve, traditionalElement.spsolve = self._sp


def overlay(input_data):
    """femMain.py:76-84 (second overlay pass; the first one only feeds plotting)."""
    meshes = input_data[3]
    coordinates = input_data[5]
    number_of_mesh = len(meshes)
    meshes, overlapping_mesh = planeSweepMeshes.elementFilter(coordinates, meshes)
    vertex_list, face_list, half_edge_list, line_list, node_elements_partial = \
        toDC.toDoublyConnected(coordinates, overlapping_mesh)
    polygons, _ = planeSweepMeshes.meshOverlay(line_list, vertex_list, half_edge_list, face_list, number_of_mesh)
    planeSweepMeshes.rhoCalculation(polygons)
    for p in polygons:
        triangulation.polygonTriangulation(p)
    return overlapping_mesh, polygons, node_elements_partial


def reference_array_level(fp: FlatProblem):
    """Reference COO values for a FlatProblem via the reference's element functions,
    following the cell loop of AMORE.py:59-123 / 328-398 / 751-821."""
    V = []
    quadratic = fp.solver != 1
    amore = stiffnessMatrix.quadraticAMORE if quadratic else stiffnessMatrix.lowerOrderAMORE
    I, J = [], []

    def coords(e):
        n = fp.elem_nn[e]
        return fp.node_xy[fp.elem_nodes[e, :n]].copy(), [int(v) for v in fp.elem_nodes[e, :n]]

    def tri_coord(t):
        if fp.tri_curved is not None and fp.tri_curved[t]:
            return np.vstack([fp.tri_xy[t], fp.tri_mid[t]])
        return fp.tri_xy[t].copy()

    for c in range(fp.n_cell):
        ce = fp.cell_elem[c]
        if fp.cell_kind[c] == 0:
            e = [x for x in ce if x >= 0][0]
            coord, num = coords(e)
            D = fp.mat_D[fp.elem_mat[e]]
            n = len(num)
            if n == 3: K = stiffnessMatrix.triFE(coord, D)
            elif n == 4: K = stiffnessMatrix.ICMFE(coord, D)[0] if fp.solver == 3 else stiffnessMatrix.quadFE(coord, D)
            elif n == 6: K = stiffnessMatrix.triQuadFE(coord, D)
            else: K = stiffnessMatrix.quadQuadFE(coord, D)
            i, j, v = stiffnessMatrix.sparsifyElementMatrix(K, num)
            I += i; J += j; V += v
            continue
        tris = range(fp.cell_tri_ptr[c], fp.cell_tri_ptr[c + 1])
        for a in range(fp.n_mesh):
            if ce[a] < 0: continue
            coord, num = coords(ce[a])
            D = fp.mat_D[fp.elem_mat[ce[a]]]
            K = np.zeros((2 * len(num), 2 * len(num)))
            for t in tris:
                K += amore(tri_coord(t), D, coord, fp.tri_rho[t, :, a].copy())
            i, j, v = stiffnessMatrix.sparsifyElementMatrix(K, num)
            I += i; J += j; V += v
            for b in range(a + 1, fp.n_mesh):
                if ce[b] < 0: continue
                coord2, num2 = coords(ce[b])
                K = np.zeros((2 * len(num), 2 * len(num2)))
                for t in tris:
                    K += amore(tri_coord(t), D, coord, fp.tri_rho[t, :, a].copy(), coord2, fp.tri_rho[t, :, b].copy())
                i, j, v = stiffnessMatrix.sparsifyElementMatrix(K, num, num2)
                I += i; J += j; V += v
                i, j, v = stiffnessMatrix.sparsifyElementMatrix(K.transpose(), num2, num)
                I += i; J += j; V += v
    return np.array(I, dtype=np.int64), np.array(J, dtype=np.int64), np.array(V, dtype=np.float64)


def pack(fp, I, J, V, extra=None):
    n = fp.n_dof
    K = scipy.sparse.coo_matrix((V, (I, J)), shape=(n, n)).tocsr()
    d = fp.to_npz_dict()
    d.update(coo_I=I.astype(np.int32), coo_J=J.astype(np.int32), coo_V=V,
             csr_indptr=K.indptr.astype(np.int32), csr_indices=K.indices.astype(np.int32), csr_data=K.data)
    if extra:
        d.update(extra)
    return d


def dump_graph(input_data, overlapping_mesh, polygons, nep):
    """The overlay object graph as plain builtins (tests/graph_fixture.py rebuilds duck-typed objects from it, so the
    drop-in entry points can be called with the reference's own argument structure where /root/reference is absent)."""
    pts, hes = {}, {}
    P, H = [], []

    def pid(p):
        k = id(p)
        if k not in pts:
            pts[k] = len(P)
            P.append([float(p.x), float(p.y), None if p.rho is None else [float(v) for v in p.rho]])
        return pts[k]

    def hid(h, follow=True):
        k = id(h)
        if k not in hes:
            hes[k] = len(H)
            H.append([pid(h.origin), pid(h.destination), -1, -1])
        rec = H[hes[k]]
        if follow and rec[2] == -1:
            rec[2] = -2                      # being expanded
            rec[2] = hid(h.next)
            if getattr(h, "twin", None) is not None:
                rec[3] = hid(h.twin, follow=False)
        return hes[k]

    face_id = {id(f): i for i, f in enumerate(polygons)}
    elem_id = {id(e): [i, j] for i, mesh in enumerate(overlapping_mesh) if mesh is not None for j, e in enumerate(mesh)}
    faces = []
    for f in polygons:
        inc = [None if e is None else elem_id[id(e)] for e in f.incidentElements]   # index into overlappingMesh
        tris = None if f.triangles is None else [hid(t) for t in f.triangles]
        faces.append([inc, tris])
    elems = []
    for mesh in overlapping_mesh:
        if mesh is None:
            elems.append(None)
            continue
        elems.append([[list(map(int, e.numbering)), [int(e.position[0]), int(e.position[1])],
                       [face_id[id(f)] for f in e.polygons]] for e in mesh])
    return {"points": P, "half_edges": H, "faces": faces, "elements": elems,
            "node_elements_partial": [[list(t) for t in lst] for lst in nep], "input_data": input_data}


def run_amore(stem, input_data=None):
    if input_data is None:
        input_data = inputF.txtInput(stem)
    solver = input_data[0][0]
    with contextlib.redirect_stdout(io.StringIO()):
        overlapping_mesh, polygons, nep = overlay(input_data)
        fp = flatten(input_data, overlapping_mesh, polygons)
        constraints_in = copy.deepcopy(input_data[-3])
        graph = dump_graph(copy.deepcopy(input_data), overlapping_mesh, polygons, nep)
        fn = {1: AMORE.lowerOrderAMORE, 2: AMORE.quadraticAMORE, 3: AMORE.quadraticICMAMORE}[solver]
        with Capture() as cap:
            out = fn(input_data, overlapping_mesh, polygons, nep)
    disp, energy = out[0], out[1]
    sample = None
    if solver in (2, 3):      # femMain.py:103,112 -- the stress / displacement sampler on the solution
        import tempfile
        from outputFunctions import writeSolution
        cwd = os.getcwd()
        with tempfile.TemporaryDirectory() as tmp:
            os.chdir(tmp)
            try:
                with contextlib.redirect_stdout(io.StringIO()):
                    writeSolution.writeSampleData2D(input_data[5], input_data[3], polygons, disp, out[2], input_data[4],
                                                    component='All', nSampling=3, solver=solver)
                sample = np.loadtxt("SampleData.txt")
            finally:
                os.chdir(cwd)
    I, J, V = cap.coo
    # cross-check: array-level path on OUR flattening reproduces the entry-point triplets
    I2, J2, V2 = reference_array_level(fp)
    assert np.array_equal(I, I2) and np.array_equal(J, J2), "emission order mismatch"
    assert np.array_equal(V, V2), "flatten() does not reproduce the reference's triplets bit-for-bit"
    Kr, fr, xr = cap.reduced
    extra = dict(displacement=disp.reshape(-1), energy=np.array(energy).reshape(-1),
                 red_indptr=Kr.indptr.astype(np.int32), red_indices=Kr.indices.astype(np.int32), red_data=Kr.data,
                 red_rhs=fr.reshape(-1), red_x=xr.reshape(-1),
                 constraints=np.array(constraints_in, dtype=np.float64),
                 graph=np.frombuffer(pickle.dumps(graph, protocol=4), dtype=np.uint8))
    if sample is not None:
        extra["sample_all"] = sample
    print(stem, "cells", fp.n_cell, "tris", fp.n_tri, "coo", len(V), "energy", float(np.array(energy).reshape(-1)[0]))
    return pack(fp, I, J, V, extra), (input_data, overlapping_mesh, polygons, nep)


def single_mesh_deck(kind, nx=5, ny=4, seed=3):
    """inputData tuple for traditionalElement.* (a cantilever strip; never written to disk)."""
    rng = np.random.default_rng(seed)
    coords, meshes = [], []
    if kind == "lower":      # mixed quad4 / tri3
        for j in range(ny + 1):
            for i in range(nx + 1):
                jit = rng.uniform(-0.1, 0.1, 2) if 0 < i < nx and 0 < j < ny else np.zeros(2)
                coords.append([i + jit[0], 0.5 * j + jit[1]])
        nid = lambda i, j: j * (nx + 1) + i
        for j in range(ny):
            for i in range(nx):
                q = [nid(i, j), nid(i + 1, j), nid(i + 1, j + 1), nid(i, j + 1)]
                if (i + j) % 3 == 0:
                    meshes += [[q[0], q[1], q[2]], [q[0], q[2], q[3]]]
                else:
                    meshes.append(q)
        params = [1, 1]
    elif kind == "icm":      # squares only (ICMFE assumes squares)
        for j in range(ny + 1):
            for i in range(nx + 1):
                coords.append([0.5 * i, 0.5 * j])
        nid = lambda i, j: j * (nx + 1) + i
        for j in range(ny):
            for i in range(nx):
                meshes.append([nid(i, j), nid(i + 1, j), nid(i + 1, j + 1), nid(i, j + 1)])
        params = [3, 2]
    else:                    # quadratic: quad9 + tri6 with some None mid-nodes
        gx, gy = 2 * nx + 1, 2 * ny + 1
        for j in range(gy):
            for i in range(gx):
                if (i % 2 == 1 or j % 2 == 1) and rng.random() < 0.5:
                    coords.append(None)      # mid node completed by getCoord
                else:
                    jit = rng.uniform(-0.04, 0.04, 2) if 0 < i < gx - 1 and 0 < j < gy - 1 and i % 2 == 0 and j % 2 == 0 else np.zeros(2)
                    coords.append([0.5 * i + jit[0], 0.25 * j + jit[1]])
        # make sure explicitly given mid nodes sit near their edge midpoint (filled in below)
        nid = lambda i, j: j * gx + i
        def corner(i, j): return np.array(coords[nid(i, j)])
        for j in range(gy):
            for i in range(gx):
                if coords[nid(i, j)] is None: continue
                if i % 2 == 1 and j % 2 == 0:
                    coords[nid(i, j)] = list(0.5 * (corner(i - 1, j) + corner(i + 1, j)) + rng.uniform(-0.01, 0.01, 2) * (0 < j < gy - 1))
                elif i % 2 == 0 and j % 2 == 1:
                    coords[nid(i, j)] = list(0.5 * (corner(i, j - 1) + corner(i, j + 1)) + rng.uniform(-0.01, 0.01, 2) * (0 < i < gx - 1))
                elif i % 2 == 1 and j % 2 == 1:
                    coords[nid(i, j)] = list(0.25 * (corner(i - 1, j - 1) + corner(i + 1, j - 1) + corner(i + 1, j + 1) + corner(i - 1, j + 1)))
        for j in range(ny):
            for i in range(nx):
                I0, J0 = 2 * i, 2 * j
                if (i + j) % 3 == 0:
                    # two 6-node triangles; the diagonal mid node is the quad centre node
                    a, b, c, d = nid(I0, J0), nid(I0 + 2, J0), nid(I0 + 2, J0 + 2), nid(I0, J0 + 2)
                    meshes.append([a, b, c, nid(I0 + 1, J0), nid(I0 + 2, J0 + 1), nid(I0 + 1, J0 + 1)])
                    meshes.append([a, c, d, nid(I0 + 1, J0 + 1), nid(I0 + 1, J0 + 2), nid(I0, J0 + 1)])
                else:
                    meshes.append([nid(I0, J0), nid(I0 + 2, J0), nid(I0 + 2, J0 + 2), nid(I0, J0 + 2),
                                   nid(I0 + 1, J0), nid(I0 + 2, J0 + 1), nid(I0 + 1, J0 + 2), nid(I0, J0 + 1),
                                   nid(I0 + 1, J0 + 1)])
        params = [2, 1]
        # diagonal mid nodes of triangles must be straight (centre nodes that are None are completed per element
        # by different rules for quad9 vs tri6, so give every centre node an explicit coordinate)
        for j in range(ny):
            for i in range(nx):
                k = nid(2 * i + 1, 2 * j + 1)
                if coords[k] is None:
                    coords[k] = list(0.25 * (corner(2 * i, 2 * j) + corner(2 * i + 2, 2 * j) + corner(2 * i + 2, 2 * j + 2) + corner(2 * i, 2 * j + 2)))
    stride = (nx + 1) if kind != "quad" else (2 * nx + 1)
    rows = (ny + 1) if kind != "quad" else (2 * ny + 1)
    constraints = [[j * stride, 1, 1, 0.0, 0.01 * (j % 2)] for j in range(rows) if coords[j * stride] is not None]
    # traction on the right edge: force pairs (node, fx, fy)
    force_pairs = []
    step = 1 if kind != "quad" else 2
    for j in range(0, rows - step, step):
        n1, n2 = j * stride + stride - 1, (j + step) * stride + stride - 1
        force_pairs += [[n1, 1.0, -0.3], [n2, 1.0 + 0.1 * j, -0.3]]
    mats = [[1000.0, 0.3]]
    material_mesh = [[0] * len(meshes)]
    return (params, None, None, [meshes], material_mesh, coords, mats, constraints, [0, 0], force_pairs)


def run_single(kind):
    input_data = single_mesh_deck(kind)
    fn = {"lower": traditionalElement.lowerOrderFE, "icm": traditionalElement.ICMFE,
          "quad": traditionalElement.quadraticFE}[kind]
    fp = flatten(input_data)
    deck = copy.deepcopy(input_data)
    with contextlib.redirect_stdout(io.StringIO()):
        with Capture() as cap:
            disp, energy = fn(input_data)
    I, J, V = cap.coo
    I2, J2, V2 = reference_array_level(fp)
    assert np.array_equal(I, I2) and np.array_equal(J, J2) and np.array_equal(V, V2)
    Kr, fr, xr = cap.reduced
    import pickle
    extra = dict(displacement=disp.reshape(-1), energy=np.array(energy).reshape(-1),
                 red_indptr=Kr.indptr.astype(np.int32), red_indices=Kr.indices.astype(np.int32), red_data=Kr.data,
                 red_rhs=fr.reshape(-1), red_x=xr.reshape(-1),
                 deck=np.frombuffer(pickle.dumps(deck, protocol=4), dtype=np.uint8))
    print("single", kind, "elems", fp.n_cell, "coo", len(V), "energy", float(np.array(energy).reshape(-1)[0]))
    return pack(fp, I, J, V, extra)


def newton_vectors(seed=11):
    rng = np.random.default_rng(seed)
    out = {}
    for n, fn in ((4, invShapeFunction.invQuad), (9, invShapeFunction.invQuadQuad), (6, invShapeFunction.invTriQuad),
                  (3, invShapeFunction.invTri)):
        C, X, ISO = [], [], []
        for _ in range(40):
            c = synthetic._element_coords(n, rng, rng.uniform(-5, 5, 2), rng.uniform(0.1, 3), 0.2)
            xy = c[:3].mean(0) * 0.6 + c.mean(0) * 0.4 + rng.uniform(-0.05, 0.05, 2)
            iso = np.array(fn(c, xy), dtype=float)
            C.append(c); X.append(xy); ISO.append(iso[:2])
        out[f"coord{n}"] = np.array(C); out[f"xy{n}"] = np.array(X); out[f"iso{n}"] = np.array(ISO)
    return out


def general_families(names=("patch16_tri", "patch16_m3s", "patch32_m3s")):
    """The general families of mofem_b200/overlay_gen.py (BASELINE config 3 variant b: patch quads split into tri3;
    config 4: staggered third mesh, cells with 1, 2 and 3 incident elements) through the reference's own plane sweep.
    The graph dump (only needed by the drop-in tests of the shipped problems) is left out to keep the files small."""
    from mofem_b200 import overlay_gen
    planeSweepMeshes.scale = [1, 1]
    for name in names:
        n0 = int(name[5:7]); fam = name.split("_")[1]
        d, _ = run_amore(name, overlay_gen.family_deck(n0, fam))
        d.pop("graph", None)
        if n0 > 16:      # the larger fixture pins counts, pattern and values; triplet lists stay out of the repo
            for k in ("coo_I", "coo_J", "coo_V", "red_indptr", "red_indices", "red_data", "red_x"):
                d.pop(k, None)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)


def main():
    os.makedirs(HERE, exist_ok=True)
    for stem in ("simpleOFE3", "AMOREBracket"):
        d, _ = run_amore(stem)
        np.savez_compressed(os.path.join(HERE, stem + ".npz"), **d)
    # synthetic patch family (mofem_b200/overlay_gen.py) pushed through the reference's own plane sweep
    from mofem_b200 import overlay_gen
    planeSweepMeshes.scale = [1, 1]
    d, _ = run_amore("patch16", overlay_gen.patch_deck(16))
    np.savez_compressed(os.path.join(HERE, "patch16.npz"), **d)
    d, _ = run_amore("patch16_m3", overlay_gen.patch_deck(16, n_mesh=3))     # three meshes (patches alternate)
    np.savez_compressed(os.path.join(HERE, "patch16_m3.npz"), **d)
    general_families()
    # the reference's shipped golden output (solution.txt, AMOREBracket, solver 3) as numbers
    from mofem_b200 import solution_io
    disp, energy, name = solution_io.read_solution(os.path.join(REF, "solution.txt"))
    np.savez_compressed(os.path.join(HERE, "shipped_solution_AMOREBracket.npz"), displacement=disp.reshape(-1),
                        energy=energy.reshape(-1), name=np.array(name))
    for kind in ("lower", "icm", "quad"):
        np.savez_compressed(os.path.join(HERE, f"single_{kind}.npz"), **run_single(kind))
    for solver in (1, 2, 3):
        fp = synthetic.zoo_problem(seed=100 + solver, solver=solver)
        I, J, V = reference_array_level(fp)
        print("zoo", solver, "cells", fp.n_cell, "tris", fp.n_tri, "coo", len(V))
        np.savez_compressed(os.path.join(HERE, f"zoo_s{solver}.npz"), **pack(fp, I, J, V))
    np.savez_compressed(os.path.join(HERE, "newton.npz"), **newton_vectors())


if __name__ == "__main__":
    main()
