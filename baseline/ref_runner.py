"""Runs the UNMODIFIED reference staged under baseline/_ref/MOFEM (baseline/stage_ref.py) -- the CPU side of bench.py and of
the end-to-end femMain test.  Always executed as a subprocess (the reference needs its own cwd, a matplotlib import stub and
its top-level package names on sys.path); prints ONE JSON line on stdout.

    python baseline/ref_runner.py sample <problem.npz> <n_proc>      array level: stiffnessMatrix.* over the cells of a flattened
                                                                      (sub)problem, following the cell loop of AMORE.py:59-123,
                                                                      split over n_proc forked processes (cells are independent)
    python baseline/ref_runner.py entry <stem> <ref|b200> [out.npz]  plane sweep + rho + triangulation (untimed preprocessing,
                                                                      reported), then the solver entry point femMain.py would
                                                                      call -- the reference's, or mofem_b200.solvers' drop-in
    python baseline/ref_runner.py femmain <stem> <ref|b200> <workdir> femMain.py itself, unmodified, fed <stem> on stdin; with
                                                                      b200 the two solver modules it imports are the drop-ins

Nothing here is imported by mofem_b200/.
"""
import contextlib
import io
import json
import os
import re
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
TREE = os.path.join(HERE, "_ref", "MOFEM")
STUB = os.path.join(HERE, "_ref", "_mplstub")


def _enter(workdir=None):
    if not os.path.isdir(TREE):
        print(json.dumps({"unavailable": "baseline/_ref/MOFEM not staged (python baseline/stage_ref.py, build container only)"}))
        raise SystemExit(0)
    sys.path.insert(0, REPO)
    sys.path.insert(0, TREE)
    sys.path.insert(0, STUB)
    os.chdir(workdir or TREE)


PHASES = (("assembly_s", r"Assembling stiffness matrix costs ([0-9.eE+-]+) seconds"),
          ("force_s", r"Calculating force term costs ([0-9.eE+-]+) seconds"),
          ("constraints_s", r"Imposing constraints costs ([0-9.eE+-]+) seconds"),
          ("solve_s", r"Solving the linear system costs ([0-9.eE+-]+) seconds"))


def _phases(text):
    out = {}
    for key, pat in PHASES:
        m = re.search(pat, text)
        if m:
            out[key] = float(m.group(1))
    return out


# ---- array level ----------------------------------------------------------------------------------------------------------
def _cells_loop(fp, lo, hi):
    """The reference's element functions over cells [lo, hi) in the order of AMORE.py:59-123 / 328-398 / 751-821."""
    import numpy as np
    from elementLibrary import stiffnessMatrix
    amore = stiffnessMatrix.lowerOrderAMORE if fp["solver"] == 1 else stiffnessMatrix.quadraticAMORE
    node_xy, en, nn = fp["node_xy"], fp["elem_nodes"], fp["elem_nn"]
    total, ntrip = 0.0, 0

    def coords(e):
        n = nn[e]
        return node_xy[en[e, :n]].copy(), [int(v) for v in en[e, :n]]
    for c in range(lo, hi):
        ce = fp["cell_elem"][c]
        if fp["cell_kind"][c] == 0:
            e = [x for x in ce if x >= 0][0]
            coord, num = coords(e)
            D = fp["mat_D"][fp["elem_mat"][e]]
            n = len(num)
            if n == 3: K = stiffnessMatrix.triFE(coord, D)
            elif n == 4: K = stiffnessMatrix.ICMFE(coord, D)[0] if fp["solver"] == 3 else stiffnessMatrix.quadFE(coord, D)
            elif n == 6: K = stiffnessMatrix.triQuadFE(coord, D)
            else: K = stiffnessMatrix.quadQuadFE(coord, D)
            _, _, v = stiffnessMatrix.sparsifyElementMatrix(K, num)
            total += float(np.sum(v)); ntrip += len(v)
            continue
        tris = range(fp["cell_tri_ptr"][c], fp["cell_tri_ptr"][c + 1])
        M = len(ce)
        for a in range(M):
            if ce[a] < 0: continue
            coord, num = coords(ce[a])
            D = fp["mat_D"][fp["elem_mat"][ce[a]]]
            K = np.zeros((2 * len(num), 2 * len(num)))
            for t in tris:
                K += amore(fp["tri_xy"][t].copy(), D, coord, fp["tri_rho"][t, :, a].copy())
            _, _, v = stiffnessMatrix.sparsifyElementMatrix(K, num)
            total += float(np.sum(v)); ntrip += len(v)
            for b in range(a + 1, M):
                if ce[b] < 0: continue
                coord2, num2 = coords(ce[b])
                K = np.zeros((2 * len(num), 2 * len(num2)))
                for t in tris:
                    K += amore(fp["tri_xy"][t].copy(), D, coord, fp["tri_rho"][t, :, a].copy(), coord2, fp["tri_rho"][t, :, b].copy())
                _, _, v = stiffnessMatrix.sparsifyElementMatrix(K, num, num2)
                _, _, v2 = stiffnessMatrix.sparsifyElementMatrix(K.transpose(), num2, num)
                total += float(np.sum(v)) + float(np.sum(v2)); ntrip += len(v) + len(v2)
    return total, ntrip


def _sample_worker(args):
    path, lo, hi = args
    import numpy as np
    fp = dict(np.load(path))
    fp["solver"] = int(fp["meta"][0])
    t0 = time.perf_counter()
    total, ntrip = _cells_loop(fp, lo, hi)
    return time.perf_counter() - t0, total, ntrip


def cmd_sample(path, n_proc):
    _enter()
    import numpy as np
    from linearSolvers import traditionalElement, AMORE  # noqa: F401  (import order of the reference: the two modules import each other)
    path = os.path.abspath(path) if os.path.isabs(path) else os.path.join(REPO, path)
    n_cell = int(np.load(path)["cell_kind"].shape[0])
    n_proc = max(1, min(int(n_proc), n_cell))
    bounds = [n_cell * k // n_proc for k in range(n_proc + 1)]
    jobs = [(path, bounds[k], bounds[k + 1]) for k in range(n_proc)]
    t0 = time.perf_counter()
    if n_proc == 1:
        res = [_sample_worker(jobs[0])]
    else:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(n_proc) as pool:
            res = pool.map(_sample_worker, jobs)
    wall = time.perf_counter() - t0
    busy = [r[0] for r in res]
    print(json.dumps({"cells": n_cell, "processes": n_proc, "wall_s": wall, "cells_per_s": n_cell / wall,
                      "cells_per_s_per_process": n_cell / sum(busy), "coo_triplets": int(sum(r[2] for r in res)),
                      "sum_values": float(sum(r[1] for r in res))}))


# ---- entry-point level -------------------------------------------------------------------------------------------------------
def _overlay(input_data):
    """femMain.py:77-84, verbatim calls."""
    import copy
    from computGeometry import planeSweepMeshes, triangulation
    from meshTools import toDC
    coordinates = input_data[5]
    meshes = input_data[3]
    number_of_mesh = len(meshes)
    (meshes, overlapping_mesh) = planeSweepMeshes.elementFilter(coordinates, meshes)
    vertex_list, face_list, half_edge_list, line_list, nep = toDC.toDoublyConnected(coordinates, overlapping_mesh)
    polygons, _ = planeSweepMeshes.meshOverlay(line_list, vertex_list, half_edge_list, face_list, number_of_mesh)
    planeSweepMeshes.rhoCalculation(polygons)
    for p in polygons:
        triangulation.polygonTriangulation(p)
    return overlapping_mesh, polygons, nep


def cmd_entry(stem, impl, out=None, repeat=1):
    _enter()
    import numpy as np
    from linearSolvers import traditionalElement, AMORE
    from inputFunctions import inputF
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        t0 = time.perf_counter()
        input_data = inputF.txtInput(stem)
        t_in = time.perf_counter() - t0
        multi = len(input_data[3]) > 1
        t0 = time.perf_counter()
        if multi:
            overlapping_mesh, polygons, nep = _overlay(input_data)
        t_ov = time.perf_counter() - t0
    solver = input_data[0][0]
    if impl == "b200":
        import mofem_b200.solvers as mod
        single = mod
    else:
        mod, single = AMORE, traditionalElement
    fn = ({1: mod.lowerOrderAMORE, 2: mod.quadraticAMORE, 3: mod.quadraticICMAMORE}[solver] if multi else
          {1: single.lowerOrderFE, 2: single.quadraticFE, 3: single.ICMFE}[solver])
    runs = []
    import copy
    for _ in range(int(repeat)):
        data = copy.deepcopy(input_data)       # the entry points sort inputData[-3] in place
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            t0 = time.perf_counter()
            res = fn(data, overlapping_mesh, polygons, nep) if multi else fn(data)
            wall = time.perf_counter() - t0
        ph = _phases(buf.getvalue())
        ph["entry_s"] = wall
        if impl == "b200":     # device-side times of the same call (CUDA events on the context's stream) and the PCG outcome
            ctx = mod._context()
            tm = ctx.timing()
            ph["device_ms"] = {k: tm[k] for k in ("symbolic_ms", "integrate_ms", "numeric_ms", "solve_ms", "h2d_ms", "d2h_ms")}
            ph["gpu_launches"] = int(sum(tm["launches"].values())) + 2
            ph["pcg"] = dict(mod.last_solve_info)
            inf = ctx.info()
            ph["nnz"], ph["coo_triplets"] = inf["nnz"], inf["n_coo"]
        runs.append(ph)
    disp, energy = np.asarray(res[0]).reshape(-1), float(np.asarray(res[1]).reshape(-1)[0])
    if out:
        np.savez(out, displacement=disp, energy=np.array([energy]))
    n_cells = len(polygons) if multi else 0
    print(json.dumps({"stem": stem, "impl": impl, "solver": int(solver), "cells": n_cells, "elements": int(sum(len(m) for m in input_data[3])),
                      "dof": int(disp.size), "input_s": t_in, "overlay_s": t_ov, "runs": runs, "energy": energy}))


def cmd_femmain(stem, impl, workdir):
    os.makedirs(workdir, exist_ok=True)
    workdir = os.path.abspath(workdir)
    link = os.path.join(workdir, "inputFiles")
    if not os.path.exists(link):
        os.symlink(os.path.join(TREE, "inputFiles"), link)
    _enter(workdir)
    import linearSolvers
    from linearSolvers import traditionalElement, AMORE  # noqa: F401
    if impl == "b200":
        # INTEGRATION.md section 1 without editing the file: the two module objects femMain.py imports are the drop-ins
        import mofem_b200.solvers as shim
        for name in ("AMORE", "traditionalElement"):
            sys.modules["linearSolvers." + name] = shim
            setattr(linearSolvers, name, shim)
        # ... and the optional sampler swap (femMain.py:103,112): same arguments, same SampleData.txt / nSampling.txt
        from outputFunctions import writeSolution
        import mofem_b200.sampler as gpu_sampler
        writeSolution.writeSampleData2D = gpu_sampler.writeSampleData2D
    import runpy
    buf = io.StringIO()
    stdin = sys.stdin
    sys.stdin = io.StringIO(stem + "\n")
    t0 = time.perf_counter()
    try:
        with contextlib.redirect_stdout(buf):
            runpy.run_path(os.path.join(TREE, "femMain.py"), run_name="__main__")
    finally:
        sys.stdin = stdin
    wall = time.perf_counter() - t0
    text = buf.getvalue()
    ph = _phases(text)
    m = re.search(r"FE solver costs ([0-9.eE+-]+) seconds in total", text)
    print(json.dumps({"stem": stem, "impl": impl, "wall_s": wall, "fe_solver_s": float(m.group(1)) if m else None, "phases": ph,
                      "files": sorted(os.listdir(workdir))}))


if __name__ == "__main__":
    cmd = sys.argv[1]
    if cmd == "sample":
        cmd_sample(sys.argv[2], sys.argv[3])
    elif cmd == "entry":
        cmd_entry(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 and sys.argv[4] != "-" else None,
                  sys.argv[5] if len(sys.argv) > 5 else 1)
    elif cmd == "femmain":
        cmd_femmain(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        raise SystemExit("unknown command")
