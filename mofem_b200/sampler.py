"""Displacement / stress sampling of a solution and the ``SampleData.txt`` writer.

Drop-in for ``outputFunctions/writeSolution.writeSampleData2D:41`` (same positional arguments, same file and prints); the
per-point evaluation of ``plotTools/plotResults.getGraphDataFE:137`` / ``getGraphDataOFE:15`` runs as one CUDA kernel over
all regular elements and integration triangles (``csrc/sample.cu``).
"""
from __future__ import annotations

import math

import numpy as np

from . import flatten as _fl
from . import solvers as _solvers

_COMPONENT = {"u": 2, "v": 3, "Sxx": 4, "Syy": 5, "Sxy": 6}


def tri_area(tri_xy):
    """writeSolution.triArea:371 on [T,3,2]."""
    c = tri_xy
    return (c[:, 1, 0] * c[:, 2, 1] - c[:, 2, 0] * c[:, 1, 1] - c[:, 0, 0] * (c[:, 2, 1] - c[:, 1, 1])
            + c[:, 0, 1] * (c[:, 2, 0] - c[:, 1, 0])) / 2


def sample_units(ctx, fp, displacement, n_sampling, solver):
    """(values [n_unit,7,nS,nS], tri_of_unit [n_unit] (-1 = regular element)) for a flattened problem."""
    ctx.clear_halo()
    ctx.set_problem(fp)
    vals = ctx.sample(n_sampling, solver, displacement)
    regular = fp.cell_kind == _fl.CELL_REGULAR
    per_cell = np.where(regular, 1, np.diff(fp.cell_tri_ptr)).astype(np.int64)
    tri_of_unit = np.full(int(per_cell.sum()), -1, dtype=np.int64)
    start = np.concatenate([[0], np.cumsum(per_cell)])[:-1]
    ov = np.flatnonzero(~regular)
    for c in ov:                                  # O(cells) host bookkeeping
        n = per_cell[c]
        tri_of_unit[start[c]:start[c] + n] = np.arange(fp.cell_tri_ptr[c], fp.cell_tri_ptr[c] + n)
    return vals, tri_of_unit


def _write_block(f, X, Y, Z, n, lim, cmin, cmax):
    """writeSolution.writeAndUpdate:26."""
    for k1 in range(n):
        for row in (X[k1], Y[k1], Z[k1]):
            f.write("".join("%s " % v for v in row))
            f.write("\n")
        for k2 in range(n):
            z = Z[k1, k2]
            if z > lim[1]:
                lim[1] = z; cmax[0] = X[k1, k2]; cmax[1] = Y[k1, k2]
            if z < lim[0]:
                lim[0] = z; cmin[0] = X[k1, k2]; cmin[1] = Y[k1, k2]


def writeSampleData2D(coordinates, meshes, polygons, displacement, materialMatrix, materialMeshList, component='u',
                      nSampling=None, solver=3, path="SampleData.txt", n_path="nSampling.txt"):
    """writeSolution.writeSampleData2D:41 -- writes ``SampleData.txt`` / ``nSampling.txt`` into the cwd and prints the
    ranges.  ``component``: 'u', 'v', 'Sxx', 'Syy', 'Sxy' or 'All' (which adds the von Mises stress and, like the
    reference, skips integration triangles with |area| < 1e-6)."""
    if nSampling is None:
        nSampling = 5
    if component not in _COMPONENT and component != "All":
        raise ValueError
    print("%s sampling points are used in each direction." % nSampling)
    # the sampler consumes the same flattened problem as the solvers (quadratic-family triangle rule)
    params = [int(solver), 1]
    input_data = (params, None, None, meshes, materialMeshList, coordinates, [[1.0, 0.0]] * len(materialMatrix), [], [0, 0], [])
    om = None
    if polygons:
        # overlapping elements are reachable from the polygons; flatten() only needs their numbering / position
        seen = {}
        for poly in polygons:
            for el in poly.incidentElements:
                if el:
                    seen[(el.position[0], el.position[1])] = el
        om = [[el for (i, j), el in sorted(seen.items()) if i == m] for m in range(len(meshes))]
    fp = _fl.flatten(input_data, om, polygons, solver=int(solver))
    fp.mat_D = np.ascontiguousarray(np.array([np.asarray(m, dtype=np.float64) for m in materialMatrix]).reshape(-1, 3, 3))
    vals, tri_of_unit = sample_units(_solvers._context(), fp, displacement, nSampling, int(solver))
    names = ["u", "v", "Sxx", "Syy", "Sxy", "Mises"] if component == "All" else [component]
    lim = {k: [math.inf, -math.inf] for k in names}
    cmin = {k: [0.0, 0.0] for k in names}
    cmax = {k: [0.0, 0.0] for k in names}
    skip = np.zeros(len(tri_of_unit), dtype=bool)
    if component == "All":
        ov = tri_of_unit >= 0
        skip[ov] = np.abs(tri_area(fp.tri_xy[tri_of_unit[ov]])) < 1e-6        # writeSolution.py:288
    with open(path, "w") as f:
        for k in range(vals.shape[0]):
            if skip[k]:
                continue
            X, Y = vals[k, 0], vals[k, 1]
            if component == "All":
                sxx, syy, sxy = vals[k, 4], vals[k, 5], vals[k, 6]
                mises = np.zeros((nSampling, nSampling))
                for ii in range(nSampling):
                    for jj in range(nSampling):
                        mises[ii, jj] = math.sqrt(sxx[ii, jj] ** 2 + syy[ii, jj] ** 2 + 3 * sxy[ii, jj] ** 2 - sxx[ii, jj] * syy[ii, jj])
                for name, Z in zip(names, (vals[k, 2], vals[k, 3], sxx, syy, sxy, mises)):
                    _write_block(f, X, Y, Z, nSampling, lim[name], cmin[name], cmax[name])
            else:
                _write_block(f, X, Y, vals[k, _COMPONENT[component]], nSampling, lim[component], cmin[component], cmax[component])
    label = {"u": "u", "v": "v", "Sxx": "Sxx", "Syy": "Syy", "Sxy": "Sxy", "Mises": "Mises stress"}
    for name in names:
        print("The range of " + (label[name] if component == "All" else component) + " is [%s,%s]." % (lim[name][0], lim[name][1]))
        print("Location of the minimum is [%s,%s]." % (cmin[name][0], cmin[name][1]))
        print("Location of the maximum is [%s,%s]." % (cmax[name][0], cmax[name][1]))
    with open(n_path, "w") as f:
        f.write(str(nSampling))
        f.write("\n")
        f.write(str(1) if component != "All" else str(6))
    return lim
