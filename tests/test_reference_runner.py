"""baseline/ref_runner.py (the CPU arm of bench.py): the reference's own element functions over a flattened cell sample give the
oracle's COO triplets -- i.e. the array-level loop in the runner follows AMORE.py:59-123 and the oracle is pinned against the very
code the bench times.  Needs the staged reference (build container; python baseline/stage_ref.py)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import REPO

TREE = os.path.join(REPO, "baseline", "_ref", "MOFEM")
pytestmark = pytest.mark.skipif(not os.path.isdir(TREE), reason="baseline/_ref/MOFEM is staged in the build container only")


@pytest.mark.parametrize("family,n_proc", [("quad", 1), ("tri", 2), ("m3s", 2)])
def test_sample_matches_the_oracle(tmp_path, family, n_proc):
    sys.path.insert(0, REPO)
    import bench
    from mofem_b200.flatten import FlatProblem
    from oracle import oracle
    fp, *_ = bench.family_problem(16, family)
    path = str(tmp_path / "s.npz")
    n = bench.write_cell_sample(fp, 120, path)
    r = bench.run_runner("sample", path, n_proc)
    d = dict(np.load(path))
    sub = FlatProblem(solver=fp.solver, n_mesh=fp.n_mesh, node_xy=d["node_xy"], elem_nodes=d["elem_nodes"], elem_nn=d["elem_nn"],
                      elem_mat=d["elem_mat"], elem_mesh=np.zeros(len(d["elem_nn"]), np.int32), mat_D=d["mat_D"], cell_elem=d["cell_elem"],
                      cell_kind=d["cell_kind"], cell_tri_ptr=d["cell_tri_ptr"], tri_xy=d["tri_xy"], tri_rho=d["tri_rho"])
    I, J, V, _, _ = oracle.assemble_coo(sub, 2)
    assert r["cells"] == n == 120 and r["coo_triplets"] == len(V)
    assert abs(r["sum_values"] - V.sum()) <= 1e-9 * np.abs(V).sum()


def test_reference_arm_line(tmp_path):
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--n0", "16",
                        "--ref-cells", "60"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "reference" and line["unit"] == "cells/s"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["config"]["sampled_cells_per_step"] == 60
