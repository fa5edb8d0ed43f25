#!/usr/bin/env python
"""bench.py -- the hot path of junbinhuang/MOFEM on B200: overlay-cell stiffness assembly + the linear solve.

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA, through the C-ABI)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU algorithm on the host cores

Workload (BASELINE.json configs[2]): synthetic 2-mesh patch-family overlay,
n0 = 664 -> 1.02 M cells, 1.82 M integration triangles, 1.38 M DOF, 2.0e8 COO triplets, 5.0e7 CSR non-zeros.
One *step* = one full pass of the hot path over that problem: CSR symbolic phase + fp64 integration of every
block + deterministic reduce into CSR values + constraint elimination + Jacobi-PCG to a true relative residual
of 1e-10 + energy.  ``value`` = overlay cells / step time (inputs resident in HBM); ``ms_per_step`` is the
K-assembly + solve wall time; ``e2e`` is the same through ``Context.assemble_solve`` with pinned HOST buffers
(host->device copies of the flattened problem and device->host copy of the displacement inside the timed region).
The plane sweep / generator that produces the flattened overlay is untimed preprocessing (reported in ``prep_s``).

One JSON line on stdout (rank 0); everything else goes to stderr.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

METRIC = "overlay-cell stiffness assembly cells/s (fp64) and K-assembly+solve wall time"
FAMILY_TEXT = {"quad": "2 meshes, quad4/quad4", "tri": "2 meshes, quad4/tri3", "m3s": "3 meshes, staggered third mesh, quad4: cells with 1/2/3 incident elements"}
UNIT = "cells/s"


_REAL_STDOUT = None


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.buffer.write(line); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line)


def measured_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device=0):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def algorithmic_bytes(info, n_dof):
    """Per-launch algorithmic bytes of the HBM-bound kernels (DESIGN.md 'Kernels and rooflines')."""
    nnz = info["nnz"]
    n_node = n_dof // 2
    n_pair = info["n_pair"]
    n_contrib = info["n_coo"] // 4
    return {
        # block CSR as stored: 8 B per value + one 4 B column index per 2x2 block + node_ptr + x, y + fixed mask
        "spmv": 8 * nnz + nnz + 8 * (n_node + 1) + 8 * n_dof + 8 * n_dof + n_dof,
        # SURVEY.md section 8d scalar-CSR figure for the same product (12 B per stored entry)
        "spmv_scalar_csr": 12 * nnz + 4 * (n_dof + 1) + 16 * n_dof,
        "vec": 8 * n_dof * 8,            # k_pcg_vec: reads q, r, dinv, p, x; writes r, x, p (p's second read comes out of L2)
        # numeric reduce (block values stored in reduction order): every contribution record once (32 B; the transposed copy of an
        # off-diagonal pair is its own record), per unique node pair its run pointer (8 B) and row node (4 B), CSR values out
        "numeric": 32 * n_contrib + 12 * (nnz // 4) + 8 * nnz,
        # the same reduce with round 1's layout (32 B per stored node pair + a 4 B source slot per contribution): kept for comparison
        "numeric_r1_layout": 32 * n_pair + 4 * n_contrib + 8 * nnz,
    }


FP64_PEAK_TFLOPS = 37.2     # 148 SM x 64 DFMA/clk x 2 x 1.965 GHz (ncu: 9472 DFMA thread-inst/clk peak)


def integration_flops(fp, info, ms, cell_kernel=True):
    """fp64 roofline of the integration kernels with SURVEY.md section 8d's flop model for quad4 overlays: per evaluation
    87 k + 167 (k Newton updates), per (point, block) contraction 304, regular quad4 element 4 x (95 + 304).
    ``algorithmic`` counts ONE evaluation per (point, incident element) -- the minimum, and what the cell kernel (k_cell_quads, the
    default for all-quad4 overlay cells with <= 3 elements and <= 5 straight triangles) executes; ``reference_order`` counts what
    the thread-per-block kernels (like the reference) do: every block re-evaluates its elements.  ``executed`` = the count of the
    kernels that ran (``cell_kernel`` = the library option)."""
    if not (np.asarray(fp.elem_nn) == 4).all():
        # the survey's flop model is written for quad4 elements; tri3 patches (closed-form inverse map, 3/4-point rules) only get counts
        return {"newton_updates_per_evaluation": info["newton_updates"] / max(info["n_eval"], 1), "fp64_peak_tflops": FP64_PEAK_TFLOPS,
                "note": "flop model of SURVEY.md section 8d covers quad4 cells only"}
    ce = np.asarray(fp.cell_elem)
    m = (ce >= 0).sum(1)
    tri = np.diff(np.asarray(fp.cell_tri_ptr))
    overlay = np.asarray(fp.cell_kind) == 1
    points = 6 * tri[overlay]                                   # 6-point rule for quad4 x quad4 (stiffnessMatrix.py:104-106)
    mo = m[overlay]
    evals_min = float((points * mo).sum())
    evals_done = float((points * mo * mo).sum())
    contractions = float((points * mo * (mo + 1) // 2).sum())
    kbar = info["newton_updates"] / max(info["n_eval"], 1)
    f_eval = 87.0 * kbar + 167.0
    regular = float((~overlay).sum()) * 4 * (95 + 304)
    alg = evals_min * f_eval + contractions * 304 + regular
    ref_order = evals_done * f_eval + contractions * 304 + regular
    if cell_kernel:     # cells outside the cell kernel's range (more than 3 elements / 5 triangles, curved) stay on the block kernels
        fast = (mo <= 3) & (tri[overlay] <= 5) & (tri[overlay] >= 1)
        if getattr(fp, "tri_curved", None) is not None and np.asarray(fp.tri_curved).any():
            fast &= False
        evals_done = float((points * mo * np.where(fast, 1, mo)).sum())
    done = evals_done * f_eval + contractions * 304 + regular
    return {"newton_updates_per_evaluation": kbar, "algorithmic_gflop": alg / 1e9, "executed_gflop": done / 1e9,
            "algorithmic_tflops": alg / (ms * 1e-3) / 1e12, "executed_tflops": done / (ms * 1e-3) / 1e12,
            "fp64_peak_tflops": FP64_PEAK_TFLOPS, "frac_fp64_algorithmic": alg / (ms * 1e-3) / 1e12 / FP64_PEAK_TFLOPS,
            "frac_fp64_executed": done / (ms * 1e-3) / 1e12 / FP64_PEAK_TFLOPS,
            "evaluations_model": evals_done, "reference_order_gflop": ref_order / 1e9, "cell_kernel": bool(cell_kernel)}


RUNNER = os.path.join(REPO, "baseline", "ref_runner.py")
REF_TREE = os.path.join(REPO, "baseline", "_ref", "MOFEM")


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def reference_staged():
    return os.path.isdir(REF_TREE)


def run_runner(*argv, timeout=1800):
    """baseline/ref_runner.py in a subprocess (the unmodified reference staged under baseline/_ref/MOFEM); BLAS pools pinned to
    one thread so that ``cores`` counts processes (the reference's 2x2 ... 18x18 LAPACK calls do not thread anyway)."""
    env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1")
    env.pop("RANK", None); env.pop("WORLD_SIZE", None)
    r = subprocess.run([sys.executable, RUNNER] + [str(a) for a in argv], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                       env=env, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError(f"ref_runner {argv} failed: {r.stderr[-2000:]}")
    return json.loads(r.stdout.strip().splitlines()[-1])


def write_cell_sample(fp, n_cells, path, seed=20260101):
    """A seeded random subsample of the workload's overlay cells as a compact flattened problem (only the elements and nodes the
    sampled cells touch), for the reference's element functions (ref_runner.py sample)."""
    rng = np.random.default_rng(seed)
    n_cells = min(int(n_cells), fp.n_cell)
    cells = np.sort(rng.choice(fp.n_cell, n_cells, replace=False))
    tp = np.asarray(fp.cell_tri_ptr)
    cnt = tp[cells + 1] - tp[cells]
    ptr = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    sel = np.repeat(tp[cells] - ptr[:-1], cnt) + np.arange(ptr[-1])
    ce = np.asarray(fp.cell_elem)[cells]
    elems, inv = np.unique(ce[ce >= 0], return_inverse=True)
    ce2 = np.full_like(ce, -1); ce2[ce >= 0] = inv
    en = np.asarray(fp.elem_nodes)[elems]
    nodes, ninv = np.unique(en[en >= 0], return_inverse=True)
    en2 = np.full_like(en, -1); en2[en >= 0] = ninv
    np.savez(path, meta=np.array([fp.solver, fp.n_mesh]), node_xy=np.asarray(fp.node_xy)[nodes], elem_nodes=en2,
             elem_nn=np.asarray(fp.elem_nn)[elems], elem_mat=np.asarray(fp.elem_mat)[elems], mat_D=np.asarray(fp.mat_D).reshape(-1, 3, 3),
             cell_elem=ce2, cell_kind=np.asarray(fp.cell_kind)[cells], cell_tri_ptr=ptr, tri_xy=np.asarray(fp.tri_xy)[sel],
             tri_rho=np.asarray(fp.tri_rho)[sel])
    return n_cells


def workload_text(cells, family):
    return f"syn-{max(1, round(cells / 1e6))}M patch-family overlay ({FAMILY_TEXT[family]}, jitter 0.12h): assembly + Jacobi-PCG"


def family_problem(n0, family, ny=None):
    from mofem_b200.overlay_gen import general_family, patch_family
    if family == "quad":
        return patch_family(n0, ny=ny, engine="native")
    return general_family(n0, family, ny=ny)


def port_figure(n0, cores):
    """The oracle C port on all host threads + scipy coo->csr + spsolve on a complete small member of the family (the round-1 CPU
    figure, kept as a second, labelled number)."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    from mofem_b200.overlay_gen import patch_family
    from oracle import oracle
    fp, rhs, fixed, val = patch_family(n0)
    free = fixed == 0
    t0 = time.perf_counter()
    I, J, V, _, _ = oracle.assemble_coo(fp, cores)
    t1 = time.perf_counter()
    K = sp.coo_matrix((V, (I, J)), shape=(fp.n_dof,) * 2).tocsr()
    t2 = time.perf_counter()
    b = (rhs - K @ val)[free]
    spla.spsolve(K[free][:, free].tocsc(), b)
    t3 = time.perf_counter()
    return {"kind": "port", "cores": cores, "value": fp.n_cell / (t3 - t0), "unit": UNIT,
            "sample": f"complete patch-family problem n0={n0}: {fp.n_cell} cells, {fp.n_dof} DOF; oracle C port of the reference's "
                      f"assembly on {cores} threads + scipy coo->csr + spsolve",
            "assembly_cells_per_s": fp.n_cell / (t1 - t0), "coo_to_csr_s": t2 - t1, "spsolve_s": t3 - t2}


def cpu_baseline(fp, args):
    """The CPU figures next to the GPU line (rank 0, N = 1).  Primary: the UNMODIFIED reference (baseline/_ref/MOFEM) -- its
    element functions stiffnessMatrix.lowerOrderAMORE / quadFE / ... over a seeded subsample of THIS workload's cells, following the
    cell loop of AMORE.py:59-123, one process (the reference has no parallelism); then the same on all host cores by forking over
    cells (cells are independent).  The reference's spsolve cannot factor the full system in the budget (1.38 M equations): the
    figure is assembly only, which favours the CPU side.  Secondary: the round-1 number (oracle C port + scipy + spsolve)."""
    import tempfile
    cores = host_cores()
    out = None
    if reference_staged():
        with tempfile.TemporaryDirectory() as td:
            path = os.path.join(td, "sample.npz")
            n1 = write_cell_sample(fp, args.ref_cells_1core, path)
            r1 = run_runner("sample", path, 1)
            nall = write_cell_sample(fp, args.ref_cells, path)
            rall = run_runner("sample", path, cores)
        out = {"value": r1["cells_per_s"], "unit": UNIT, "cores": 1, "kind": "reference",
               "sample": f"{n1} of the workload's {fp.n_cell} overlay cells (seeded random subsample), integrated by the unmodified reference's "
                         f"stiffnessMatrix.* + sparsifyElementMatrix in the order of AMORE.py:59-123, one Python process: {r1['wall_s']:.1f} s; "
                         "assembly only (coo->csr and spsolve of the full 1.38 M-equation system are not run: the direct solver does not fit "
                         "the time budget), so the figure favours the CPU side",
               "coo_triplets_sampled": r1["coo_triplets"],
               "all_cores": {"value": rall["cells_per_s"], "unit": UNIT, "cores": rall["processes"], "kind": "reference",
                             "sample": f"{nall} cells of the same subsample family forked over {rall['processes']} processes: {rall['wall_s']:.1f} s"}}
    if not args.no_port_figure:
        port = port_figure(args.ref_n0, cores)
        if out is None:
            out = dict(port, note="baseline/_ref/MOFEM not staged on this box: oracle port only")
        else:
            out["port"] = port
    return out


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores, on the GPU arm's config.  One step =
    a bounded sample of the workload (args.ref_cells of its overlay cells) integrated by the UNMODIFIED reference's element functions
    (baseline/_ref/MOFEM), forked over all host cores; value = sampled cells / wall time.  The linear solve is not part of the
    step: spsolve on the full 1.38 M-equation system does not fit the budget ("n/a"), which favours this arm.  Without the staged
    reference (it cannot be staged outside the build container) the oracle C port runs the same sample."""
    import tempfile
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    # N > 1: the GPU arm's problem is a stack of N blocks (weak scaling) or the one block cut into strips (--strong)
    fp, rhs, fixed, val = family_problem(args.n0, args.family, ny=None if (args.strong or args.gpus == 1) else args.n0 * args.gpus)
    walls = []
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "sample.npz")
        n = write_cell_sample(fp, args.ref_cells, path)
        if reference_staged():
            kind = "reference"
            for k in range(args.warmup + args.steps):
                r = run_runner("sample", path, cores)
                if k >= args.warmup:
                    walls.append(r["wall_s"])
            how = (f"unmodified reference (baseline/_ref/MOFEM): stiffnessMatrix.* + sparsifyElementMatrix over the cell loop of "
                   f"AMORE.py:59-123, forked over {r['processes']} processes")
        else:
            kind = "port"
            from mofem_b200.flatten import FlatProblem
            from oracle import oracle
            d = dict(np.load(path))
            sub = FlatProblem(solver=fp.solver, n_mesh=fp.n_mesh, node_xy=d["node_xy"], elem_nodes=d["elem_nodes"], elem_nn=d["elem_nn"],
                              elem_mat=d["elem_mat"], elem_mesh=np.zeros(len(d["elem_nn"]), np.int32), mat_D=d["mat_D"], cell_elem=d["cell_elem"],
                              cell_kind=d["cell_kind"], cell_tri_ptr=d["cell_tri_ptr"], tri_xy=d["tri_xy"], tri_rho=d["tri_rho"])
            for k in range(args.warmup + args.steps):
                t0 = time.perf_counter()
                oracle.assemble_coo(sub, cores)
                if k >= args.warmup:
                    walls.append(time.perf_counter() - t0)
            how = f"baseline/_ref/MOFEM not staged on this box: oracle C port of the same functions on {cores} threads"
    dt = sum(walls) / len(walls)
    v = n / dt
    sample = (f"{n} of the workload's {fp.n_cell} overlay cells (seeded random subsample) per step; {how}; assembly only -- coo->csr + "
              "spsolve of the full system: n/a (direct solver does not fit the budget), which favours this arm")
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong" if (args.strong and args.gpus > 1) else "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": workload_text(fp.n_cell / (1 if args.strong else args.gpus), args.family), "n0": args.n0, "cells": fp.n_cell, "triangles": fp.n_tri,
                      "dof": fp.n_dof, "cells_global": fp.n_cell, "sampled_cells_per_step": n},
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(out)


SHIPPED = {"bracket": ("AMOREBracket", "AMOREBracket plane-stress bracket with hole (inputFiles/AMOREBracket.txt, solver 3: quadraticICMAMORE), BASELINE configs[1]"),
           "ofe3": ("simpleOFE3", "simpleOFE3 patch test (3 overlapping meshes, inputFiles/simpleOFE3.txt, solver 2: quadraticAMORE), BASELINE configs[0]")}


def run_shipped(args):
    """--config bracket | ofe3: the reference's two shipped problems (BASELINE configs 2 / 1), 1 GPU.
    value  : cells / step, step = symbolic + integrate + numeric + constraints + PCG + energy on the flattened problem resident in HBM
             (the committed fixture of the reference's own flattened overlay, tests/golden/<stem>.npz), CUDA events;
    e2e    : cells / wall time of the reference-facing ENTRY POINT (mofem_b200.solvers.quadraticICMAMORE / quadraticAMORE) called with
             the reference's own inputData / overlappingMesh / polygons from its parser and plane sweep -- Python flattening of the
             object graph, host->device copies, kernels, device->host copy, force vector and constraint glue all inside;
    cpu_baseline : the UNMODIFIED reference's entry point on the same objects, timed in full, one thread, with its four phase times
                   (AMORE.py:829,1052,1099,1112)."""
    import torch
    from mofem_b200.device import Context
    from mofem_b200.flatten import FlatProblem
    stem, text = SHIPPED[args.config]
    d = dict(np.load(os.path.join(REPO, "tests", "golden", stem + ".npz"), allow_pickle=True))
    fp = FlatProblem.from_npz_dict(d)
    fixed = np.zeros(fp.n_dof, np.uint8); val = np.zeros(fp.n_dof)
    for c in d["constraints"]:
        nd = int(c[0])
        if c[1]: fixed[2 * nd] = 1; val[2 * nd] = c[3]
        if c[2]: fixed[2 * nd + 1] = 1; val[2 * nd + 1] = c[4]
    free = fixed == 0
    rhs = np.zeros(fp.n_dof); rhs[free] = d["red_rhs"].reshape(-1)       # fixed values are zero in both decks: f_free = reduced rhs
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (mofem_b200 has no CPU fallback)")
    dev = torch.device("cuda", 0)
    ctx = Context(0)
    ctx.set_profile(0 if args.no_profile else 8)      # events around every 8th iteration (sampled iterations lose their launch overlap)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    keys = [k for k in FlatProblem.ARRAYS if getattr(fp, k) is not None and k != "elem_mesh"]
    import copy
    fp_dev = copy.copy(fp)
    for k in keys:
        setattr(fp_dev, k, torch.from_numpy(np.ascontiguousarray(getattr(fp, k))).to(dev))
    rhs_d, fixed_d, val_d = (torch.from_numpy(a).to(dev) for a in (rhs, fixed, val))
    u_d = torch.empty(fp.n_dof, dtype=torch.float64, device=dev)
    last = {}

    def step():
        ctx.set_problem(fp_dev).symbolic().assemble()
        ctx.set_system(rhs_d, fixed_d, val_d)
        _, last["solve"] = ctx.solve(rtol=1e-12, maxit=200000, out=u_d)
        last["energy"] = ctx.energy()
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(0); sampler.start()
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    phases = {"symbolic_ms": 0.0, "integrate_ms": 0.0, "numeric_ms": 0.0, "solve_ms": 0.0}
    prof = {"spmv_ms": 0.0, "vec_ms": 0.0, "iterations": 0}
    launches = 0
    steps = max(args.steps, 20)          # a step is tens of milliseconds: enough of them for the clock sampler to see load
    ev0.record(stream)
    for _ in range(steps):
        step()
        tm = ctx.timing()
        for k in phases:
            phases[k] += tm[k]
        launches += sum(tm["launches"].values()) + 2
        pp = ctx.pcg_profile()
        for k in prof:
            prof[k] += pp[k]
    ev1.record(stream)
    torch.cuda.synchronize()
    clocks = sampler.stop()
    step_ms = ev0.elapsed_time(ev1) / steps
    info = ctx.info()
    si = last["solve"]
    u_fix = u_d.cpu().numpy()
    n_cells = int((np.asarray(fp.cell_kind) >= 0).sum()) - fp.n_mesh_regular
    n_regular = fp.n_mesh_regular
    ctx.close()

    parity = {"fixture_max_rel_du": float(np.abs(u_fix - d["displacement"].reshape(-1)).max() / np.abs(d["displacement"]).max()),
              "fixture_energy_rel": float(abs(last["energy"] - float(d["energy"].reshape(-1)[0])) / abs(float(d["energy"].reshape(-1)[0])))}
    e2e = cpu = None
    if reference_staged():
        import tempfile
        with tempfile.TemporaryDirectory() as td:
            rep = 5
            g = run_runner("entry", stem, "b200", os.path.join(td, "gpu.npz"), rep)
            r = run_runner("entry", stem, "ref", os.path.join(td, "ref.npz"), 1)
            ug, ur = np.load(os.path.join(td, "gpu.npz")), np.load(os.path.join(td, "ref.npz"))
        runs = g["runs"][2:]                 # the first calls create the context and warm the memory pool
        entry_s = sum(x["entry_s"] for x in runs) / len(runs)
        h2d = sum(np.asarray(getattr(fp, k)).nbytes for k in keys) + rhs.nbytes + fixed.nbytes + val.nbytes
        e2e = {"value": n_cells / entry_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(8 * fp.n_dof + 8),
               "ms_per_step": entry_s * 1e3,
               "what": f"mofem_b200.solvers entry point on the reference's own objects (parser + plane sweep of baseline/_ref/MOFEM, untimed: "
                       f"{g['overlay_s']:.2f} s); mean of {len(runs)} calls after 2 warm-ups",
               "phases_s": {k: sum(x.get(k, 0.0) for x in runs) / len(runs) for k in ("assembly_s", "force_s", "constraints_s", "solve_s")},
               "device_ms": {k: sum(x["device_ms"][k] for x in runs) / len(runs) for k in runs[0]["device_ms"]}}
        rr = r["runs"][0]
        cpu = {"value": r["cells"] / rr["entry_s"], "unit": UNIT, "cores": 1, "kind": "reference",
               "sample": f"the whole problem, in full: unmodified reference entry point ({'AMORE.quadraticICMAMORE' if stem == 'AMOREBracket' else 'AMORE.quadraticAMORE'}) "
                         f"on the same objects, one thread: {rr['entry_s']:.2f} s",
               "phases_s": {k: rr.get(k) for k in ("assembly_s", "force_s", "constraints_s", "solve_s")},
               "preprocessing_s": {"parser": r["input_s"], "plane_sweep_rho_triangulation": r["overlay_s"]}}
        du = np.abs(ug["displacement"] - ur["displacement"]).max() / np.abs(ur["displacement"]).max()
        parity.update({"entry_max_rel_du_vs_reference": float(du),
                       "entry_energy_rel_vs_reference": float(abs(ug["energy"][0] - ur["energy"][0]) / abs(ur["energy"][0])),
                       "tolerance": 1e-9})
        assert du <= 1e-9, f"drop-in displacement differs from the live reference: {du}"
    else:
        log("baseline/_ref/MOFEM not staged: no entry-point e2e / reference timing on this box")
        h_t0 = time.perf_counter()
        with Context(0) as c2:
            for _ in range(3):
                c2.assemble_solve(fp, rhs, fixed, val, rtol=1e-12, maxit=200000)
            h_t0 = time.perf_counter()
            for _ in range(5):
                c2.assemble_solve(fp, rhs, fixed, val, rtol=1e-12, maxit=200000)
            es = (time.perf_counter() - h_t0) / 5
        e2e = {"value": n_cells / es, "unit": UNIT, "h2d_bytes_per_step": int(sum(np.asarray(getattr(fp, k)).nbytes for k in keys)),
               "d2h_bytes_per_step": int(8 * fp.n_dof + 8), "ms_per_step": es * 1e3,
               "what": "Context.assemble_solve with host buffers (reference not staged on this box: entry-point timing unavailable)"}

    peaks, peak_src = measured_peaks()
    ab = algorithmic_bytes(info, fp.n_dof)
    roofline = None
    if prof["iterations"]:
        ms = prof["spmv_ms"] / prof["iterations"]
        ach = ab["spmv"] / (ms * 1e-3) / 1e9
        roofline = {"kernel": "block-CSR SpMV of the Jacobi-PCG loop", "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"],
                    "peak_source": peak_src + " copy bandwidth (MEASURED_PEAKS.json hbm_gbs)", "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
                    "traffic": None, "algorithmic_bytes_per_launch": ab["spmv"], "avg_launch_ms": ms,
                    "note": f"a {info['nnz']}-entry matrix ({ab['spmv'] / 1e6:.2f} MB) lives in L2 and a launch is latency-bound: the HBM "
                            "roofline is not the limiter at this size; it is for the syn-1M workload (default config)"}
    out = {"metric": METRIC, "value": n_cells / (step_ms * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": steps, "warmup": max(args.warmup, 3),
           "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
           "data": "the reference's shipped input deck",
           "config": {"workload": text, "cells": n_cells, "regular_elements": n_regular, "triangles": fp.n_tri, "dof": fp.n_dof,
                      "coo_triplets": info["n_coo"], "nnz": info["nnz"], "rtol_true_residual": 1e-12,
                      "l2": "problem smaller than L2 (the workload is what the reference ships); launch-latency bound", "parallelism": "1 GPU"},
           "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
           "phases_ms_per_step": {k: v / steps for k, v in phases.items()},
           "pcg": {"iterations": si["iterations"], "converged": si["converged"], "relres_true": si["relres_true"], "restarts": si["restarts"],
                   "iterations_per_s": si["iterations"] / (phases["solve_ms"] / steps * 1e-3)},
           "energy": last["energy"], "parity_check": parity}
    emit(out)


def rank_problem(n0, family, strong, rank, world, host_overlay=False):
    """What rank ``rank`` of ``world`` holds of the N-GPU problem: its strip of the overlay (host build of the overlay core, before
    CUDA / NCCL are up), partitioned by node owner.  Weak scaling: a stack of N blocks of n0 x n0 mesh-0 elements, one corner load
    per block; strong: ONE n0 x n0 problem cut into N strips.  Returns (LocalPart, rhs, fixed, fixed_value in local numbering,
    number of cells of the global problem this rank accounts for)."""
    from mofem_b200.overlay_gen import patch_family
    from mofem_b200.partition import partition_problem
    px = n0 // 8                                  # patches per patch row
    if strong:
        if px % world:
            raise SystemExit("--strong needs n0 / 8 to be a multiple of the number of GPUs")
        ny, per = n0, px // world                 # one n0 x n0 problem, `per` patch rows (8 * per element rows) per rank
    else:
        ny, per = n0 * world, px                  # a stack of N blocks of n0 x n0
    rows_per_rank = 8 * per
    n_node0 = (n0 + 1) * (ny + 1)
    if family == "quad":
        strip, rhs_g, fixed_g, val_g = patch_family(n0, ny=ny, patch_rows=(rank * per, (rank + 1) * per),
                                                    engine="numpy" if host_overlay else "native")
        owner = np.empty(strip.n_node, np.int32)
        owner[:n_node0] = np.minimum((np.arange(n_node0) // (n0 + 1)) // rows_per_rank, world - 1)
        owner[n_node0:] = ((np.arange(strip.n_node - n_node0) // 36) // px) // per
        # cells of the global problem = owned-strip cells, not counting the redundantly integrated boundary row
        own_cells = strip.n_cell - (n0 if rank > 0 else 0)      # (one mesh-0 element row per strip boundary)
    else:
        # rank r owns the nodes whose nominal row lies in its element rows; it integrates every cell that touches one
        # of them (the generator emits a margin of 3 element rows, the partitioner keeps what is needed)
        from mofem_b200.overlay_gen import general_strip
        strip, rhs_g, fixed_g, val_g, owner, own_cells = general_strip(n0, family, ny, rank, world)
    # weak scaling keeps the physics per block fixed: every block carries the corner load of the 1-GPU problem on
    # its own bottom-right node (a single load on a domain N times taller would need ~N times more CG iterations
    # just to propagate across the height); strong scaling solves the one problem as it is
    stack_loads(rhs_g, n0, 1 if strong else world)
    part = partition_problem(strip, owner, rank, world)
    return part, part.local_vector(rhs_g), part.local_vector(fixed_g), part.local_vector(val_g), own_cells


def scaling_decomposition(args, world, iterations, it_per_s):
    """N > 1, default weak-scaling workload: the two causes of value(N) / (N value(1)) < 1 kept apart.  The 1-GPU side comes
    from the committed 1-GPU line of the same block (profiles/r2_bench_1gpu.json), it is NOT measured in this run."""
    ref = os.path.join(REPO, "profiles", "r2_bench_1gpu.json")
    if world == 1 or args.strong or args.family != "quad" or args.n0 != 664 or not os.path.exists(ref):
        return None
    with open(ref) as f:
        one = json.loads(f.read().strip().splitlines()[-1])["pcg"]
    return {"pcg_iterations": iterations, "pcg_iterations_1gpu": one["iterations"],
            "iteration_growth": iterations / one["iterations"],
            "iterations_per_s": it_per_s, "iterations_per_s_1gpu": one["iterations_per_s"],
            "per_iteration_efficiency": it_per_s / one["iterations_per_s"],
            "note": "weak scaling stacks N blocks of n0 x n0: Jacobi-PCG needs more iterations on the taller domain (algorithmic); the "
                    "per-iteration rate is what the transport costs.  1-GPU figures: profiles/r2_bench_1gpu.json (committed, not measured here)"}


def stack_loads(rhs_g, n0, blocks):
    for r in range(1, blocks):
        nd = r * n0 * (n0 + 1) + n0
        rhs_g[2 * nd] = 1.0; rhs_g[2 * nd + 1] = -0.5


def global_problem(n0, family, strong, world):
    """The same N-GPU problem as ONE flattened problem (for the single-GPU side of the parity check)."""
    from mofem_b200.overlay_gen import general_family, patch_family
    ny = n0 if strong else n0 * world
    fp, rhs, fixed, val = patch_family(n0, ny=ny, engine="native") if family == "quad" else general_family(n0, family, ny=ny)
    stack_loads(rhs, n0, 1 if strong else world)
    return fp, rhs, fixed, val


def multi_gpu_parity(args, rank, world, local, transport):
    """Untimed, after the timed steps (N > 1): a reduced problem of the same family and layout (n0 = 64) solved on the N ranks with the
    run's transport and on rank 0 alone; the gathered displacement and the energy must agree to 1e-9."""
    import torch.distributed as dist
    from mofem_b200.device import Context
    from mofem_b200.distributed import assemble_solve_partitioned, gather_global, init_comm
    from mofem_b200.partition import exchange_send_lists
    n0 = 64
    part, rhs, fixed, val, _ = rank_problem(n0, args.family, args.strong, rank, world)
    exchange_send_lists(part)
    with Context(local) as ctx:
        init_comm(ctx)
        u_loc, info, energy = assemble_solve_partitioned(ctx, part, rhs, fixed, val, rtol=1e-12, transport=transport)
        if transport == "peer":
            dist.barrier()
            ctx.p2p_close()
    n_node = int(part.local2global.max()) + 1
    t = [None] * world
    dist.all_gather_object(t, n_node)
    n_node = max(t)
    u = gather_global(part, u_loc, n_node)
    if rank != 0:
        return None
    fp, rhs_g, fixed_g, val_g = global_problem(n0, args.family, args.strong, world)
    assert fp.n_node == n_node
    with Context(local) as ctx:
        ctx.set_problem(fp).symbolic().assemble()
        ctx.set_system(rhs_g, fixed_g, val_g)
        u1, info1 = ctx.solve(rtol=1e-12)
        e1 = ctx.energy()
    du = float(np.abs(u - u1).max() / np.abs(u1).max())
    de = float(abs(energy - e1) / abs(e1))
    return {"problem": f"{args.family} family, n0 = {n0}, {'one block cut into' if args.strong else 'a stack of'} {world} {'strips' if args.strong else 'blocks'}: "
                       f"{fp.n_cell} cells, {fp.n_dof} DOF; {world} ranks ({transport}) against rank 0 alone, rtol 1e-12",
            "max_rel_du": du, "energy_rel": de, "n_ranks": world, "tolerance": 1e-9, "passed": bool(du <= 1e-9 and de <= 1e-9),
            "iterations": {"ranks": info["iterations"], "single": info1["iterations"]},
            "relres_true": {"ranks": info["relres_true"], "single": info1["relres_true"]}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n0", type=int, default=664, help="mesh-0 elements per side (664 -> ~1M overlay cells)")
    ap.add_argument("--rtol", type=float, default=1e-10)
    ap.add_argument("--ref-n0", type=int, default=256, help="size of the bounded CPU sample")
    ap.add_argument("--ref-cells", type=int, default=20000, help="cells of the workload the unmodified reference integrates per sample (all cores)")
    ap.add_argument("--ref-cells-1core", type=int, default=4000, help="cells of the single-process reference sample (cpu_baseline, cores = 1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-port-figure", action="store_true", help="skip the secondary CPU figure (oracle C port + scipy + spsolve at --ref-n0)")
    ap.add_argument("--config", default="syn", choices=["syn", "bracket", "ofe3"],
                    help="syn = the synthetic workload (default; BASELINE configs 3-5 by --n0 / --family / --gpus); bracket / ofe3 = the "
                         "reference's shipped problems through its own parser and plane sweep, solver entry point swapped for the drop-in "
                         "(BASELINE configs 2 / 1), timed against the unmodified reference in full")
    ap.add_argument("--no-parity-check", action="store_true", help="N > 1: skip the untimed N-ranks-against-one parity solve after the timed steps")
    ap.add_argument("--no-profile", action="store_true", help="do not record per-kernel events inside the PCG loop")
    ap.add_argument("--profile-every", type=int, default=16,
                    help="record CUDA events around the PCG kernels of every N-th iteration of the timed steps (1 = all; the "
                         "sampled iterations lose their launch overlap, so N = 1 costs ~10 %% of the step)")
    ap.add_argument("--l2-persist", type=int, default=0, help="L2 persistence window over the PCG vectors (library default: 0)")
    ap.add_argument("--transport", default="peer", choices=["peer", "nccl"],
                    help="N > 1: halo exchange / dot-product reduction over NVLink peer memory inside the kernels, or NCCL calls")
    ap.add_argument("--family", default="quad", choices=["quad", "tri", "m3s"],
                    help="synthetic overlay family: quad = 2 meshes, quad4 patches (BASELINE config 3 variant a, the headline workload); "
                         "tri = patch quads split into tri3 (variant b); m3s = staggered third mesh, cells with 1/2/3 incident "
                         "elements (config 4: --n0 1608 gives ~10 M cells)")
    ap.add_argument("--host-overlay", action="store_true",
                    help="build the overlay with the NumPy generator instead of the CUDA kernels (1 GPU) / their host build (N > 1 strips); "
                         "same bits, untimed either way")
    ap.add_argument("--strong", action="store_true",
                    help="N > 1: ONE n0 x n0 problem cut into N strips (BASELINE configs 4 and 5) instead of a stack of N blocks "
                         "of n0 x n0 (weak scaling, the default); n0 / 8 must be a multiple of N")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: anything else a library prints to fd 1 (e.g. NCCL's version banner) goes
    # to stderr for the rest of the run
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)
    if args.config != "syn":
        return run_shipped(args)

    import torch
    import torch.distributed as dist
    from mofem_b200.device import Context
    from mofem_b200.flatten import FlatProblem
    from mofem_b200.overlay_gen import patch_family

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (mofem_b200 has no CPU fallback)")

    # ---- untimed preprocessing: the flattened overlay (what the plane sweep + flatten() hand to the solver) ------
    # N > 1 (weak scaling): the domain is a stack of N blocks of n0 x n0 mesh-0 elements; rank r owns block r (its
    # mesh-0 node rows and the patches inside) and generates only the cells that touch its nodes.
    # Strips come from the overlay core built for the host (liboverlay_host.so: the source of the CUDA overlay kernels, same
    # bits as the NumPy generator, no forked workers), before CUDA / NCCL are initialised; --host-overlay: NumPy generator.
    t0 = time.perf_counter()
    part = None
    n_cells_global = None
    overlay_ms = None
    n_meshes = 3 if args.family == "m3s" else 2
    if world == 1 and args.family != "quad":
        # general families: clipping per mesh-0 element on the host (csrc/overlay_general.h), pinned against the reference's sweep
        from mofem_b200.overlay_gen import general_family
        fp, rhs, fixed, val = general_family(args.n0, args.family)
    elif world == 1 and not args.host_overlay:
        # overlay by pairwise convex clipping on the GPU (csrc/overlay.cu; bit-identical to the NumPy generator)
        from mofem_b200.overlay_gen import patch_family_gpu
        torch.cuda.set_device(local)
        with Context(local) as gen_ctx:
            fp, rhs, fixed, val, overlay_ms = patch_family_gpu(gen_ctx, args.n0, to_host=True)
    elif world == 1:
        fp, rhs, fixed, val = patch_family(args.n0)
    else:
        from mofem_b200.distributed import init_comm
        from mofem_b200.partition import exchange_send_lists
        part, rhs, fixed, val, own_cells = rank_problem(args.n0, args.family, args.strong, rank, world, args.host_overlay)
        fp = part.fp

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        exchange_send_lists(part)
        t = torch.tensor([own_cells], dtype=torch.int64, device=dev)
        dist.all_reduce(t)
        n_cells_global = int(t.item())
    prep_s = time.perf_counter() - t0
    log(f"[rank {rank}] generated n0={args.n0}: {fp.n_cell} cells, {fp.n_tri} triangles, {fp.n_dof} local DOF in {prep_s:.1f}s"
        + (f" (overlay kernels {overlay_ms:.2f} ms)" if overlay_ms is not None else ""))

    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.from_numpy(a[:0].copy()).dtype, pin_memory=True)
        t.numpy()[...] = a
        return t
    host = {k: pinned(getattr(fp, k)) for k in FlatProblem.ARRAYS if getattr(fp, k) is not None and k != "elem_mesh"}
    host_sys = {"rhs": pinned(rhs), "fixed": pinned(fixed), "val": pinned(val)}
    h2d_bytes = sum(t.numel() * t.element_size() for t in host.values()) + \
        sum(t.numel() * t.element_size() for t in host_sys.values())
    d2h_bytes = 8 * fp.n_dof + 8

    import copy
    fp_host = copy.copy(fp)
    for k, t in host.items():
        setattr(fp_host, k, t.numpy())
    fp_dev = copy.copy(fp)
    for k, t in host.items():
        setattr(fp_dev, k, t.to(dev))
    rhs_d, fixed_d, val_d = (host_sys[k].to(dev) for k in ("rhs", "fixed", "val"))
    u_d = torch.empty(fp.n_dof, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()

    ctx = Context(local)
    if world > 1:
        init_comm(ctx)
    ctx.set_profile(0 if args.no_profile else max(1, args.profile_every))
    ctx.set_option("l2_persist", args.l2_persist)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    maxit = 400000
    last = {}

    transport = "none"
    if part is not None:
        from mofem_b200.distributed import setup_peer_transport
        ctx.set_halo(part)           # the halo plan (and the peer transport) persist across problems of the same shape
        transport = "nccl"
        if args.transport == "peer" and setup_peer_transport(ctx, part):
            transport = "peer"
        log(f"[rank {rank}] transport: {transport}")

    def step_resident():
        ctx.set_problem(fp_dev).symbolic().assemble()
        ctx.set_system(rhs_d, fixed_d, val_d)
        _, si = ctx.solve(rtol=args.rtol, maxit=maxit, out=u_d)
        last["solve"] = si
        last["energy"] = ctx.energy()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    phases = {"symbolic_ms": 0.0, "integrate_ms": 0.0, "numeric_ms": 0.0, "solve_ms": 0.0}
    prof = {"spmv_ms": 0.0, "vec_ms": 0.0, "iterations": 0}
    launches = 0
    ev0.record(stream)
    for _ in range(args.steps):
        step_resident()
        tm = ctx.timing()
        for k in phases:
            phases[k] += tm[k]
        launches += sum(tm["launches"].values()) + 2    # + energy SpMV and its reduction
        pp = ctx.pcg_profile()
        for k in prof:
            prof[k] += pp[k]
    ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    if prof["iterations"]:
        log(f"[rank {rank}] per-iteration us: spmv {prof['spmv_ms'] / prof['iterations'] * 1e3:.1f} "
            f"vector kernel {prof['vec_ms'] / prof['iterations'] * 1e3:.1f}")
    step_ms = ev0.elapsed_time(ev1) / args.steps
    if world > 1:
        t = torch.tensor([step_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        step_ms = float(t.item())
    si = last["solve"]

    # ---- end to end through the public API with pinned host buffers -----------------------------------------------
    e2e_steps = max(1, min(args.steps, 3))
    e2e_warm = 1 if args.warmup > 0 else 0       # one untimed pass of this path (its staging buffers come out of the pool afterwards)
    for k_e2e in range(e2e_warm + e2e_steps):
        if k_e2e == e2e_warm:
            barrier()
            t0 = time.perf_counter()
        if part is None:
            u_h, energy_h, si_h = ctx.assemble_solve(fp_host, host_sys["rhs"].numpy(), host_sys["fixed"].numpy(),
                                                     host_sys["val"].numpy(), rtol=args.rtol, maxit=maxit)
        else:   # same calls, host buffers in / host displacement out, on every rank
            ctx.set_problem(fp_host).symbolic().assemble()
            ctx.set_system(host_sys["rhs"].numpy(), host_sys["fixed"].numpy(), host_sys["val"].numpy())
            u_h, si_h = ctx.solve(rtol=args.rtol, maxit=maxit)
            energy_h = ctx.energy()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    assert np.array_equal(u_h, u_d.cpu().numpy()), "host-buffer path and resident path must give identical bits"

    info = ctx.info()
    parity = None
    if world > 1 and not args.no_parity_check:
        if transport == "peer":
            barrier()
            ctx.p2p_close()
        ctx.clear_halo()
        parity = multi_gpu_parity(args, rank, world, local, transport)
        barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_src = measured_peaks()
    ab = algorithmic_bytes(info, fp.n_dof)
    iters = max(prof["iterations"], 1)
    spmv_ms = prof["spmv_ms"] / iters if prof["iterations"] else None
    roofline = None
    traffic = None    # dram__bytes_read + dram__bytes_write per SpMV launch, from the committed ncu capture of this workload
    tpath = os.path.join(REPO, "profiles", "r2_spmv_traffic.json")
    if world == 1 and args.n0 == 664 and args.family == "quad" and os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f)["dram_bytes_per_launch"]
    if spmv_ms:
        ach = ab["spmv"] / (spmv_ms * 1e-3) / 1e9
        roofline = {"kernel": ("k_spmv_ring (block-CSR SpMV of the Jacobi-PCG loop, matrix streamed through shared memory by TMA)" if world == 1
                               else "k_spmv_ring_p2p (same ring, halo exchange over NVLink peer memory inside the kernel)" if transport == "peer"
                               else "k_spmv (block-CSR SpMV inside Jacobi-PCG, row kernel) + NCCL halo exchange"),
                    "bound": "hbm", "achieved": ach,
                    "peak": peaks["hbm_gbs"], "peak_source": peak_src + " copy bandwidth (MEASURED_PEAKS.json hbm_gbs)",
                    "unit": "GB/s", "frac": ach / peaks["hbm_gbs"], "traffic": traffic,
                    "traffic_source": (f"committed ncu --set full capture of this kernel on this workload ({os.path.relpath(tpath, REPO)}: dram__bytes_read.sum + "
                                       "dram__bytes_write.sum of one launch), not measured in this run") if traffic is not None else None,
                    "algorithmic_bytes_per_launch": ab["spmv"], "avg_launch_ms": spmv_ms,
                    "scalar_csr_equivalent_gbs": ab["spmv_scalar_csr"] / (spmv_ms * 1e-3) / 1e9,
                    "share_of_step": spmv_ms * si["iterations"] / step_ms,
                    "launches_timed": prof["iterations"], "sampled_every": (0 if args.no_profile else max(1, args.profile_every))}
    kernels = {}
    if prof["iterations"]:
        ms = prof["vec_ms"] / iters
        kernels["k_pcg_vec"] = {"avg_launch_ms": ms, "gbs": ab["vec"] / (ms * 1e-3) / 1e9,
                                "frac_hbm": ab["vec"] / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"]}
    nm = phases["numeric_ms"] / args.steps
    kernels["k_numeric"] = {"avg_launch_ms": nm, "gbs": ab["numeric"] / (nm * 1e-3) / 1e9,
                            "frac_hbm": ab["numeric"] / (nm * 1e-3) / 1e9 / peaks["hbm_gbs"],
                            "algorithmic_bytes": ab["numeric"],
                            "frac_hbm_r1_layout_bytes": ab["numeric_r1_layout"] / (nm * 1e-3) / 1e9 / peaks["hbm_gbs"]}
    im = phases["integrate_ms"] / args.steps
    kernels["integrate (all classes)"] = {"ms": im, "cells_per_s": fp.n_cell / (im * 1e-3),
                                          "newton_updates": info["newton_updates"], "evaluations": info["n_eval"]}
    kernels["integrate (all classes)"].update(integration_flops(fp, info, im))
    if clocks and clocks.get("sm_mhz") and clocks.get("sm_max_mhz") and "frac_fp64_algorithmic" in kernels["integrate (all classes)"]:
        # the assembly of a step starts right behind ~0.6 s of PCG at the power cap, so it runs at the capped SM clock: the same
        # fraction against the fp64 peak AT THAT CLOCK (median SM clock sampled during the timed region) next to the nominal one
        k = kernels["integrate (all classes)"]
        k["sm_mhz_during_run"] = clocks["sm_mhz"]
        k["frac_fp64_algorithmic_at_run_clock"] = k["frac_fp64_algorithmic"] * 1965.0 / clocks["sm_mhz"]

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        log("timing the CPU baseline (unmodified reference on a subsample of the workload; oracle port + scipy) ...")
        cpu = cpu_baseline(fp, args)

    cells_total = fp.n_cell if world == 1 else n_cells_global
    out = {
        "metric": METRIC, "value": cells_total / (step_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong" if (args.strong and world > 1) else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"syn-{max(1, round(cells_total / (1 if args.strong else world) / 1e6))}M patch-family overlay ({FAMILY_TEXT[args.family]}, jitter 0.12h): assembly + Jacobi-PCG",
                   "n0": args.n0, "cells": cells_total, "triangles": fp.n_tri, "dof": fp.n_dof, "coo_triplets": info["n_coo"], "nnz": info["nnz"],
                   "per_rank_fields": None if world == 1 else "triangles, dof, coo_triplets, nnz are rank 0's share (cells is the global count)",
                   "rtol_true_residual": args.rtol,
                   "l2": f"inputs larger than L2 (per GPU: CSR {9 * info['nnz'] / 1e9:.2f} GB, block values {32 * info['n_pair'] / 1e9:.2f} GB vs 126 MB L2)",
                   "parallelism": "1 GPU" if world == 1 else
                   f"{world} GPUs: cells sharded by strip (no assembly exchange), row-partitioned PCG, halo "
                   f"exchange + all-reduce over {transport}; global problem "
                   + (f"{args.n0} x {args.n0} mesh-0 elements cut into {world} strips, one corner load" if args.strong else
                      f"{args.n0} x {args.n0 * world} mesh-0 elements, one corner load per block"),
                   "cells_global": cells_total},
        "e2e": {"value": cells_total / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "ms_per_step": e2e_s * 1e3},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "phases_ms_per_step": {k: v / args.steps for k, v in phases.items()},
        "assembly_cells_per_s": cells_total / ((phases["symbolic_ms"] + phases["integrate_ms"] + phases["numeric_ms"]) / args.steps * 1e-3),
        "pcg": {"iterations": si["iterations"], "converged": si["converged"], "relres_true": si["relres_true"],
                "restarts": si["restarts"], "iterations_per_s": si["iterations"] / (phases["solve_ms"] / args.steps * 1e-3)},
        "kernels": kernels,
        "scaling_decomposition": scaling_decomposition(args, world, si["iterations"], si["iterations"] / (phases["solve_ms"] / args.steps * 1e-3)),
        "energy": last["energy"],
        "parity_check": parity,
        "prep_s": prep_s,
        "overlay": ({"where": "gpu (csrc/overlay.cu, pairwise convex clipping)", "kernels_ms": overlay_ms} if overlay_ms is not None
                    else {"where": "host, NumPy generator (mofem_b200/overlay_gen.py)" if args.host_overlay or world == 1
                          else "host build of the overlay core (liboverlay_host.so), per-rank strips"}),
    }
    emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
