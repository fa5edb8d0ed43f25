"""Timing of the pre-/post-processing kernels at the bench size, next to the oracle on the host:

    python tools/bench_aux.py [n0=664] > profiles/r1_aux_kernels.json

  rho      k_rho on the overlay vertices of a jittered n x n quad mesh pair (2 meshes, every vertex inside both): CUDA-event
           kernel time from mofem_rho_eval, resident inputs; oracle (C port, one thread) on a 200 k-vertex sample
  sampler  k_sample on the syn-1M solution: 3 x 3 points per regular element / integration triangle, device output
"""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from mofem_b200 import _lib
from mofem_b200.device import Context
from mofem_b200.flatten import FlatProblem
from mofem_b200.overlay_gen import patch_family
from oracle import oracle

n0 = int(sys.argv[1]) if len(sys.argv) > 1 else 664
dev = torch.device("cuda", 0)
ctx = Context(0)
out = {}

# ---- rho ------------------------------------------------------------------------------------------------------
rng = np.random.default_rng(5)
n = 1024
gx, gy = np.meshgrid(np.arange(n + 1.0), np.arange(n + 1.0), indexing="xy")
node = np.stack([gx, gy], -1)
node[1:-1, 1:-1] += rng.uniform(-0.12, 0.12, (n - 1, n - 1, 2))
ex, ey = np.meshgrid(np.arange(n), np.arange(n), indexing="xy")
corners = np.stack([node[ey, ex], node[ey, ex + 1], node[ey + 1, ex + 1], node[ey + 1, ex]], 2).reshape(-1, 4, 2)
wnode = np.ones((n + 1, n + 1)); wnode[0] = wnode[-1] = 0; wnode[:, 0] = wnode[:, -1] = 0
w1 = np.stack([wnode[ey, ex], wnode[ey, ex + 1], wnode[ey + 1, ex + 1], wnode[ey + 1, ex]], 2).reshape(-1, 4)
E = n * n
elem_xy = np.concatenate([corners, corners]); elem_w = np.concatenate([np.ones((E, 4)), w1])
V = 2 * E
e = np.sort(rng.integers(0, E, V))
t = rng.uniform(0.05, 0.95, (V, 2))
c = corners[e]
xy = ((1 - t[:, :1]) * (1 - t[:, 1:]) * c[:, 0] + t[:, :1] * (1 - t[:, 1:]) * c[:, 1] + t[:, :1] * t[:, 1:] * c[:, 2]
      + (1 - t[:, :1]) * t[:, 1:] * c[:, 3])
ve = np.stack([e, e + E], 1).astype(np.int32)
nv = np.full(2 * E, 4, np.int32)
args = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (xy, ve, nv, elem_xy, elem_w)]
rho_d = torch.empty((V, 2), dtype=torch.float64, device=dev)
ms = []
for _ in range(6):
    _, nw, k_ms = ctx.rho_eval(*args, out=rho_d)
    ms.append(k_ms)
k_ms = float(np.median(ms[2:]))
sub = np.arange(0, V, V // 200000)[:200000]
t0 = time.perf_counter()
want, _ = oracle.rho_eval(xy[sub], ve[sub], nv, elem_xy, elem_w)
cpu_s = time.perf_counter() - t0
assert np.array_equal(rho_d.cpu().numpy()[sub], want)
alg_bytes = V * (16 + 8 + 16) + 2 * E * (64 + 32 + 4)
out["rho"] = {"vertices": V, "meshes": 2, "newton_updates": nw, "kernel_ms": k_ms, "vertices_per_s": V / (k_ms * 1e-3),
              "algorithmic_bytes": alg_bytes, "gbs": alg_bytes / (k_ms * 1e-3) / 1e9,
              "oracle_cpu": {"vertices": len(sub), "seconds": cpu_s, "vertices_per_s": len(sub) / cpu_s, "threads": 1},
              "parity": "bit-identical to the oracle on the sample"}

# ---- sampler ----------------------------------------------------------------------------------------------------
fp, rhs, fixed, val = patch_family(n0, engine="native")
for k in FlatProblem.ARRAYS:
    a = getattr(fp, k)
    if a is not None and k != "elem_mesh":
        setattr(fp, k, torch.from_numpy(np.ascontiguousarray(a)).to(dev))
ctx.set_problem(fp).symbolic().assemble()
ctx.set_system(torch.from_numpy(rhs).to(dev), torch.from_numpy(fixed).to(dev), torch.from_numpy(val).to(dev))
u_d = torch.empty(fp.n_dof, dtype=torch.float64, device=dev)
ctx.solve(rtol=1e-6, maxit=100000, out=u_d, check=False)
L = _lib.load()
nS = 3
grid = np.ascontiguousarray(np.linspace(-1, 1, nS))
nu = C.c_int64(0)
ctx._ck(L.mofem_sample(ctx._h, None, nS, 2, None, None, 0, C.byref(nu)))
out_d = torch.empty((nu.value, 7, nS, nS), dtype=torch.float64, device=dev)
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
times = []
for _ in range(5):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    ctx._ck(L.mofem_sample(ctx._h, _lib.ptr(u_d), nS, 2, _lib.ptr(grid), _lib.ptr(out_d), int(nu.value), C.byref(nu)))
    e1.record(stream)
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
s_ms = float(np.median(times[1:]))
vals = out_d[:1000].cpu().numpy()
out["sampler"] = {"units": int(nu.value), "points": int(nu.value) * nS * nS, "call_ms": s_ms,
                  "points_per_s": int(nu.value) * nS * nS / (s_ms * 1e-3), "output_bytes": int(out_d.numel()) * 8,
                  "output_gbs": int(out_d.numel()) * 8 / (s_ms * 1e-3) / 1e9, "finite": bool(np.isfinite(vals).all()),
                  "note": "whole mofem_sample call with device-resident u and output (unit count + fill + sample kernels); "
                          "the reference's getGraphDataOFE runs ~250 sample points/s on one core (AMOREBracket: 4.5 s for 1140 units)"}
print(json.dumps(out, indent=1))
