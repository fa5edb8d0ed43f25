"""Assemble + a fixed number of PCG iterations on the bench problem, for ncu / tuning runs (development tool).

    python tools/profile_run.py [n0=664] [iters=20] [key=value ...]     keys: l2_persist, spmv_ring, pdl, profile, reps
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from mofem_b200.device import Context
from mofem_b200.flatten import FlatProblem
from mofem_b200.overlay_gen import patch_family

pos = [a for a in sys.argv[1:] if "=" not in a]
kv = dict(a.split("=") for a in sys.argv[1:] if "=" in a)
n0 = int(pos[0]) if len(pos) > 0 else 664
iters = int(pos[1]) if len(pos) > 1 else 20
fp, rhs, fixed, val = patch_family(n0, engine="native")
dev = torch.device("cuda", 0)
for k in FlatProblem.ARRAYS:
    a = getattr(fp, k)
    if a is not None and k != "elem_mesh":
        setattr(fp, k, torch.from_numpy(np.ascontiguousarray(a)).to(dev))
pr = torch.cuda.get_device_properties(0)
print("device", pr.name, "L2", pr.L2_cache_size, {k: getattr(pr, k) for k in dir(pr) if "persist" in k.lower() or "window" in k.lower()})
ctx = Context(0)
ctx.set_problem(fp).symbolic().assemble()
ctx.set_system(torch.from_numpy(rhs).to(dev), torch.from_numpy(fixed).to(dev), torch.from_numpy(val).to(dev))
configs = kv.get("configs")
runs = []
if configs:      # e.g. configs=ring:pdl,ring,pdl,none   (each: the switches that are ON)
    for c in configs.split(","):
        on = set(c.split(":"))
        win = [int(t[1:]) for t in on if t[:1] == "w" and t[1:].isdigit()]
        runs.append(dict(spmv_ring=int("ring" in on), pdl=int("pdl" in on), l2_persist=int("l2" in on), profile=int("prof" in on),
                         l2_window=win[0] if win else 5))
else:
    runs.append(dict(spmv_ring=int(kv.get("spmv_ring", 1)), pdl=int(kv.get("pdl", 1)), l2_persist=int(kv.get("l2_persist", 0)),
                     profile=int(kv.get("profile", 0)), l2_window=int(kv.get("l2_window", 5))))
for opt in runs:
    for k in ("spmv_ring", "pdl", "l2_persist", "l2_window"):
        ctx.set_option(k, opt[k])
    ctx.set_profile(bool(opt["profile"]))
    best = None
    for rep in range(int(kv.get("reps", 3))):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        u, info = ctx.solve(rtol=1e-10, maxit=iters, check=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        ms = ctx.timing()["solve_ms"]
        best = ms if best is None else min(best, ms)
    line = {"opt": opt, "iterations": info["iterations"], "solve_ms": round(best, 3), "us_per_iteration": round(best / max(info["iterations"], 1) * 1e3, 2),
            "relres_rec": info["relres_rec"], "relres_true": info["relres_true"]}
    if opt["profile"]:
        pp = ctx.pcg_profile()
        line["kernel_us"] = {k: round(v / max(pp["iterations"], 1) * 1e3, 2) for k, v in pp.items() if k != "iterations"}
    print(line, flush=True)
print("ok", ctx.info())
