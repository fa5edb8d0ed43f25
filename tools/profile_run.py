"""Short run of every kernel of the hot path at the bench workload, for ncu (a few PCG iterations only).

    python tools/profile_run.py [n0] [pcg_iterations]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from mofem_b200.device import Context
from mofem_b200.flatten import FlatProblem
from mofem_b200.overlay_gen import patch_family

n0 = int(sys.argv[1]) if len(sys.argv) > 1 else 664
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
l2 = int(sys.argv[3]) if len(sys.argv) > 3 else 1
fp, rhs, fixed, val = patch_family(n0)
dev = torch.device("cuda", 0)
for k in FlatProblem.ARRAYS:
    a = getattr(fp, k)
    if a is not None and k != "elem_mesh":
        setattr(fp, k, torch.from_numpy(np.ascontiguousarray(a)).to(dev))
ctx = Context(0)
ctx.set_option("l2_persist", l2)
ctx.set_profile(True)
for rep in range(2):
    ctx.set_problem(fp).symbolic().assemble()
    ctx.set_system(torch.from_numpy(rhs).to(dev), torch.from_numpy(fixed).to(dev), torch.from_numpy(val).to(dev))
    u, info = ctx.solve(rtol=1e-10, maxit=iters, check=False)
pp = ctx.pcg_profile()
print("per-iteration us:", {k: round(v / max(pp["iterations"], 1) * 1e3, 1) for k, v in pp.items() if k != "iterations"}, "l2_persist", l2)
print("ok", ctx.info(), info, ctx.timing())
