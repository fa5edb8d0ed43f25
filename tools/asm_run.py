"""Assembly phases on the bench problem with the block kernels and the cell kernel (development tool).

    python tools/asm_run.py [n0=664] [family=quad|tri|m3s] [reps=5] [key=value ...]   (options passed to mofem_set_option)

Prints per option set: best symbolic / integrate / numeric ms, and whether block values and CSR data are bit-identical
between the option sets.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from mofem_b200.device import Context
from mofem_b200.flatten import FlatProblem
from mofem_b200.overlay_gen import patch_family, general_family

pos = [a for a in sys.argv[1:] if "=" not in a]
kv = dict(a.split("=") for a in sys.argv[1:] if "=" in a)
n0 = int(pos[0]) if pos else 664
family = kv.pop("family", "quad")
reps = int(kv.pop("reps", 5))
variants = kv.pop("variants", "cell_kernel:0,cell_kernel:1")
if family == "quad":
    fp, rhs, fixed, val = patch_family(n0, engine="native")
else:
    fp, rhs, fixed, val = general_family(n0, family)
dev = torch.device("cuda", 0)
for k in FlatProblem.ARRAYS:
    a = getattr(fp, k)
    if a is not None and k != "elem_mesh":
        setattr(fp, k, torch.from_numpy(np.ascontiguousarray(a)).to(dev))
ctx = Context(0)
for k, v in kv.items():
    ctx.set_option(k, int(v))
ref = None
for var in variants.split(","):
    opts = dict(o.split(":") for o in var.split("+"))
    for k, v in opts.items():
        ctx.set_option(k, int(v))
    best = {}
    for r in range(reps):
        ctx.set_problem(fp).symbolic().assemble()
        torch.cuda.synchronize()
        t = ctx.timing()
        for k in ("symbolic_ms", "integrate_ms", "numeric_ms"):
            best[k] = min(best.get(k, 1e30), t[k])
    info = ctx.info()
    V = ctx.coo_values()
    _, _, data = ctx.csr()
    same = None
    if ref is None:
        ref = (V, data, info["newton_updates"])
    else:
        same = bool(np.array_equal(ref[0], V) and np.array_equal(ref[1], data) and ref[2] == info["newton_updates"])
    print(json.dumps({"family": family, "n0": n0, "cells": int(fp.cell_kind.shape[0]), "opts": opts, **{k: round(v, 4) for k, v in best.items()},
                      "newton": info["newton_updates"], "n_eval": info.get("n_eval"), "identical_to_first": same}), flush=True)
