"""Assemble the bench problem on the GPU and dump its node-level block CSR for tools/spmv_lab (development tool)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from mofem_b200.device import Context
from mofem_b200.overlay_gen import patch_family

n0 = int(sys.argv[1]) if len(sys.argv) > 1 else 664
out = sys.argv[2] if len(sys.argv) > 2 else "/tmp/spmv_lab"
os.makedirs(out, exist_ok=True)
fp, rhs, fixed, val = patch_family(n0, engine="native")
ctx = Context(0)
ctx.set_problem(fp).symbolic().assemble()
indptr, indices, data = ctx.csr()
n_node = fp.n_node
node_ptr = (indptr[0::2].astype(np.int64)) // 4
starts = indptr[0:-1:2].astype(np.int64)
deg = node_ptr[1:] - node_ptr[:-1]
# column nodes: every second entry of the even scalar rows
sel = np.concatenate([np.arange(s, s + 2 * d, 2) for s, d in zip(starts[:0], deg[:0])]) if False else None
row_of = np.repeat(np.arange(n_node), deg)
k_in_row = np.arange(node_ptr[-1]) - node_ptr[row_of]
node_idx = (indices[starts[row_of] + 2 * k_in_row] // 2).astype(np.int32)
node_ptr.tofile(out + "/node_ptr.bin"); node_idx.tofile(out + "/node_idx.bin"); data.tofile(out + "/data.bin")
np.random.default_rng(1).standard_normal(2 * n_node).tofile(out + "/x.bin")
print("dumped", n_node, len(node_idx), "deg min/mean/max", deg.min(), deg.mean(), deg.max())
