"""Row partition of a flattened problem for the multi-GPU path (one process per GPU).

Ownership is per NODE.  Rank ``r`` gets a *local problem*: every cell that touches a node it owns (cells on a
partition boundary are integrated redundantly by both sides, so the CSR rows of owned nodes are complete without
any exchange of contributions), with nodes renumbered ``[owned (ascending global id) | ghosts grouped by owner rank
(ascending rank, ascending global id)]``.  The PCG then needs, per iteration, the ghost entries of the search
direction from their owners (halo exchange) and an all-reduce of the dot products -- see csrc/pcg.cu.

All of this is host-side integer work on numpy arrays; :func:`exchange_send_lists` is the only collective (run once
at set-up, over any ``torch.distributed`` backend -- ``gloo`` in the CPU tests, ``nccl`` on the GPUs).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from .flatten import FlatProblem


@dataclass
class LocalPart:
    rank: int
    world: int
    fp: FlatProblem                 # local problem, local node / element numbering
    n_owned_node: int
    local2global: np.ndarray        # [n_local_node] global node ids (owned first, then ghosts)
    neighbors: np.ndarray           # [n_nb] ranks that own ghosts of this rank (== ranks that need nodes from it)
    recv_ptr: np.ndarray            # [n_nb+1] ghost-segment offsets (nodes) per neighbour
    send_ptr: Optional[np.ndarray] = None    # [n_nb+1], filled by exchange_send_lists()
    send_idx: Optional[np.ndarray] = None    # local ids (owned) to send, neighbour by neighbour
    cell_global: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int64))   # global cell ids kept

    @property
    def n_local_node(self):
        return int(self.local2global.shape[0])

    def ghost_globals(self, k):
        a, b = self.n_owned_node + self.recv_ptr[k], self.n_owned_node + self.recv_ptr[k + 1]
        return self.local2global[a:b]

    def local_vector(self, v_global):
        """Restrict a global DOF vector (2 entries per node) to the local numbering."""
        v = np.asarray(v_global).reshape(-1, 2)
        return np.ascontiguousarray(v[self.local2global].reshape(-1))


def strip_owner(n_node, bounds):
    """Owner array for contiguous global node ranges ``bounds[r] <= node < bounds[r+1]``."""
    owner = np.empty(n_node, dtype=np.int32)
    for r in range(len(bounds) - 1):
        owner[bounds[r]:bounds[r + 1]] = r
    return owner


def partition_problem(fp: FlatProblem, owner: np.ndarray, rank: int, world: int) -> LocalPart:
    """Local problem of ``rank``.  ``fp`` may already be restricted to a superset of the needed cells (e.g. one
    strip of :func:`mofem_b200.overlay_gen.patch_family`); node and element ids in it are global."""
    owner = np.asarray(owner)
    nn = fp.elem_nn
    en = fp.elem_nodes
    valid = en >= 0
    # elements with at least one owned node
    elem_owned = ((owner[np.where(valid, en, 0)] == rank) & valid).any(1)
    ce = fp.cell_elem
    cell_keep = (np.where(ce >= 0, elem_owned[np.where(ce >= 0, ce, 0)], False)).any(1)
    cells = np.flatnonzero(cell_keep)
    ce_k = ce[cells]
    elems = np.unique(ce_k[ce_k >= 0])
    nodes = np.unique(en[elems][valid[elems]])
    owned_all = np.flatnonzero(owner == rank)                     # every owned node, used or not
    ghosts = nodes[owner[nodes] != rank]
    order = np.lexsort((ghosts, owner[ghosts]))
    ghosts = ghosts[order]
    local2global = np.concatenate([owned_all, ghosts]).astype(np.int64)
    n_owned = len(owned_all)
    g_owner = owner[ghosts]
    neighbors, counts = np.unique(g_owner, return_counts=True)
    recv_ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)

    g2l = np.full(fp.n_node, -1, dtype=np.int64)
    g2l[local2global] = np.arange(len(local2global))
    e2l = np.full(fp.n_elem, -1, dtype=np.int64)
    e2l[elems] = np.arange(len(elems))

    en_l = en[elems].astype(np.int64)
    en_l = np.where(en_l >= 0, g2l[np.where(en_l >= 0, en_l, 0)], -1).astype(np.int32)
    ce_l = np.where(ce_k >= 0, e2l[np.where(ce_k >= 0, ce_k, 0)], -1).astype(np.int32)
    tp = fp.cell_tri_ptr
    cnt = (tp[cells + 1] - tp[cells]).astype(np.int64)
    tri_ptr_l = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    if cnt.sum():
        tri_sel = np.repeat(tp[cells] - tri_ptr_l[:-1], cnt) + np.arange(tri_ptr_l[-1])
    else:
        tri_sel = np.zeros(0, dtype=np.int64)
    curved = fp.tri_curved is not None
    fl = FlatProblem(
        solver=fp.solver, n_mesh=fp.n_mesh, node_xy=np.ascontiguousarray(fp.node_xy[local2global]),
        elem_nodes=np.ascontiguousarray(en_l), elem_nn=np.ascontiguousarray(nn[elems]),
        elem_mat=np.ascontiguousarray(fp.elem_mat[elems]), elem_mesh=np.ascontiguousarray(fp.elem_mesh[elems]),
        mat_D=fp.mat_D, cell_elem=np.ascontiguousarray(ce_l), cell_kind=np.ascontiguousarray(fp.cell_kind[cells]),
        cell_tri_ptr=tri_ptr_l, tri_xy=np.ascontiguousarray(fp.tri_xy[tri_sel]),
        tri_rho=np.ascontiguousarray(fp.tri_rho[tri_sel]),
        tri_mid=np.ascontiguousarray(fp.tri_mid[tri_sel]) if curved else None,
        tri_curved=np.ascontiguousarray(fp.tri_curved[tri_sel]) if curved else None,
        n_mesh_regular=0, mesh_offset=[0])
    return LocalPart(rank=rank, world=world, fp=fl, n_owned_node=n_owned, local2global=local2global,
                     neighbors=neighbors.astype(np.int32), recv_ptr=recv_ptr, cell_global=cells)


def send_lists_from_requests(part: LocalPart, requests):
    """``requests[k]`` = global node ids neighbour ``part.neighbors[k]`` needs from this rank (its ghost list for
    this owner, in its ghost order).  Fills ``send_ptr`` / ``send_idx``."""
    owned = part.local2global[:part.n_owned_node]
    idx = []
    ptr = [0]
    for req in requests:
        req = np.asarray(req, dtype=np.int64)
        loc = np.searchsorted(owned, req)
        if len(req) and (loc.max() >= len(owned) or not np.array_equal(owned[loc], req)):
            raise ValueError("a neighbour asked for a node this rank does not own")
        idx.append(loc.astype(np.int32))
        ptr.append(ptr[-1] + len(req))
    part.send_ptr = np.array(ptr, dtype=np.int64)
    part.send_idx = np.concatenate(idx).astype(np.int32) if idx else np.zeros(0, np.int32)
    return part


def exchange_send_lists(part: LocalPart, group=None):
    """Collective (any torch.distributed backend): every rank learns which of its owned nodes its neighbours need.
    Ghost relations are symmetric for FE stencils; a rank that needs nothing from a neighbour that needs something
    from it gets an empty receive range for that neighbour."""
    import torch.distributed as dist
    mine = {int(nb): part.ghost_globals(k).tolist() for k, nb in enumerate(part.neighbors)}
    gathered = [None] * part.world
    dist.all_gather_object(gathered, mine, group=group)
    want_from_me = {r: np.array(g[part.rank], dtype=np.int64) for r, g in enumerate(gathered)
                    if r != part.rank and part.rank in g}
    nbs = sorted(set(want_from_me) | set(int(x) for x in part.neighbors))
    # rebuild the neighbour list as the union (keeps the recv ranges of the original neighbours)
    old = {int(nb): (part.recv_ptr[k], part.recv_ptr[k + 1]) for k, nb in enumerate(part.neighbors)}
    recv_ptr = [0]
    for nb in nbs:
        a, b = old.get(nb, (recv_ptr[-1], recv_ptr[-1]))
        assert a == recv_ptr[-1], "ghost segment must stay grouped by ascending owner rank"
        recv_ptr.append(b)
    part.neighbors = np.array(nbs, dtype=np.int32)
    part.recv_ptr = np.array(recv_ptr, dtype=np.int64)
    return send_lists_from_requests(part, [want_from_me.get(nb, np.zeros(0, np.int64)) for nb in nbs])
