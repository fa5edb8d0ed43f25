"""Multi-GPU driver: one process per GPU, ``torch.distributed`` for the set-up plumbing, NCCL inside the C-ABI
library for the per-iteration halo exchange and all-reduce (csrc/pcg.cu).

    part = partition_problem(fp_strip, owner, rank, world)       # host, integer work
    exchange_send_lists(part)                                     # one collective at set-up
    init_comm(ctx)                                                # NCCL communicator inside the library
    u_local, info, energy = assemble_solve_partitioned(ctx, part, rhs, fixed, val)
"""
from __future__ import annotations

import numpy as np

from .device import Context
from .partition import LocalPart


def init_comm(ctx: Context, group=None):
    """Create the library's NCCL communicator over the ranks of ``group`` (default: the world)."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    box = [ctx.dist_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    ctx.dist_init(box[0], rank, world)
    return rank, world


def setup_peer_transport(ctx: Context, part: LocalPart, group=None):
    """Switch the PCG loop of ``ctx`` from NCCL to the NVLink peer-memory transport (after ``ctx.set_halo(part)``).
    Collective over ``group``; returns False (and leaves NCCL in place) when CUDA IPC is not available."""
    import torch.distributed as dist
    handle, n_ghost = ctx.p2p_export()
    mine = (handle, int(part.n_owned_node), {int(nb): int(part.recv_ptr[k]) for k, nb in enumerate(part.neighbors)})
    gathered = [None] * part.world
    dist.all_gather_object(gathered, mine, group=group)
    send_remote_off = [gathered[int(nb)][2].get(part.rank, 0) for nb in part.neighbors]
    ok = True
    try:
        ctx.p2p_import([g[0] for g in gathered], [g[1] for g in gathered], send_remote_off)
    except Exception:
        ok = False
    flags = [None] * part.world
    dist.all_gather_object(flags, ok, group=group)
    if not all(flags):          # all ranks or none
        ctx.p2p_close()
        return False
    return True


def assemble_solve_partitioned(ctx: Context, part: LocalPart, rhs_local, fixed_local, val_local, rtol=1e-10,
                               maxit=400000, out=None, transport="nccl", group=None):
    """Assembly of the local problem (no exchange) + row-partitioned Jacobi-PCG.  The three system arrays are in
    LOCAL numbering (``part.local_vector(global)``), ghosts included.  ``transport``: "nccl" or "peer" (NVLink peer
    memory inside the kernels; collective set-up).  Returns (u_local, info, global energy)."""
    ctx.set_problem(part.fp).symbolic().assemble()
    ctx.set_halo(part)
    if transport == "peer" and part.world > 1:
        if not setup_peer_transport(ctx, part, group):
            raise RuntimeError("CUDA IPC peer transport could not be set up")
    ctx.set_system(rhs_local, fixed_local, val_local)
    u, info = ctx.solve(rtol=rtol, maxit=maxit, out=out)
    return u, info, ctx.energy()


def gather_global(part: LocalPart, u_local, n_node_global, group=None):
    """Assemble the global displacement vector on every rank from the owned parts (host, for checks)."""
    import torch.distributed as dist
    own = np.asarray(u_local).reshape(-1, 2)[:part.n_owned_node]
    gathered = [None] * part.world
    dist.all_gather_object(gathered, (part.local2global[:part.n_owned_node], own), group=group)
    u = np.zeros((n_node_global, 2))
    for ids, vals in gathered:
        u[ids] = vals
    return u.reshape(-1)
