"""Host-side logic of the multi-GPU path on CPU: partition, halo plan exchange over a real 2-process ``gloo``
rendezvous, and a row-partitioned Jacobi-PCG written in numpy (test code) that uses exactly the plan the CUDA path
consumes.  The local matrices come from the oracle."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from mofem_b200 import overlay_gen
from mofem_b200.partition import partition_problem, send_lists_from_requests

NX, NY, WORLD = 16, 32, 2


def _owner(fp, world, nx=NX, ny=NY):
    """Strips of equal height: mesh-0 node rows and the patches inside them."""
    n_node0 = (nx + 1) * (ny + 1)
    rows_per = ny // world
    owner = np.zeros(fp.n_node, np.int32)
    j = np.arange(n_node0) // (nx + 1)
    owner[:n_node0] = np.minimum(j // rows_per, world - 1)
    patch = np.arange(fp.n_node - n_node0) // 36
    owner[n_node0:] = (patch // (nx // 8)) // (rows_per // 8)
    return owner


def _global_reference():
    from oracle import oracle
    fp, rhs, fixed, val = overlay_gen.patch_family(NX, ny=NY)
    I, J, V, _, _ = oracle.assemble_coo(fp)
    n = fp.n_dof
    K = sp.coo_matrix((V, (I, J)), shape=(n, n)).tocsr()
    free = fixed == 0
    u = np.zeros(n)
    u[free] = spla.spsolve(K[free][:, free].tocsc(), rhs[free])
    return fp, K, rhs, fixed, val, u


def test_local_rows_are_complete_without_exchange():
    from oracle import oracle
    fp, K, rhs, fixed, val, _ = _global_reference()
    owner = _owner(fp, 4, ny=NY) if NY % 32 == 0 else _owner(fp, 2)
    world = int(owner.max()) + 1
    x = np.random.default_rng(0).standard_normal(fp.n_dof)
    y = K @ x
    seen = np.zeros(fp.n_node, bool)
    for r in range(world):
        part = partition_problem(fp, owner, r, world)
        Il, Jl, Vl, _, _ = oracle.assemble_coo(part.fp)
        nl = part.fp.n_dof
        Kl = sp.coo_matrix((Vl, (Il, Jl)), shape=(nl, nl)).tocsr()
        yl = (Kl @ part.local_vector(x))[:2 * part.n_owned_node]
        own = part.local2global[:part.n_owned_node]
        assert not seen[own].any()
        seen[own] = True
        assert np.abs(yl - y.reshape(-1, 2)[own].reshape(-1)).max() <= 1e-13 * np.abs(y).max()
        # ghosts are grouped by ascending owner rank
        g_owner = owner[part.local2global[part.n_owned_node:]]
        assert (np.diff(g_owner) >= 0).all()
    assert seen.all()


@pytest.mark.parametrize("family", ["m3s", "tri"])
def test_general_family_strips_give_complete_rows(family):
    """Strips of the general families (staggered mesh-2 patches straddle the strip boundary): every rank's local matrix has
    the complete rows of the nodes it owns, and the owned sets tile the node set."""
    from oracle import oracle
    world = 2
    fp, rhs, fixed, val = overlay_gen.general_family(NX, family, ny=NY)
    I, J, V, _, _ = oracle.assemble_coo(fp)
    n = fp.n_dof
    K = sp.coo_matrix((V, (I, J)), shape=(n, n)).tocsr()
    x = np.random.default_rng(1).standard_normal(n)
    y = K @ x
    seen = np.zeros(fp.n_node, bool)
    cells = 0
    for r in range(world):
        strip, _, _, _, owner, own_cells = overlay_gen.general_strip(NX, family, NY, r, world)
        cells += own_cells
        part = partition_problem(strip, owner, r, world)
        Il, Jl, Vl, _, _ = oracle.assemble_coo(part.fp)
        nl = part.fp.n_dof
        Kl = sp.coo_matrix((Vl, (Il, Jl)), shape=(nl, nl)).tocsr()
        yl = (Kl @ part.local_vector(x))[:2 * part.n_owned_node]
        own = part.local2global[:part.n_owned_node]
        assert not seen[own].any()
        seen[own] = True
        assert np.abs(yl - y.reshape(-1, 2)[own].reshape(-1)).max() <= 1e-13 * np.abs(y).max()
    assert seen.all() and cells == fp.n_cell


def test_strip_generation_matches_partition_of_the_full_problem():
    fp, *_ = overlay_gen.patch_family(NX, ny=NY)
    owner = _owner(fp, WORLD)
    for r in range(WORLD):
        full = partition_problem(fp, owner, r, WORLD)
        strip, *_ = overlay_gen.patch_family(NX, ny=NY, patch_rows=(r * 2, r * 2 + 2))
        part = partition_problem(strip, owner, r, WORLD)
        assert np.array_equal(full.local2global, part.local2global)
        assert full.fp.n_cell == part.fp.n_cell and full.fp.n_tri == part.fp.n_tri
        assert np.array_equal(np.sort(full.fp.tri_xy.reshape(-1)), np.sort(part.fp.tri_xy.reshape(-1)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    from mofem_b200.partition import exchange_send_lists
    from oracle import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        strip, rhs, fixed, val = overlay_gen.patch_family(NX, ny=NY, patch_rows=(rank * 2, rank * 2 + 2))
        owner = _owner(strip, world)
        part = exchange_send_lists(partition_problem(strip, owner, rank, world))
        Il, Jl, Vl, _, _ = oracle.assemble_coo(part.fp)
        nl = part.fp.n_dof
        K = sp.coo_matrix((Vl, (Il, Jl)), shape=(nl, nl)).tocsr()
        no = 2 * part.n_owned_node
        A = K[:no]
        f, fx, uf = part.local_vector(rhs), part.local_vector(fixed), part.local_vector(val)
        free = (fx[:no] == 0)

        def halo(v):
            reqs = []
            for k, nb in enumerate(part.neighbors):
                sb = torch.from_numpy(v.reshape(-1, 2)[part.send_idx[part.send_ptr[k]:part.send_ptr[k + 1]]].copy())
                rb = torch.empty((int(part.recv_ptr[k + 1] - part.recv_ptr[k]), 2), dtype=torch.float64)
                reqs.append((dist.isend(sb, int(nb)), dist.irecv(rb, int(nb)), rb, k))
            for s, r, rb, k in reqs:
                s.wait(); r.wait()
                v.reshape(-1, 2)[part.n_owned_node + part.recv_ptr[k]:part.n_owned_node + part.recv_ptr[k + 1]] = rb.numpy()

        def gsum(*vals):
            t = torch.tensor(vals, dtype=torch.float64)
            dist.all_reduce(t)
            return t.tolist()

        b = np.where(free, f[:no] - (A @ uf), 0.0)
        dinv = np.where(free, 1.0 / np.where(A.diagonal() != 0, A.diagonal(), 1.0), 0.0)
        x = np.zeros(nl); r = b.copy(); p = np.zeros(nl); p[:no] = dinv * r
        rz, bb = gsum(float(r @ (dinv * r)), float(r @ r))
        for it in range(5000):
            halo(p)
            q = np.where(free, A @ p, 0.0)
            (pq,) = gsum(float(p[:no] @ q))
            alpha = rz / pq
            x[:no] += alpha * p[:no]; r -= alpha * q
            z = dinv * r
            rz_new, rr = gsum(float(r @ z), float(r @ r))
            if rr <= 1e-24 * bb:
                break
            p[:no] = z + (rz_new / rz) * p[:no]
            rz = rz_new
        halo(x)
        u = x + np.where(fx, uf, 0.0)
        out[rank] = (part.local2global[:part.n_owned_node].copy(), u.reshape(-1, 2)[:part.n_owned_node].copy(), it)
    finally:
        dist.destroy_process_group()


def test_row_partitioned_pcg_over_gloo_world_size_2():
    import torch.multiprocessing as mp
    fp, K, rhs, fixed, val, u_ref = _global_reference()
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(WORLD, port, out), nprocs=WORLD, join=True)
        u = np.zeros((fp.n_node, 2))
        for r in range(WORLD):
            ids, vals, it = out[r]
            u[ids] = vals
    assert np.abs(u.reshape(-1) - u_ref).max() <= 1e-9 * np.abs(u_ref).max()


def test_send_list_validation():
    fp, *_ = overlay_gen.patch_family(NX, ny=NY)
    owner = _owner(fp, WORLD)
    part = partition_problem(fp, owner, 0, WORLD)
    with pytest.raises(ValueError):
        send_lists_from_requests(part, [part.ghost_globals(0)])     # rank 0 does not own its own ghosts
