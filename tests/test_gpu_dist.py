"""Row-partitioned assembly + Jacobi-PCG on 2 GPUs (NCCL halo exchange and all-reduce inside the library) against the
single-GPU solve and scipy.  Needs >= 2 CUDA devices; skipped otherwise."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

NX, NY, WORLD = 16, 32, 2
NY4 = 64          # 4-rank variant: middle ranks have two neighbours


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out, transport, ny=NY):
    import torch
    import torch.distributed as dist
    from mofem_b200 import overlay_gen
    from mofem_b200.device import Context
    from mofem_b200.distributed import assemble_solve_partitioned, gather_global, init_comm
    from mofem_b200.partition import exchange_send_lists, partition_problem
    from test_partition import _owner
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        per = ny // 8 // world
        strip, rhs, fixed, val = overlay_gen.patch_family(NX, ny=ny, patch_rows=(rank * per, (rank + 1) * per))
        owner = _owner(strip, world, ny=ny)
        part = exchange_send_lists(partition_problem(strip, owner, rank, world))
        with Context(rank) as ctx:
            init_comm(ctx)
            u_loc, info, energy = assemble_solve_partitioned(ctx, part, part.local_vector(rhs), part.local_vector(fixed),
                                                             part.local_vector(val), rtol=1e-12, transport=transport)
            u = gather_global(part, u_loc, strip.n_node)
            # ghost entries of the returned local vector are current
            assert np.array_equal(u_loc, part.local_vector(u))
        out[rank] = (u, info, energy)
        out[("n_owned", rank)] = int(part.n_owned_node)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("transport", ["nccl", "peer"])
def test_two_gpu_solve_matches_single_gpu(gpu_ctx, transport):
    import torch
    if torch.cuda.device_count() < WORLD:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from mofem_b200 import overlay_gen
    fp, rhs, fixed, val = overlay_gen.patch_family(NX, ny=NY)
    gpu_ctx.clear_halo()
    gpu_ctx.set_problem(fp).symbolic().assemble()
    gpu_ctx.set_system(rhs, fixed, val)
    u1, info1 = gpu_ctx.solve(rtol=1e-12)
    e1 = gpu_ctx.energy()
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(WORLD, port, out, transport), nprocs=WORLD, join=True)
        res = [out[r] for r in range(WORLD)]
        owned = [out[("n_owned", r)] for r in range(WORLD)]
    # a rank with an ODD owned-node count: its last owned node and its first ghost node share one 32-byte sector of p, the case in
    # which a stale non-coherent read of the ghost would show (the halo rows read ghosts with ld.global.cg)
    assert any(n % 2 for n in owned), owned
    for u, info, energy in res:
        assert info["converged"] == 1
        assert np.abs(u - u1).max() <= 1e-9 * np.abs(u1).max()
        assert abs(energy - e1) <= 1e-9 * abs(e1)
    assert np.array_equal(res[0][0], res[1][0])
    # same Krylov process: iteration counts agree to within one polling chunk
    assert abs(res[0][1]["iterations"] - info1["iterations"]) <= 50


@pytest.mark.parametrize("transport", ["nccl", "peer"])
def test_four_gpu_solve_matches_single_gpu(gpu_ctx, transport):
    import torch
    if torch.cuda.device_count() < 4:
        pytest.skip("needs 4 GPUs")
    import torch.multiprocessing as mp
    from mofem_b200 import overlay_gen
    fp, rhs, fixed, val = overlay_gen.patch_family(NX, ny=NY4)
    gpu_ctx.clear_halo()
    gpu_ctx.set_problem(fp).symbolic().assemble()
    gpu_ctx.set_system(rhs, fixed, val)
    u1, info1 = gpu_ctx.solve(rtol=1e-12)
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(4, port, out, transport, NY4), nprocs=4, join=True)
        res = [out[r] for r in range(4)]
    for u, info, energy in res:
        assert info["converged"] == 1
        assert np.abs(u - u1).max() <= 1e-9 * np.abs(u1).max()
    for r in range(1, 4):
        assert np.array_equal(res[0][0], res[r][0])


@pytest.mark.parametrize("transport", ["peer"])
def test_eight_gpu_solve_matches_single_gpu(gpu_ctx, transport):
    import torch
    if torch.cuda.device_count() < 8:
        pytest.skip("needs 8 GPUs")
    import torch.multiprocessing as mp
    from mofem_b200 import overlay_gen
    ny = 128
    fp, rhs, fixed, val = overlay_gen.patch_family(NX, ny=ny)
    gpu_ctx.clear_halo()
    gpu_ctx.set_problem(fp).symbolic().assemble()
    gpu_ctx.set_system(rhs, fixed, val)
    u1, info1 = gpu_ctx.solve(rtol=1e-12)
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(8, port, out, transport, ny), nprocs=8, join=True)
        res = [out[r] for r in range(8)]
    for u, info, energy in res:
        assert info["converged"] == 1
        assert np.abs(u - u1).max() <= 1e-9 * np.abs(u1).max()
    for r in range(1, 8):
        assert np.array_equal(res[0][0], res[r][0])


def test_halo_plan_with_no_neighbours_is_the_single_gpu_path(gpu_ctx):
    """world = 1: the partitioned entry points reduce to the plain solve."""
    from mofem_b200 import overlay_gen
    from mofem_b200.partition import partition_problem, send_lists_from_requests
    fp, rhs, fixed, val = overlay_gen.patch_family(NX)
    gpu_ctx.clear_halo()
    gpu_ctx.set_problem(fp).symbolic().assemble()
    gpu_ctx.set_system(rhs, fixed, val)
    u1, _ = gpu_ctx.solve(rtol=1e-12)
    part = send_lists_from_requests(partition_problem(fp, np.zeros(fp.n_node, np.int32), 0, 1), [])
    assert part.n_owned_node == fp.n_node and len(part.neighbors) == 0
    gpu_ctx.set_problem(part.fp).symbolic().assemble()
    gpu_ctx.set_halo(part)
    gpu_ctx.set_system(part.local_vector(rhs), part.local_vector(fixed), part.local_vector(val))
    u2, _ = gpu_ctx.solve(rtol=1e-12)
    gpu_ctx.clear_halo()
    assert np.array_equal(u1, u2)
