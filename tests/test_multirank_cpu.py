"""CPU tier, world_size 2 over gloo: the host-side logic of the multi-GPU path -- slab partition,
partial sums, one all-reduce of two doubles, finalisation -- with the oracle standing in for the
kernel (no GPU here).  The GPU tier runs the same driver code with the real kernel."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import oracle
    from oracle import Grid
    from phys_autodiff_b200.ops import slab_for_rank
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P = oracle.port()
    g = Grid(10, 6, 7, 1, 1, 1, 2e-3, True)       # nz = 7: uneven slabs
    w = P.mlp_random_init(16, 42, 0.5)
    full = P.fused_loss(g, w, 0.25, 2e-3, 1.3, 0.7, want_residuals=True)
    z0, z1 = slab_for_rank(g.nz, rank, world)
    plane = g.nx * g.ny
    a_s, a_u = P.sumsq(full["R"], z0 * plane, z1 * plane)      # what this rank's kernel would produce
    acc = torch.tensor([a_s, a_u], dtype=torch.float64)
    dist.all_reduce(acc, op=dist.ReduceOp.SUM)
    ls = np.float32(np.float64(np.float32(1.3)) * acc[0].item() * (1.0 / g.N))
    lu = np.float32(np.float64(np.float32(0.7)) * acc[1].item() * (1.0 / g.N))
    q.put((rank, (z0, z1), float(ls), float(lu), float(full["loss_sigma"]), float(full["loss_u"])))
    dist.destroy_process_group()


def test_slab_partition_covers_grid():
    from phys_autodiff_b200.ops import slab_for_rank
    for nz in (1, 7, 24, 256):
        for world in (1, 2, 3, 4, 8):
            slabs = [slab_for_rank(nz, r, world) for r in range(world)]
            assert slabs[0][0] == 0 and slabs[-1][1] == nz
            assert all(a[1] == b[0] for a, b in zip(slabs, slabs[1:]))
            sizes = [b - a for a, b in slabs]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_allreduce_of_partial_sums():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert res[0][1] == (0, 3) and res[1][1] == (3, 7)
    for _, _, ls, lu, fs, fu in res:
        # summation order differs (two partial sums), so allow the documented loss tolerance
        assert abs(ls - fs) <= 1e-4 * abs(fs) and abs(lu - fu) <= 1e-4 * abs(fu)
    assert res[0][2] == res[1][2] and res[0][3] == res[1][3]   # every rank ends with the same loss
