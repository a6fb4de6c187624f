"""GPU tier, needs >= 2 GPUs: the real multi-rank path -- one process per GPU, z-slab per rank, one NCCL
all-reduce of two doubles -- launched with torch.distributed.run exactly as bench.py is."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
from phys_autodiff_b200 import Grid, MLPConfig, PhysWeights, ops
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
g = Grid(96, 80, 37, 1, 1, 1, 2e-3, True)          # 37 planes: uneven slabs
w = ops.mlp_random_init(64, 777, 0.25)
ctx = ops.Context(local)
ctx.set_weights(MLPConfig(4, 64, 4, True), *w)
ls, lu, R = ctx.fused_loss(g, PhysWeights(1.3, 0.7), 0.25, 2e-3, want_residuals=True)      # NCCL all-reduce
assert ctx.connect_peers()
p2p = [ctx.fused_loss(g, PhysWeights(1.3, 0.7), 0.25, 2e-3) for _ in range(5)]             # in-kernel exchange, 5 epochs
host_step = ctx.prepare_step_host(g, MLPConfig(4, 64, 4, True), *w, PhysWeights(1.3, 0.7), 0.25, 2e-3,
                                  slab=ops.slab_for_rank(g.nz, rank, world))
p2p += [host_step() for _ in range(3)]                                                     # one-call host form
ctx.disconnect_peers()
# stage-wise path on z-sharded, externally supplied fields: every rank generates only ITS slab of the six fields,
# boundary planes are all-gathered, stencil + reduction per slab, one all-reduce
fl = ctx.mlp_generate_fields(g, 0.25, 2e-3, slab=ops.slab_for_rank(g.nz, rank, world))
sh = ctx.phys_loss_sharded(g, PhysWeights(1.3, 0.7), fl, want_residuals=True)
z0, z1 = ops.slab_for_rank(g.nz, rank, world)
# single-GPU answer for the same grid on this rank's device (no process group involved)
whole = ctx.fused_loss_acc(g, 0.25, 2e-3).cpu().numpy()
Rw = [torch.empty(g.N, device="cuda") for _ in range(4)]
ctx.fused_loss_acc(g, 0.25, 2e-3, residuals=Rw)
plane = g.nx * g.ny
same = all(torch.equal(a, b[z0 * plane:z1 * plane]) for a, b in zip(R, Rw))
same = same and all(torch.equal(a, b[z0 * plane:z1 * plane]) for a, b in zip(sh[2], Rw))
l1 = ctx.finalize(whole, PhysWeights(1.3, 0.7), g.N)
# closed loop: slab gradients + one all-reduce of 9H+6 doubles == the single-GPU gradient
gl = ctx.fused_loss_grad(g, PhysWeights(1.3, 0.7), 0.25, 2e-3)
ga1, gg1 = ctx.fused_loss_grad_acc(g, PhysWeights(1.3, 0.7), 0.25, 2e-3)
gg1 = gg1.cpu().numpy()
grad_err = float(np.abs(gl[2] - gg1).max() / np.abs(gg1).max())
out = dict(grad_err=grad_err, grad_ls=float(gl[0]), grad_lu=float(gl[1]), rank=rank, sharded=[float(sh[0]), float(sh[1])], p2p=[[float(a), float(b)] for a, b in p2p], ls=float(ls), lu=float(lu), ls1=float(l1[0]), lu1=float(l1[1]), same=bool(same), slab=[z0, z1])
gathered = [None] * world
dist.all_gather_object(gathered, out)
if rank == 0:
    print("RESULT " + json.dumps(gathered))
dist.destroy_process_group()
'''


def test_two_rank_fused_loss_matches_single_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    n = min(torch.cuda.device_count(), 8)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29731", str(script)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")][-1]
    res = json.loads(line[len("RESULT "):])
    assert len(res) == n
    for d in res:
        assert d["same"], d                                  # sharded residuals == the whole-grid ones, bitwise
        assert d["ls"] == res[0]["ls"] and d["lu"] == res[0]["lu"]   # every rank ends with the same loss
        # peer-memory exchange: same value on every rank and every epoch, and equal to the NCCL result
        # up to the order of the additions (rank order vs NCCL's)
        assert all(p == res[0]["p2p"][0] for p in d["p2p"]), d["p2p"]
        assert d["sharded"] == res[0]["sharded"]
        assert abs(d["sharded"][0] - d["ls1"]) <= 1e-6 * abs(d["ls1"]) and abs(d["sharded"][1] - d["lu1"]) <= 1e-6 * abs(d["lu1"])
        assert abs(d["p2p"][0][0] - d["ls"]) <= 1e-6 * abs(d["ls"]) and abs(d["p2p"][0][1] - d["lu"]) <= 1e-6 * abs(d["lu"])
        assert abs(d["ls"] - d["ls1"]) <= 1e-6 * abs(d["ls1"]) and abs(d["lu"] - d["lu1"]) <= 1e-6 * abs(d["lu1"])
        # closed loop over the ranks == single GPU (different summation order only)
        assert d["grad_err"] <= 2e-6, d["grad_err"]   # fp32 32-point batch sums fall on different points
        assert abs(d["grad_ls"] - d["ls1"]) <= 1e-6 * abs(d["ls1"]) and abs(d["grad_lu"] - d["lu1"]) <= 1e-6 * abs(d["lu1"])
