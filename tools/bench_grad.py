"""Times the closed-loop step (loss + gradient w.r.t. the MLP weights): CUDA events around the launches of one
step (fields, residuals + sums, stencil adjoint, MLP backward [+ the all-reduce of 9H+6 doubles under torchrun]).
  python tools/bench_grad.py [--n 256] [--hidden 64] [--steps 20]
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_grad.py   (z-slab per rank)"""
import argparse
import json
import sys, os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from phys_autodiff_b200 import Grid, MLPConfig, PhysWeights, ops


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--hidden", type=int, default=64)
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    g = Grid(a.n, a.n, a.n, 1, 1, 1, 2e-3, True)
    if "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        return distributed(a, g)
    ctx = ops.Context()
    ctx.set_weights(MLPConfig(4, a.hidden, 4, True), *ops.mlp_random_init(a.hidden, 777, 0.25))
    pw = PhysWeights(1, 1)
    acc, grad = ctx.fused_loss_grad_acc(g, pw, 0.25, 2e-3)
    for _ in range(3):
        ctx.fused_loss_grad_acc(g, pw, 0.25, 2e-3, acc, grad)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    for e0, e1 in ev:
        e0.record()
        ctx.fused_loss_grad_acc(g, pw, 0.25, 2e-3, acc, grad)
        e1.record()
    torch.cuda.synchronize()
    ms = sorted(e0.elapsed_time(e1) for e0, e1 in ev)
    fwd = []
    for _ in range(a.steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ctx.fused_loss_acc(g, 0.25, 2e-3, acc=acc)
        e1.record()
        torch.cuda.synchronize()
        fwd.append(e0.elapsed_time(e1))
    fwd.sort()
    med = ms[len(ms) // 2]
    # backward-kernel work: per point and hidden unit 41 lane-ops (see grad_kernels.cuh) -> report the step only
    print(json.dumps({"workload": f"{a.n}^3 H={a.hidden} loss+grad", "ms_median": med, "ms_min": ms[0],
                      "gpts_per_s": g.N / med / 1e6, "forward_only_fused_ms": fwd[len(fwd) // 2],
                      "grad_norm": float(grad.norm()), "acc": acc.tolist()}))


def distributed(a, g):
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = ops.Context(local)
    ctx.set_weights(MLPConfig(4, a.hidden, 4, True), *ops.mlp_random_init(a.hidden, 777, 0.25))
    pw = PhysWeights(1, 1)
    slab = ops.slab_for_rank(g.nz, rank, world)
    buf = ctx.fused_loss_grad_slab_acc(g, pw, 0.25, 2e-3, slab)

    def step():
        ctx.fused_loss_grad_slab_acc(g, pw, 0.25, 2e-3, slab, buf)
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / a.steps], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"workload": f"{a.n}^3 H={a.hidden} loss+grad, z-slab per rank, 1 all-reduce of 9H+6 doubles",
                          "n_gpus": world, "ms_per_step_max_over_ranks": t.item(), "gpts_per_s": g.N / t.item() / 1e6,
                          "grad_norm": float(buf[2:].norm()), "acc": buf[:2].tolist()}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
