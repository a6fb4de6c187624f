#!/usr/bin/env python
"""Per-block timeline of the fused kernel (physad_set_fused_trace): where a slab launch spends its time.

    python tools/trace_fused.py [--ranks 8] [--planes P] [--hidden 64] [--variant -1] [--out file.json]

Runs rank 0's slab of an N-rank run on ONE GPU, a few warm launches, then one traced launch, and prints a JSON
summary: launch span, prologue, per-plane step time, cost of a segment's first (halo) plane, idle tail of every
block (span end - block exit), and the blocks on the critical path.  Diagnostics only; nothing here is a bench value.
"""
import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ranks", type=int, default=8)
    ap.add_argument("--planes", type=int, default=0)
    ap.add_argument("--grid", type=int, default=256)
    ap.add_argument("--hidden", type=int, default=64)
    ap.add_argument("--variant", type=int, default=-1)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import torch
    from phys_autodiff_b200 import Grid, MLPConfig, ops
    from phys_autodiff_b200.ops import slab_for_rank
    n, H = args.grid, args.hidden
    g = Grid(n, n, n, 1.0, 1.0, 1.0, 2e-3, True)
    ctx = ops.Context(0)
    if args.variant >= 0:
        ctx.set_fused_variant(args.variant)
    ctx.set_weights(MLPConfig(4, H, 4, True), *ops.mlp_random_init(H, 777, 0.25))
    slab = slab_for_rank(n, 0, args.ranks)
    if args.planes:
        slab = (0, args.planes)
    acc = torch.zeros(2, dtype=torch.float64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(5):
        ctx.fused_loss_acc(g, 0.25, 2e-3, slab=slab, acc=acc)
    torch.cuda.synchronize()
    NB, S = 1024, 18
    runs = []
    for rep in range(5):
        buf = torch.zeros(NB * S * 2, dtype=torch.int64, device="cuda")
        ctx.set_fused_trace(buf)
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ctx.fused_loss_acc(g, 0.25, 2e-3, slab=slab, acc=acc); e1.record()
        torch.cuda.synchronize()
        ctx.set_fused_trace(None)
        t = buf.cpu().numpy().reshape(NB, S, 2)
        used = [b for b in range(NB) if t[b, 0, 0] != 0]
        runs.append((e0.elapsed_time(e1), t[used]))
    ms, t = sorted(runs, key=lambda r: r[0])[len(runs) // 2]
    gt = t[:, :, 0].astype("float64") * 1e-3     # us, global timer
    ck = t[:, :, 1].astype("float64")            # SM clocks
    t0 = gt[:, 0].min()
    start, pro, done, exit_ = gt[:, 0] - t0, gt[:, 1] - gt[:, 0], gt[:, 14] - t0, gt[:, 16] - t0
    span = exit_.max()
    planes = t[:, 15, 1].astype("int64")
    segs = []
    for b in range(t.shape[0]):
        for s in range(4):
            if t[b, 2 + 3 * s, 0] == 0:
                break
            segs.append((b, s, gt[b, 2 + 3 * s] - t0, gt[b, 3 + 3 * s] - gt[b, 2 + 3 * s], gt[b, 4 + 3 * s] - gt[b, 3 + 3 * s],
                         ck[b, 4 + 3 * s] - ck[b, 2 + 3 * s]))
    nseg = [sum(1 for x in segs if x[0] == b) for b in range(t.shape[0])]
    march = done - (start + pro)
    # per-plane step: regress block march time on (planes, segments)
    import numpy as np
    A = np.stack([planes.astype("float64"), np.array(nseg, dtype="float64")], axis=1)
    coef, *_ = np.linalg.lstsq(A, march, rcond=None)
    order = np.argsort(-exit_)
    out = {
        "event_ms": ms, "slab": slab, "hidden": H, "variant": args.variant, "blocks": int(t.shape[0]),
        "span_us": span, "event_minus_span_us": ms * 1e3 - span,
        "block_start_us": {"min": float(start.min()), "median": float(np.median(start)), "max": float(start.max())},
        "prologue_us": {"median": float(np.median(pro)), "max": float(pro.max())},
        "march_done_us": {"min": float(done.min()), "median": float(np.median(done)), "max": float(done.max())},
        "exit_us": {"min": float(exit_.min()), "median": float(np.median(exit_)), "max": float(exit_.max())},
        "reduce_tail_us_after_last_march": float(span - done.max()),
        "fit_us_per_plane": float(coef[0]), "fit_us_per_segment": float(coef[1]),
        "first_halo_plane_us": {"median": statistics.median(x[3] for x in segs), "max": max(x[3] for x in segs)},
        "planes_per_block": {"min": int(planes.min()), "max": int(planes.max()), "mean": float(planes.mean())},
        "segments": len(segs),
        "idle_before_span_end_us": {"mean": float((span - exit_).mean()), "max": float((span - exit_).max())},
        "sm_clock_mhz_est": float(np.median([(ck[b, 14] - ck[b, 1]) / max(1e-9, (gt[b, 14] - gt[b, 1])) for b in range(t.shape[0])])),
        "critical_blocks": [{"block": int(b), "sm": int(t[b, 15, 0]), "planes": int(planes[b]), "segments": nseg[b],
                             "start": float(start[b]), "march_done": float(done[b]), "exit": float(exit_[b])} for b in order[:6]],
        "fastest_blocks": [{"block": int(b), "planes": int(planes[b]), "segments": nseg[b], "march_done": float(done[b])} for b in order[-4:]],
    }
    s = json.dumps(out)
    print(s)
    if args.out:
        with open(args.out, "w") as fh:
            fh.write(s + "\n")
    ctx.close()


if __name__ == "__main__":
    main()
