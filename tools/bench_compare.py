#!/usr/bin/env python
"""bench_compare -- the report tool the reference plans in REQUIREMENT.md:138-151 (M5) and never ships: one CSV row per
(grid, field dtype, advection scheme) with the fused and the stage-wise time of the path, their ratio, peak device memory
and error metrics against the fp32 / central / stage-wise result.

    python tools/bench_compare.py [--grids 64,128,256] [--hidden 64] [--iters 10] [--out report/]

Writes <out>/bench_compare.csv and <out>/summary.txt (and prints the CSV).  Timing: CUDA events, median of --iters, a
256 MiB L2 flush between iterations.  "fused" = the one-kernel metric path (fp32 fields never leave the chip; central
scheme only); "stage-wise" = generate_fields + phys_loss on HBM-resident fields of the given dtype and scheme."""
import argparse
import csv
import io
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grids", default="64,128,256")
    ap.add_argument("--hidden", type=int, default=64)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--out", default="report")
    a = ap.parse_args()
    import torch
    from phys_autodiff_b200 import Grid, MLPConfig, PhysWeights, ops
    ctx = ops.Context(0)
    H = a.hidden
    ctx.set_weights(MLPConfig(4, H, 4, True), *ops.mlp_random_init(H, 777, 0.25))
    pw = PhysWeights(1, 1)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def timeit(fn):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(a.iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        return statistics.median(ts)

    rows = []
    for n in [int(v) for v in a.grids.split(",")]:
        g = Grid(n, n, n, 1, 1, 1, 2e-3, True)
        f32 = ctx.mlp_generate_fields(g, 0.25, 2e-3)
        acc_ref, R_ref = ctx.phys_loss_acc(g, f32, want_residuals=True)
        l_ref = ctx.finalize(acc_ref.cpu().numpy(), pw, g.N)
        rmax = [float(r.abs().max()) for r in R_ref]
        for dtype in ("f32", "f16", "bf16"):
            for scheme in ("central", "upwind"):
                if scheme == "upwind" and dtype != "f32":
                    continue          # the 16-bit kernels implement the central scheme
                torch.cuda.reset_peak_memory_stats()
                ctx.set_advection(scheme == "upwind")
                try:
                    if dtype == "f32":
                        gen = lambda: ctx.mlp_generate_fields(g, 0.25, 2e-3)
                        fld = gen()
                        loss = lambda: ctx.phys_loss_acc(g, fld)
                        acc, R = ctx.phys_loss_acc(g, fld, want_residuals=True)
                    else:
                        gen = lambda: ctx.mlp_generate_fields_lp(g, 0.25, 2e-3, dtype)
                        fld = gen()
                        loss = lambda: ctx.phys_loss_lp_acc(g, fld, dtype)
                        acc, R = ctx.phys_loss_lp_acc(g, fld, dtype, want_residuals=True)
                    t_gen, t_loss = timeit(gen), timeit(loss)
                finally:
                    ctx.set_advection(False)
                l = ctx.finalize(acc.cpu().numpy(), pw, g.N)
                t_fused = timeit(lambda: ctx.fused_loss_acc(g, 0.25, 2e-3)) if (dtype == "f32" and scheme == "central") else None
                rows.append({
                    "case": f"{n}^3/H{H}/{dtype}/{scheme}", "nx": n, "ny": n, "nz": n, "hidden": H, "dtype": dtype, "scheme": scheme,
                    "T_fused_ms": "" if t_fused is None else f"{t_fused:.4f}",
                    "T_fields_ms": f"{t_gen:.4f}", "T_phys_loss_ms": f"{t_loss:.4f}", "T_nonfused_ms": f"{t_gen + t_loss:.4f}",
                    "speedup_fused_vs_nonfused": "" if t_fused is None else f"{(t_gen + t_loss) / t_fused:.3f}",
                    "gpts_per_s_best": f"{g.N / (min(t_fused or 1e9, t_gen + t_loss) * 1e-3) / 1e9:.3f}",
                    "peak_mem_MiB": f"{torch.cuda.max_memory_allocated() / 2**20:.1f}",
                    "loss_sigma": f"{float(l[0]):.9g}", "loss_u": f"{float(l[1]):.9g}",
                    "loss_sigma_rel_err_vs_f32_central": f"{abs(float(l[0]) - float(l_ref[0])) / float(l_ref[0]):.3e}",
                    "loss_u_rel_err_vs_f32_central": f"{abs(float(l[1]) - float(l_ref[1])) / float(l_ref[1]):.3e}",
                    "residual_max_err_over_max_vs_f32_central": f"{max(float((x - y).abs().max()) / m for x, y, m in zip(R, R_ref, rmax)):.3e}",
                })
                del fld, R
        del f32, R_ref
    buf = io.StringIO()
    wr = csv.DictWriter(buf, fieldnames=list(rows[0].keys()))
    wr.writeheader()
    wr.writerows(rows)
    text = buf.getvalue()
    os.makedirs(a.out, exist_ok=True)
    with open(os.path.join(a.out, "bench_compare.csv"), "w") as fh:
        fh.write(text)
    best = [r for r in rows if r["T_fused_ms"]]
    with open(os.path.join(a.out, "summary.txt"), "w") as fh:
        for r in best:
            fh.write(f"{r['case']}: fused {r['T_fused_ms']} ms vs stage-wise {r['T_nonfused_ms']} ms "
                     f"(x{r['speedup_fused_vs_nonfused']}), peak memory of the stage-wise path {r['peak_mem_MiB']} MiB\n")
        fh.write("16-bit fields: see loss_*_rel_err columns (the time difference multiplies the field rounding by 1/(2 dt) = 250)\n")
    print(text, end="")


if __name__ == "__main__":
    main()
