#!/usr/bin/env python
"""What 16-bit field I/O costs and buys on the stage-wise path (GPU; writes one JSON object to stdout).

Drift: fields of the benchmark network (seed 777, H = 64, 64^3) stored as fp16 / bf16 instead of fp32; error of the
residuals (max |dR| / max |R|) and of the two losses against the fp32 path, for the reference's dt = 2e-3 and for
larger time steps -- the central time difference multiplies the rounding error of the fields by 1 / (2 dt).
Speed: generate_fields and the loss-only stencil at 256^3, fp32 vs 16-bit fields (48 vs 24 B/point read)."""
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import torch
    from phys_autodiff_b200 import Grid, MLPConfig, PhysWeights, ops
    ctx = ops.Context(0)
    ctx.set_weights(MLPConfig(4, 64, 4, True), *ops.mlp_random_init(64, 777, 0.25))
    out = {"network": "4-64-4, seed 777, scale 0.25, t 0.25", "drift_64cubed": [], "speed_256cubed": {}}
    for dt in (2e-3, 1e-2, 1e-1):
        g = Grid(64, 64, 64, 1, 1, 1, dt, True)
        f = ctx.mlp_generate_fields(g, 0.25, dt)
        acc, R = ctx.phys_loss_acc(g, f, want_residuals=True)
        ref_l = ctx.finalize(acc.cpu().numpy(), PhysWeights(1, 1), g.N)
        for dtype in ("f16", "bf16"):
            fl = ctx.mlp_generate_fields_lp(g, 0.25, dt, dtype)
            a2, R2 = ctx.phys_loss_lp_acc(g, fl, dtype, want_residuals=True)
            l2 = ctx.finalize(a2.cpu().numpy(), PhysWeights(1, 1), g.N)
            rerr = max(float((x - y).abs().max() / y.abs().max()) for x, y in zip(R2, R))
            out["drift_64cubed"].append({"dt": dt, "fields": dtype, "residual_max_err_over_max": rerr,
                                         "loss_sigma_rel_err": abs(float(l2[0]) - float(ref_l[0])) / float(ref_l[0]),
                                         "loss_u_rel_err": abs(float(l2[1]) - float(ref_l[1])) / float(ref_l[1])})
    g = Grid(256, 256, 256, 1, 1, 1, 2e-3, True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def timeit(fn, n=10):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(n):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        return statistics.median(ts)
    f = ctx.mlp_generate_fields(g, 0.25, 2e-3)
    sp = out["speed_256cubed"]
    sp["generate_fields_fp32_ms"] = timeit(lambda: ctx.mlp_generate_fields(g, 0.25, 2e-3))
    sp["phys_loss_only_fp32_ms"] = timeit(lambda: ctx.phys_loss_acc(g, f))
    sp["phys_loss_only_fp32_frac_of_hbm_peak"] = 48 * g.N / (sp["phys_loss_only_fp32_ms"] * 1e-3) / 6547.5e9
    del f
    for dtype in ("f16", "bf16"):
        fl = ctx.mlp_generate_fields_lp(g, 0.25, 2e-3, dtype)
        sp[f"generate_fields_{dtype}_ms"] = timeit(lambda: ctx.mlp_generate_fields_lp(g, 0.25, 2e-3, dtype))
        ms = timeit(lambda: ctx.phys_loss_lp_acc(g, fl, dtype))
        sp[f"phys_loss_only_{dtype}_ms"] = ms
        sp[f"phys_loss_only_{dtype}_frac_of_hbm_peak"] = 24 * g.N / (ms * 1e-3) / 6547.5e9
        del fl
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
