#!/usr/bin/env python
"""Fast (tensor-core, three-term bf16) vs strict deep MLP: output / residual / loss differences and time.  Prints JSON."""
import argparse, json, os, statistics, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=128)
    ap.add_argument("--cases", default="32:2,32:3,64:2,64:3,64:5,128:2,128:3")
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    import numpy as np
    import torch
    from phys_autodiff_b200 import Grid, MLPConfig, PhysWeights, ops
    n = a.grid
    g = Grid(n, n, n, 1.0, 1.0, 1.0, 2e-3, True)
    ctx = ops.Context(0)
    rng = np.random.default_rng(0)
    pw = PhysWeights(1.0, 1.0)

    def timeit(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(a.iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        return statistics.median(ts)

    rows = []
    for case in a.cases.split(","):
        H, L = (int(v) for v in case.split(":"))
        W1, b1, W2, b2 = ops.mlp_random_init(H, 777, 0.25)
        Wh = rng.uniform(-0.2, 0.2, (L - 1) * H * H).astype(np.float32)
        bh = rng.uniform(-0.2, 0.2, (L - 1) * H).astype(np.float32)
        ctx.set_weights_deep(MLPConfig(4, H, 4, True), L, W1, b1, Wh, bh, W2, b2)
        ctx.set_deep_mode(0)
        fs = ctx.mlp_generate_fields_deep(g, 0.25, 2e-3)
        ms_s = timeit(lambda: ctx.mlp_generate_fields_deep(g, 0.25, 2e-3))
        acc_s, Rs = ctx.phys_loss_acc(g, fs, want_residuals=True)
        ctx.set_deep_mode(1)
        ff = ctx.mlp_generate_fields_deep(g, 0.25, 2e-3)
        ms_f = timeit(lambda: ctx.mlp_generate_fields_deep(g, 0.25, 2e-3))
        acc_f, Rf = ctx.phys_loss_acc(g, ff, want_residuals=True)
        rmax = max(float(r.abs().max()) for r in Rs)
        rerr = max(float((x - y).abs().max()) for x, y in zip(Rs, Rf)) / rmax
        ctx.set_deep_mode(0)
        errs = []
        for x, y in zip(fs, ff):
            errs.append(float((x - y).abs().max() / x.abs().max()))
        ls, lf = ctx.finalize(acc_s.cpu().numpy(), pw, g.N), ctx.finalize(acc_f.cpu().numpy(), pw, g.N)
        rows.append({"H": H, "hidden_layers": L, "ms_strict": ms_s, "ms_fast": ms_f, "speedup": ms_s / ms_f,
                     "max_rel_output_err": max(errs), "max_residual_err_over_max_residual": rerr, "loss_strict": [float(v) for v in ls], "loss_fast": [float(v) for v in lf],
                     "loss_rel_err": [abs(float(x) - float(y)) / abs(float(x)) for x, y in zip(ls, lf)]})
        print(json.dumps(rows[-1]), flush=True)
    json.dump({"grid": [n, n, n], "rows": rows}, open(os.path.join(ROOT, "gpurun_out", "deep_tc_check.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
