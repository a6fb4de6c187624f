#!/usr/bin/env python
"""Which deep-MLP arithmetic is closer to the real-number network?  fp64 numpy evaluation of the network and of the
physics loss on a small periodic grid, against the strict fp32 kernel and the tensor-core fast mode.  Prints JSON."""
import argparse, json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=48)
    ap.add_argument("--cases", default="32:3,64:3,64:5,128:2,128:3,128:5")
    a = ap.parse_args()
    import numpy as np
    from phys_autodiff_b200 import Grid, MLPConfig, PhysWeights, ops
    n = a.grid
    dt = 2e-3
    g = Grid(n, n, n, 1.0, 1.0, 1.0, dt, True)
    ctx = ops.Context(0)
    rng = np.random.default_rng(0)
    pw = PhysWeights(1.0, 1.0)
    ax = (2.0 * (np.arange(n, dtype=np.float32) / np.float32(n - 1)) - 1.0).astype(np.float32).astype(np.float64)   # the kernel's fp32 coordinates
    Z, Y, X = np.meshgrid(ax, ax, ax, indexing="ij")

    def net64(t, W1, b1, Wh, bh, W2, b2, H, L):
        c = np.stack([X.ravel(), Y.ravel(), Z.ravel(), np.full(X.size, float(np.float32(t)))], 1)
        act = np.maximum(c @ W1.reshape(H, 4).astype(np.float64).T + b1.astype(np.float64), 0)
        for l in range(L - 1):
            act = np.maximum(act @ Wh[l * H * H:(l + 1) * H * H].reshape(H, H).astype(np.float64).T + bh[l * H:(l + 1) * H].astype(np.float64), 0)
        return act @ W2.reshape(4, H).astype(np.float64).T + b2.astype(np.float64)

    def loss64(ym, y0, yp):
        f = lambda y, c: y[:, c].reshape(n, n, n)
        inv2h, inv2dt = 0.5, 1.0 / (2.0 * float(np.float32(dt)))
        d = lambda q, axis: (np.roll(q, -1, axis) - np.roll(q, 1, axis)) * inv2h     # axis 2 = x, 1 = y, 0 = z
        s, u = f(y0, 0), [f(y0, 1), f(y0, 2), f(y0, 3)]
        Rs = (f(yp, 0) - f(ym, 0)) * inv2dt + u[0] * d(s, 2) + u[1] * d(s, 1) + u[2] * d(s, 0) + s * (d(u[0], 2) + d(u[1], 1) + d(u[2], 0))
        Ru = [(f(yp, 1 + k) - f(ym, 1 + k)) * inv2dt + u[0] * d(u[k], 2) + u[1] * d(u[k], 1) + u[2] * d(u[k], 0) for k in range(3)]
        return float((Rs ** 2).mean()), float(sum((r ** 2).mean() for r in Ru))

    rows = []
    for case in a.cases.split(","):
        H, L = (int(v) for v in case.split(":"))
        W1, b1, W2, b2 = ops.mlp_random_init(H, 777, 0.25)
        Wh = rng.uniform(-0.2, 0.2, (L - 1) * H * H).astype(np.float32)
        bh = rng.uniform(-0.2, 0.2, (L - 1) * H).astype(np.float32)
        t0 = np.float32(0.25)
        ts = [t0 - np.float32(dt), t0, t0 + np.float32(dt)]
        y64 = [net64(t, W1, b1, Wh, bh, W2, b2, H, L) for t in ts]
        l64 = loss64(*y64)
        ctx.set_weights_deep(MLPConfig(4, H, 4, True), L, W1, b1, Wh, bh, W2, b2)
        out = {"H": H, "hidden_layers": L, "loss_fp64": l64}
        for mode, name in ((0, "strict"), (1, "fast")):
            ctx.set_deep_mode(mode)
            f = ctx.mlp_generate_fields_deep(g, 0.25, dt)
            y = f[1].cpu().numpy().astype(np.float64)
            ls = ctx.finalize(ctx.phys_loss_acc(g, f).cpu().numpy(), pw, g.N)
            out[name] = {"max_output_err_vs_fp64": float(np.abs(y - y64[1][:, 0]).max() / np.abs(y64[1]).max()),
                         "loss_rel_err_vs_fp64": [abs(float(ls[0]) - l64[0]) / l64[0], abs(float(ls[1]) - l64[1]) / l64[1]]}
        ctx.set_deep_mode(0)
        rows.append(out)
        print(json.dumps(out), flush=True)
    json.dump({"grid": [n, n, n], "rows": rows}, open(os.path.join(ROOT, "gpurun_out", "deep_tc_truth.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
