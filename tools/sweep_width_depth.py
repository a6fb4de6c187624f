#!/usr/bin/env python
"""BASELINE config 5: MLP width/depth sweep at 256^3 (strict fp32, stage-wise: three time slices -> six fields,
then the physics loss on those fields).  H in {32, 64, 128}, hidden layers L in {1..5}; L = 1 also shows the fused
kernel, L >= 2 also the tensor-core mode.  Reports ms, Gpts/s and the fraction of the measured strict FMUL+FADD peak for the ALGORITHMIC flops
3 * (2*4H + (L-1)*2H^2 + 2*4H) + 3*L*H (ReLU) per point.  Prints JSON."""
import argparse, json, os, statistics, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=256)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--hs", default="32,64,128")
    ap.add_argument("--ls", default="1,2,3,4,5")
    a = ap.parse_args()
    import numpy as np
    import torch
    from phys_autodiff_b200 import Grid, MLPConfig, PhysWeights, ops
    n = a.grid
    g = Grid(n, n, n, 1.0, 1.0, 1.0, 2e-3, True)
    ctx = ops.Context(0)
    try:
        strict = json.load(open(os.path.join(ROOT, "profiles", "r01_microbench_fp32_long.json")))["strict"]["tflops"]
    except Exception:
        strict = 37.2
    rng = np.random.default_rng(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def timeit(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(a.iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        return statistics.mean(ts)

    out = {"grid": [n, n, n], "strict_fp32_peak_tflops": strict, "rows": []}
    for H in [int(v) for v in a.hs.split(',')]:
        W1, b1, W2, b2 = ops.mlp_random_init(H, 777, 0.25)
        for L in [int(v) for v in a.ls.split(',')]:
            Wh = rng.uniform(-0.2, 0.2, (L - 1) * H * H).astype(np.float32)
            bh = rng.uniform(-0.2, 0.2, (L - 1) * H).astype(np.float32)
            ctx.set_weights_deep(MLPConfig(4, H, 4, True), L, W1, b1, Wh, bh, W2, b2)
            f = ctx.mlp_generate_fields_deep(g, 0.25, 2e-3)
            ms_mlp = timeit(lambda: ctx.mlp_generate_fields_deep(g, 0.25, 2e-3))
            ms_phys = timeit(lambda: ctx.phys_loss_acc(g, f))
            flops = 3 * (2 * 4 * H + (L - 1) * 2 * H * H + 2 * 4 * H) + 3 * L * H
            row = {"H": H, "hidden_layers": L, "ms_fields": ms_mlp, "ms_phys_loss": ms_phys,
                   "gpts_per_s_total": g.N / (ms_mlp + ms_phys) / 1e6, "flops_per_point_mlp": flops,
                   "mlp_tflops": flops * g.N / (ms_mlp * 1e-3) / 1e12, "mlp_frac_of_strict_fp32": flops * g.N / (ms_mlp * 1e-3) / 1e12 / strict}
            if L >= 2:     # the tensor-core mode of the same network (tcgen05, three-term bf16 operands; not bit-exact)
                ctx.set_deep_mode(1)
                try:
                    ms_fast = timeit(lambda: ctx.mlp_generate_fields_deep(g, 0.25, 2e-3))
                    row["ms_fields_tensor_cores"] = ms_fast
                    row["tensor_core_speedup"] = ms_mlp / ms_fast
                    row["bf16_tflops_issued"] = 3 * (L - 1) * 6 * 2 * H * H * g.N / (ms_fast * 1e-3) / 1e12
                finally:
                    ctx.set_deep_mode(0)
            if L == 1:
                ctx.set_weights(MLPConfig(4, H, 4, True), W1, b1, W2, b2)
                row["ms_fused_kernel"] = timeit(lambda: ctx.fused_loss_acc(g, 0.25, 2e-3))
                row["ms_fields_one_layer_kernel"] = timeit(lambda: ctx.mlp_generate_fields(g, 0.25, 2e-3))
            out["rows"].append(row)
            del f
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
