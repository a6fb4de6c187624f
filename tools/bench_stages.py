#!/usr/bin/env python
"""Secondary measurements ("stage-wise, device-resident" rows): each stage-wise
operator on device-resident buffers, CUDA events, L2 flushed between iterations, against the HBM
roofline (MEASURED_PEAKS.json hbm_gbs, else the 6.65 TB/s fallback) or the strict fp32 pipe.
    python tools/bench_stages.py [--grid 256] [--hidden 64] [--iters 20]
Prints one JSON object."""
import argparse, json, os, statistics, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=256)
    ap.add_argument("--hidden", type=int, default=64)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    import torch
    from phys_autodiff_b200 import Grid, MLPConfig, PhysWeights, ops
    n, H = a.grid, a.hidden
    g = Grid(n, n, n, 1.0, 1.0, 1.0, 2e-3, True)
    N = g.N
    ctx = ops.Context(0)
    ctx.set_weights(MLPConfig(4, H, 4, True), *ops.mlp_random_init(H, 777, 0.25))
    pw = PhysWeights(1.0, 1.0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    try:
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]; hbm_src = "measured (MEASURED_PEAKS.json)"
    except Exception:
        hbm, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        strict = json.load(open(os.path.join(ROOT, "profiles", "r01_microbench_fp32_long.json")))["strict"]["tflops"]
    except Exception:
        strict = 37.2

    def timeit(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(a.iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        return statistics.mean(ts), min(ts)

    out = {"grid": [n, n, n], "hidden": H, "points": N, "hbm_peak_gbs": hbm, "hbm_peak_source": hbm_src,
           "strict_fp32_peak_tflops": strict, "stages": {}}

    def add(name, fn, bytes_per_pt=None, flops_per_pt=None):
        mean, mn = timeit(fn)
        d = {"ms_mean": mean, "ms_min": mn, "gpts_per_s": N / mean / 1e6}
        if bytes_per_pt:
            d.update(bytes_per_point=bytes_per_pt, achieved_gbs=bytes_per_pt * N / (mean * 1e-3) / 1e9,
                     frac_of_hbm=bytes_per_pt * N / (mean * 1e-3) / 1e9 / hbm)
        if flops_per_pt:
            d.update(flops_per_point=flops_per_pt, achieved_tflops=flops_per_pt * N / (mean * 1e-3) / 1e12,
                     frac_of_strict_fp32=flops_per_pt * N / (mean * 1e-3) / 1e12 / strict)
        out["stages"][name] = d

    fields = ctx.mlp_generate_fields(g, 0.25, 2e-3)
    R = ctx.phys_residuals(g, fields)
    cg, cw = g.c(), pw.c()
    import ctypes as C
    from phys_autodiff_b200.capi import ptr, check
    lib, h = ctx._lib, ctx._h
    st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
    y = torch.empty(N * 4, device="cuda")
    s = [torch.empty(N, device="cuda") for _ in range(3)]; u = [torch.empty(3 * N, device="cuda") for _ in range(3)]
    G = [torch.empty(N, device="cuda") for _ in range(4)]
    acc = torch.zeros(2, dtype=torch.float64, device="cuda")
    fp = [ptr(f) for f in fields]; Rp = [ptr(r) for r in R]; Gp = [ptr(x) for x in G]

    add("mlp_grid_infer (1 slice, AoS out)", lambda: check(lib.physad_mlp_grid_infer_dev(h, C.byref(cg), None, C.c_float(0.25), ptr(y), st())),
        bytes_per_pt=16, flops_per_pt=17 * H)
    add("mlp_generate_fields (3 slices, 6 fields out)", lambda: check(lib.physad_mlp_generate_fields_dev(
        h, C.byref(cg), None, C.c_float(0.25), C.c_float(2e-3), *[ptr(t) for t in s], *[ptr(t) for t in u], st())),
        bytes_per_pt=48, flops_per_pt=51 * H)
    add("phys_residuals (48 B in + 16 B out)", lambda: check(lib.physad_phys_residuals_dev(h, C.byref(cg), *fp, *Rp, st())), bytes_per_pt=64)
    add("phys_loss, loss only (48 B in)", lambda: check(lib.physad_phys_loss_dev(h, C.byref(cg), *fp, ptr(acc), None, None, None, None, st())), bytes_per_pt=48)
    add("phys_loss + residuals (64 B)", lambda: check(lib.physad_phys_loss_dev(h, C.byref(cg), *fp, ptr(acc), *Rp, st())), bytes_per_pt=64)
    add("phys_backward from residuals (32 B)", lambda: check(lib.physad_phys_backward_dev(h, C.byref(cg), C.byref(cw), *Rp, *Gp, st())), bytes_per_pt=32)
    add("phys_backward from fields (64 B)", lambda: check(lib.physad_phys_backward_from_fields_dev(h, C.byref(cg), C.byref(cw), *fp, *Gp, st())), bytes_per_pt=64)
    add("fused MLP+phys loss (metric path)", lambda: ctx.fused_loss_acc(g, 0.25, 2e-3, acc=acc), flops_per_pt=51 * H + 68)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
