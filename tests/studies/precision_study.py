#!/usr/bin/env python
"""Why the MLP is evaluated in strict fp32 (and not with FFMA contraction or on tensor cores): the residual error
each cheaper arithmetic would cause, measured on the CPU against the reference arithmetic.

For the benchmark network (seed 777, scale 0.25, 64^3, dt = 2e-3, periodic) the MLP is re-evaluated with
  strict   : separate fp32 multiply and add, reference order          (what this repository's kernels do)
  ffma     : fused multiply-add, same order                           (what nvcc's default contraction gives; the
                                                                       reference's own CUDA kernels)
  tf32     : inputs of every product rounded to 10 mantissa bits, fp32 accumulate   (tcgen05 kind::tf32)
  bf16     : inputs rounded to bf16 (7 bits), fp32 accumulate                        (tcgen05 kind::f16, bf16)
and the physics residuals of each are compared with the strict ones: max|dR| / max|R|  (gate: 1e-5).
Runs on the CPU only (numpy + the oracle); prints JSON."""
import json, os, sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402  (test infrastructure: this study lives under tests/ because it uses the checker)
from oracle import Grid  # noqa: E402


def round_mantissa(a, bits):
    """Round fp32 values to `bits` explicit mantissa bits (round-to-nearest-even on the bit pattern)."""
    u = a.astype(np.float32).view(np.uint32).astype(np.uint64)
    drop = 23 - bits
    half = (1 << (drop - 1)) - 1 + ((u >> drop) & 1)
    return (((u + half) >> drop) << drop).astype(np.uint32).view(np.float32)


def mlp(coords, W1, b1, W2, b2, mode):
    H = b1.size
    W1 = W1.reshape(H, 4); W2 = W2.reshape(4, H)
    f32 = np.float32
    cast = {"strict": lambda v: v, "ffma": lambda v: v, "tf32": lambda v: round_mantissa(v, 10), "bf16": lambda v: round_mantissa(v, 7)}[mode]

    def mac(s, w, x):  # s + w*x in the mode's arithmetic; s, x arrays, w scalar
        if mode == "ffma":
            return (s.astype(np.float64) + np.float64(w) * x.astype(np.float64)).astype(f32)
        p = (cast(np.full(1, w, f32))[0] * cast(x)).astype(f32) if mode in ("tf32", "bf16") else (f32(w) * x).astype(f32)
        if mode in ("tf32", "bf16"):  # tensor cores: exact products of the rounded inputs, fp32 accumulation
            p = (np.float64(cast(np.full(1, w, f32))[0]) * cast(x).astype(np.float64)).astype(f32)
        return (s + p).astype(f32)
    a = np.empty((coords.shape[0], H), f32)
    for h in range(H):
        s = np.full(coords.shape[0], b1[h], f32)
        for k in range(4):
            s = mac(s, W1[h, k], coords[:, k])
        a[:, h] = np.maximum(s, f32(0))
    y = np.empty((coords.shape[0], 4), f32)
    for o in range(4):
        s = np.full(coords.shape[0], b2[o], f32)
        for h in range(H):
            s = mac(s, W2[o, h], a[:, h])
        y[:, o] = s
    return y


def main():
    n, H, t, dt = 64, 64, 0.25, 2e-3
    P = oracle.port()
    g = Grid(n, n, n, 1, 1, 1, dt, True)
    w = P.mlp_random_init(H, 777, 0.25)
    N = g.N
    out = {"grid": [n, n, n], "hidden": H, "dt": dt, "gate": 1e-5, "modes": {}}
    ref_R = None
    for mode in ("strict", "ffma", "tf32", "bf16"):
        fields = []
        ys = []
        for tt in (np.float32(t) - np.float32(dt), np.float32(t), np.float32(t) + np.float32(dt)):
            c = P.make_grid_coords(g, float(tt)).reshape(N, 4)
            ys.append(mlp(c, *w, mode))
        s = [np.ascontiguousarray(y[:, 0]) for y in ys]
        u = [np.ascontiguousarray(np.concatenate([y[:, 1], y[:, 2], y[:, 3]])) for y in ys]
        R = P.phys_residuals(g, (s[0], s[1], s[2], u[0], u[1], u[2]))
        if mode == "strict":
            ref_R, ref_y = R, ys[1]
            chk = P.mlp_grid_infer(g, w, t).reshape(N, 4)
            assert np.array_equal(chk, ys[1]), "numpy strict emulation must equal the oracle bit for bit"
        err_R = max(float(np.max(np.abs(a.astype(np.float64) - b))) for a, b in zip(R, ref_R)) / \
            max(float(np.max(np.abs(b))) for b in ref_R)
        err_y = float(np.max(np.abs(ys[1].astype(np.float64) - ref_y))) / float(np.max(np.abs(ref_y)))
        out["modes"][mode] = {"max_output_err_rel_to_max": err_y, "max_residual_err_rel_to_max": err_R,
                              "passes_1e-5_residual_gate": bool(err_R <= 1e-5)}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
