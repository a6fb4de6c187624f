"""Shared test helpers: error norms in the reference's own terms and manufactured fields."""
import numpy as np


def max_rel_to_max(a, b):
    """max|a-b| / max|b| -- the per-array form of the north-star tolerance (SURVEY.md section 7:
    elementwise relative error is ill-defined where residuals cross zero)."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    d = np.max(np.abs(a - b)) if a.size else 0.0
    m = np.max(np.abs(b)) if b.size else 0.0
    return d / m if m > 0 else d


def rel_l2(a, b):
    """The reference tests' metric (e.g. test/test_phys_cpu_ref.cpp:73-86)."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    n = np.sqrt(np.sum(b * b))
    d = np.sqrt(np.sum((a - b) ** 2))
    return d / n if n > 0 else d


def bits_equal(a, b):
    """BIT equality of fp32 arrays: the uint32 views are compared, so +0 != -0 and NaN payloads count."""
    a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
    return a.shape == b.shape and bool(np.array_equal(a.view(np.uint32), b.view(np.uint32)))


def manufactured_fields(g, t, kx=1, ky=1, kz=1, const_u=True):
    """sigma = sin(kx x + ky y + kz z - t) on a 2*pi-periodic box, u = (1,1,1) or
    (sin z, cos x, sin y): the fields of test/test_phys_cpu_ref.cpp:32-48 and
    test/test_phys_cuda_fused_vs_nonfused.cpp:43-51 (float sinf/cosf like the reference)."""
    nx, ny, nz = g.nx, g.ny, g.nz
    x = (np.arange(nx, dtype=np.float32) * np.float32(g.hx))[None, None, :]
    y = (np.arange(ny, dtype=np.float32) * np.float32(g.hy))[None, :, None]
    z = (np.arange(nz, dtype=np.float32) * np.float32(g.hz))[:, None, None]
    N = g.N
    out_s, out_u = [], []
    for tt in (np.float32(t) - np.float32(g.dt), np.float32(t), np.float32(t) + np.float32(g.dt)):
        ph = (np.float32(kx) * x + np.float32(ky) * y + np.float32(kz) * z - tt).astype(np.float32)
        out_s.append(np.sin(ph).astype(np.float32).reshape(N))
        u = np.empty(3 * N, np.float32)
        if const_u:
            u[:] = 1.0
        else:
            u[0:N] = np.broadcast_to(np.sin(z), (nz, ny, nx)).reshape(N)
            u[N:2 * N] = np.broadcast_to(np.cos(x), (nz, ny, nx)).reshape(N)
            u[2 * N:] = np.broadcast_to(np.sin(y), (nz, ny, nx)).reshape(N)
        out_u.append(u)
    return out_s[0], out_s[1], out_s[2], out_u[0], out_u[1], out_u[2]
