"""CPU tier: pin the oracle.  (a) the plain-C port against the golden vectors generated from the
unmodified reference, (b) the port against the reference library itself where it is present,
(c) the reference's own known-answer test (test/test_phys_cpu_ref.cpp) restated on the oracle."""
import math
import os
import subprocess

import numpy as np
import pytest

from helpers import bits_equal, manufactured_fields, max_rel_to_max, rel_l2
from oracle import Grid


def _case_grid(c):
    return Grid(*c["g"], *c["h"], c["dt"], c["periodic"])


def test_weights_match_golden(port, golden):
    arr, meta = golden
    for c in meta["cases"]:
        if c["kind"] != "weights":
            continue
        w = port.mlp_random_init(c["H"], c["seed"], c["scale"])
        for name, a in zip(["W1", "b1", "W2", "b2"], w):
            assert np.array_equal(a, arr[f"{c['key']}_{name}"]), (c, name)
    assert float(port.mlp_random_init(64, 777, 0.25)[0][0]) == float(np.float32(-0.173668131))


def test_full_path_matches_golden(port, golden):
    arr, meta = golden
    for c in meta["cases"]:
        if c["kind"] != "path":
            continue
        g = _case_grid(c)
        n = c["name"]
        w = port.mlp_random_init(c["H"], c["seed"], c["scale"])
        assert bits_equal(port.mlp_grid_infer(g, w, c["t"], c["m1p1"]), arr[n + "_y"]), n
        f = port.generate_fields(g, w, c["t"], c["dt"], c["m1p1"])
        for k, a in zip(["sm", "s0", "sp", "um", "u0", "up"], f):
            assert bits_equal(a, arr[f"{n}_{k}"]), (n, k)
        ls, lu, R = port.phys_loss_forward(g, c["w"][0], c["w"][1], f, want_residuals=True)
        for k, a in zip(["Rs", "Rx", "Ry", "Rz"], R):
            assert bits_equal(a, arr[f"{n}_{k}"]), (n, k)
        assert float(ls) == float(c["loss_sigma"]) and float(lu) == float(c["loss_u"]), n
        G = port.phys_loss_backward(g, c["w"][0], c["w"][1], R)
        for k, a in zip(["gs", "gx", "gy", "gz"], G):
            assert bits_equal(a, arr[f"{n}_{k}"]), (n, k)
        # composed entry point == staged calls
        fl = port.fused_loss(g, w, c["t"], c["dt"], c["w"][0], c["w"][1], c["m1p1"], want_residuals=True)
        assert float(fl["loss_sigma"]) == float(ls) and float(fl["loss_u"]) == float(lu)


def test_mlp_backward_matches_golden(port, golden):
    arr, meta = golden
    c = [c for c in meta["cases"] if c["kind"] == "backward"][0]
    got = port.mlp_backward(arr["bwd_x"], arr["bwd_t"], arr["bwd_W1"], arr["bwd_b1"], arr["bwd_W2"], arr["bwd_b2"],
                            c["B"], c["In"], c["H"], c["Out"])
    for k, a in zip(["dW1", "db1", "dW2", "db2"], got):
        assert bits_equal(a, arr["bwd_" + k]), k


def test_deep_oracle_with_one_hidden_layer_is_the_pinned_forward(port):
    """oracle_mlp_forward_deep is unpinned for L > 1; its L = 1 case must equal the reference-pinned forward."""
    rng = np.random.default_rng(5)
    B, In, H, Out = 200, 4, 64, 4
    x = rng.uniform(-1, 1, B * In).astype(np.float32)
    W1 = rng.uniform(-.4, .4, H * In).astype(np.float32); b1 = rng.uniform(-.4, .4, H).astype(np.float32)
    W2 = rng.uniform(-.4, .4, Out * H).astype(np.float32); b2 = rng.uniform(-.4, .4, Out).astype(np.float32)
    assert bits_equal(port.mlp_forward_deep(x, 1, W1, b1, None, None, W2, b2, B, In, H, Out),
                      port.mlp_forward(x, W1, b1, W2, b2, B, In, H, Out))
    # L = 2 with an identity-like middle layer (W = I, b = 0) reproduces L = 1: relu(a) = a for a >= 0
    Wh = np.eye(H, dtype=np.float32).reshape(-1); bh = np.zeros(H, np.float32)
    assert bits_equal(port.mlp_forward_deep(x, 2, W1, b1, Wh, bh, W2, b2, B, In, H, Out),
                      port.mlp_forward(x, W1, b1, W2, b2, B, In, H, Out))


def test_anchor_losses_64cubed(port, golden):
    """Scalar anchors at a BASELINE size (also listed in SURVEY.md section 7 / BASELINE.md section 2)."""
    _, meta = golden
    a = meta["anchors"]["64c_h32"]
    g = Grid(64, 64, 64, 1, 1, 1, 2e-3, True)
    w = port.mlp_random_init(32, 777, 0.25)
    r = port.fused_loss(g, w, 0.25, 2e-3, want_residuals=True)
    assert float(r["loss_sigma"]) == float(a["loss_sigma"]) == float(np.float32(0.00194079021))
    assert float(r["loss_u"]) == float(a["loss_u"]) == float(np.float32(0.0100833438))
    assert float(r["R"][0][0]) == float(a["R_sigma0"])


def test_port_equals_reference_library(port, ref):
    if ref is None:
        pytest.skip("oracle/_ref not built here")
    rng = np.random.default_rng(0)
    for seed in (0, 1, 7, 2**32 - 1):
        for H in (8, 64, 100):
            a, b = port.mlp_random_init(H, seed, 0.3), ref.mlp_random_init(H, seed, 0.3)
            assert all(np.array_equal(x, y) for x, y in zip(a, b))
    for (nx, ny, nz, per, m1p1) in [(17, 9, 5, True, True), (8, 8, 8, False, False), (3, 1, 2, True, True)]:
        g = Grid(nx, ny, nz, 0.7, 1.3, 0.9, 3e-3, per)
        w = ref.mlp_random_init(24, 5, 0.4)
        assert bits_equal(port.make_grid_coords(g, 0.1, m1p1), ref.make_grid_coords(g, 0.1, m1p1))
        fp, fr = port.generate_fields(g, w, 0.1, 3e-3, m1p1), ref.generate_fields(g, w, 0.1, 3e-3, m1p1)
        assert all(bits_equal(x, y) for x, y in zip(fp, fr))
        # random (non-MLP) fields through the physics
        f = [rng.standard_normal(g.N).astype(np.float32) for _ in range(3)] + \
            [rng.standard_normal(3 * g.N).astype(np.float32) for _ in range(3)]
        rp, rr = port.phys_loss_forward(g, 1.1, 0.8, f, True), ref.phys_loss_forward(g, 1.1, 0.8, f, True)
        assert float(rp[0]) == float(rr[0]) and float(rp[1]) == float(rr[1])
        assert all(bits_equal(x, y) for x, y in zip(rp[2], rr[2]))
    # generic dims through mlp_forward
    B, In, H, Out = 37, 5, 19, 3
    x = rng.standard_normal(B * In).astype(np.float32)
    W1 = rng.standard_normal(H * In).astype(np.float32); b1 = rng.standard_normal(H).astype(np.float32)
    W2 = rng.standard_normal(Out * H).astype(np.float32); b2 = rng.standard_normal(Out).astype(np.float32)
    assert bits_equal(port.mlp_forward(x, W1, b1, W2, b2, B, In, H, Out), ref.mlp_forward(x, W1, b1, W2, b2, B, In, H, Out))
    t = rng.standard_normal(B * Out).astype(np.float32)
    for a, b in zip(port.mlp_backward(x, t, W1, b1, W2, b2, B, In, H, Out), ref.mlp_backward(x, t, W1, b1, W2, b2, B, In, H, Out)):
        assert bits_equal(a, b)


def test_reference_mt_driver_is_bit_identical(ref):
    """The threaded chunk driver in oracle/ref_shim.cpp must equal the reference's one-call path."""
    if ref is None:
        pytest.skip("oracle/_ref not built here")
    g = Grid(20, 12, 9, 1, 1, 1, 2e-3, True)
    w = ref.mlp_random_init(32, 777, 0.25)
    f = ref.generate_fields(g, w, 0.25, 2e-3)
    ls, lu, R = ref.phys_loss_forward(g, 1.0, 1.0, f, True)
    r = ref.fused_loss(g, w, 0.25, 2e-3, threads=3, want_residuals=True)
    assert float(r["loss_sigma"]) == float(ls) and float(r["loss_u"]) == float(lu)
    assert all(bits_equal(a, b) for a, b in zip(r["R"], R))


def test_known_answer_manufactured_solution(checker):
    """test/test_phys_cpu_ref.cpp: sigma = sin(x+y+z-t), u = (1,1,1) on 64x64x32 =>
    R_sigma = cos(phi) * (-sin(dt)/dt + sum_d sin(h_d)/h_d), R_u = 0; then g = 2 w / N * R."""
    nx, ny, nz = 64, 64, 32
    g = Grid(nx, ny, nz, 2 * math.pi / nx, 2 * math.pi / ny, 2 * math.pi / nz, 1e-3, True)
    t = 1.2345
    f = manufactured_fields(g, t)
    Rs, Rx, Ry, Rz = checker.phys_residuals(g, f)
    x = (np.arange(nx) * g.hx)[None, None, :]; y = (np.arange(ny) * g.hy)[None, :, None]; z = (np.arange(nz) * g.hz)[:, None, None]
    phi = x + y + z - t
    coef = -math.sin(g.dt) / g.dt + math.sin(g.hx) / g.hx + math.sin(g.hy) / g.hy + math.sin(g.hz) / g.hz
    exact = (np.cos(phi) * coef).reshape(-1)
    assert rel_l2(Rs, exact) <= 3e-4 and np.max(np.abs(Rs - exact)) <= 1e-3        # :87
    assert max(np.max(np.abs(Rx)), np.max(np.abs(Ry)), np.max(np.abs(Rz))) <= 1e-6  # :76
    G = checker.phys_loss_backward(g, 1.7, 0.9, (Rs, Rx, Ry, Rz))
    want = (np.float32(2.0) * np.float32(1.7) / np.float32(g.N)) * Rs
    assert rel_l2(G[0], want) <= 1e-7 and np.max(np.abs(G[0] - want)) <= 1e-6         # :120


def test_reference_own_test_binary_passes():
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "test_phys_cpu_ref")
    if not os.path.exists(exe):
        pytest.skip("reference test binary not built here")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "[PASS]" in r.stdout, r.stdout + r.stderr


def test_edge_cases_on_oracle(port):
    """Degenerate extents: n=1 axes give coordinate 0 and zero differences; clamp halves edges."""
    g = Grid(1, 1, 1, 1, 1, 1, 2e-3, True)
    w = port.mlp_random_init(16, 42, 0.5)
    r = port.fused_loss(g, w, 0.25, 2e-3, want_residuals=True)
    f = port.generate_fields(g, w, 0.25, 2e-3)
    dts = (np.float64(f[2][0]) - np.float64(f[0][0])) * (1.0 / (2.0 * np.float64(np.float32(2e-3))))
    assert float(r["R"][0][0]) == float(np.float32(dts))  # all spatial terms vanish
    # clamp: a field linear in x has interior slope a, edge slope a/2 (divisor stays 2h)
    g = Grid(6, 1, 1, 1, 1, 1, 1.0, False)
    s = np.arange(6, dtype=np.float32) * 2
    u = np.concatenate([np.ones(6, np.float32), np.zeros(12, np.float32)])
    Rs, *_ = port.phys_residuals(g, (s, s, s, u, u, u))
    assert list(Rs) == [1.0, 2.0, 2.0, 2.0, 2.0, 1.0]


def test_upwind_switch_properties(port):
    """The additive upwind-advection switch (REQUIREMENT.md:123-134, never shipped by the reference -- parity unpinned,
    oracle.c is its only statement).  The plan's own acceptance criteria: (1) consistent with the central scheme for
    small velocities -- the two differ only in u . grad(f), so the difference is linear in the velocity scale;
    (2) first-order accurate: on a smooth manufactured field the deviation from the central (second-order) residual
    halves with the grid spacing; (3) finite on large random velocity fields."""
    rng = np.random.default_rng(3)
    g = Grid(24, 20, 16, 0.5, 0.25, 1.0, 1e-2, True)
    N = g.N
    base = [rng.uniform(-1, 1, n).astype(np.float32) for n in (N, N, N, 3 * N, 3 * N, 3 * N)]

    def diff(eps):
        f = [base[0], base[1], base[2]] + [(eps * a).astype(np.float32) for a in base[3:]]
        c, u = port.phys_residuals(g, f), port.phys_residuals_upwind(g, f)
        return max(float(np.max(np.abs(a.astype(np.float64) - b))) for a, b in zip(c, u))
    d1, d2, d3 = diff(1.0), diff(1e-1), diff(1e-2)
    assert d1 > 0 and d2 <= 0.2 * d1 and d3 <= 0.2 * d2          # -> 0 with the velocity (linear for sigma, quadratic for u)
    errs = []
    for n in (32, 64):
        h = 2 * math.pi / n
        gm = Grid(n, 8, 8, h, h, h, 1e-3, True)
        f = manufactured_fields(gm, 0.3, kx=1, ky=0, kz=0, const_u=True)   # varies along x only: periodic for every n
        c, u = port.phys_residuals(gm, f), port.phys_residuals_upwind(gm, f)
        errs.append(float(np.max(np.abs(c[0].astype(np.float64) - u[0]))))
    assert 0.4 <= errs[1] / errs[0] <= 0.6, errs                  # O(h)
    big = [base[0], base[1], base[2]] + [(1e6 * a).astype(np.float32) for a in base[3:]]
    assert all(np.all(np.isfinite(r)) for r in port.phys_residuals_upwind(g, big))
    # clamped faces: a clamped neighbour is the point itself, the one-sided difference vanishes there
    gc = Grid(5, 4, 3, 1, 1, 1, 1e-2, False)
    f = [rng.uniform(-1, 1, n).astype(np.float32) for n in (60, 60, 60, 180, 180, 180)]
    assert all(np.all(np.isfinite(r)) for r in port.phys_residuals_upwind(gc, f))


def test_tangent_loss_checker_agrees_with_finite_differences_where_the_network_is_affine(port):
    """The analytic (forward-mode) loss is an additive, non-parity mode; its checker (oracle_tangent_loss) is pinned the
    only way it can be: where every hidden unit is active over the whole domain the network is affine in (x,y,z,t), central
    differences of it are exact, and the tangent residuals must equal the reference-pinned finite-difference residuals
    (physical spacing h consistent with the normalisation: dc/dx * 2h = c(i+1) - c(i-1)).  With the random-init network the
    two losses differ by the discretisation error at the ReLU kinks -- a few per cent at 48^3, shrinking with the spacing."""
    rng = np.random.default_rng(2)
    H = 16
    W1 = rng.uniform(-0.2, 0.2, H * 4).astype(np.float32)
    b1 = np.full(H, 5.0, np.float32)                                   # all units active on [-1, 1]^3 x t
    W2 = rng.uniform(-0.3, 0.3, 4 * H).astype(np.float32)
    b2 = rng.uniform(-0.3, 0.3, 4).astype(np.float32)
    g = Grid(20, 12, 9, 0.7, 1.1, 0.9, 0.05, False)
    tan = port.tangent_loss(g, (W1, b1, W2, b2), 0.25, True, want_residuals=True)
    fd = port.fused_loss(g, (W1, b1, W2, b2), 0.25, 0.05, 1.0, 1.0, True, want_residuals=True)
    nx, ny, nz = g.nx, g.ny, g.nz
    inner = np.zeros((nz, ny, nx), bool)
    inner[1:-1, 1:-1, 1:-1] = True                                     # clamped faces use one-sided halves: compare the interior
    inner = inner.reshape(-1)
    for a, b in zip(tan["R"], fd["R"]):
        assert np.max(np.abs(a[inner].astype(np.float64) - b[inner])) <= 1e-3 * np.max(np.abs(b[inner]))   # fp32 noise of the differenced outputs
    # random-init network of the benchmark: same PDE, different discretisation
    w = port.mlp_random_init(64, 777, 0.25)
    errs = []
    for n in (24, 48):
        h = 2.0 / (n - 1)                                              # physical spacing = coordinate spacing: dc/dx = 1
        gg = Grid(n, n, n, h, h, h, 2e-3, False)
        t1 = port.tangent_loss(gg, w, 0.25, True)
        f1 = port.fused_loss(gg, w, 0.25, 2e-3, 1.0, 1.0, True)
        errs.append(abs(t1["acc_u"] - f1["acc_u"]) / f1["acc_u"])
    assert errs[1] < 0.2 and errs[1] < errs[0] + 0.02, errs
