"""TEST INFRASTRUCTURE ONLY -- ctypes loaders for the two CPU checkers.

* ``port()``      -> oracle/_build/liboracle.so, the plain-C restatement in oracle/oracle.c.
* ``reference()`` -> oracle/_ref/libphysref.so, the UNMODIFIED reference CPU sources behind
  oracle/ref_shim.cpp (built in the dev container where /root/reference exists; the prebuilt
  .so travels to the GPU box).  Returns None when it is not there.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` / ``--impl reference`` legs
may import this package.  The product (phys_autodiff_b200) never does; tests/test_boundary.py
checks that by grepping the package sources.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_F = C.POINTER(C.c_float)
_D = C.POINTER(C.c_double)


class CGrid(C.Structure):
    """phys::GridSpec with the bool widened to int (reference include/phys.h:8-13)."""

    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int),
                ("hx", C.c_float), ("hy", C.c_float), ("hz", C.c_float),
                ("dt", C.c_float), ("periodic", C.c_int)]


@dataclass
class Grid:
    nx: int
    ny: int
    nz: int
    hx: float = 1.0
    hy: float = 1.0
    hz: float = 1.0
    dt: float = 1.0
    periodic: bool = True

    @property
    def N(self) -> int:
        return self.nx * self.ny * self.nz

    def c(self) -> CGrid:
        return CGrid(self.nx, self.ny, self.nz, self.hx, self.hy, self.hz, self.dt, int(self.periodic))


def _fp(a: np.ndarray | None):
    if a is None:
        return C.cast(None, _F)
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_F)


def build(ref_too: bool = True) -> None:
    """Compile oracle.c (always) and, when /root/reference is present, oracle/_ref."""
    subprocess.run(["make", "-C", _HERE, "port"], check=True, capture_output=True)
    if ref_too:
        subprocess.run(["make", "-C", _HERE, "ref"], check=True, capture_output=True)


class _Oracle:
    """Common numpy-facing surface over either library (function-name prefix differs)."""

    kind = "port"

    def __init__(self, lib: C.CDLL, prefix: str):
        self.lib, self.p = lib, prefix

    def _fn(self, name):
        return getattr(self.lib, self.p + name)

    # -- weights -------------------------------------------------------------------------
    def mlp_random_init(self, H: int, seed: int = 42, scale: float = 0.5, In: int = 4, Out: int = 4):
        W1 = np.empty(H * In, np.float32); b1 = np.empty(H, np.float32)
        W2 = np.empty(Out * H, np.float32); b2 = np.empty(Out, np.float32)
        f = self._fn("mlp_random_init"); f.restype = None
        f(C.c_int(In), C.c_int(H), C.c_int(Out), C.c_uint(seed), C.c_float(scale), _fp(W1), _fp(b1), _fp(W2), _fp(b2))
        return W1, b1, W2, b2

    # -- coords / MLP --------------------------------------------------------------------
    def make_grid_coords(self, g: Grid, t: float, m1p1: bool = True) -> np.ndarray:
        out = np.empty(g.N * 4, np.float32)
        f = self._fn("make_grid_coords"); f.restype = None
        f(C.byref(g.c()), C.c_float(t), C.c_int(int(m1p1)), _fp(out))
        return out

    def mlp_forward(self, x, W1, b1, W2, b2, B, In, H, Out) -> np.ndarray:
        y = np.empty(B * Out, np.float32)
        f = self._fn("mlp_forward" if self.kind == "port" else "mlp_forward_cpu"); f.restype = None
        f(_fp(x), _fp(W1), _fp(b1), _fp(W2), _fp(b2), _fp(y), C.c_size_t(B), C.c_size_t(In), C.c_size_t(H), C.c_size_t(Out))
        return y

    def mlp_backward(self, x, y_target, W1, b1, W2, b2, B, In, H, Out):
        dW1 = np.empty(H * In, np.float32); db1 = np.empty(H, np.float32)
        dW2 = np.empty(Out * H, np.float32); db2 = np.empty(Out, np.float32)
        f = self._fn("mlp_backward" if self.kind == "port" else "mlp_backward_cpu"); f.restype = None
        f(_fp(x), _fp(y_target), _fp(W1), _fp(b1), _fp(W2), _fp(b2), _fp(dW1), _fp(db1), _fp(dW2), _fp(db2),
          C.c_size_t(B), C.c_size_t(In), C.c_size_t(H), C.c_size_t(Out))
        return dW1, db1, dW2, db2

    def mlp_grid_infer(self, g: Grid, w, t: float, m1p1: bool = True) -> np.ndarray:
        W1, b1, W2, b2 = w
        H = b1.size
        out = np.empty(g.N * 4, np.float32)
        f = self._fn("mlp_grid_infer" if self.kind == "port" else "mlp_grid_infer_cpu"); f.restype = None
        f(C.byref(g.c()), C.c_int(4), C.c_int(H), C.c_int(4), C.c_int(int(m1p1)), _fp(W1), _fp(b1), _fp(W2), _fp(b2),
          C.c_float(t), _fp(out))
        return out

    def generate_fields(self, g: Grid, w, t: float, dt: float, m1p1: bool = True):
        W1, b1, W2, b2 = w
        H = b1.size
        N = g.N
        s = [np.empty(N, np.float32) for _ in range(3)]
        u = [np.empty(3 * N, np.float32) for _ in range(3)]
        f = self._fn("generate_fields" if self.kind == "port" else "mlp_generate_fields_cpu"); f.restype = None
        f(C.byref(g.c()), C.c_int(4), C.c_int(H), C.c_int(4), C.c_int(int(m1p1)), _fp(W1), _fp(b1), _fp(W2), _fp(b2),
          C.c_float(t), C.c_float(dt), _fp(s[0]), _fp(s[1]), _fp(s[2]), _fp(u[0]), _fp(u[1]), _fp(u[2]))
        return s[0], s[1], s[2], u[0], u[1], u[2]

    # -- physics -------------------------------------------------------------------------
    def phys_residuals(self, g: Grid, fields):
        N = g.N
        R = [np.empty(N, np.float32) for _ in range(4)]
        f = self._fn("phys_residuals"); f.restype = None
        f(C.byref(g.c()), *[_fp(a) for a in fields], *[_fp(r) for r in R])
        return tuple(R)

    def phys_loss_forward(self, g: Grid, w_sigma: float, w_u: float, fields, want_residuals: bool = False):
        N = g.N
        ls, lu = C.c_float(), C.c_float()
        R = [np.empty(N, np.float32) for _ in range(4)] if want_residuals else [None] * 4
        f = self._fn("phys_loss_forward"); f.restype = None
        f(C.byref(g.c()), C.c_float(w_sigma), C.c_float(w_u), *[_fp(a) for a in fields], C.byref(ls), C.byref(lu),
          *[_fp(r) for r in R])
        return (np.float32(ls.value), np.float32(lu.value)) + ((tuple(R),) if want_residuals else ())

    def phys_loss_backward(self, g: Grid, w_sigma: float, w_u: float, R):
        N = g.N
        G = [np.empty(N, np.float32) for _ in range(4)]
        f = self._fn("phys_loss_backward"); f.restype = None
        f(C.byref(g.c()), C.c_float(w_sigma), C.c_float(w_u), *[_fp(r) for r in R], *[_fp(x) for x in G])
        return tuple(G)


class PortOracle(_Oracle):
    kind = "port"

    def tangent_loss(self, g: Grid, w, t, m1p1=True, want_residuals=False):
        """Analytic forward-mode physics loss (additive, parity unpinned) -> dict(acc_sigma, acc_u[, R])."""
        W1, b1, W2, b2 = w
        a_s, a_u = C.c_double(), C.c_double()
        R = [np.empty(g.N, np.float32) for _ in range(4)] if want_residuals else [None] * 4
        f = self.lib.oracle_tangent_loss; f.restype = None
        f(C.byref(g.c()), C.c_int(b1.size), C.c_int(int(m1p1)), _fp(W1), _fp(b1), _fp(W2), _fp(b2), C.c_float(t),
          C.byref(a_s), C.byref(a_u), *[_fp(r) for r in R])
        out = dict(acc_sigma=a_s.value, acc_u=a_u.value)
        if want_residuals:
            out["R"] = tuple(R)
        return out

    def phys_residuals_upwind(self, g: Grid, fields):
        """First-order upwind advection (additive switch, parity unpinned: oracle.c is the only statement of it)."""
        N = g.N
        R = [np.empty(N, np.float32) for _ in range(4)]
        f = self.lib.oracle_phys_residuals_upwind; f.restype = None
        f(C.byref(g.c()), *[_fp(a) for a in fields], *[_fp(r) for r in R])
        return tuple(R)

    def fused_loss(self, g: Grid, w, t, dt, w_sigma=1.0, w_u=1.0, m1p1=True, want_residuals=False):
        """Whole path on one thread -> dict(loss_sigma, loss_u, acc_sigma, acc_u[, R])."""
        W1, b1, W2, b2 = w
        H = b1.size
        ls, lu, a_s, a_u = C.c_float(), C.c_float(), C.c_double(), C.c_double()
        R = [np.empty(g.N, np.float32) for _ in range(4)] if want_residuals else [None] * 4
        f = self.lib.oracle_fused_loss; f.restype = C.c_int
        rc = f(C.byref(g.c()), C.c_int(4), C.c_int(H), C.c_int(4), C.c_int(int(m1p1)), _fp(W1), _fp(b1), _fp(W2), _fp(b2),
               C.c_float(t), C.c_float(dt), C.c_float(w_sigma), C.c_float(w_u), C.byref(ls), C.byref(lu),
               C.byref(a_s), C.byref(a_u), *[_fp(r) for r in R])
        assert rc == 0
        out = dict(loss_sigma=np.float32(ls.value), loss_u=np.float32(lu.value), acc_sigma=a_s.value, acc_u=a_u.value)
        if want_residuals:
            out["R"] = tuple(R)
        return out

    def mlp_forward_deep(self, x, L, W1, b1, Wh, bh, W2, b2, B, In, H, Out) -> np.ndarray:
        """Deeper MLP (parity unpinned for L > 1, see oracle.c)."""
        y = np.empty(B * Out, np.float32)
        f = self.lib.oracle_mlp_forward_deep; f.restype = None
        f(_fp(x), C.c_int(L), _fp(W1), _fp(b1), _fp(Wh if Wh is not None and Wh.size else None),
          _fp(bh if bh is not None and bh.size else None), _fp(W2), _fp(b2), _fp(y), C.c_size_t(B), C.c_size_t(In),
          C.c_size_t(H), C.c_size_t(Out))
        return y

    def mlp_grid_infer_deep(self, g: Grid, H, L, W1, b1, Wh, bh, W2, b2, t: float, m1p1: bool = True) -> np.ndarray:
        out = np.empty(g.N * 4, np.float32)
        f = self.lib.oracle_mlp_grid_infer_deep; f.restype = None
        f(C.byref(g.c()), C.c_int(H), C.c_int(L), C.c_int(int(m1p1)), _fp(W1), _fp(b1),
          _fp(Wh if Wh is not None and Wh.size else None), _fp(bh if bh is not None and bh.size else None), _fp(W2), _fp(b2),
          C.c_float(t), _fp(out))
        return out

    def phys_loss_grad(self, g: Grid, w, t, dt, w_sigma=1.0, w_u=1.0, m1p1=True, all_double=False):
        """d(L_sigma + L_u)/d(weights) (oracle_grad.c; parity unpinned) -> dict(loss_sigma, loss_u, grad[9H+4] f64)."""
        W1, b1, W2, b2 = w
        H = b1.size
        ls, lu = C.c_double(), C.c_double()
        grad = np.empty(9 * H + 4, np.float64)
        f = self.lib.oracle_phys_loss_grad; f.restype = C.c_int
        rc = f(C.byref(g.c()), C.c_int(H), C.c_int(int(m1p1)), _fp(W1), _fp(b1), _fp(W2), _fp(b2), C.c_float(t),
               C.c_float(dt), C.c_float(w_sigma), C.c_float(w_u), C.c_int(int(all_double)), C.byref(ls), C.byref(lu),
               grad.ctypes.data_as(C.POINTER(C.c_double)))
        assert rc == 0
        return dict(loss_sigma=ls.value, loss_u=lu.value, grad=grad)

    def phys_loss_double(self, g: Grid, theta: np.ndarray, H: int, t, dt, w_sigma=1.0, w_u=1.0, m1p1=True):
        """All-double L_sigma, L_u of double weights theta[9H+4] (the function finite differences probe)."""
        theta = np.ascontiguousarray(theta, np.float64)
        assert theta.size == 9 * H + 4
        ls, lu = C.c_double(), C.c_double()
        f = self.lib.oracle_phys_loss_double; f.restype = C.c_int
        rc = f(C.byref(g.c()), C.c_int(H), C.c_int(int(m1p1)), theta.ctypes.data_as(C.POINTER(C.c_double)), C.c_float(t),
               C.c_float(dt), C.c_float(w_sigma), C.c_float(w_u), C.byref(ls), C.byref(lu))
        assert rc == 0
        return ls.value, lu.value

    def sumsq(self, R, i0: int, i1: int):
        a_s, a_u = C.c_double(), C.c_double()
        f = self.lib.oracle_sumsq; f.restype = None
        f(*[_fp(r) for r in R], C.c_size_t(i0), C.c_size_t(i1), C.byref(a_s), C.byref(a_u))
        return a_s.value, a_u.value


class RefOracle(_Oracle):
    kind = "reference"

    def hardware_threads(self) -> int:
        f = self.lib.ref_hardware_threads; f.restype = C.c_int
        return int(f())

    def fused_loss(self, g: Grid, w, t, dt, w_sigma=1.0, w_u=1.0, m1p1=True, want_residuals=False, threads=1):
        """Reference mlp_infer_cpu on `threads` host threads + reference cpu_phys_loss_forward."""
        W1, b1, W2, b2 = w
        H = b1.size
        ls, lu = C.c_float(), C.c_float()
        R = [np.empty(g.N, np.float32) for _ in range(4)] if want_residuals else [None] * 4
        f = self.lib.ref_fused_loss_mt; f.restype = C.c_int
        rc = f(C.byref(g.c()), C.c_int(4), C.c_int(H), C.c_int(4), C.c_int(int(m1p1)), _fp(W1), _fp(b1), _fp(W2), _fp(b2),
               C.c_float(t), C.c_float(dt), C.c_float(w_sigma), C.c_float(w_u), C.c_int(threads), C.byref(ls),
               C.byref(lu), *[_fp(r) for r in R])
        assert rc == 0
        out = dict(loss_sigma=np.float32(ls.value), loss_u=np.float32(lu.value))
        if want_residuals:
            out["R"] = tuple(R)
        return out


_PORT = None
_REF = None


def port() -> PortOracle:
    global _PORT
    if _PORT is None:
        path = os.path.join(_HERE, "_build", "liboracle.so")
        if not os.path.exists(path) or os.path.getmtime(path) < max(
                os.path.getmtime(os.path.join(_HERE, f)) for f in ("oracle.c", "oracle_grad.c")):
            build(ref_too=False)
        _PORT = PortOracle(C.CDLL(path), "oracle_")
    return _PORT


def reference() -> RefOracle | None:
    global _REF
    if _REF is None:
        path = os.path.join(_HERE, "_ref", "libphysref.so")
        if not os.path.exists(path):
            return None
        _REF = RefOracle(C.CDLL(path), "ref_")
    return _REF
