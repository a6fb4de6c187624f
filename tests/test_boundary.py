"""CPU tier: the drop-in boundary.  The C-ABI library loads without a GPU, exports every symbol
include/physad_b200.h declares (and the reference's C++ symbols), reports
errors instead of computing on the CPU, and the product never touches oracle/."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "physad_b200.h")


@pytest.fixture(scope="module")
def libpath():
    from phys_autodiff_b200 import capi
    p = capi.library_path()
    if not os.path.exists(p):
        capi.build_library()
    return p


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(physad_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported(libpath):
    from phys_autodiff_b200 import capi
    declared = _declared_symbols()
    assert len(declared) >= 25
    assert sorted(capi.EXPORTS) == declared, "capi.EXPORTS out of sync with include/physad_b200.h"
    lib = C.CDLL(libpath)
    for s in declared:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    assert lib.physad_abi_version() == 1


def test_reference_cxx_symbols_exported(libpath):
    """Itanium-mangled names a replacement CUDA backend must define."""
    out = subprocess.run(["nm", "-D", "--defined-only", libpath], capture_output=True, text=True, check=True).stdout
    need = [
        "_Z11mlp_forwardI8ExecCudaEvPKfS2_S2_S2_S2_Pfmmmm",
        "_Z12mlp_backwardI8ExecCudaEvPKfS2_S2_S2_S2_S2_PfS3_S3_S3_mmmm",
        "cuda_phys_residuals_nonfusedE", "cuda_phys_residuals_nonfused_timedE", "cuda_phys_loss_forward_nonfusedE",
        "cuda_phys_loss_backward_nonfusedE", "cuda_phys_residuals_fusedE", "cuda_phys_residuals_fused_timedE",
        "cuda_phys_loss_backward_fusedE", "mlp_infer_cudaE", "mlp_grid_infer_cudaE", "mlp_generate_fields_cudaE",
        "mlp_random_initE", "make_grid_coordsE", "cuda_phys_loss_forward_fusedE", "mlp_phys_loss_fused_cudaE", "mlp_phys_loss_grad_cudaE",
    ]
    for n in need:
        assert n in out, n


def test_no_cpu_fallback_without_device(libpath):
    """Here (no GPU) context creation must fail with a CUDA status, not fall back to anything."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = C.CDLL(libpath)
    lib.physad_last_error.restype = C.c_char_p
    h = C.c_void_p()
    rc = lib.physad_ctx_create(C.byref(h), C.c_int(-1))
    assert rc != 0 and not h.value
    assert len(lib.physad_last_error()) > 0


def test_host_weight_init_matches_reference_stream(libpath, golden):
    """physad_mlp_random_init is host-only code, so it can be pinned on the CPU tier."""
    import numpy as np
    from phys_autodiff_b200 import ops
    arr, meta = golden
    for c in meta["cases"]:
        if c["kind"] == "weights":
            w = ops.mlp_random_init(c["H"], c["seed"], c["scale"])
            for name, a in zip(["W1", "b1", "W2", "b2"], w):
                assert np.array_equal(a, arr[f"{c['key']}_{name}"])


def test_finalize_loss_is_reference_formula(libpath):
    """L = float(w * acc * (1.0/N)) (reference src/phys_cpu.cpp:146-148); pure host arithmetic."""
    import numpy as np
    lib = C.CDLL(libpath)
    lib.physad_finalize_loss.restype = None

    class W(C.Structure):
        _fields_ = [("a", C.c_float), ("b", C.c_float)]
    acc = (C.c_double * 2)(2670.937891475786, 7551.66)
    ls, lu = C.c_float(), C.c_float()
    lib.physad_finalize_loss(acc, C.byref(W(1.7, 0.9)), C.c_size_t(262144), C.byref(ls), C.byref(lu))
    assert ls.value == float(np.float32(np.float64(np.float32(1.7)) * 2670.937891475786 * (1.0 / 262144)))
    assert lu.value == float(np.float32(np.float64(np.float32(0.9)) * 7551.66 * (1.0 / 262144)))


def test_product_never_uses_the_oracle():
    """Only tests/, smoke() and bench.py's cpu_baseline/reference legs may touch oracle/."""
    pkg = os.path.join(ROOT, "phys_autodiff_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                for bad in ("import oracle", "from oracle", "liboracle", "libphysref", "oracle/"):
                    assert bad not in txt, (f, bad)
    out = subprocess.run(["ldd", os.path.join(pkg, "libphysad_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out and "physref" not in out


def test_strict_kernels_have_no_contracted_fma_in_mlp(libpath):
    """SASS check: the MLP grid kernel (pure MLP, no stencil) must contain FMUL/FADD and no FFMA
    (FFMA contraction breaks residual parity)."""
    r = subprocess.run(["cuobjdump", "-sass", libpath], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    cur, counts = None, {}
    for line in r.stdout.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = {"FFMA": 0, "FMUL": 0, "FADD": 0, "FFMA2": 0, "FMUL2": 0, "FADD2": 0,
                           "UTCHMMA": 0, "LDTM": 0, "STTM": 0, "UBLKCP": 0}
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            if op in counts[cur]:
                counts[cur][op] += 1
    # packed layer 2: FMUL2 and FADD2 must stay separate instructions in every FORWARD kernel that has them.  The
    # closed-loop backward kernel (k_phys_grad, no reference counterpart) may fuse everything except the recomputed
    # hidden pre-activation: per point that is 3 strict packed products + the first term of the two W2^T A sums.
    # The analytic tangent kernel (k_tangent_loss) is additive and explicitly NOT a parity path: its Jacobian
    # propagation fuses; only its pre-activation stays strict.
    may_fuse = ("k_phys_grad", "k_tangent_loss", "k_mlp_deep_tc")
    assert all(v["FFMA2"] == 0 for k, v in counts.items() if not any(n in k for n in may_fuse)), \
        {k: v for k, v in counts.items() if v["FFMA2"] and not any(n in k for n in may_fuse)}
    # the tensor-core fast mode of the deep MLP (explicitly not bit-exact): its layer 1 must still be the strict one,
    # its only fused multiply-adds are the packed 2 x 16 of each unrolled output-layer chunk (one or two chunks per thread)
    tcs = {k: v for k, v in counts.items() if "k_mlp_deep_tc" in k}
    assert len(tcs) == 10 and all(v["FMUL2"] >= 3 and v["FADD2"] >= v["FMUL2"] and v["FFMA"] == 0 and v["FFMA2"] in (32, 64)
                                 for v in tcs.values()), tcs
    # ... and they really are tcgen05 kernels: tensor-core MMAs with tensor-memory operands (UTCHMMA), tensor-memory loads and
    # stores in the epilogue (LDTM / STTM), bulk copies of the weight images (UBLKCP); nothing else in the library uses them
    assert all(v["UTCHMMA"] >= 24 and v["LDTM"] >= 1 and v["STTM"] >= 6 and v["UBLKCP"] >= 1 for v in tcs.values()), tcs
    assert all(v["UTCHMMA"] == 0 and v["LDTM"] == 0 for k, v in counts.items() if "k_mlp_deep_tc" not in k)
    tang = {k: v for k, v in counts.items() if "k_tangent_loss" in k}
    assert len(tang) == 3 and all(v["FMUL2"] > 0 and v["FADD2"] > 0 for v in tang.values()), tang
    gradk = {k: v for k, v in counts.items() if "k_phys_grad" in k}
    assert len(gradk) == 3 and all(v["FMUL2"] >= 5 and v["FADD2"] >= 6 and v["FFMA2"] > 0 for v in gradk.values()), gradk
    fused = {k: v for k, v in counts.items() if "k_fused_mlp_phys_loss" in k}
    assert fused and all(v["FMUL"] + v.get("FMUL2", 0) > 0 for v in fused.values())
    grid = {k: v for k, v in counts.items() if "k_mlp_grid" in k or "k_mlp_forward_4x4" in k or "k_strict_gemm" in k
            or ("k_mlp_deep" in k and "k_mlp_deep_tc" not in k)}
    assert grid and any("k_strict_gemm" in k for k in grid) and any("k_mlp_deep" in k for k in grid), "MLP kernels not found in SASS"
    for k, v in grid.items():
        assert v["FMUL"] + v["FMUL2"] > 0 and v["FADD"] + v["FADD2"] > 0, (k, v)
        assert v["FADD2"] >= v["FMUL2"], (k, v)  # every packed product is followed by its own packed add
        # the only FFMAs allowed are the Newton steps of the three IEEE coordinate divisions
        # (__fdiv_rn, 9 each) in the grid kernels; a contracted MLP loop would add one per MAC
        limit = 27 if "k_mlp_grid" in k else 0
        assert v["FFMA"] <= limit and v["FFMA2"] == 0, (k, v)


def test_fused_work_partition_properties(libpath):
    """Host logic of the persistent launch (capi.cu: balanced_ranges): contiguous cover of the tile-plane
    sequence, no more blocks than slots, and near-equal cost (planes + 0.9 per started z-segment)."""
    lib = C.CDLL(libpath)
    for tiles, planes, slots in [(32, 256, 148), (64, 256, 296), (32, 32, 148), (64, 32, 148), (4, 64, 296),
                                 (1, 1, 148), (1, 5, 3), (7, 3, 2), (128, 16, 148), (3, 1000, 296)]:
        out = (C.c_int * (slots + 2))()
        n = lib.physad_plan_ranges(tiles, planes, slots, out, slots + 2)
        assert n >= 2, (tiles, planes, slots, n)
        r = list(out[:n])
        assert r[0] == 0 and r[-1] == tiles * planes
        assert all(b > a for a, b in zip(r, r[1:]))          # non-empty, increasing
        assert n - 1 <= min(slots, tiles * planes)
        costs = []
        for a, b in zip(r, r[1:]):
            segs = (b - 1) // planes - a // planes + 1        # tiles touched by [a, b)
            costs.append((b - a) + 0.9 * segs)
        ideal = (tiles * planes + 0.9 * max(tiles, n - 1)) / (n - 1)
        assert max(costs) <= ideal + 2.9, (tiles, planes, slots, max(costs), ideal)  # one plane + one extra segment of slack


@pytest.mark.parametrize("H", [32, 64, 128])
def test_tensor_core_layer_image_is_an_exact_three_term_split_in_the_documented_layout(libpath, H):
    """physad_deep_tc_pack_layer (host code, no GPU): the three bf16 terms add up to the fp32 weight to 2^-24 and sit where
    include/physad_b200.h says -- the layout the MMA's shared-memory descriptor (SBO 128 B, LBO 16 H B) walks."""
    import numpy as np
    lib = C.CDLL(libpath)
    lib.physad_deep_tc_layer_bytes.restype = C.c_size_t
    assert lib.physad_deep_tc_layer_bytes(H) == 3 * H * H * 2 and lib.physad_deep_tc_layer_bytes(48) == 0
    rng = np.random.default_rng(H)
    W = rng.uniform(-0.2, 0.2, (H, H)).astype(np.float32)
    W[0, 0], W[1, 1], W[2, 2] = 0.0, 1.0, np.float32(-3.1415927)
    img = np.zeros(3 * H * H, dtype=np.uint16)
    assert lib.physad_deep_tc_pack_layer(H, W.ctypes.data_as(C.c_void_p), img.ctypes.data_as(C.c_void_p)) == 0
    g, h = np.meshgrid(np.arange(H), np.arange(H), indexing="ij")
    off = (h // 8) * (H // 8) * 64 + (g // 8) * 64 + (g % 8) * 8 + (h % 8)
    assert sorted(off.ravel().tolist()) == list(range(H * H))                 # a permutation of the tile
    terms = [(img[p * H * H + off].astype(np.uint32) << 16).view(np.float32).astype(np.float64) for p in range(3)]
    total = terms[0] + terms[1] + terms[2]
    assert np.all(np.abs(total - W.astype(np.float64)) <= 2.0 ** -24 * np.abs(W) + 1e-45)
    assert np.all(np.abs(terms[1]) <= 2.0 ** -8 * np.abs(terms[0]) + 1e-45) and np.all(np.abs(terms[2]) <= 2.0 ** -16 * np.abs(terms[0]) + 1e-45)
    assert lib.physad_deep_tc_pack_layer(48, W.ctypes.data_as(C.c_void_p), img.ctypes.data_as(C.c_void_p)) != 0
