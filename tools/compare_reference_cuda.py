#!/usr/bin/env python
"""Run the reference's own CSV benchmark programs twice on this GPU -- once linked against the REFERENCE'S CUDA
sources (tests/refprogs/_bin/refcuda_*, compiled unmodified for sm_100a) and once against this repository's
library (tests/refprogs/_bin/test_*) -- and print both tables side by side as JSON.
Same harness, same host-pointer API, same grids (docs/BENCHMARK_REPORT.md shapes); only the CUDA backend differs."""
import csv, io, json, os, subprocess, sys

BIN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "refprogs", "_bin")


def run(name):
    exe = os.path.join(BIN, name)
    if not os.path.exists(exe):
        return None
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    if r.returncode != 0:
        return {"error": r.stderr[-500:]}
    rows = list(csv.DictReader(io.StringIO(r.stdout)))
    return rows


def main():
    out = {}
    for prog in ("test_mlp_phys_perf", "test_phys_perf"):
        ref, ours = run("refcuda_" + prog), run(prog)
        out[prog] = {"reference_cuda_sm100a": ref, "this_repo": ours}
        if isinstance(ref, list) and isinstance(ours, list):
            cmp_rows = []
            for a, b in zip(ref, ours):
                row = {k: a[k] for k in a if k in ("mode", "nx", "ny", "nz")}
                for k in a:
                    if k.startswith("ms"):
                        row[k] = {"reference": float(a[k]), "ours": float(b[k]), "speedup": float(a[k]) / max(float(b[k]), 1e-12)}
                cmp_rows.append(row)
            out[prog]["comparison"] = cmp_rows
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
