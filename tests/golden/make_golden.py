#!/usr/bin/env python
"""Regenerates tests/golden/c1_default_records.json: the per-check records of BASELINE config C1 (the reference's
default wafer.yaml: 50^3 Harmonic, ThreePoint, Boolean IC, tol 1e-4, check every 1000 steps) as computed by the CPU
oracle, plus the excited state started from the driver's deterministic seed.

Provenance: the reference itself cannot be executed here (Rust, no toolchain), so this fixture is produced by
oracle/wafer_oracle.cpp — whose sweep agrees bit-for-bit with the independent numpy restatement and whose results
match the survey's scratch numbers (converges at step 18000, E0 = 3.56925, r_rms = 16.09).  It pins the oracle against
silent drift and gives the GPU tests a stored target.

usage: python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import binding as oracle  # noqa: E402


def main():
    g = oracle.make_grid(50, 50, 50, ext=1, dn=0.01, dt=3e-5, mass=15.9994)
    v = oracle.potential(g, "Harmonic")
    a, b = oracle.build_ab(v, g.dt)
    phi = oracle.initial_condition(g, "Boolean")
    conv, rec = oracle.solve(g, v, a, b, phi, tolerance=1e-4, screen_update=1000)
    p1 = oracle.seed_from_state(g, phi)
    conv1, rec1 = oracle.solve(g, v, a, b, p1, lowers=[phi], tolerance=1e-4, screen_update=1000, max_records=400)
    out = {
        "config": "tests/golden/wafer_default.yaml",
        "state0": {"converged": conv, "records": [{k: r[k] for k in ("step", "tau", "E", "norm2", "r2")} for r in rec],
                   "phi_sha256": hashlib.sha256(phi.tobytes()).hexdigest()},
        "state1_seeded": {"converged": conv1, "n_records": len(rec1), "final_step": rec1[-1]["step"], "E": rec1[-1]["E"]},
    }
    with open(os.path.join(ROOT, "tests", "golden", "c1_default_records.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("state 0: %d checks, E0 = %.12f; state 1: %d checks, E1 = %.12f" % (len(rec), rec[-1]["E"], len(rec1), rec1[-1]["E"]))


if __name__ == "__main__":
    main()
