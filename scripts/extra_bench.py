"""Secondary measurements on one GPU (not the headline bench): 5/7-point sweeps, excited-state steps, checks.
Prints one JSON object.  Device-timed with the library's CUDA events."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wafer_b200  # noqa: E402


def timed(lat, fn, reps):
    fn()
    lat.synchronize()
    lat.timer_begin()
    for _ in range(reps):
        fn()
    return lat.timer_end() / reps


def main():
    out = {}
    n = int(os.environ.get("N", "512"))
    dn = 10.24 / n
    for cd in ("ThreePoint", "FivePoint", "SevenPoint"):
        for flags, tag in ((0, "default"), (4, "simple"), (1, "ab_arrays"), (12 if cd == "ThreePoint" else 8, "tma1")):
            if cd != "ThreePoint" and flags == 4:
                continue
            if tag == "tma1" and os.environ.get("NO_TMA1"):
                continue
            with wafer_b200.Lattice((n,) * 3, cd, dn=dn, dt=0.1 * dn * dn, mass=1.0, flags=flags) as lat:
                lat.generate_potential("Harmonic")
                lat.set_initial_conditions("Boolean")
                lat.check(0)
                ms = timed(lat, lambda: lat.evolve(0, 20), 5)
                out["%s_%s_glups_%d" % (cd, tag, n)] = n ** 3 * 20 / ms / 1e6
                if flags == 0:
                    out["%s_check_ms_%d" % (cd, n)] = timed(lat, lambda: lat.check(0), 5)
    m = int(os.environ.get("M", "256"))
    dm = 10.24 / m
    for flags, tag in ((0, ""), (8, "_tma1")):
        if flags and os.environ.get("NO_TMA1"):
            continue
        with wafer_b200.Lattice((m,) * 3, "ThreePoint", dn=dm, dt=0.1 * dm * dm, mass=1.0, flags=flags) as lat:
            lat.generate_potential("Harmonic")
            lat.set_initial_conditions("Boolean")
            lat.check(0)
            for k in (1, 2, 3):
                lat.push_lower()  # any stored state will do for timing
                lat.phi_seed_from_lower(0)
                lat.check(k)
                ms = timed(lat, lambda: lat.evolve(k, 10), 5)
                out["excited_k%d%s_glups_%d" % (k, tag, m)] = m ** 3 * 10 / ms / 1e6
                if not flags:
                    out["excited_k%d_GBps_algorithmic_%d" % (k, m)] = (48 + 16 * k) * m ** 3 * 10 / ms / 1e6
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
