"""One pass over every hot kernel at 512^3 for `ncu --set full` (scripts/gpu_round.sh stage ncu2): ground-state sweeps
(time-tiled + the fused-check tail), a check, excited-state steps with 1 and 3 stored states, 5- and 7-point sweeps."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wafer_b200  # noqa: E402

n = int(os.environ.get("N", "512"))
dn = 10.24 / n
for cd, steps in (("ThreePoint", 70), ("FivePoint", 2), ("SevenPoint", 2)):
    with wafer_b200.Lattice((n,) * 3, cd, dn=dn, dt=0.1 * dn * dn, mass=1.0) as lat:
        lat.generate_potential("Harmonic")
        lat.set_initial_conditions("Boolean")
        lat.check(0)
        lat.evolve(0, steps)       # ThreePoint: 34 time-tiled launches + 2 one-step launches, the last with the check sums
        lat.check(0)               # energy-only observables pass + normalise
        if cd == "ThreePoint":
            lat.set_pot_sub(2.0)
            lat.compute_observables()  # full observables pass (pot_sub scalar)
            for k in (1, 2, 3):
                lat.push_lower()
                lat.phi_seed_from_lower(0)
                lat.check(k)
                lat.evolve(k, 2)
        lat.synchronize()
print("ncu workload done")
