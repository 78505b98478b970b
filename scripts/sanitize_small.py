"""Tiny end-to-end run for compute-sanitizer (memcheck / racecheck): every kernel family once, small lattices."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wafer_b200  # noqa: E402

rng = np.random.default_rng(0)
for cd, ext, shape in (("ThreePoint", 1, (37, 33, 70)), ("FivePoint", 2, (12, 9, 10)), ("SevenPoint", 3, (9, 8, 7))):
    P = tuple(s + 2 * ext for s in shape)
    v = rng.normal(size=P)
    phi = np.zeros(P)
    phi[ext:-ext, ext:-ext, ext:-ext] = rng.normal(size=shape)
    with wafer_b200.Lattice(shape, cd, dn=0.1, dt=2e-3, mass=1.0) as lat:
        lat.set_potential(v)
        lat.set_phi(phi)
        lat.evolve(0, 5)
        lat.check(0)
        lat.push_lower()
        lat.phi_seed_from_lower(0)
        lat.check(1)
        lat.evolve(1, 2)
        lat.get_phi()
        lat.generate_potential("Harmonic")
        lat.set_initial_conditions("Boolean")
        lat.solve(0, 1e-3, max_steps=20, screen_update=4)
        if ext == 1:
            # round 2: a multi-segment persistent launch (several tile columns per CTA would need a big lattice; a tall thin
            # one gives every CTA bulk + tail segments), the fused check sums (>= 64 steps), more stored states than one pass
            lat.evolve(0, 66)
            lat.check(0)
            for _ in range(4):
                lat.push_lower()
            lat.phi_seed_from_lower(0)
            lat.evolve(5, 2)
            lat.check(5)
            lat.phi_checksum()
            lat.set_phi_owned(lat.get_phi_slab())
# many tile columns on few planes: the in-order dispatcher hands out bulk items and static tail shares
with wafer_b200.Lattice((24, 400, 500), "ThreePoint", dn=0.1, dt=2e-3, mass=1.0) as lat:
    lat.generate_potential("Harmonic")
    lat.set_initial_conditions("Boolean")
    lat.evolve(0, 4)
    lat.check(0)
print("sanitize run complete")
