#!/usr/bin/env python
"""Test stand-in for a user potential script (protocol of the reference's gen_potential.py / input.rs:186-248):
JSON {"grid": {"x","y","z","dn"}} on stdin, one float per line on stdout, work area in x-major / z-fastest order.
Prints the harmonic potential V = dn^2 r^2 / 2 about the index-space centre used by potential.rs:270-274."""
import json
import math
import sys

g = json.load(sys.stdin)["grid"]
nx, ny, nz, dn = g["x"], g["y"], g["z"], g["dn"]
out = []
for i in range(nx):
    dx = (i + 1) - (nx + 1.0) / 2.0  # work index i sits at padded index i + 1 (ThreePoint)
    for j in range(ny):
        dy = (j + 1) - (ny + 1.0) / 2.0
        for k in range(nz):
            dz = (k + 1) - (nz + 1.0) / 2.0
            r = dn * math.sqrt(dx * dx + dy * dy + dz * dz)
            out.append(repr(r * r / 2.0))
sys.stdout.write("\n".join(out) + "\n")
