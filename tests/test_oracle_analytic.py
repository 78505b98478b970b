"""Analytic known answers for the two functions the reference leaves untested (evolve,
compute_observables) — SURVEY.md §8(c).  They anchor the oracle on physics, not on itself.  CPU only."""
import numpy as np
import pytest

import np_restatement as npr


def _box_mode(g, n):
    """psi = prod_a sin(pi n_a i_a/(N_a+1)), i_a = padded index (ext=1): exact 3-pt eigenvector."""
    px, py, pz = g.padded_shape
    i = np.arange(px)[:, None, None]
    j = np.arange(py)[None, :, None]
    k = np.arange(pz)[None, None, :]
    return np.ascontiguousarray(np.sin(np.pi * n[0] * i / (g.nx + 1)) * np.sin(np.pi * n[1] * j / (g.ny + 1)) *
                                np.sin(np.pi * n[2] * k / (g.nz + 1)))


@pytest.mark.parametrize("mode", [(1, 1, 1), (2, 1, 3)])
def test_box_mode_energy_and_decay(oracle, mode):
    g = oracle.make_grid(50, 50, 50, ext=1, dn=0.01, dt=3e-5, mass=15.9994)
    v = oracle.potential(g, "NoPotential")
    a, b = oracle.build_ab(v, g.dt)
    assert np.all(a == 1.0) and np.all(b == 1.0)
    psi = _box_mode(g, mode)
    psi[0], psi[-1], psi[:, 0], psi[:, -1], psi[:, :, 0], psi[:, :, -1] = 0, 0, 0, 0, 0, 0
    e_exact = sum(2 - 2 * np.cos(np.pi * m / 51) for m in mode) / (2 * g.mass * g.dn ** 2)
    if mode == (1, 1, 1):
        assert e_exact == pytest.approx(3.5563919827, rel=1e-10)
    oracle.set_sum_mode(1)
    obs = oracle.observables(g, psi, v)
    oracle.set_sum_mode(0)
    assert obs["energy"] / obs["norm2"] == pytest.approx(e_exact, rel=1e-12)
    before = psi.copy()
    oracle.evolve(g, psi, a, b, 1)
    w0, w1 = npr.work(before, 1), npr.work(psi, 1)
    big = np.abs(w0) > 1e-3
    assert np.allclose(w1[big] / w0[big], 1 - g.dt * e_exact, rtol=1e-11, atol=0)


def test_harmonic_levels_three_point(oracle):
    """E = w(n+3/2) - (dn^2 m w^2/32) sum_a(2n_a^2+2n_a+1) + O(dn^4), w = 1/sqrt(m); ground state via solve."""
    n, dn, mass = 32, 0.4, 1.0
    dt = dn * dn / 4
    g = oracle.make_grid(n, n, n, ext=1, dn=dn, dt=dt, mass=mass)
    v = oracle.potential(g, "Harmonic")
    a, b = oracle.build_ab(v, dt)
    phi = oracle.initial_condition(g, "Constant")
    conv, rec = oracle.solve(g, v, a, b, phi, tolerance=1e-12, screen_update=200)
    assert conv
    e0 = rec[-1]["E"]
    formula = 1.5 - dn * dn / 32 * 3
    # imaginary-time fixed point of the Strickland update carries an O(dt) shift on top of O(dn^4)
    assert e0 == pytest.approx(formula, abs=3e-3)
    # first excited shell from a deterministic odd seed
    lowers = [phi.copy()]
    seed = phi * (np.arange(g.padded_shape[0])[:, None, None] - (n + 1) / 2)
    conv, rec1 = oracle.solve(g, v, a, b, seed, lowers=lowers, tolerance=1e-12, screen_update=200)
    assert conv
    formula1 = 2.5 - dn * dn / 32 * (5 + 1 + 1)
    assert rec1[-1]["E"] == pytest.approx(formula1, abs=5e-3)
    assert abs((seed * lowers[0]).sum()) < 1e-12


def test_simple_cornell_binding_energy(oracle):
    """potential.rs:360: pot_sub = 4 m  =>  binding = (energy - v_inf)/norm2 = E - 4m exactly (output.rs:544)"""
    g = oracle.make_grid(16, 16, 16, ext=1, dn=0.2, dt=0.01, mass=0.75)
    v = oracle.potential(g, "SimpleCornell", sig=0.223)
    phi = oracle.initial_condition(g, "Boolean")
    ps = oracle.potential_sub(g, "SimpleCornell")
    obs = oracle.observables(g, phi, v, ps)
    assert (obs["energy"] - obs["v_infinity"]) / obs["norm2"] == pytest.approx(obs["energy"] / obs["norm2"] - 3.0,
                                                                               rel=1e-13)


def test_solve_max_steps_semantics(oracle):
    """grid.rs:211-213: strict '>' => evolve runs while step <= max_steps: (floor(S/u)+1)*u steps in total."""
    g = oracle.make_grid(8, 8, 8, ext=1, dn=0.1, dt=1e-3, mass=1.0)
    v = oracle.potential(g, "Harmonic")
    a, b = oracle.build_ab(v, g.dt)
    phi = oracle.initial_condition(g, "Boolean")
    conv, rec = oracle.solve(g, v, a, b, phi, tolerance=1e-300, max_steps=25, screen_update=10)
    assert not conv
    assert [r["step"] for r in rec] == [0, 10, 20, 30]
    ref = oracle.initial_condition(g, "Boolean")
    for _ in range(3):
        o = npr.observables(ref, v, 1, 0.1, 1.0)
        ref = npr.normalise(ref, oracle.observables(g, ref, v)["norm2"])
        for _ in range(10):
            ref = npr.sweep(ref, a, b, 1, 0.1, 1e-3, 1.0)
    ref = npr.normalise(ref, oracle.observables(g, ref, v)["norm2"])
    assert np.array_equal(ref, phi)


def test_snapshot_double_normalise_quirk(oracle):
    """grid.rs:137-139: with snap_update set the state is divided by sqrt(norm2) twice at snapshot checks."""
    g = oracle.make_grid(6, 6, 6, ext=1, dn=0.1, dt=1e-3, mass=1.0)
    v = oracle.potential(g, "Harmonic")
    a, b = oracle.build_ab(v, g.dt)
    p1 = oracle.initial_condition(g, "Constant")
    p2 = p1.copy()
    n2 = oracle.observables(g, p1, v)["norm2"]
    oracle.solve(g, v, a, b, p1, tolerance=float('inf'), screen_update=5)
    oracle.solve(g, v, a, b, p2, tolerance=float('inf'), screen_update=5, snap_update=5)
    assert np.array_equal(p2, p1 / np.sqrt(n2))


def test_sweep_is_linear(oracle):
    rng = np.random.default_rng(0)
    g = oracle.make_grid(10, 8, 12, ext=2, dn=0.1, dt=2e-3, mass=1.0)
    v = oracle.potential(g, "Harmonic")
    a, b = oracle.build_ab(v, g.dt)
    x = np.zeros(g.padded_shape)
    y = np.zeros(g.padded_shape)
    npr.work(x, 2)[...] = rng.normal(size=g.work_shape)
    npr.work(y, 2)[...] = rng.normal(size=g.work_shape)
    z = 2.0 * x - 0.5 * y
    for arr in (x, y, z):
        oracle.evolve(g, arr, a, b, 3)
    assert np.allclose(z, 2.0 * x - 0.5 * y, rtol=0, atol=1e-13)


def test_default_config_golden_records(oracle):
    """tests/golden/c1_default_records.json (made by tests/golden/make_golden.py): BASELINE config C1 ground state."""
    import hashlib
    import json
    import os
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "c1_default_records.json")))
    g = oracle.make_grid(50, 50, 50, ext=1, dn=0.01, dt=3e-5, mass=15.9994)
    v = oracle.potential(g, "Harmonic")
    a, b = oracle.build_ab(v, g.dt)
    phi = oracle.initial_condition(g, "Boolean")
    conv, rec = oracle.solve(g, v, a, b, phi, tolerance=1e-4, screen_update=1000)
    assert conv and len(rec) == len(gold["state0"]["records"]) == 19
    for r, gr in zip(rec, gold["state0"]["records"]):
        assert r["step"] == gr["step"] and r["E"] == pytest.approx(gr["E"], rel=1e-13)
    assert rec[-1]["step"] == 18000 and rec[-1]["E"] == pytest.approx(3.56925, abs=1e-5)
    assert np.sqrt(rec[-1]["r2"] / rec[-1]["norm2"]) == pytest.approx(16.09, abs=5e-3)
    if oracle.num_threads() == gold.get("threads", oracle.num_threads()):
        # sums are per-plane, so the state is independent of the thread count: the hash must match
        assert hashlib.sha256(phi.tobytes()).hexdigest() == gold["state0"]["phi_sha256"]
