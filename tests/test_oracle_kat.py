"""Pins the CPU oracle against every known-answer test the reference holds for the hot path
(SURVEY.md §4 / §8c), and against an independent numpy restatement.  CPU only."""
import numpy as np
import pytest

import np_restatement as npr


def _ijk(shape, f):
    return np.fromfunction(lambda i, j, k: f(i, j, k), shape, dtype=np.float64)


# ---- reference inline tests -------------------------------------------------------------------
def test_gram_schmidt_kat(oracle):
    """grid.rs:721-746"""
    ground = _ijk((2, 2, 2), lambda i, j, k: i + j + k)
    test = _ijk((2, 2, 2), lambda i, j, k: -i - j - k)
    oracle.orthogonalise(test, [ground])
    expect = np.array([0., 23., 23., 46., 23., 46., 46., 69.]).reshape(2, 2, 2)
    assert np.allclose(test, expect, atol=0.01, rtol=0)
    assert np.array_equal(test, expect)  # all values are small integers: exact


def test_work_area_kat(oracle):
    """grid.rs:748-756 and 758-778"""
    g = oracle.make_grid(3, 6, 5, ext=1)
    assert g.padded_shape == (5, 8, 7)
    test = np.zeros((5, 8, 7))
    w = oracle.get_work_area(g, test)
    assert w.shape == (3, 6, 5)
    oracle.set_work_area(g, test, np.ones((3, 6, 5)))
    i, j, k = np.indices((5, 8, 7))
    ring = (i == 0) | (i == 4) | (j == 0) | (j == 7) | (k == 0) | (k == 6)
    assert np.array_equal(test, np.where(ring, 0.0, 1.0))


def test_norm2_kat(oracle):
    """grid.rs:780-786: sum over the ext=1 work area of (i*j*k)^2 on (5,8,7) = 70070"""
    g = oracle.make_grid(3, 6, 5, ext=1)
    test = _ijk((5, 8, 7), lambda i, j, k: i * j * k)
    assert abs(oracle.norm2_work(g, test) - 70070.0) < 1e-6
    assert abs(oracle.norm2_flat(oracle.get_work_area(g, test)) - 70070.0) < 1e-6


def test_wfn_normalise_kat(oracle):
    """grid.rs:788-799"""
    test = _ijk((3, 2, 5), lambda i, j, k: i * j * k)
    expect = test / 1.1091
    oracle.normalise(test, 1.23)
    assert np.allclose(test, expect, atol=0.01, rtol=0)
    assert np.array_equal(test, _ijk((3, 2, 5), lambda i, j, k: i * j * k) / np.sqrt(1.23))


def test_distance_squared_kat(oracle):
    """potential.rs:434-443"""
    assert abs(oracle.calculate_r2((3, 3, 3), (5, 6, 3)) - 1.25) < 1e-6


def test_running_coupling_and_debye_mass_kat(oracle):
    """potential.rs:445-454"""
    assert abs(oracle.alphas(3.2) - 6.189593433886306) < 1e-14
    assert abs(oracle.mu(5.2) - 2.604838027702063) < 1e-14


# ---- oracle vs independent numpy restatement ---------------------------------------------------
@pytest.mark.parametrize("ext", [1, 2, 3])
@pytest.mark.parametrize("shape", [(12, 9, 10), (8, 8, 8)])
def test_sweep_bitwise_vs_numpy(oracle, ext, shape):
    rng = np.random.default_rng(7 + ext)
    dn, dt, mass = 0.05, 6.25e-4, 1.3
    g = oracle.make_grid(*shape, ext=ext, dn=dn, dt=dt, mass=mass)
    v = rng.normal(size=g.padded_shape) * 3.0
    a, b = oracle.build_ab(v, dt)
    a2, b2 = npr.build_ab(v, dt)
    assert np.array_equal(a, a2) and np.array_equal(b, b2)
    phi = np.zeros(g.padded_shape)
    npr.work(phi, ext)[...] = rng.normal(size=shape)
    ref = phi.copy()
    for _ in range(3):
        ref = npr.sweep(ref, a, b, ext, dn, dt, mass)
    oracle.evolve(g, phi, a, b, 3)
    assert np.array_equal(phi, ref)
    ring = np.ones(g.padded_shape, bool)
    npr.work(ring, ext)[...] = False
    assert not phi[ring].any()


@pytest.mark.parametrize("ext", [1, 2, 3])
def test_observables_vs_numpy(oracle, ext):
    rng = np.random.default_rng(11)
    dn, mass = 0.1, 0.7
    shape = (9, 11, 10)
    g = oracle.make_grid(*shape, ext=ext, dn=dn, dt=1e-3, mass=mass)
    v = rng.normal(size=g.padded_shape)
    phi = np.zeros(g.padded_shape)
    npr.work(phi, ext)[...] = rng.normal(size=shape)
    ps = rng.uniform(0.5, 2.0, size=shape)
    for mode in (0, 1):
        oracle.set_sum_mode(mode)
        for potsub, nps in ((None, None), (2.5, 2.5), (ps, ps)):
            o = oracle.observables(g, phi, v, potsub)
            r = npr.observables(phi, v, ext, dn, mass, nps)
            for key in ("energy", "norm2", "v_infinity", "r2"):
                assert o[key] == pytest.approx(r[key], rel=1e-13, abs=1e-13), (mode, key)
    oracle.set_sum_mode(0)


def test_evolve_excited_vs_numpy(oracle):
    """grid.rs:674-681: per-step norm2 -> normalise -> MGS"""
    rng = np.random.default_rng(3)
    ext, shape, dn, dt, mass = 1, (10, 10, 10), 0.1, 2e-3, 1.0
    g = oracle.make_grid(*shape, ext=ext, dn=dn, dt=dt, mass=mass)
    v = npr.harmonic(shape, ext, dn)
    a, b = npr.build_ab(v, dt)
    lowers = []
    for _ in range(2):
        q = np.zeros(g.padded_shape)
        npr.work(q, ext)[...] = rng.normal(size=shape)
        q = npr.orthogonalise(q, lowers)
        q = q / np.sqrt((q * q).sum())
        lowers.append(np.ascontiguousarray(q))
    phi = np.zeros(g.padded_shape)
    npr.work(phi, ext)[...] = rng.normal(size=shape)
    ref = phi.copy()
    for _ in range(4):
        ref = npr.sweep(ref, a, b, ext, dn, dt, mass)
        n2 = float((npr.work(ref, ext) ** 2).astype(np.longdouble).sum())
        ref = npr.orthogonalise(npr.normalise(ref, n2), lowers)
    oracle.set_sum_mode(1)
    oracle.evolve(g, phi, a, b, 4, lowers=lowers)
    oracle.set_sum_mode(0)
    assert np.linalg.norm(phi - ref) / np.linalg.norm(ref) < 1e-14
    for q in lowers:
        assert abs((q * phi).sum()) < 1e-14


def test_potentials_and_ic_vs_numpy(oracle):
    for ext in (1, 2, 3):
        g = oracle.make_grid(10, 12, 8, ext=ext, dn=0.05, dt=1e-4, mass=1.0)
        assert np.array_equal(oracle.potential(g, "Harmonic"), npr.harmonic((10, 12, 8), ext, 0.05))
        assert np.array_equal(oracle.initial_condition(g, "Boolean"), npr.boolean_ic((10, 12, 8), ext))
        c = oracle.initial_condition(g, "Constant")
        assert np.array_equal(npr.work(c, ext), np.full((10, 12, 8), 0.1))
        assert c.sum() == pytest.approx(0.1 * 960)
        assert not oracle.potential(g, "NoPotential").any()
    with pytest.raises(ValueError):
        oracle.potential(g, "FromFile")
    with pytest.raises(ValueError):
        oracle.initial_condition(g, "Gaussian")


def test_simple_cornell_and_potsub(oracle):
    """potential.rs:241-249, 346-363"""
    g = oracle.make_grid(8, 8, 8, ext=1, dn=0.2, dt=1e-3, mass=0.75)
    v = oracle.potential(g, "SimpleCornell", sig=0.223)
    r = 0.2 * np.sqrt(oracle.calculate_r2((2, 3, 7), (8, 8, 8)))
    assert v[2, 3, 7] == (-0.5 * (4. / 3.)) / r + 0.223 * r + 4. * 0.75
    assert oracle.potential_sub(g, "SimpleCornell") == 3.0
    assert oracle.potential_sub(g, "Harmonic") is None
    assert oracle.potential_sub(g, "ElipticalCoulomb") == 5.0
    assert oracle.potential_sub(g, "FullCornell", sig=0.223).shape == (8, 8, 8)
    g2 = oracle.make_grid(9, 9, 9, ext=1, dn=0.2, dt=1e-3, mass=0.75)  # odd N: a site sits at r = 0 < dn
    v2 = oracle.potential(g2, "SimpleCornell", sig=0.223)
    assert v2[5, 5, 5] == 3.0
    c = oracle.potential(g2, "Coulomb")
    assert c[5, 5, 5] == -1. / 0.2 and c[5, 5, 7] == -1. / (0.2 * 2.0)


def test_poschl_teller_matches_script(oracle):
    """The oracle's vectorised formula vs gen_potential.py:45-60 restated with the script's own numpy calls."""
    n, dn = (6, 7, 8), 0.3
    g = oracle.make_grid(*n, ext=2, dn=dn)
    v = oracle.potential(g, "PoschlTeller")
    ext_ = [(dn * m - dn) / 2 for m in n]
    sx, sy, sz = (np.linspace(-x, x, m) for x, m in zip(ext_, n))
    x, y, z = np.meshgrid(sx, sy, sz, indexing="ij")
    sech = lambda u: 1 / np.cosh(u)
    coeff = -(6 * (6 + 1)) / 2
    expect = coeff * (sech(x) * sech(x)) + coeff * (sech(y) * sech(y)) + coeff * (sech(z) * sech(z))
    assert np.allclose(npr.work(v, 2), expect, rtol=1e-15, atol=0)
    ring = np.ones(g.padded_shape, bool)
    npr.work(ring, 2)[...] = False
    assert not v[ring].any()
