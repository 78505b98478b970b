"""Property tests (hypothesis) of the CPU oracle: size-independent invariants of the reference algorithm that the GPU
tests also rely on at full size.  CPU only, small lattices, bounded example counts."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import np_restatement as npr

from oracle import binding as oracle

shapes = st.tuples(st.integers(3, 9), st.integers(3, 9), st.integers(3, 9))
exts = st.sampled_from([1, 2, 3])
seeds = st.integers(0, 2 ** 31 - 1)
FAST = settings(max_examples=25, deadline=None)


def _state(shape, ext, seed, scale=1.0):
    rng = np.random.default_rng(seed)
    g = oracle.make_grid(*shape, ext=ext, dn=0.1, dt=1e-3, mass=1.0)
    v = rng.normal(size=g.padded_shape)
    phi = np.zeros(g.padded_shape)
    npr.work(phi, ext)[...] = rng.normal(size=shape) * scale
    return g, v, phi


@FAST
@given(shapes, exts, seeds)
def test_sweep_matches_numpy_restatement_bitwise(shape, ext, seed):
    g, v, phi = _state(shape, ext, seed)
    a, b = oracle.build_ab(v, g.dt)
    ref = npr.sweep(npr.sweep(phi, a, b, ext, g.dn, g.dt, g.mass), a, b, ext, g.dn, g.dt, g.mass)
    oracle.evolve(g, phi, a, b, 2)
    assert np.array_equal(phi, ref)


@FAST
@given(shapes, exts, seeds, st.floats(-3, 3), st.floats(-3, 3))
def test_sweep_is_linear(shape, ext, seed, ca, cb):
    g, v, x = _state(shape, ext, seed)
    _, _, y = _state(shape, ext, seed + 1)
    a, b = oracle.build_ab(v, g.dt)
    z = ca * x + cb * y
    for arr in (x, y, z):
        oracle.evolve(g, arr, a, b, 2)
    assert np.allclose(z, ca * x + cb * y, rtol=0, atol=1e-12 * (1 + abs(ca) + abs(cb)))


@FAST
@given(shapes, exts, seeds, st.floats(0.1, 10))
def test_observables_scale_quadratically_and_energy_is_scale_free(shape, ext, seed, c):
    g, v, phi = _state(shape, ext, seed)
    o1 = oracle.observables(g, phi, v, 2.5)
    o2 = oracle.observables(g, np.ascontiguousarray(c * phi), v, 2.5)
    for k in o1:
        assert o2[k] == pytest.approx(c * c * o1[k], rel=1e-11, abs=1e-11)
    assert o2["energy"] / o2["norm2"] == pytest.approx(o1["energy"] / o1["norm2"], rel=1e-10, abs=1e-10)
    assert o1["v_infinity"] == pytest.approx(2.5 * o1["norm2"], rel=1e-12)  # scalar pot_sub: grid.rs:419-424


@FAST
@given(shapes, exts, seeds)
def test_normalise_then_norm_is_one_and_ring_stays_zero(shape, ext, seed):
    g, v, phi = _state(shape, ext, seed)
    n2 = oracle.norm2_work(g, phi)
    oracle.normalise(phi, n2)
    assert oracle.norm2_work(g, phi) == pytest.approx(1.0, rel=1e-13)
    ring = np.ones(g.padded_shape, bool)
    npr.work(ring, ext)[...] = False
    assert not phi[ring].any()


@FAST
@given(shapes, seeds, st.integers(1, 3))
def test_gram_schmidt_leaves_state_orthogonal_to_orthonormal_lowers(shape, seed, k):
    g, v, phi = _state(shape, 1, seed)
    rng = np.random.default_rng(seed + 7)
    lowers = []
    for _ in range(k):
        q = np.zeros(g.padded_shape)
        npr.work(q, 1)[...] = rng.normal(size=shape)
        q = npr.orthogonalise(q, lowers)
        lowers.append(np.ascontiguousarray(q / np.sqrt((q * q).sum())))
    oracle.orthogonalise(phi, lowers)
    for q in lowers:
        assert abs((q * phi).sum()) < 1e-12 * max(1.0, np.abs(phi).max() * phi.size ** 0.5)
    before = phi.copy()
    oracle.orthogonalise(phi, lowers)  # idempotent to rounding
    assert np.allclose(phi, before, rtol=0, atol=1e-13 * max(1.0, np.abs(before).max()))


@FAST
@given(st.integers(1, 5000), st.integers(1, 16))
def test_slab_partition_tiles_the_axis(nx, world):
    import wafer_b200
    if world > nx:
        return
    edges = [wafer_b200.slab_partition(nx, world, r) for r in range(world)]
    assert edges[0][0] == 0 and edges[-1][1] == nx
    assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
    sizes = [b - a for a, b in edges]
    assert max(sizes) - min(sizes) <= 1
