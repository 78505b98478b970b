"""ctypes binding of the CPU oracle (oracle/wafer_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under wafer_b200/ imports this module.

All arrays are the reference's padded C-order Array3 layout, shape (nx+2e, ny+2e, nz+2e),
x slowest / z contiguous, dtype float64.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libwafer_oracle.so")

POTENTIALS = {
    "NoPotential": 0, "Cube": 1, "QuadWell": 2, "Periodic": 3, "Coulomb": 4, "ComplexCoulomb": 5,
    "ElipticalCoulomb": 6, "SimpleCornell": 7, "FullCornell": 8, "Harmonic": 9, "ComplexHarmonic": 10,
    "Dodecahedron": 11, "FromFile": 12, "FromScript": 13, "PoschlTeller": 100,
}
INITIAL_CONDITIONS = {"FromFile": 0, "Gaussian": 1, "Coulomb": 2, "Constant": 3, "Boolean": 4}
EXT = {"ThreePoint": 1, "FivePoint": 2, "SevenPoint": 3}


class Grid(C.Structure):
    _fields_ = [("nx", C.c_uint64), ("ny", C.c_uint64), ("nz", C.c_uint64), ("ext", C.c_uint32),
                ("dn", C.c_double), ("dt", C.c_double), ("mass", C.c_double)]

    @property
    def padded_shape(self):
        e = self.ext
        return (self.nx + 2 * e, self.ny + 2 * e, self.nz + 2 * e)

    @property
    def work_shape(self):
        return (self.nx, self.ny, self.nz)


class Record(C.Structure):
    _fields_ = [("step", C.c_uint64), ("tau", C.c_double), ("diff", C.c_double), ("energy", C.c_double),
                ("norm2", C.c_double), ("v_infinity", C.c_double), ("r2", C.c_double)]


def build(force=False):
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(
            os.path.join(_HERE, "wafer_oracle.cpp")):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _LIB


_lib = None
_dp = C.POINTER(C.c_double)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        gp = C.POINTER(Grid)
        L.wo_calculate_r2.restype = C.c_double
        L.wo_calculate_r2.argtypes = [C.c_uint64] * 6
        L.wo_alphas.restype = C.c_double
        L.wo_alphas.argtypes = [C.c_double]
        L.wo_mu.restype = C.c_double
        L.wo_mu.argtypes = [C.c_double]
        L.wo_build_ab.argtypes = [_dp, C.c_double, _dp, _dp, C.c_uint64]
        L.wo_get_work_area.argtypes = [gp, _dp, _dp]
        L.wo_set_work_area.argtypes = [gp, _dp, _dp]
        L.wo_norm2_work.restype = C.c_double
        L.wo_norm2_work.argtypes = [gp, _dp]
        L.wo_norm2_flat.restype = C.c_double
        L.wo_norm2_flat.argtypes = [_dp, C.c_uint64]
        L.wo_normalise.argtypes = [_dp, C.c_uint64, C.c_double]
        L.wo_orthogonalise.argtypes = [_dp, C.POINTER(_dp), C.c_uint32, C.c_uint64, C.c_uint64]
        L.wo_evolve.argtypes = [gp, _dp, _dp, _dp, C.POINTER(_dp), C.c_uint32, C.c_uint64]
        L.wo_evolve_ws.argtypes = [gp, _dp, _dp, _dp, C.POINTER(_dp), C.c_uint32, C.c_uint64, _dp]
        L.wo_observables.argtypes = [gp, _dp, _dp, C.c_int, C.c_double, _dp, _dp]
        L.wo_potential.restype = C.c_int
        L.wo_potential.argtypes = [gp, C.c_int, C.c_double, _dp]
        L.wo_potential_sub.restype = C.c_int
        L.wo_potential_sub.argtypes = [gp, C.c_int, _dp]
        L.wo_potential_sub_array.argtypes = [gp, C.c_double, _dp]
        L.wo_zero_ring.argtypes = [gp, _dp]
        L.wo_initial_condition.restype = C.c_int
        L.wo_initial_condition.argtypes = [gp, C.c_int, _dp]
        L.wo_seed_from_state.argtypes = [gp, _dp, _dp]
        L.wo_solve.restype = C.c_int
        L.wo_solve.argtypes = [gp, _dp, _dp, _dp, C.c_int, C.c_double, _dp, _dp, C.POINTER(_dp), C.c_uint32,
                               C.c_double, C.c_int64, C.c_uint64, C.c_uint64, C.POINTER(Record), C.c_uint64,
                               C.POINTER(C.c_uint64)]
        L.wo_set_sum_mode.argtypes = [C.c_int]
        L.wo_num_threads.restype = C.c_int
        L.wo_set_num_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def _p(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"], "oracle arrays are C-order float64"
    return a.ctypes.data_as(_dp)


def _lowers(lowers):
    n = len(lowers)
    arr = (_dp * max(n, 1))()
    for i, q in enumerate(lowers):
        arr[i] = _p(q)
    return arr


def make_grid(nx, ny, nz, ext=1, dn=0.01, dt=3e-5, mass=1.0):
    return Grid(nx, ny, nz, ext, dn, dt, mass)


def set_sum_mode(mode):
    """0: per-plane partial sums (deterministic); 1: long double accumulation."""
    lib().wo_set_sum_mode(mode)


def num_threads():
    return lib().wo_num_threads()


def set_num_threads(n):
    lib().wo_set_num_threads(n)


def calculate_r2(idx, size):
    return lib().wo_calculate_r2(idx[0], idx[1], idx[2], size[0], size[1], size[2])


def alphas(mu):
    return lib().wo_alphas(mu)


def mu(t):
    return lib().wo_mu(t)


def build_ab(v, dt):
    a = np.empty_like(v)
    b = np.empty_like(v)
    lib().wo_build_ab(_p(v), dt, _p(a), _p(b), v.size)
    return a, b


def get_work_area(g, padded):
    out = np.empty(g.work_shape)
    lib().wo_get_work_area(C.byref(g), _p(padded), _p(out))
    return out


def set_work_area(g, padded, work):
    lib().wo_set_work_area(C.byref(g), _p(padded), _p(np.ascontiguousarray(work)))


def norm2_work(g, phi):
    return lib().wo_norm2_work(C.byref(g), _p(phi))


def norm2_flat(w):
    w = np.ascontiguousarray(w)
    return lib().wo_norm2_flat(_p(w), w.size)


def normalise(w, norm2):
    lib().wo_normalise(_p(w), w.size, norm2)


def orthogonalise(w, lowers, wnum=None):
    wnum = len(lowers) if wnum is None else wnum
    lib().wo_orthogonalise(_p(w), _lowers(lowers), wnum, w.size, w.shape[0])


def evolve(g, phi, a, b, steps, lowers=(), wnum=None, work=None):
    """work: optional pre-allocated work-sized scratch array (grid.rs:560); the timed baseline passes one so that a
    bounded sample is not charged the per-call allocation the reference amortises over screen_update sweeps"""
    wnum = len(lowers) if wnum is None else wnum
    if work is None:
        lib().wo_evolve(C.byref(g), _p(phi), _p(a), _p(b), _lowers(lowers), wnum, steps)
    else:
        assert work.shape == tuple(g.work_shape)
        lib().wo_evolve_ws(C.byref(g), _p(phi), _p(a), _p(b), _lowers(lowers), wnum, steps, _p(work))


def observables(g, phi, v, potsub=None):
    """potsub: None | float | ndarray(work shape).  Returns dict energy/norm2/v_infinity/r2 (raw sums)."""
    out = np.zeros(4)
    mode, sc, arr = 0, 0.0, None
    if isinstance(potsub, np.ndarray):
        mode, arr = 2, _p(potsub)
    elif potsub is not None and potsub > 0.0:
        mode, sc = 1, float(potsub)
    lib().wo_observables(C.byref(g), _p(phi), _p(v), mode, sc, arr, _p(out))
    return dict(energy=out[0], norm2=out[1], v_infinity=out[2], r2=out[3])


def potential(g, kind, sig=1.0):
    v = np.zeros(g.padded_shape)
    rc = lib().wo_potential(C.byref(g), POTENTIALS[kind] if isinstance(kind, str) else kind, sig, _p(v))
    if rc:
        raise ValueError("PotentialNotAvailable: %r" % (kind,))
    return v


def potential_sub(g, kind, sig=1.0):
    """Returns None, a float, or a work-sized array, like Potentials.pot_sub (potential.rs:115-153)."""
    sc = C.c_double(0.0)
    mode = lib().wo_potential_sub(C.byref(g), POTENTIALS[kind], C.byref(sc))
    if mode == 0:
        return None
    if mode == 1:
        return sc.value
    out = np.zeros(g.work_shape)
    lib().wo_potential_sub_array(C.byref(g), sig, _p(out))
    return out


def initial_condition(g, kind):
    w = np.zeros(g.padded_shape)
    rc = lib().wo_initial_condition(C.byref(g), INITIAL_CONDITIONS[kind], _p(w))
    if rc:
        raise ValueError("initial condition %r is not reproducible in the oracle" % (kind,))
    return w


def seed_from_state(g, q):
    """the product driver's deterministic excited-state start (generators.cuh seed_poly)"""
    w = np.zeros(g.padded_shape)
    lib().wo_seed_from_state(C.byref(g), _p(q), _p(w))
    return w


def solve(g, v, a, b, phi, potsub=None, lowers=(), wnum=None, tolerance=1e-4, max_steps=None, screen_update=1000,
          snap_update=None, max_records=4096):
    """Restates grid.rs:50-246.  phi is updated in place.  Returns (converged, [record dicts])."""
    wnum = len(lowers) if wnum is None else wnum
    recs = (Record * max_records)()
    n = C.c_uint64(0)
    mode, sc, arr = 0, 0.0, None
    if isinstance(potsub, np.ndarray):
        mode, arr = 2, _p(potsub)
    elif potsub is not None and potsub > 0.0:
        mode, sc = 1, float(potsub)
    conv = lib().wo_solve(C.byref(g), _p(v), _p(a), _p(b), mode, sc, arr, _p(phi), _lowers(lowers), wnum, tolerance,
                          -1 if max_steps is None else int(max_steps), screen_update,
                          0 if snap_update is None else int(snap_update), recs, max_records, C.byref(n))
    out = []
    for i in range(min(n.value, max_records)):
        r = recs[i]
        out.append(dict(step=r.step, tau=r.tau, diff=r.diff, energy=r.energy, norm2=r.norm2,
                        v_infinity=r.v_infinity, r2=r.r2, E=r.energy / r.norm2))
    return bool(conv), out
