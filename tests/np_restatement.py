"""Second, independent restatement of the reference hot path in vectorised numpy.

Written directly from /root/reference/src/grid.rs and potential.rs (not from oracle/wafer_oracle.cpp).
numpy evaluates each ufunc separately in IEEE double (no FMA contraction), and the expressions
below keep the reference's left-to-right association, so the sweep must agree with the C++ oracle
BIT FOR BIT.  Used only by tests.
"""
import numpy as np

_COEF = {1: 2.0, 2: 24.0, 3: 360.0}


def work(a, e):
    """grid.rs:505-513"""
    return a[e:a.shape[0] - e, e:a.shape[1] - e, e:a.shape[2] - e]


def _sh(phi, e, dx, dy, dz):
    nx, ny, nz = phi.shape
    return phi[e + dx:nx - e + dx, e + dy:ny - e + dy, e + dz:nz - e + dz]


def lap_sum(phi, e):
    w = work(phi, e)
    if e == 1:  # grid.rs:582-588
        s = _sh(phi, e, 1, 0, 0) + _sh(phi, e, -1, 0, 0)
        s = s + _sh(phi, e, 0, 1, 0)
        s = s + _sh(phi, e, 0, -1, 0)
        s = s + _sh(phi, e, 0, 0, 1)
        s = s + _sh(phi, e, 0, 0, -1)
        return s - 6.0 * w
    if e == 2:  # grid.rs:608-620
        s = None
        for ax in range(3):
            d = [0, 0, 0]
            for off, c in ((2, -1.0), (1, 16.0), (-1, 16.0), (-2, -1.0)):
                d[ax] = off
                arr = _sh(phi, e, *d)
                if s is None:
                    s = -arr
                elif c < 0:
                    s = s - arr
                else:
                    s = s + c * arr
        return s - 90.0 * w
    # grid.rs:642-659
    s = None
    for ax in range(3):
        d = [0, 0, 0]
        for off, c in ((3, 2.0), (2, -27.0), (1, 270.0), (-1, 270.0), (-2, -27.0), (-3, 2.0)):
            d[ax] = off
            t = abs(c) * _sh(phi, e, *d)
            if s is None:
                s = t
            elif c > 0:
                s = s + t
            else:
                s = s - t
    return s - 1470.0 * w


def denominator(e, dn, mass):
    return _COEF[e] * dn * dn * mass


def build_ab(v, dt):
    """potential.rs:104-110"""
    b = 1.0 / (1.0 + dt * v / 2.0)
    a = (1.0 - dt * v / 2.0) * b
    return a, b


def sweep(phi, a, b, e, dn, dt, mass):
    """One Jacobi step of grid.rs:567-673; returns a new padded array."""
    den = denominator(e, dn, mass)
    w = work(phi, e)
    new = w * work(a, e) + work(b, e) * dt * lap_sum(phi, e) / den
    out = phi.copy()
    work(out, e)[...] = new
    return out


def calculate_r2_grid(n):
    """potential.rs:366-371 on work indices (grid.rs:432-433)"""
    i = np.arange(n[0], dtype=np.float64)[:, None, None] - (float(n[0]) + 1.0) / 2.0
    j = np.arange(n[1], dtype=np.float64)[None, :, None] - (float(n[1]) + 1.0) / 2.0
    k = np.arange(n[2], dtype=np.float64)[None, None, :] - (float(n[2]) + 1.0) / 2.0
    return i * i + j * j + k * k


def observables(phi, v, e, dn, mass, potsub=None):
    """grid.rs:303-445, sums in long double (exactly-rounded reference value for the tests)."""
    den = denominator(e, dn, mass)
    w = work(phi, e)
    integrand = work(v, e) * w * w - w * lap_sum(phi, e) / den
    ld = np.longdouble
    out = dict(energy=float(integrand.astype(ld).sum()), norm2=float((w * w).astype(ld).sum()))
    if potsub is None:
        out["v_infinity"] = 0.0
    else:
        out["v_infinity"] = float((w * w * potsub).astype(ld).sum())
    out["r2"] = float((w * w * calculate_r2_grid(w.shape)).astype(ld).sum())
    return out


def normalise(phi, norm2):
    """grid.rs:465-468"""
    return phi / np.sqrt(norm2)


def orthogonalise(phi, lowers):
    """grid.rs:477-492 (sequential / modified Gram-Schmidt)"""
    w = phi.copy()
    for q in lowers:
        s = float((q * w).astype(np.longdouble).sum())
        w = w - q * s
    return w


def harmonic(n, e, dn):
    """potential.rs:270-274 at padded indices"""
    p = [x + 2 * e for x in n]
    i = np.arange(p[0], dtype=np.float64)[:, None, None] - (float(n[0]) + 1.0) / 2.0
    j = np.arange(p[1], dtype=np.float64)[None, :, None] - (float(n[1]) + 1.0) / 2.0
    k = np.arange(p[2], dtype=np.float64)[None, None, :] - (float(n[2]) + 1.0) / 2.0
    r = dn * np.sqrt(i * i + j * j + k * k)
    return r * r / 2.0


def boolean_ic(n, e):
    """config.rs:676-683 + ring zeroing 597-622"""
    p = [x + 2 * e for x in n]
    i = np.arange(p[0], dtype=np.float64)[:, None, None]
    j = np.arange(p[1], dtype=np.float64)[None, :, None]
    k = np.arange(p[2], dtype=np.float64)[None, None, :]
    w = np.fmod(np.fmod(np.fmod(i, 2.0) * j, 2.0) * k, 2.0)
    out = np.zeros(p)
    work(out, e)[...] = work(w, e)
    return out
