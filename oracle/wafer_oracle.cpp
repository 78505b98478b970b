// wafer_oracle.cpp — CPU restatement of Wafer's imaginary-time FDTD hot path.
//
// THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load it.  The product library
// (wafer_b200/libwafer_b200.so) never links, loads or calls anything in this directory.
//
// What it restates (citations are file:line under /root/reference, the Libbum/Wafer tree):
//   evolve                      src/grid.rs:544-687   (3/5/7-point Jacobi sweep + per-step norm/GS for wnum>0)
//   compute_observables         src/grid.rs:303-445
//   get_norm_squared            src/grid.rs:454-457
//   normalise_wavefunction      src/grid.rs:465-468
//   orthogonalise_wavefunction  src/grid.rs:477-492   (modified Gram-Schmidt, sequential)
//   get_work_area               src/grid.rs:505-513
//   solve (driver loop)         src/grid.rs:50-246
//   A/B ancillary arrays        src/potential.rs:101-110
//   potential / potential_sub   src/potential.rs:188-363
//   calculate_r2, alphas, mu    src/potential.rs:366-398
//   initial conditions + ring   src/config.rs:577-683
//   Poschl-Teller script        gen_potential.py:45-60
//
// Pinning status.  The reference cannot be built here (no cargo/rustc; crates not vendored), so the
// oracle is pinned against the reference's own inline unit tests only:
//   gram_schmidt (grid.rs:721-746), norm2 = 70070 (grid.rs:780-786), wfn_normalise (grid.rs:788-799),
//   work_area / mut_work_area (grid.rs:748-778), distance_squared = 1.25 (potential.rs:434-443),
//   running_coupling (potential.rs:445-449), debye_screening_mass (potential.rs:450-454).
// The reference holds NO test, fixture or golden vector for `evolve` or `compute_observables`:
// for those two functions PARITY IS UNPINNED by reference data.  They are anchored instead on
// (i) a line-by-line restatement with the reference's exact floating-point evaluation order,
// (ii) an independent numpy restatement (tests/np_restatement.py) that must agree bit-for-bit, and
// (iii) analytic known answers (discrete box modes, harmonic oscillator levels) in tests/.
//
// Floating point: Rust never contracts a*b+c into an FMA and evaluates `*` `/` left to right.
// Build with -ffp-contract=off (see oracle/Makefile); every expression below is written in the
// reference's association order.
//
// Layout: every array is the reference's padded C-order Array3 (x slowest, z contiguous) of shape
// (nx+2e, ny+2e, nz+2e), e = ext in {1,2,3} (config.rs:222-239), unless stated "work" (nx,ny,nz).

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

extern "C" {

typedef struct {
    uint64_t nx, ny, nz;  // WORK sizes: config.grid.size.{x,y,z}   (config.rs:16-23)
    uint32_t ext;         // 1|2|3: central_difference.ext()          (config.rs:232-238)
    double dn, dt, mass;  // grid.dn, grid.dt, mass                   (config.rs:20-22,319)
} wo_grid;

typedef struct {
    uint64_t step;
    double tau, diff;
    double energy, norm2, v_infinity, r2;  // raw Observables (grid.rs:17-28); E = energy/norm2
} wo_record;

// potential kinds, in the order of PotentialType (config.rs:74-104); 100 = gen_potential.py formula
enum {
    WO_NOPOTENTIAL = 0, WO_CUBE, WO_QUADWELL, WO_PERIODIC, WO_COULOMB, WO_COMPLEXCOULOMB,
    WO_ELIPTICALCOULOMB, WO_SIMPLECORNELL, WO_FULLCORNELL, WO_HARMONIC, WO_COMPLEXHARMONIC,
    WO_DODECAHEDRON, WO_FROMFILE, WO_FROMSCRIPT, WO_POSCHLTELLER = 100
};

static int g_sum_mode = 0;  // 0: per-x-plane partials then sequential (deterministic, thread-count independent)
                            // 1: long double accumulation (bounds reduction-order noise)
void wo_set_sum_mode(int m) { g_sum_mode = m; }
int wo_get_sum_mode(void) { return g_sum_mode; }
int wo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void wo_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

}  // extern "C"

namespace {

struct Dims {
    size_t nx, ny, nz, e, px, py, pz;
    explicit Dims(const wo_grid* g)
        : nx(g->nx), ny(g->ny), nz(g->nz), e(g->ext), px(g->nx + 2 * g->ext), py(g->ny + 2 * g->ext),
          pz(g->nz + 2 * g->ext) {}
    size_t padded() const { return px * py * pz; }
    size_t work() const { return nx * ny * nz; }
    // padded linear index of padded coords
    size_t p(size_t i, size_t j, size_t k) const { return (i * py + j) * pz + k; }
    // work linear index
    size_t w(size_t i, size_t j, size_t k) const { return (i * ny + j) * nz + k; }
};

// Sum of f(plane) over x-planes: each plane partial is summed sequentially in memory order, the
// planes are then combined sequentially.  The reference's rayon sums (grid.rs:405,407,417,436,456,487)
// have no defined order, so any fixed order is an equally valid representative.
template <class PlaneFn>
double sum_planes(size_t nplanes, PlaneFn f) {
    if (g_sum_mode == 1) {
        std::vector<long double> part(nplanes);
#pragma omp parallel for schedule(static)
        for (long long i = 0; i < (long long)nplanes; ++i) part[i] = f((size_t)i, true);
        long double t = 0.0L;
        for (size_t i = 0; i < nplanes; ++i) t += part[i];
        return (double)t;
    }
    std::vector<double> part(nplanes);
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)nplanes; ++i) part[i] = (double)f((size_t)i, false);
    double t = 0.0;
    for (size_t i = 0; i < nplanes; ++i) t += part[i];
    return t;
}

// Unnormalised Laplacian sum S at padded coords (i,j,k), reference association order.
// grid.rs:582-588 / 608-620 / 642-659 (evolve) and 326-331 / 350-362 / 382-399 (observables).
template <int E>
inline double lap_sum(const double* phi, const Dims& d, size_t i, size_t j, size_t k, double w);

template <>
inline double lap_sum<1>(const double* phi, const Dims& d, size_t i, size_t j, size_t k, double w) {
    const size_t sx = d.py * d.pz, sy = d.pz, c = d.p(i, j, k);
    double s = phi[c + sx] + phi[c - sx];
    s = s + phi[c + sy];
    s = s + phi[c - sy];
    s = s + phi[c + 1];
    s = s + phi[c - 1];
    s = s - 6. * w;
    return s;
}

template <>
inline double lap_sum<2>(const double* phi, const Dims& d, size_t i, size_t j, size_t k, double w) {
    const size_t sx = d.py * d.pz, sy = d.pz, c = d.p(i, j, k);
    double s = -phi[c + 2 * sx];
    s = s + 16. * phi[c + sx];
    s = s + 16. * phi[c - sx];
    s = s - phi[c - 2 * sx];
    s = s - phi[c + 2 * sy];
    s = s + 16. * phi[c + sy];
    s = s + 16. * phi[c - sy];
    s = s - phi[c - 2 * sy];
    s = s - phi[c + 2];
    s = s + 16. * phi[c + 1];
    s = s + 16. * phi[c - 1];
    s = s - phi[c - 2];
    s = s - 90. * w;
    return s;
}

template <>
inline double lap_sum<3>(const double* phi, const Dims& d, size_t i, size_t j, size_t k, double w) {
    const size_t sx = d.py * d.pz, sy = d.pz, c = d.p(i, j, k);
    double s = 2. * phi[c + 3 * sx] - 27. * phi[c + 2 * sx];
    s = s + 270. * phi[c + sx];
    s = s + 270. * phi[c - sx];
    s = s - 27. * phi[c - 2 * sx];
    s = s + 2. * phi[c - 3 * sx];
    s = s + 2. * phi[c + 3 * sy];
    s = s - 27. * phi[c + 2 * sy];
    s = s + 270. * phi[c + sy];
    s = s + 270. * phi[c - sy];
    s = s - 27. * phi[c - 2 * sy];
    s = s + 2. * phi[c - 3 * sy];
    s = s + 2. * phi[c + 3];
    s = s - 27. * phi[c + 2];
    s = s + 270. * phi[c + 1];
    s = s + 270. * phi[c - 1];
    s = s - 27. * phi[c - 2];
    s = s + 2. * phi[c - 3];
    s = s - 1470. * w;
    return s;
}

inline double denominator(const wo_grid* g) {
    // grid.rs:569 / 594 / 626:  r64(c) * dn * dn * mass, left to right
    const double c = g->ext == 1 ? 2. : (g->ext == 2 ? 24. : 360.);
    return c * g->dn * g->dn * g->mass;
}

template <int E>
void sweep_into_work(const wo_grid* g, const Dims& d, const double* phi, const double* a, const double* b,
                     double* work) {
    const double den = denominator(g);
    const double dt = g->dt;
#pragma omp parallel for collapse(2) schedule(static)
    for (long long i = 0; i < (long long)d.nx; ++i)
        for (long long j = 0; j < (long long)d.ny; ++j) {
            const size_t pi = i + E, pj = j + E;
            for (size_t k = 0; k < d.nz; ++k) {
                const size_t c = d.p(pi, pj, k + E);
                const double w = phi[c];
                const double s = lap_sum<E>(phi, d, pi, pj, k + E, w);
                // grid.rs:580-589:  w*pa + pb*dt*S/den  ==  (w*pa) + (((pb*dt)*S)/den)
                work[d.w(i, j, k)] = w * a[c] + b[c] * dt * s / den;
            }
        }
}

template <int E>
long double energy_plane(const wo_grid* g, const Dims& d, const double* phi, const double* v, size_t i,
                         bool wide) {
    const double den = denominator(g);
    long double tl = 0.0L;
    double td = 0.0;
    const size_t pi = i + E;
    for (size_t j = 0; j < d.ny; ++j)
        for (size_t k = 0; k < d.nz; ++k) {
            const size_t c = d.p(pi, j + E, k + E);
            const double w = phi[c];
            const double s = lap_sum<E>(phi, d, pi, j + E, k + E, w);
            // grid.rs:325-332:  v*w*w - w*S/den  ==  ((v*w)*w) - ((w*S)/den)
            const double t = v[c] * w * w - w * s / den;
            if (wide) tl += t; else td += t;
        }
    return wide ? tl : (long double)td;
}

}  // namespace

extern "C" {

// ---------------------------------------------------------------- potential.rs:366-371
double wo_calculate_r2(uint64_t ix, uint64_t iy, uint64_t iz, uint64_t nx, uint64_t ny, uint64_t nz) {
    const double dx = (double)ix - ((double)nx + 1.) / 2.;
    const double dy = (double)iy - ((double)ny + 1.) / 2.;
    const double dz = (double)iz - ((double)nz + 1.) / 2.;
    return dx * dx + dy * dy + dz * dz;
}

// ---------------------------------------------------------------- potential.rs:374-398
double wo_alphas(double mu) {
    const double nf = 2.0;
    const double b0 = 11. - 2. * nf / 3.;
    const double b1 = 51. - 19. * nf / 3.;
    const double b2 = 2857. - 5033. * nf / 9. + 325. * nf * nf / 27.;
    const double r = 2.3;
    const double l = 2. * std::log(mu / r);
    return 4. * M_PI *
           (1. - 2. * b1 * std::log(l) / (b0 * b0 * l) +
            4. * b1 * b1 *
                ((std::log(l) - 0.5) * (std::log(l) - 0.5) + b2 * b0 / (8. * b1 * b1) - 5.0 / 4.0) /
                (b0 * b0 * b0 * b0 * l * l)) /
           (b0 * l);
}

double wo_mu(double t) {
    const double nf = 2.0, tc = 0.2;
    return 1.4 * std::sqrt((1. + nf / 6.) * 4. * M_PI * wo_alphas(2. * M_PI * t)) * t * tc;
}

// ---------------------------------------------------------------- potential.rs:101-110
void wo_build_ab(const double* v, double dt, double* a, double* b, uint64_t n) {
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)n; ++i) {
        b[i] = 1. / (1. + dt * v[i] / 2.);
        a[i] = (1. - dt * v[i] / 2.) * b[i];
    }
}

// ---------------------------------------------------------------- grid.rs:505-534
void wo_get_work_area(const wo_grid* g, const double* padded, double* work) {
    const Dims d(g);
    for (size_t i = 0; i < d.nx; ++i)
        for (size_t j = 0; j < d.ny; ++j)
            std::memcpy(work + d.w(i, j, 0), padded + d.p(i + d.e, j + d.e, d.e), d.nz * sizeof(double));
}
void wo_set_work_area(const wo_grid* g, double* padded, const double* work) {
    const Dims d(g);
    for (size_t i = 0; i < d.nx; ++i)
        for (size_t j = 0; j < d.ny; ++j)
            std::memcpy(padded + d.p(i + d.e, j + d.e, d.e), work + d.w(i, j, 0), d.nz * sizeof(double));
}

// ---------------------------------------------------------------- grid.rs:454-457 over the work area
double wo_norm2_work(const wo_grid* g, const double* phi) {
    const Dims d(g);
    return sum_planes(d.nx, [&](size_t i, bool wide) -> long double {
        long double tl = 0.0L;
        double td = 0.0;
        for (size_t j = 0; j < d.ny; ++j)
            for (size_t k = 0; k < d.nz; ++k) {
                const double w = phi[d.p(i + d.e, j + d.e, k + d.e)];
                if (wide) tl += w * w; else td += w * w;
            }
        return wide ? tl : (long double)td;
    });
}

// grid.rs:454-457 over an arbitrary dense view of n elements
double wo_norm2_flat(const double* w, uint64_t n) {
    if (g_sum_mode == 1) {
        long double t = 0.0L;
        for (uint64_t i = 0; i < n; ++i) t += w[i] * w[i];
        return (double)t;
    }
    double t = 0.0;
    for (uint64_t i = 0; i < n; ++i) t += w[i] * w[i];
    return t;
}

// ---------------------------------------------------------------- grid.rs:465-468
void wo_normalise(double* w, uint64_t n, double norm2) {
    const double norm = std::sqrt(norm2);
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)n; ++i) w[i] /= norm;
}

// ---------------------------------------------------------------- grid.rs:477-492
// n = number of elements of the (padded) arrays; planes = chunking used for the deterministic sum.
void wo_orthogonalise(double* w, const double* const* lowers, uint32_t wnum, uint64_t n, uint64_t planes) {
    if (planes == 0 || n % planes != 0) planes = 1;
    const size_t per = n / planes;
    for (uint32_t l = 0; l < wnum; ++l) {
        const double* q = lowers[l];
        const double s = sum_planes(planes, [&](size_t i, bool wide) -> long double {
            long double tl = 0.0L;
            double td = 0.0;
            for (size_t m = i * per; m < (i + 1) * per; ++m) {
                if (wide) tl += q[m] * w[m]; else td += q[m] * w[m];
            }
            return wide ? tl : (long double)td;
        });
#pragma omp parallel for schedule(static)
        for (long long m = 0; m < (long long)n; ++m) w[m] -= q[m] * s;
    }
}

// ---------------------------------------------------------------- grid.rs:544-687
// Same pass structure as the reference (stencil into `work`, copy back, per-step norm/normalise/GS
// when wnum>0), so this function is also the timed CPU baseline.
// `work` is the scratch array of grid.rs:560 (work-sized).  The reference allocates it once per evolve call, i.e. once
// per screen_update (1000) sweeps; wo_evolve_ws lets the timed baseline hand in a pre-faulted buffer so that a bounded
// sample of a few sweeps is not charged an allocation the real run amortises over a thousand.
void wo_evolve_ws(const wo_grid* g, double* phi, const double* a, const double* b, const double* const* lowers,
                  uint32_t wnum, uint64_t steps, double* work) {
    const Dims d(g);
    uint64_t done = 0;
    for (;;) {
        switch (g->ext) {
            case 1: sweep_into_work<1>(g, d, phi, a, b, work); break;
            case 2: sweep_into_work<2>(g, d, phi, a, b, work); break;
            default: sweep_into_work<3>(g, d, phi, a, b, work); break;
        }
        // grid.rs:666-673
#pragma omp parallel for collapse(2) schedule(static)
        for (long long i = 0; i < (long long)d.nx; ++i)
            for (long long j = 0; j < (long long)d.ny; ++j)
                std::memcpy(phi + d.p(i + d.e, j + d.e, d.e), work + d.w(i, j, 0), d.nz * sizeof(double));
        if (wnum > 0) {  // grid.rs:674-681
            const double n2 = wo_norm2_work(g, phi);
            wo_normalise(phi, d.padded(), n2);
            wo_orthogonalise(phi, lowers, wnum, d.padded(), d.px);
        }
        done += 1;  // grid.rs:682-685: do-while, so steps==0 still performs one sweep
        if (done >= steps) break;
    }
}

void wo_evolve(const wo_grid* g, double* phi, const double* a, const double* b, const double* const* lowers,
               uint32_t wnum, uint64_t steps) {
    const Dims d(g);
    std::vector<double> work(d.work());  // grid.rs:560
    wo_evolve_ws(g, phi, a, b, lowers, wnum, steps, work.data());
}

// ---------------------------------------------------------------- grid.rs:303-445
// potsub_mode: 0 none (_ => 0.), 1 scalar (None,Some(c)), 2 array (Some(arr),None) of WORK size.
// out = {energy, norm2, v_infinity, r2}
void wo_observables(const wo_grid* g, const double* phi, const double* v, int potsub_mode, double potsub_scalar,
                    const double* potsub_arr, double* out) {
    const Dims d(g);
    out[0] = sum_planes(d.nx, [&](size_t i, bool wide) -> long double {
        switch (g->ext) {
            case 1: return energy_plane<1>(g, d, phi, v, i, wide);
            case 2: return energy_plane<2>(g, d, phi, v, i, wide);
            default: return energy_plane<3>(g, d, phi, v, i, wide);
        }
    });
    out[1] = wo_norm2_work(g, phi);
    if (potsub_mode == 0) {
        out[2] = 0.;
    } else {
        out[2] = sum_planes(d.nx, [&](size_t i, bool wide) -> long double {
            long double tl = 0.0L;
            double td = 0.0;
            for (size_t j = 0; j < d.ny; ++j)
                for (size_t k = 0; k < d.nz; ++k) {
                    const double w = phi[d.p(i + d.e, j + d.e, k + d.e)];
                    const double ps = potsub_mode == 2 ? potsub_arr[d.w(i, j, k)] : potsub_scalar;
                    const double t = w * w * ps;  // grid.rs:415 / 421
                    if (wide) tl += t; else td += t;
                }
            return wide ? tl : (long double)td;
        });
    }
    out[3] = sum_planes(d.nx, [&](size_t i, bool wide) -> long double {
        long double tl = 0.0L;
        double td = 0.0;
        for (size_t j = 0; j < d.ny; ++j)
            for (size_t k = 0; k < d.nz; ++k) {
                const double w = phi[d.p(i + d.e, j + d.e, k + d.e)];
                // grid.rs:432-434: WORK-area indices against the (N+1)/2 centre
                const double t = w * w * wo_calculate_r2(i, j, k, d.nx, d.ny, d.nz);
                if (wide) tl += t; else td += t;
            }
        return wide ? tl : (long double)td;
    });
}

// ---------------------------------------------------------------- potential.rs:188-319 (+ gen_potential.py:45-60)
// Evaluated at PADDED indices 0..N+2e-1 (potential.rs:46-62).  Returns 0, or 1 for kinds with no formula.
static int potential_at(const wo_grid* g, int kind, double sig, size_t ix, size_t iy, size_t iz, double* out) {
    const double nx = (double)g->nx, ny = (double)g->ny, nz = (double)g->nz, dn = g->dn, mass = g->mass;
    const uint64_t ux = g->nx, uy = g->ny, uz = g->nz;
    switch (kind) {
        case WO_NOPOTENTIAL: *out = 0.0; return 0;
        case WO_CUBE:
            *out = ((ix > ux / 4 && ix <= 3 * ux / 4) && (iy > uy / 4 && iy <= 3 * uy / 4) &&
                    (iz > uz / 4 && iz <= 3 * uz / 4)) ? -10.0 : 0.0;
            return 0;
        case WO_QUADWELL:
            *out = ((ix > ux / 4 && ix <= 3 * ux / 4) && (iy > uy / 4 && iy <= 3 * uy / 4) &&
                    (iz > 3 * uz / 8 && iz <= 5 * uz / 8)) ? -10.0 : 0.0;
            return 0;
        case WO_PERIODIC: {
            double t = std::sin(2. * M_PI * ((double)ix - 1.) / (nx - 1.)) * std::sin(2. * M_PI * ((double)ix - 1.) / (nx - 1.));
            t *= std::sin(2. * M_PI * ((double)iy - 1.) / (ny - 1.)) * std::sin(2. * M_PI * ((double)iy - 1.) / (ny - 1.));
            t *= std::sin(2. * M_PI * ((double)iz - 1.) / (nz - 1.)) * std::sin(2. * M_PI * ((double)iz - 1.) / (nz - 1.));
            *out = -t + 1.;
            return 0;
        }
        case WO_COULOMB:
        case WO_COMPLEXCOULOMB: {
            const double r = dn * std::sqrt(wo_calculate_r2(ix, iy, iz, ux, uy, uz));
            *out = r < dn ? -1. / dn : -1. / r;
            return 0;
        }
        case WO_ELIPTICALCOULOMB: {
            const double dx = (double)ix - (nx + 1.) / 2.;
            const double dy = (double)iy - (ny + 1.) / 2.;
            const double dz = ((double)iz - (nz + 1.) / 2.) * 2.;
            const double r = dn * std::sqrt(dx * dx + dy * dy + dz * dz);
            *out = r < dn ? 0.0 : -1. / r + 1. / dn;
            return 0;
        }
        case WO_SIMPLECORNELL: {
            const double r = dn * std::sqrt(wo_calculate_r2(ix, iy, iz, ux, uy, uz));
            if (r < dn) *out = 4. * mass;
            else *out = (-0.5 * (4. / 3.)) / r + sig * r + 4. * mass;
            return 0;
        }
        case WO_FULLCORNELL: {
            const double t = 1.0, xi = 0.0;
            const double dz = (double)iz - (nz + 1.) / 2.;
            const double r = dn * std::sqrt(wo_calculate_r2(ix, iy, iz, ux, uy, uz));
            const double md = wo_mu(t) * (1. + (0.07 * std::pow(xi, 0.2)) * (1. - dn * dn * dz * dz / (r * r))) *
                              std::pow(1. + xi, -0.29);
            if (r < dn) *out = 4. * mass;
            else
                *out = (-wo_alphas(2. * M_PI * t) * (4. / 3.)) * std::exp(-md * r) / r +
                       sig * (1. - std::exp(-md * r)) / md - (0.8 * sig) / (4. * mass * mass * r) + 4. * mass;
            return 0;
        }
        case WO_HARMONIC:
        case WO_COMPLEXHARMONIC: {
            const double r = dn * std::sqrt(wo_calculate_r2(ix, iy, iz, ux, uy, uz));
            *out = r * r / 2.;
            return 0;
        }
        case WO_DODECAHEDRON: {
            const double x = ((double)ix - (nx + 1.) / 2.) / ((nx - 1.) / 2.);
            const double y = ((double)iy - (ny + 1.) / 2.) / ((ny - 1.) / 2.);
            const double z = ((double)iz - (nz + 1.) / 2.) / ((nz - 1.) / 2.);
            // twelve half-spaces of potential.rs:283-308, constants named for readability
            const double c0 = 12.70820393249937, c1 = 11.210068307552588, c2 = 14.674169922690343;
            const double c3 = 5.605034153776295, c4 = 3.23606797749979, c5 = 1.2360679774997896;
            const double c6 = 4.23606797749979, c7 = 5.23606797749979, c8 = 18.1382715378281;
            const double c9 = 3.464101615137755, c10 = 9.06913576891405, c11 = 15.70820393249937;
            const double c12 = 9.70820393249937, c13 = 5.605034153776294, c14 = 6.47213595499958;
            const double c15 = 25.41640786499874, c16 = 1.7320508075688772, c17 = 8.47213595499958;
            const bool in = c0 + c1 * x >= c2 * z && c1 * x <= c0 + c2 * z &&
                            c3 * (c4 * x - c5 * z) <= 6. * (c6 + c7 * y) && c8 * x + c9 * z <= c0 &&
                            c10 * x + c11 * y <= c0 + c9 * z && c12 * y <= c0 + c13 * x + c2 * z &&
                            c0 + c13 * x + c12 * y + c2 * z >= 0. && c11 * y + c9 * z <= c0 + c10 * x &&
                            c3 * (-c14 * x - c5 * z) <= c15 && c9 * z <= c10 * x + 3. * (c6 + c7 * y) &&
                            c16 * (c4 * x + c17 * z) <= 3. * (c6 + c4 * y) && c13 * x + c12 * y + c2 * z <= c0;
            *out = in ? -100. : 0.0;
            return 0;
        }
        default: return 1;  // FromFile / FromScript: PotentialNotAvailable (potential.rs:315-317)
    }
}

int wo_potential(const wo_grid* g, int kind, double sig, double* v) {
    const Dims d(g);
    if (kind == WO_POSCHLTELLER) {
        // gen_potential.py:45-60 fills the WORK area (script_potential embeds it in padded zeros,
        // input.rs:239-246); lam = 6; sx = linspace(-extent, extent, n).
        std::memset(v, 0, d.padded() * sizeof(double));
        const double lam = 6., coeff = -(lam * (lam + 1.)) / 2.;
        auto axis = [&](size_t n, size_t i) {
            const double extent = (g->dn * (double)n - g->dn) / 2.;
            if (n == 1) return -extent;
            const double step = (extent - (-extent)) / (double)(n - 1);
            return i == n - 1 ? extent : -extent + (double)i * step;  // numpy.linspace semantics
        };
        auto sech2 = [](double u) { const double s = 1. / std::cosh(u); return s * s; };
        // per-axis terms coeff*sech^2 (each site's value is (tx + ty) + tz, the script's left-to-right sum)
        std::vector<double> tx(d.nx), ty(d.ny), tz(d.nz);
        for (size_t i = 0; i < d.nx; ++i) tx[i] = coeff * sech2(axis(d.nx, i));
        for (size_t j = 0; j < d.ny; ++j) ty[j] = coeff * sech2(axis(d.ny, j));
        for (size_t k = 0; k < d.nz; ++k) tz[k] = coeff * sech2(axis(d.nz, k));
#pragma omp parallel for collapse(2) schedule(static)
        for (long long i = 0; i < (long long)d.nx; ++i)
            for (long long j = 0; j < (long long)d.ny; ++j)
                for (size_t k = 0; k < d.nz; ++k) v[d.p(i + d.e, j + d.e, k + d.e)] = tx[i] + ty[j] + tz[k];
        return 0;
    }
    double probe;
    if (potential_at(g, kind, sig, 0, 0, 0, &probe)) return 1;
#pragma omp parallel for collapse(2) schedule(static)
    for (long long i = 0; i < (long long)d.px; ++i)
        for (long long j = 0; j < (long long)d.py; ++j)
            for (size_t k = 0; k < d.pz; ++k) potential_at(g, kind, sig, i, j, k, &v[d.p(i, j, k)]);
    return 0;
}

// potential.rs:346-363 scalar pot_sub; returns mode (0 none, 1 scalar, 2 array needed) per potential.rs:134-153
int wo_potential_sub(const wo_grid* g, int kind, double* scalar) {
    *scalar = 0.0;
    if (kind == WO_FULLCORNELL) return 2;
    if (kind == WO_ELIPTICALCOULOMB) *scalar = 1. / g->dn;
    else if (kind == WO_SIMPLECORNELL) *scalar = 4.0 * g->mass;
    return *scalar > 0.0 ? 1 : 0;  // potential.rs:148-152
}

// potential.rs:326-341, evaluated at WORK indices (potential.rs:135-142); out has work size
void wo_potential_sub_array(const wo_grid* g, double sig, double* out) {
    const Dims d(g);
    const double t = 1.0, xi = 0.0;
    for (size_t i = 0; i < d.nx; ++i)
        for (size_t j = 0; j < d.ny; ++j)
            for (size_t k = 0; k < d.nz; ++k) {
                const double dz = (double)k - ((double)g->nz + 1.) / 2.;
                const double r = g->dn * std::sqrt(wo_calculate_r2(i, j, k, g->nx, g->ny, g->nz));
                const double md = wo_mu(t) * 1. + (0.07 * std::pow(xi, 0.2)) * (1. - g->dn * g->dn * dz * dz / (r * r)) *
                                                      std::pow(1. + xi, -0.29);
                out[d.w(i, j, k)] = sig / md + 4. * g->mass;
            }
}

// ---------------------------------------------------------------- config.rs:597-622
void wo_zero_ring(const wo_grid* g, double* w) {
    const Dims d(g);
    for (size_t i = 0; i < d.px; ++i)
        for (size_t j = 0; j < d.py; ++j)
            for (size_t k = 0; k < d.pz; ++k) {
                const bool ring = i < d.e || i >= d.px - d.e || j < d.e || j >= d.py - d.e || k < d.e || k >= d.pz - d.e;
                if (ring) w[d.p(i, j, k)] = 0.;
            }
}

// config.rs:586-595 + ring; kind: 2 Coulomb (650-669), 3 Constant (593), 4 Boolean (676-683)
int wo_initial_condition(const wo_grid* g, int kind, double* w) {
    const Dims d(g);
    if (kind < 2 || kind > 4) return 1;  // FromFile / Gaussian(thread_rng) are not reproducible here
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < d.px; ++i)
        for (size_t j = 0; j < d.py; ++j)
            for (size_t k = 0; k < d.pz; ++k) {
                double val;
                if (kind == 3) {
                    val = 0.1;
                } else if (kind == 4) {
                    // ((((i % 2) * j) % 2) * k) % 2 in f64  (config.rs:680)
                    val = std::fmod(std::fmod(std::fmod((double)i, 2.) * (double)j, 2.) * (double)k, 2.);
                } else {
                    const double dx = (double)i - (double)d.px / 2.;
                    const double dy = (double)j - (double)d.py / 2.;
                    const double dz = (double)k - (double)d.pz / 2.;
                    const double r = g->dn * std::sqrt(dx * dx + dy * dy + dz * dz);
                    const double costheta = g->dn * dz / r;
                    const double cosphi = g->dn * dx / r;
                    const double mr2 = std::exp(-g->mass * r / 2.);
                    val = std::exp(-g->mass * r) + (2. - g->mass * r) * mr2 + g->mass * r * mr2 * costheta +
                          g->mass * r * mr2 * std::sqrt(1. - costheta * costheta) * cosphi;
                }
                w[d.p(i, j, k)] = val;
            }
    wo_zero_ring(g, w);
    return 0;
}

// Deterministic excited-state seed used by the product's driver instead of the reference's noise-seeded clone
// (grid.rs:95): w = q * f(u,v,w).  Restated here so that whole excited-state runs can be compared.
void wo_seed_from_state(const wo_grid* g, const double* q, double* w) {
    const Dims d(g);
    std::memset(w, 0, d.padded() * sizeof(double));
    for (size_t i = 0; i < d.nx; ++i)
        for (size_t j = 0; j < d.ny; ++j)
            for (size_t k = 0; k < d.nz; ++k) {
                const double u = (2. * (double)i - ((double)d.nx - 1.)) / (double)d.nx;
                const double v = (2. * (double)j - ((double)d.ny - 1.)) / (double)d.ny;
                const double ww = (2. * (double)k - ((double)d.nz - 1.)) / (double)d.nz;
                double f = 1. + u;
                f = f + 0.5 * v;
                f = f + 0.25 * ww;
                f = f + 0.7 * (u * v);
                f = f + 0.4 * (v * ww);
                f = f + 0.3 * (u * ww);
                f = f + 0.2 * (u * u);
                f = f - 0.1 * (v * v);
                const size_t c = d.p(i + d.e, j + d.e, k + d.e);
                w[c] = q[c] * f;
            }
}

// ---------------------------------------------------------------- grid.rs:50-246
// phi: in = initial condition (set_initial_conditions result, or a seed / clone of w_store[wnum-1]);
//      out = state at loop exit.  max_steps < 0 means None; snap_update == 0 means None.
// Returns 1 if converged (the caller then pushes phi to w_store, grid.rs:241), 0 for Err(MaxStep).
int wo_solve(const wo_grid* g, const double* v, const double* a, const double* b, int potsub_mode,
             double potsub_scalar, const double* potsub_arr, double* phi, const double* const* lowers, uint32_t wnum,
             double tolerance, int64_t max_steps, uint64_t screen_update, uint64_t snap_update, wo_record* records,
             uint64_t max_records, uint64_t* n_records) {
    const Dims d(g);
    uint64_t step = 0, nrec = 0;
    double last_energy = 1.7976931348623157e308;  // f64::MAX, grid.rs:124
    int converged = 0;
    for (;;) {
        double obs[4];
        wo_observables(g, phi, v, potsub_mode, potsub_scalar, potsub_arr, obs);  // grid.rs:127
        const double norm_energy = obs[0] / obs[1];                               // grid.rs:128
        const double tau = (double)step * g->dt;                                  // grid.rs:129
        wo_normalise(phi, d.padded(), obs[1]);                                    // grid.rs:130
        if (wnum > 0) wo_orthogonalise(phi, lowers, wnum, d.padded(), d.px);      // grid.rs:133-135
        if (snap_update != 0 && step % snap_update == 0)                          // grid.rs:137-139 (NotConstrained)
            wo_normalise(phi, d.padded(), obs[1]);
        const double diff = std::fabs(norm_energy - last_energy);                 // grid.rs:161
        if (nrec < max_records) {
            wo_record& r = records[nrec];
            r.step = step; r.tau = tau; r.diff = diff;
            r.energy = obs[0]; r.norm2 = obs[1]; r.v_infinity = obs[2]; r.r2 = obs[3];
        }
        nrec++;
        if (diff < tolerance) { converged = 1; break; }                           // grid.rs:162-192
        last_energy = norm_energy;                                                // grid.rs:194
        if (max_steps >= 0 && step > (uint64_t)max_steps) break;                  // grid.rs:211-213
        wo_evolve(g, phi, a, b, lowers, wnum, screen_update);                     // grid.rs:216
        step += screen_update;                                                    // grid.rs:220
    }
    if (n_records) *n_records = nrec;
    return converged;
}

}  // extern "C"
