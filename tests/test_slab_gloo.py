"""World-size-2 (and 3) CPU runs over gloo of the x-slab decomposition protocol the CUDA library uses for N > 1
(wafer_b200.cu: wafer_slab_partition, exchange(), owned-plane reductions, element-wise passes over ghost planes).

Each rank holds planes [x0-e, x1+e) of every field, sweeps its owned planes with the numpy restatement of the
reference sweep, swaps `e` boundary planes with its neighbours after every sweep and all-reduces partial sums.
The result must equal the single-domain run: bit-for-bit for the sweep, to rounding for the sums."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _exchange(dist, torch, loc, e, rank, world):
    """ghost-plane swap of wafer_b200.cu::exchange: send owned boundary planes, receive into ghost planes"""
    reqs, bufs = [], []
    L = loc.shape[0] - 2 * e
    if rank > 0:
        s = torch.from_numpy(np.ascontiguousarray(loc[e:2 * e]))
        r = torch.empty_like(s)
        reqs += [dist.isend(s, rank - 1), dist.irecv(r, rank - 1)]
        bufs.append((slice(0, e), r))
    if rank < world - 1:
        s = torch.from_numpy(np.ascontiguousarray(loc[L:L + e]))
        r = torch.empty_like(s)
        reqs += [dist.isend(s, rank + 1), dist.irecv(r, rank + 1)]
        bufs.append((slice(L + e, L + 2 * e), r))
    for q in reqs:
        q.wait()
    for sl, r in bufs:
        loc[sl] = r.numpy()


def _worker(rank, world, port, ext, shape, steps, nlow, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    import np_restatement as npr
    import wafer_b200

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    e, (nx, ny, nz) = ext, shape
    dn, dt, mass = 0.1, 2e-3, 1.0
    rng = np.random.default_rng(5)  # same global fields on every rank
    P = (nx + 2 * e, ny + 2 * e, nz + 2 * e)
    v = rng.normal(size=P)
    phi = np.zeros(P)
    npr.work(phi, e)[...] = rng.normal(size=shape)
    lowers = []
    for _ in range(nlow):
        q = np.zeros(P)
        npr.work(q, e)[...] = rng.normal(size=shape)
        lowers.append(q / np.sqrt((q * q).sum()))
    a, b = npr.build_ab(v, dt)
    x0, x1 = wafer_b200.slab_partition(nx, world, rank)
    sl = slice(x0, x1 + 2 * e)  # padded planes [x0, x1+2e) = owned + e ghosts each side
    loc, la, lb, lv = phi[sl].copy(), a[sl], b[sl], v[sl]
    llow = [q[sl].copy() for q in lowers]
    own = slice(e, e + (x1 - x0))
    # Gram matrix of the stored states, measured once when they join the store (wafer_b200.cu::register_lower)
    gram = torch.zeros((max(nlow, 1), max(nlow, 1)), dtype=torch.float64)
    for i in range(nlow):
        for j in range(i):
            gram[i, j] = float((llow[i][own] * llow[j][own]).astype(np.longdouble).sum())
    dist.all_reduce(gram)
    for _ in range(steps):
        loc = npr.sweep(loc, la, lb, e, dn, dt, mass)  # writes planes [e, L+e) only: exactly the owned planes
        _exchange(dist, torch, loc, e, rank, world)
        if nlow:
            # wafer_b200.cu::gs_apply: ONE all-reduce per step of [sum psi'^2, sum q_i psi'] taken on the un-normalised
            # psi', then s_i = d_i - sum_{j<i} G_ij s_j with d_i = raw_i / sqrt(norm2)  (== modified Gram-Schmidt)
            raw = [float((npr.work(loc, e) ** 2).astype(np.longdouble).sum())]
            raw += [float((q[own] * loc[own]).astype(np.longdouble).sum()) for q in llow]
            raw = torch.tensor(raw, dtype=torch.float64)
            dist.all_reduce(raw)
            norm = np.sqrt(raw[0].item())
            coef = []
            for i in range(nlow):
                s = raw[1 + i].item() / norm
                for j in range(i):
                    s -= gram[i, j].item() * coef[j]
                coef.append(s)
            loc = loc / norm  # element-wise over ghosts too: no exchange needed afterwards
            for q, s in zip(llow, coef):
                loc = loc - q * s
    # observables with GLOBAL work indices for r2 (grid.rs:432-433) on owned planes
    den = npr.denominator(e, dn, mass)
    w = npr.work(loc, e)
    integrand = npr.work(lv, e) * w * w - w * npr.lap_sum(loc, e) / den
    r2 = npr.calculate_r2_grid(shape)[x0:x1]
    part = torch.tensor([float(integrand.astype(np.longdouble).sum()), float((w * w).astype(np.longdouble).sum()),
                         float((w * w * r2).astype(np.longdouble).sum())], dtype=torch.float64)
    dist.all_reduce(part)
    np.save(os.path.join(out_dir, "slab_%d.npy" % rank), loc)
    if rank == 0:
        np.save(os.path.join(out_dir, "obs.npy"), part.numpy())
    dist.barrier()
    dist.destroy_process_group()


def _exchange_depth(dist, torch, loc, gx, rank, world):
    """depth-gx ghost swap (the ThreePoint slabs keep gx = 2 ghost planes for the two-steps-per-pass sweep)"""
    reqs, bufs = [], []
    L = loc.shape[0] - 2 * gx
    if rank > 0:
        s = torch.from_numpy(np.ascontiguousarray(loc[gx:2 * gx]))
        r = torch.empty_like(s)
        reqs += [dist.isend(s, rank - 1), dist.irecv(r, rank - 1)]
        bufs.append((slice(0, gx), r))
    if rank < world - 1:
        s = torch.from_numpy(np.ascontiguousarray(loc[L:L + gx]))
        r = torch.empty_like(s)
        reqs += [dist.isend(s, rank + 1), dist.irecv(r, rank + 1)]
        bufs.append((slice(L + gx, L + 2 * gx), r))
    for q in reqs:
        q.wait()
    for sl, r in bufs:
        loc[sl] = r.numpy()


def _worker_two_step(rank, world, port, shape, passes, tail, out_dir):
    """Protocol of the time-tiled path (wafer_b200.cu::wafer_evolve, `two` passes): ghost depth 2, TWO sweeps per halo
    exchange — the first sweep is also evaluated on the inner ghost plane (redundantly, from the depth-2 ghosts), the
    second on the owned planes only; planes outside the global lattice stay exactly zero."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    import np_restatement as npr
    import wafer_b200

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    e, gx, (nx, ny, nz) = 1, 2, shape
    dn, dt, mass = 0.1, 2e-3, 1.0
    rng = np.random.default_rng(9)
    P = (nx + 2, ny + 2, nz + 2)
    v = rng.normal(size=P)
    phi = np.zeros(P)
    npr.work(phi, e)[...] = rng.normal(size=shape)
    a, b = npr.build_ab(v, dt)
    x0, x1 = wafer_b200.slab_partition(nx, world, rank)
    L = x1 - x0

    def slab(arr):  # local planes [-gx, L+gx) = padded planes [x0+1-gx, x1+1+gx), zero outside the array
        out = np.zeros((L + 2 * gx,) + P[1:])
        lo, hi = x0 + 1 - gx, x1 + 1 + gx
        clo, chi = max(lo, 0), min(hi, P[0])
        out[clo - lo:chi - lo] = arr[clo:chi]
        return out

    loc, la, lb = slab(phi), slab(a), slab(b)
    inside = np.array([0 <= x0 + i - gx < nx for i in range(L + 2 * gx)])  # local plane inside the global lattice

    def sweep_planes(src, lo, hi):
        """one step on local planes [lo, hi) (indices into the slab array), others copied"""
        new = npr.sweep(src, la, lb, e, dn, dt, mass)  # updates planes [1, n-1)
        out = src.copy()
        out[lo:hi] = new[lo:hi]
        out[~inside] = 0.0  # the ring never changes
        return out

    for _ in range(passes):
        lvl1 = sweep_planes(loc, gx - 1, gx + L + 1)   # first step incl. one ghost plane each side
        loc = sweep_planes(lvl1, gx, gx + L)           # second step on owned planes
        _exchange_depth(dist, torch, loc, gx, rank, world)
    for _ in range(tail):                              # odd tail: one step, still a depth-2 exchange
        loc = sweep_planes(loc, gx, gx + L)
        _exchange_depth(dist, torch, loc, gx, rank, world)
    np.save(os.path.join(out_dir, "slab2_%d.npy" % rank), loc)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,shape,passes,tail", [(2, (10, 6, 7), 3, 1), (3, (13, 5, 6), 2, 0), (2, (9, 4, 5), 2, 2)])
def test_two_steps_per_exchange_protocol(tmp_path, world, shape, passes, tail):
    import torch.multiprocessing as mp

    import np_restatement as npr
    import wafer_b200

    port = 31500 + (os.getpid() + 11 * world + passes) % 2000
    mp.spawn(_worker_two_step, args=(world, port, shape, passes, tail, str(tmp_path)), nprocs=world, join=True)
    nx, ny, nz = shape
    rng = np.random.default_rng(9)
    P = (nx + 2, ny + 2, nz + 2)
    v = rng.normal(size=P)
    phi = np.zeros(P)
    npr.work(phi, 1)[...] = rng.normal(size=shape)
    a, b = npr.build_ab(v, 2e-3)
    for _ in range(2 * passes + tail):
        phi = npr.sweep(phi, a, b, 1, 0.1, 2e-3, 1.0)
    for r in range(world):
        x0, x1 = wafer_b200.slab_partition(nx, world, r)
        loc = np.load(tmp_path / ("slab2_%d.npy" % r))
        assert np.array_equal(loc[2:2 + x1 - x0], phi[x0 + 1:x1 + 1])          # owned planes: bit-identical
        for gpl, row in ((x0 - 1, loc[0]), (x0, loc[1]), (x1 + 1, loc[-2]), (x1 + 2, loc[-1])):  # depth-2 ghosts
            expect = phi[gpl] if 0 <= gpl < P[0] else np.zeros(P[1:])
            assert np.array_equal(row, expect)


@pytest.mark.parametrize("world,ext,shape,nlow", [(2, 1, (10, 6, 7), 0), (2, 2, (9, 6, 8), 0), (3, 3, (11, 8, 8), 0),
                                                   (2, 1, (8, 6, 6), 2)])
def test_slab_decomposition_matches_single_domain(tmp_path, world, ext, shape, nlow):
    import torch.multiprocessing as mp

    import np_restatement as npr
    import wafer_b200

    steps = 4
    port = 29500 + (os.getpid() + 7 * world + ext) % 2000
    mp.spawn(_worker, args=(world, port, ext, shape, steps, nlow, str(tmp_path)), nprocs=world, join=True)

    e, (nx, ny, nz) = ext, shape
    dn, dt, mass = 0.1, 2e-3, 1.0
    rng = np.random.default_rng(5)
    P = (nx + 2 * e, ny + 2 * e, nz + 2 * e)
    v = rng.normal(size=P)
    phi = np.zeros(P)
    npr.work(phi, e)[...] = rng.normal(size=shape)
    lowers = []
    for _ in range(nlow):
        q = np.zeros(P)
        npr.work(q, e)[...] = rng.normal(size=shape)
        lowers.append(q / np.sqrt((q * q).sum()))
    a, b = npr.build_ab(v, dt)
    for _ in range(steps):
        phi = npr.sweep(phi, a, b, e, dn, dt, mass)
        if nlow:
            n2 = float((npr.work(phi, e) ** 2).astype(np.longdouble).sum())
            phi = npr.orthogonalise(npr.normalise(phi, n2), lowers)
    got = np.zeros(P)
    covered = 0
    for r in range(world):
        x0, x1 = wafer_b200.slab_partition(nx, world, r)
        loc = np.load(tmp_path / ("slab_%d.npy" % r))
        assert loc.shape[0] == x1 - x0 + 2 * e
        got[x0 + e:x1 + e] = loc[e:e + x1 - x0]
        # ghost planes hold the neighbour's boundary planes (or the zero ring at the ends)
        assert np.array_equal(loc[:e], phi[x0:x0 + e]) if nlow == 0 else np.allclose(loc[:e], phi[x0:x0 + e], atol=1e-14)
        covered += x1 - x0
    assert covered == nx
    if nlow == 0:
        assert np.array_equal(got, phi)  # the sweep has no reductions: bit-identical to the single domain
    else:
        assert np.linalg.norm(got - phi) / np.linalg.norm(phi) < 1e-13
    ref = npr.observables(phi, v, e, dn, mass)
    obs = np.load(tmp_path / "obs.npy")
    for val, key in zip(obs, ("energy", "norm2", "r2")):
        assert val == pytest.approx(ref[key], rel=1e-12)


def test_partition_rule():
    import wafer_b200
    for nx in (1, 7, 50, 1024, 2048):
        for world in (1, 2, 3, 4, 8):
            if world > nx:
                continue
            edges = [wafer_b200.slab_partition(nx, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == nx
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    with pytest.raises(wafer_b200.WaferError):
        wafer_b200.slab_partition(8, 2, 2)
