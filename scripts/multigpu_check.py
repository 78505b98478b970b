"""Multi-rank parity check (run under torchrun, one rank per GPU): the x-slab decomposed run must equal the
single-GPU run of the same lattice — bit-for-bit after ground-state sweeps, to rounding for excited-state steps
and observables.  Rank 0 prints one JSON line; exit code 1 on mismatch."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist

    import wafer_b200

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def fresh_nccl_id():
        """a ncclUniqueId is good for exactly one communicator: make and broadcast a new one per Lattice"""
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.tensor(list(wafer_b200.nccl_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(idt, 0)
        return bytes(idt.cpu().tolist())

    use_p2p = os.environ.get("WAFER_P2P", "1") == "1"

    def connect(lat):
        mine = torch.tensor(list(lat.p2p_export()), dtype=torch.uint8, device="cuda")
        blobs = [torch.zeros(192, dtype=torch.uint8, device="cuda") for _ in range(world)]
        dist.all_gather(blobs, mine)
        blobs = [bytes(b.cpu().tolist()) for b in blobs]
        lat.p2p_connect(blobs[rank - 1] if rank > 0 else None, blobs[rank + 1] if rank < world - 1 else None)

    results = {}
    ok = True
    for ext, cd, shape in ((1, "ThreePoint", (96, 40, 72)), (2, "FivePoint", (50, 33, 40)), (3, "SevenPoint", (41, 24, 30))):
        dn, dt, mass = 0.1, 2e-3, 1.0
        rng = np.random.default_rng(11)
        P = tuple(s + 2 * ext for s in shape)
        v = rng.normal(size=P)
        phi = np.zeros(P)
        phi[ext:-ext, ext:-ext, ext:-ext] = rng.normal(size=shape)
        q = np.zeros(P)
        q[ext:-ext, ext:-ext, ext:-ext] = rng.normal(size=shape)
        q /= np.sqrt((q * q).sum())
        kw = dict(dn=dn, dt=dt, mass=mass, device=local)
        single = wafer_b200.Lattice(shape, cd, **kw)
        multi = wafer_b200.Lattice(shape, cd, rank=rank, world=world, nccl_id=fresh_nccl_id(), **kw)
        if use_p2p:
            connect(multi)
        outs = []
        for lat in (single, multi):
            lat.set_potential(v)
            lat.set_phi(phi)
            lat.evolve(0, 7)
            lat.evolve(0, 3)     # odd + even tails back to back: the single-step pass must hand over depth-2 ghosts
            lat.evolve(0, 4)
            g = lat.get_phi()    # before the check: its normalise divides by a sum whose last bit depends on the
            o1 = lat.check(0)    # reduction order (per-rank partials + all-reduce), the sweep itself does not
            lat.push_lower(q)
            lat.evolve(1, 3)
            o2 = lat.check(1)
            e = lat.get_phi()
            outs.append((o1, g, o2, e))
        (s1, sg, s2, se), (m1, mg, m2, me) = outs
        x0, x1 = multi.slab
        own = slice(x0 + ext, x1 + ext)
        bit = bool(np.array_equal(sg[own], mg[own]))
        l2 = float(np.linalg.norm(se[own] - me[own]) / max(np.linalg.norm(se[own]), 1e-300))
        de = max(abs(s1[k] - m1[k]) / max(abs(s1[k]), 1e-300) for k in s1)
        de2 = max(abs(s2[k] - m2[k]) / max(abs(s2[k]), 1e-300) for k in s2)
        flags = torch.tensor([float(bit), l2, de, de2], dtype=torch.float64, device="cuda")
        allf = [torch.zeros_like(flags) for _ in range(world)]
        dist.all_gather(allf, flags)
        allf = torch.stack(allf).cpu().numpy()
        res = dict(ground_bitwise=bool(allf[:, 0].all()), excited_l2=float(allf[:, 1].max()),
                   obs_rel=float(allf[:, 2].max()), obs_rel_excited=float(allf[:, 3].max()))
        results[cd] = res
        ok = ok and res["ground_bitwise"] and res["excited_l2"] < 1e-12 and res["obs_rel"] < 1e-12 and res["obs_rel_excited"] < 1e-11
        single.close()
        multi.close()
    # ---- ordering under drift: one rank's halo stream is stalled before every boundary pass, so
    # its neighbours run ahead.  The check after evolve reads the ghost planes (energy stencil) and normalises them in
    # place; a second evolve then starts from those ghosts.  Everything is compared with a single-GPU run of the same
    # lattice through the slab-independent checksum (bit-for-bit) and the observables (<= 1e-12).
    shape = (int(os.environ.get("WAFER_DRIFT_NX", 48 * world)), 160, 192)
    sweeps = int(os.environ.get("WAFER_DRIFT_SWEEPS", 300))
    dn = 10.24 / shape[1]
    kw = dict(dn=dn, dt=0.3 * dn * dn, mass=1.0, device=local)
    single = wafer_b200.Lattice(shape, "ThreePoint", **kw)
    multi = wafer_b200.Lattice(shape, "ThreePoint", rank=rank, world=world, nccl_id=fresh_nccl_id(), **kw)
    if use_p2p:
        connect(multi)
    if rank == 1 % world:
        multi.debug_halo_delay(int(os.environ.get("WAFER_DRIFT_DELAY_NS", 300000)))
    x0, x1 = multi.slab
    drift = dict(bitwise=[], obs_rel=[])
    for lat in (single, multi):
        lat.generate_potential("PoschlTeller")
        lat.set_initial_conditions("Boolean")
    norm2 = None
    for rnd in range(3):
        per = []
        for lat in (single, multi):
            if rnd == 2:
                # the stateless drop-in round trip: owned planes out to the host and back in, ghosts re-fetched
                lat.set_phi_owned(lat.get_phi_slab()) if lat is multi else lat.set_phi(lat.get_phi())
            lat.evolve(0, sweeps + rnd)  # rnd = 1: odd count -> the one-step tail pass takes part too
            # the observables kernel is queued straight behind the last sweep, with no host synchronisation in between:
            # it reads the ghost planes the (stalled) neighbour is still about to store
            obs = lat.compute_observables()
            ck = lat.phi_checksum(x0, x1)
            per.append((ck, obs))
        (cs, os_), (cm, om) = per
        drift["bitwise"].append(cs == cm)
        drift["obs_rel"].append(max(abs(os_[k] - om[k]) / max(abs(os_[k]), 1e-300) for k in os_))
        norm2 = os_["norm2"]  # the SAME divisor on both, so that the next round stays bit-comparable
        single.normalise_wavefunction(norm2)
        multi.normalise_wavefunction(norm2)
    t = torch.tensor([float(all(drift["bitwise"])), max(drift["obs_rel"])], dtype=torch.float64, device="cuda")
    allt = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allt, t)
    allt = torch.stack(allt).cpu().numpy()
    results["drift"] = dict(shape=list(shape), sweeps=sweeps, bitwise=bool(allt[:, 0].all()), obs_rel=float(allt[:, 1].max()),
                            final_wait=os.environ.get("WAFER_DEBUG_SKIP_FINAL_WAIT", "0") != "1")
    ok = ok and results["drift"]["bitwise"] and results["drift"]["obs_rel"] < 1e-12
    single.close()
    multi.close()
    if rank == 0:
        print(json.dumps({"world": world, "ok": ok, "halo": "p2p" if use_p2p else "nccl", "results": results}), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
