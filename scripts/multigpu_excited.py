"""Excited-state step throughput on N GPUs (run under torchrun): 512^3 per GPU (weak scaling along x), k = 1..3 stored
states, NCCL halo exchange + one all-reduce of 1+k doubles per step (grid.rs:674-681).  Rank 0 prints one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import wafer_b200

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt = torch.tensor(list(wafer_b200.nccl_unique_id()), dtype=torch.uint8, device="cuda")
    dist.broadcast(idt, 0)
    n = int(os.environ.get("N", "512"))
    shape = (n * world, n, n)
    dn = 10.24 / n
    out = {"world": world, "lattice": list(shape)}
    with wafer_b200.Lattice(shape, "ThreePoint", dn=dn, dt=0.1 * dn * dn, mass=1.0, device=local, rank=rank, world=world,
                            nccl_id=bytes(idt.cpu().tolist())) as lat:
        lat.generate_potential("Harmonic")
        lat.set_initial_conditions("Boolean")
        lat.check(0)
        for k in (1, 2, 3):
            lat.push_lower()
            lat.phi_seed_from_lower(0)
            lat.check(k)
            lat.evolve(k, 5)
            lat.synchronize()
            dist.barrier()
            lat.timer_begin()
            lat.evolve(k, 20)
            ms = lat.timer_end()
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            out["excited_k%d_glups" % k] = shape[0] * n * n * 20 / (t.item() * 1e-3) / 1e9
            out["excited_k%d_frac_of_roofline" % k] = out["excited_k%d_glups" % k] * (48 + 16 * k) / (6552.0 * world)
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
