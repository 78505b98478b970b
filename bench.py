#!/usr/bin/env python
"""bench.py — f64 lattice updates/s (GLUPS) of the imaginary-time FDTD hot path.

  python bench.py --gpus N --steps K --warmup W          # B200 arm (this repo's CUDA path)
  python bench.py --impl reference --gpus N ...          # reference arm: the CPU restatement of the
                                                         # reference's rayon path on the host cores

Workload (BASELINE.json configs[3]): 1024^3, ThreePoint, gen_potential.py's Poschl-Teller potential, Boolean
initial condition, ground state; x-slab decomposed over N GPUs of one box (strong scaling).  One "step" is one
`evolve(wnum=0, screen_update)` call = SWEEPS lattice sweeps (grid.rs:544-687; default 1000 = wafer.yaml:98).  `value` counts
nx*ny*nz*SWEEPS*K updates over the max-over-ranks device time; `e2e` adds, every step, the host->device copy of
psi from pinned memory before evolve, one observables check (grid.rs:127-135) and the device->host copy of the evolved
psi after it (what a stateless drop-in of the reference's loop body has to do).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_PER_UPDATE = 32.0  # read psi, A, B + write psi', 8 B each
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--grid", type=int, default=1024, help="lattice edge N (N^3 work sites)")
    ap.add_argument("--nx", type=int, default=0,
                    help="x extent (the decomposed axis) if not N: e.g. --grid 2048 --nx 256 is one GPU's share of C5")
    ap.add_argument("--sweeps", type=int, default=1000,
                    help="lattice sweeps per step = output.screen_update; 1000 is the reference's default (wafer.yaml:98)")
    ap.add_argument("--stencil", default="ThreePoint", choices=["ThreePoint", "FivePoint", "SevenPoint"])
    ap.add_argument("--flags", type=int, default=0, help="wafer_params.flags (1 = A/B arrays, 4 = simple sweep)")
    ap.add_argument("--no-p2p", action="store_true", help="multi-GPU: NCCL send/recv halos instead of fused peer stores")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-512", action="store_true")
    ap.add_argument("--cpu-grid", type=int, default=512)
    return ap.parse_args()


def physical_params(n):
    """dn, dt, mass for the workload: box of width 10.24 like BASELINE C4 (dn 0.01 at 1024), dt = dn^2/3.33."""
    dn = 10.24 / n
    return dn, 0.3 * dn * dn, 1.0


# ------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU during the timed region (NVML, 100 ms period)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self._stop, self._t = [], set(), None, threading.Event(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.1)

    def start(self):
        if self.nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join()
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------------- CPU arm
def cpu_arm(args, steps, warmup, all_threads=True, budget_s=20.0):
    """The reference's CPU path (its Rust cannot be built here — SURVEY F2/F3 — so: the C++ restatement with the
    reference's own pass structure, oracle/wafer_oracle.cpp) timed on the host cores on a bounded sample."""
    from oracle import binding as oracle
    n = args.cpu_grid
    dn, dt, mass = physical_params(args.grid)
    ext = {"ThreePoint": 1, "FivePoint": 2, "SevenPoint": 3}[args.stencil]
    cores = os.cpu_count() or 1
    if all_threads:
        oracle.set_num_threads(cores)
    g = oracle.make_grid(n, n, n, ext=ext, dn=dn, dt=dt, mass=mass)
    v = oracle.potential(g, "PoschlTeller")
    a, b = oracle.build_ab(v, dt)
    phi = oracle.initial_condition(g, "Boolean")
    t0 = time.perf_counter()
    oracle.evolve(g, phi, a, b, 1)
    t1 = time.perf_counter() - t0
    per_step = max(1, min(50, int(budget_s / max(steps + warmup, 1) / max(t1, 1e-3))))
    for _ in range(warmup):
        oracle.evolve(g, phi, a, b, per_step)
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle.evolve(g, phi, a, b, per_step)
    el = time.perf_counter() - t0
    glups = n ** 3 * per_step * steps / el / 1e9
    return {"value": glups, "unit": "GLUPS", "cores": oracle.num_threads(), "kind": "port",
            "sample": "%d^3 sub-lattice of the workload, %d steps x %d sweeps of oracle evolve (stencil into work + "
                      "copy-back, grid.rs:560-673), OpenMP on %d threads" % (n, steps, per_step, oracle.num_threads()),
            "ms_per_step": el / steps * 1e3, "sweeps_per_step": per_step}


def reference_main(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_arm(args, args.steps, args.warmup, all_threads=True, budget_s=60.0)
    line = {
        "impl": "reference", "metric": "f64 lattice updates/s", "value": cb["value"], "unit": "GLUPS",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": "GLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    n = args.grid
    nx = args.nx or n
    return {"workload": "C4: %dx%dx%d %s, Poschl-Teller (gen_potential.py formula, lam=6), Boolean IC, ground state; "
                        "step = evolve(wnum=0, screen_update=%d)" % (nx, n, n, args.stencil, args.sweeps),
            "grid": [nx, n, n], "stencil": args.stencil, "sweeps_per_step": args.sweeps,
            "decomposition": ("x-slab x%d, halo: %s" % (world, "NCCL send/recv" if args.no_p2p else "fused peer stores (CUDA IPC over NVLink)"))
            if world > 1 else "single GPU",
            "l2": "inputs larger than L2 (%.1f GB per field per GPU vs 126 MB)" % (nx * n * n * 8 / world / 1e9)}


# ------------------------------------------------------------------------------------------------- B200 arm
def measure(lat, args, n, world, dist, steps, warmup, sampler=None):
    """W untimed + K timed evolve(0, sweeps) calls, device-timed with CUDA events on the library's stream."""
    for _ in range(warmup):
        lat.evolve(0, args.sweeps)
    lat.synchronize()
    if dist is not None:
        dist.barrier()
    if sampler:
        sampler.start()
    launches0 = lat.kernel_launches
    lat.timer_begin()
    for _ in range(steps):
        lat.evolve(0, args.sweeps)
    ms = lat.timer_end()
    lat.synchronize()
    launches = lat.kernel_launches - launches0
    clocks = sampler.stop() if sampler else None
    if dist is not None:
        import torch
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        ms = float(t.item())
    glups = sites(args, n) * args.sweeps * steps / (ms * 1e-3) / 1e9
    return glups, ms, launches, clocks


def sites(args, n):
    return (args.nx or n) * n * n if n == args.grid else n ** 3


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic(variant, updates_per_launch):
    """dram__bytes_read+write per launch of the sweep kernel, from the committed `ncu --set full` capture of the
    same kernel (profiles/ncu_traffic.json: measured at 512^3, scaled by updates per launch)."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            for rec in json.load(f):
                if variant.startswith(rec["variant"]):
                    return rec["dram_bytes_per_update"] * updates_per_launch, rec["note"]
    except Exception:
        pass
    return None, None


def b200_main(args):
    import numpy as np

    import wafer_b200

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    nccl_id = None
    if world > 1:
        import torch
        import torch.distributed as td
        torch.cuda.set_device(local_rank)
        td.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = td
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.tensor(list(wafer_b200.nccl_unique_id()), dtype=torch.uint8, device="cuda")
        td.broadcast(idt, 0)
        nccl_id = bytes(idt.cpu().tolist())
    if world != args.gpus and rank == 0:
        sys.stderr.write("warning: --gpus %d but WORLD_SIZE %d; using WORLD_SIZE\n" % (args.gpus, world))

    n = args.grid
    dn, dt, mass = physical_params(n)
    nsites = sites(args, n)
    lat = wafer_b200.Lattice((args.nx or n, n, n), args.stencil, dn=dn, dt=dt, mass=mass, device=local_rank, rank=rank, world=world,
                             nccl_id=nccl_id, flags=args.flags)
    if world > 1 and not args.no_p2p:
        # fused halo: boundary CTAs store into the neighbours' ghost planes through CUDA-IPC mapped peer memory
        import torch
        mine = torch.tensor(list(lat.p2p_export()), dtype=torch.uint8, device="cuda")
        blobs = [torch.zeros(192, dtype=torch.uint8, device="cuda") for _ in range(world)]
        dist.all_gather(blobs, mine)
        blobs = [bytes(b.cpu().tolist()) for b in blobs]
        lat.p2p_connect(blobs[rank - 1] if rank > 0 else None, blobs[rank + 1] if rank < world - 1 else None)
    lat.generate_potential("PoschlTeller")
    lat.set_initial_conditions("Boolean")
    lat.check(0)  # normalise once so that thousands of sweeps stay in range

    sampler = ClockSampler(local_rank) if rank == 0 else None
    glups, ms, launches, clocks = measure(lat, args, n, world, dist, args.steps, args.warmup, sampler)
    sweeps_total = args.sweeps * args.steps

    peak, peak_src = hbm_peak()
    # dominant kernel = the sweep.  The time-tiled kernel advances TWO lattice steps per launch, the plain one a
    # single step; multi-GPU runs add two small boundary-plane launches per pass on the halo stream.  The figure is
    # per main-stream launch: the timed region holds nothing but back-to-back sweep launches.
    steps_per_launch = 2 if lat.sweep_variant.startswith("tb2") else 1
    n_main = sweeps_total // steps_per_launch + sweeps_total % steps_per_launch
    launch_ms = ms / n_main
    updates_per_launch = nsites / world * sweeps_total / n_main
    achieved = BYTES_PER_UPDATE * updates_per_launch / (launch_ms * 1e-3) / 1e9
    traffic, traffic_note = ncu_traffic(lat.sweep_variant, updates_per_launch)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_src, "kernel": lat.sweep_variant,
                "algorithmic_bytes_per_update": BYTES_PER_UPDATE, "updates_per_launch": updates_per_launch,
                "steps_per_launch": steps_per_launch, "launch_ms": launch_ms,
                "dram_frac_of_peak": (traffic / (launch_ms * 1e-3) / 1e9 / peak) if traffic else None,
                "frac_of_nominal_8TBps": achieved / 8000.0}

    # ---- e2e: host buffers in and out of every step (pinned), through the C ABI
    e2e = None
    if not args.no_e2e:
        p0, p1 = lat.slab_planes(0)
        q0, q1 = lat.slab_planes(1)
        shape_yz = lat.padded_shape[1:]
        planes = max(p1 - p0, q1 - q0)
        host = wafer_b200.pinned_empty((planes,) + shape_yz)
        h_in = host[:p1 - p0]
        h_out = host[:q1 - q0]
        # the chunk this rank would hold of the global array: current psi + ghost planes (zero ring at the ends)
        h_in[...] = 0.0
        tmp = lat.get_phi_slab()
        h_in[q0 - p0:q0 - p0 + (q1 - q0)] = tmp
        del tmp
        e_steps, e_warm = min(args.steps, 3), 1
        tot = 0.0
        for it in range(e_warm + e_steps):
            if dist is not None:
                dist.barrier()
            t0 = time.perf_counter()
            lat.timer_begin()
            lat.set_phi_slab(h_in)
            lat.evolve(0, args.sweeps)
            obs = lat.check(0)  # the loop body of grid.rs:126-221: observables + normalise, 4 scalars back to the host
            lat.get_phi_slab(h_out)
            ms_e = lat.timer_end()
            wall = (time.perf_counter() - t0) * 1e3
            ms_e = max(ms_e, wall)  # the copies synchronise the host: count whichever clock saw more
            if dist is not None:
                import torch
                t = torch.tensor([ms_e], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms_e = float(t.item())
            if it >= e_warm:
                tot += ms_e
            # feed the evolved state back in (keeps the ring zero and the values finite)
            h_in[q0 - p0:q0 - p0 + (q1 - q0)] = h_out
        e2e = {"value": nsites * args.sweeps * e_steps / (tot * 1e-3) / 1e9, "unit": "GLUPS",
               "h2d_bytes_per_step": int(h_in.size * 8 * world), "d2h_bytes_per_step": int(h_out.size * 8 * world),
               "steps": e_steps, "ms_per_step": tot / e_steps,
               "call": "wafer_set_phi_slab(pinned) -> wafer_evolve(0, %d) -> wafer_check(0) -> wafer_get_phi_slab(pinned)"
                       % args.sweeps, "last_energy": obs["energy"] / obs["norm2"]}
        wafer_b200.pinned_free(host)

    info = lat.device_info()
    variant = lat.sweep_variant
    lat.close()

    extra = {}
    if rank == 0 and world == 1 and not args.no_512 and n != 512 and not args.nx:
        # BASELINE metric's other quoted point: 512^3 on one GPU
        a2 = argparse.Namespace(**vars(args))
        a2.grid = 512
        dn2, dt2, m2 = physical_params(512)
        with wafer_b200.Lattice((512,) * 3, args.stencil, dn=dn2, dt=dt2, mass=m2, device=local_rank, flags=args.flags) as l2:
            l2.generate_potential("PoschlTeller")
            l2.set_initial_conditions("Boolean")
            l2.check(0)
            g512, ms512, _, _ = measure(l2, a2, 512, 1, None, max(args.steps, 3), 3)
        extra["glups_512cubed_1gpu"] = g512
        extra["hbm_frac_512cubed"] = g512 * BYTES_PER_UPDATE / peak

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_arm(args, steps=2, warmup=1, all_threads=True, budget_s=15.0)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {
            "metric": "f64 lattice updates/s", "value": glups, "unit": "GLUPS", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "device": info["name"], "sweep_variant": variant, "extra": extra,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        reference_main(a)
    else:
        b200_main(a)
