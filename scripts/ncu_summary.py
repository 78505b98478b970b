#!/usr/bin/env python
"""Summarise a gpurun session's ncu output into profiles/ (tracked).
usage: python scripts/ncu_summary.py gpurun_out/<label> profiles/<name>
Writes <name>_launches.txt (per-kernel share of the step from the gpu__time_duration launch list) and
<name>_sweep.txt (the ncu --set full metrics the roofline line quotes) and prints the dram bytes per launch."""
import collections
import csv
import json
import os
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__block_size", "launch__grid_size",
        "launch__waves_per_multiprocessor", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg.per_second",
        "dram__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum", "smsp__warps_eligible.avg.per_cycle_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]


def main():
    src, dst = sys.argv[1], sys.argv[2]
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    lc = os.path.join(src, "launches.csv")
    if os.path.exists(lc):
        rows = [r for r in csv.reader(open(lc)) if len(r) > 10]
        hdr = rows[0]
        ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
        agg = collections.OrderedDict()
        for r in rows[1:]:
            try:
                v = float(r[vi].replace(",", ""))
            except ValueError:
                continue
            a = agg.setdefault(r[ki].split("(")[0][:70], [0, 0.0])
            a[0] += 1
            a[1] += v
        tot = sum(v[1] for v in agg.values())
        with open(dst + "_launches.txt", "w") as f:
            f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
            f.write("# command: see scripts/gpu_round.sh stage 'ncu'\n")
            f.write("%-72s %6s %12s %7s %10s\n" % ("kernel", "n", "total_ms", "share", "avg_us"))
            for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
                f.write("%-72s %6d %12.3f %6.1f%% %10.1f\n" % (k, v[0], v[1] / 1e6, 100 * v[1] / tot, v[1] / v[0] / 1e3))
    reps = [f for f in os.listdir(src) if f.endswith(".ncu-rep")]
    out = {}
    for rep in reps:
        raw = subprocess.run(["ncu", "-i", os.path.join(src, rep), "--page", "raw", "--csv"], capture_output=True,
                             text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        with open(dst + "_" + rep.replace(".ncu-rep", "") + ".txt", "w") as f:
            f.write("# ncu --set full --clock-control none --import-source on (one capture per launch listed)\n")
            for r in rows[2:]:
                f.write("\n== %s  grid %s block %s\n" % (r[hdr.index("Kernel Name")].split("(")[0], r[hdr.index("Grid Size")],
                                                        r[hdr.index("Block Size")]))
                for k in KEYS:
                    if k in hdr:
                        f.write("%-75s %22s %s\n" % (k, r[hdr.index(k)], units[hdr.index(k)]))
                for i, h in enumerate(hdr):
                    if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                        try:
                            if float(r[i]) > 0.2:
                                f.write("%-75s %22s %s\n" % (h.replace("smsp__average_warps_issue_stalled_", "stall:"), r[i], units[i]))
                        except ValueError:
                            pass
                rd = float(r[hdr.index("dram__bytes_read.sum")])
                wr = float(r[hdr.index("dram__bytes_write.sum")])
                scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
                tot = rd * scale[units[hdr.index("dram__bytes_read.sum")]] + wr * scale[units[hdr.index("dram__bytes_write.sum")]]
                f.write("dram bytes per launch (read+write)                                           %22.0f byte\n" % tot)
                out[r[hdr.index("Kernel Name")].split("(")[0]] = tot
    print(json.dumps(out))


if __name__ == "__main__":
    main()
