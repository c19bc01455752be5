"""Summarise an ncu report (--set full) into the text kept under profiles/: per launch the duration, DRAM traffic,
DRAM / L2 / SM / tensor-pipe utilisation, occupancy limits and the top warp-stall reasons.
Usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x_summary.txt   (runs here, no GPU needed)"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of peak"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("launch__registers_per_thread", "registers/thread"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__cluster_size", "cluster"), ("launch__occupancy_limit_shared_mem", "CTAs/SM (smem limit)"),
        ("launch__waves_per_multiprocessor", "waves/SM"), ("smsp__issue_active.avg.per_cycle_active", "issue slots busy")]
print("# ncu --set full --clock-control none (cold caches, serialised replays: compare shares, not absolutes)")
for r in data:
    print("\n== %s" % r[col["Kernel Name"]])
    for key, name in want:
        if key in col:
            print("   %-30s %14s %s" % (name, r[col[key]], units[col[key]]))
    stalls = []
    for h, i in col.items():
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
            try:
                stalls.append((float(r[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    stalls.sort(reverse=True)
    print("   top stalls (warps per issue): " + ", ".join("%s %.2f" % (n, v) for v, n in stalls[:5]))
