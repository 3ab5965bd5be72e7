"""Summaries of the ncu captures brought back in gpurun_out/ -> profiles/ (tracked).
    python profiles/scripts/summarize_ncu.py launches gpurun_out/r02_launches.csv > profiles/r02_launch_summary.txt
    python profiles/scripts/summarize_ncu.py full gpurun_out/r02_coupling_tc5.ncu-rep > profiles/r02_coupling_tc5_ncu_summary.txt"""
import csv, io, subprocess, sys
from collections import OrderedDict

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "dram__bytes_write.sum.per_second", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.per_cycle_active", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]

mode, path = sys.argv[1], sys.argv[2]
if mode == "launches":
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        name = r[ki][:60]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", "")) / 1000.0     # ns -> us
    tot = sum(v[1] for v in agg.values())
    print("ncu launch list (per-launch times are cold-cache and serialised under ncu: compare SHARES, not absolutes)\n")
    print("%-62s %8s %12s %8s" % ("kernel", "launches", "total us", "share"))
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-62s %8d %12.1f %7.1f%%" % (name, n, us, 100 * us / tot))
    print("%-62s %8d %12.1f" % ("TOTAL", sum(v[0] for v in agg.values()), tot))
else:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for key in ["Kernel Name"] + KEYS:
        if key in hdr:
            i = hdr.index(key)
            print("%-88s %-16s %s" % (key, units[i], " | ".join(r[i] for r in rows[2:])))
