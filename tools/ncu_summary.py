#!/usr/bin/env python
"""summarise an .ncu-rep: key raw metrics per kernel + (optionally) the opcode mix of the source page."""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("kernel:", d.get("Kernel Name", "?")[:90])
    for w in want:
        if w in d:
            print("  %-85s %s" % (w, d[w]))
if len(sys.argv) > 2:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    # find header row
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    idx = {h: i for i, h in enumerate(rows[hi])}
    data = [r for r in rows[hi + 1:] if len(r) == len(rows[hi]) and r[0] != "Address"]
    tot = sum(int(r[idx["# Samples"]]) for r in data) or 1
    g = collections.defaultdict(lambda: [0, 0, 0, 0])
    for r in data:
        toks = r[idx["Source"]].split()
        if not toks: continue
        op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
        parts = op.split(".")
        key = parts[0] + ("." + ".".join(parts[1:3]) if parts[0] in ("ATOMS", "LDS", "LDGSTS", "STS", "LDG", "STG", "RED", "ATOMG", "SHFL", "REDUX") else "")
        a = g[key]
        a[0] += int(r[idx["# Samples"]]); a[1] += int(r[idx["Instructions Executed"]]); a[2] += int(r[idx["L1 Wavefronts Shared"]]); a[3] += int(r[idx["L1 Wavefronts Shared Ideal"]])
    print("opcode mix (samples share, warp instructions, shared wavefronts, ideal):")
    for k, a in sorted(g.items(), key=lambda kv: -kv[1][0])[:22]:
        print("  %-18s %5.1f%% inst %11d wf %11d ideal %11d" % (k, 100 * a[0] / tot, a[1], a[2], a[3]))
