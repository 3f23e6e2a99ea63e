"""Digest of an .ncu-rep: headline metrics + hottest SASS lines by stall samples (run on the CPU box)."""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum", "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_selected",
        "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_branch_resolving",
        "smsp__pcsamp_warps_issue_stalled_lg_throttle", "smsp__pcsamp_warps_issue_stalled_sleeping", "smsp__pcsamp_warps_issue_stalled_membar"]
for i, h in enumerate(hdr):
    if h in want:
        print(f"{h} [{rows[1][i]}]: {[r[i][:90] for r in rows[2:]]}")
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
hi = [i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r]
if hi:
    h = rows[hi[0]]; si = h.index("Source"); ni = h.index("# Samples")
    end = hi[1] - 1 if len(hi) > 1 else len(rows)
    data = rows[hi[0] + 1:end]
    tot = sum(int(r[ni]) for r in data if len(r) > ni and r[ni].isdigit())
    print("total samples", tot)
    for s, idx, src in sorted([(int(r[ni]), k, r[si]) for k, r in enumerate(data) if len(r) > ni and r[ni].isdigit()], reverse=True)[:14]:
        print(f"  {s:6d} {100*s/tot:5.1f}%  #{idx:5d}  {src[:100]}")
