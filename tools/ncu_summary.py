"""Summarise an ncu report: headline metrics, opcode mix and hot SASS of the top kernel.
usage: python tools/ncu_summary.py <report.ncu-rep> <updates per launch> [--sass]"""
import collections, csv, io, subprocess, sys

rep, updates = sys.argv[1], float(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
        "lts__t_sectors_srcunit_tex_op_red.sum", "lts__t_sectors_srcunit_tex_op_atom.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio"]
for k in keys:
    if k in m:
        print(f"{k:95s} {m[k][0]:>18s} {m[k][1]}")
inst = float(m["smsp__inst_executed.sum"][0].replace(",", ""))
print(f"warp instructions per 32 updates: {inst / (updates / 32):.1f}  (thread-instr per update if all lanes active)")

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
isrc, iexe, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = rows[2:]
tot = sum(int(r[iexe]) for r in data); ts = sum(int(r[ismp]) for r in data)
hist, samp = collections.Counter(), collections.Counter()
for r in data:
    t = r[isrc].strip().split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    hist[op] += int(r[iexe]); samp[op] += int(r[ismp])
print("opcode      share   per-32-updates   stall-sample share")
for op, c in hist.most_common(24):
    print(f"{op:10s} {c / tot * 100:6.2f}%  {c / (updates / 32):8.2f}  {samp[op] / ts * 100:6.1f}%")
if "--sass" in sys.argv:
    thr = 0.3 * max(int(r[iexe]) for r in data)
    for i, r in enumerate(data):
        if int(r[iexe]) > thr:
            print(i, f"{int(r[iexe]) / (updates / 256):6.2f}", r[ismp], r[isrc].strip()[:100])
