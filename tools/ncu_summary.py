#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): key metrics + executed-instruction mix per pixel-warp.
usage: tools/ncu_summary.py gpurun_out/prof_stack.ncu-rep [npix]"""
import collections, csv, subprocess, sys

rep = sys.argv[1]
npix = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 4096 * 4096
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
        "sm__cycles_elapsed.avg.per_second", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warp_latency_per_inst_issued.ratio"]
print("== kernel:", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "")
if "--traffic-json" in sys.argv:  # per-launch DRAM traffic for bench.py's roofline.traffic
    import json
    def _bytes(name):
        i = hdr.index(name)
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
        return float(vals[i].replace(",", "")) * mult
    rd, wr = _bytes("dram__bytes_read.sum"), _bytes("dram__bytes_write.sum")
    out = sys.argv[sys.argv.index("--traffic-json") + 1]
    json.dump({"kernel": vals[hdr.index("Kernel Name")], "dram_bytes_read": int(rd),
               "dram_bytes_write": int(wr), "dram_bytes_per_launch": int(rd + wr),
               "source": "ncu --set full --clock-control none, one launch of bench.py's step"},
              open(out, "w"), indent=1)
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w:85s} {vals[i]:>16s} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter()
stall_by_op = collections.Counter()
tot = 0
samples = 0
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    toks = r[ix["Source"]].split()
    op = toks[0] if not toks[0].startswith("@") else toks[1]
    op = op.split(".")[0]
    n = int(r[ix["Instructions Executed"]])
    ops[op] += n
    tot += n
    s = int(r[ix["# Samples"]])
    stall_by_op[op] += s
    samples += s
npw = npix / 32
print(f"== executed warp-instructions per 32-pixel warp: {tot / npw:.1f}")
for op, n in ops.most_common(22):
    print(f"   {op:10s} {n / npw:8.1f}   stall-samples {100 * stall_by_op[op] / max(samples, 1):5.1f}%")
stalls = collections.Counter()
for h in hdr:
    if h.startswith("stall_") and "Not Issued" not in h:
        stalls[h] = sum(int(r[ix[h]]) for r in rows[2:] if len(r) >= len(hdr))
tot_s = sum(stalls.values())
print("== stall reasons (all samples):", ", ".join(f"{k[6:]} {100 * v / tot_s:.1f}%" for k, v in stalls.most_common(8)))
