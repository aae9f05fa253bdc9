#!/usr/bin/env python
"""Summarise an ncu report (read here, not on the GPU box): per kernel launch the headline metrics, the SASS opcode
histogram weighted by executed instructions and stall samples, and the hottest stall addresses.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [samples_per_launch_for_per_sample_column]
"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.sum", "sm__inst_executed.sum.per_cycle_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum", "l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum",
        "smsp__inst_executed_op_global_red.sum", "lts__t_bytes.sum.per_second", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
name_i = hdr.index("Kernel Name")
for r in rows[2:]:
    print("==", r[name_i][:90])
    for i, h in enumerate(hdr):
        if h in WANT:
            print(f"   {h:78s} {r[i]} {rows[1][i]}")
    stalls = [(float(r[i]), h) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and r[i]]
    print("   stalls/issue:", ", ".join(f"{h[34:-23]} {v:.2f}" for v, h in sorted(stalls, reverse=True)[:7]))

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
blocks = src.split('"Kernel Name"')
for blk in blocks[1:]:
    rr = list(csv.reader(io.StringIO('"Kernel Name"' + blk)))
    kname = rr[0][1]
    hi = [i for i, r in enumerate(rr) if r and r[0] == "Address"][0]
    h = rr[hi]
    iI, iS = h.index("Instructions Executed"), h.index("# Samples")
    ins = [(r[1].strip(), int(r[iI]), int(r[iS])) for r in rr[hi + 1:] if len(r) > iI and r[iI].isdigit()]
    tot, ts = sum(i[1] for i in ins), max(1, sum(i[2] for i in ins))
    agg = collections.defaultdict(lambda: [0, 0])
    for s, c, sm in ins:
        op = re.sub(r'^@!?U?P\d+\s+', '', s).split()[0]
        agg[".".join(op.split(".")[:2]) if op.startswith(("RED", "ATOM", "LDG", "UTC", "LDTM", "STTM", "UBLKCP")) else op.split(".")[0]][0] += c
        agg[".".join(op.split(".")[:2]) if op.startswith(("RED", "ATOM", "LDG", "UTC", "LDTM", "STTM", "UBLKCP")) else op.split(".")[0]][1] += sm
    print("==", kname[:90], "warp instructions", tot, "stall samples", ts)
    for op, (c, sm) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:22]:
        print(f"   {op:14s} {c / tot * 100:6.2f}% inst {sm / ts * 100:6.2f}% samp {c:12d}")
    print("   hottest:")
    for s, c, sm in sorted(ins, key=lambda x: -x[2])[:12]:
        print(f"   {sm / ts * 100:5.2f}% samp {c:10d} exec  {s[:100]}")
