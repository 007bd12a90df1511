#!/usr/bin/env python
"""Summarise an ncu report: headline metrics (raw page) and warp-stall samples per source line (source page).

    python tools/ncu_stalls.py report.ncu-rep [top_n]
"""
import csv, io, subprocess, sys, collections, re
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warp_latency_issue_stalled_barrier.ratio", "smsp__warp_issue_stalled_barrier_per_warp_active.pct"]
for r in rows[2:]:
    for w in want:
        if w in hdr:
            print(f"{w} = {r[hdr.index(w)]} {rows[1][hdr.index(w)]}")
    for i, h in enumerate(hdr):
        if "warp_issue_stalled" in h and "per_warp_active" in h:
            try:
                v = float(r[i])
            except ValueError:
                continue
            if v > 3: print(f"  stall {h.split('stalled_')[1].split('_per')[0]} = {v:.1f} %")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
h = rows[hi]
samp = h.index("# Samples"); inst = h.index("Instructions Executed")
stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
conf = h.index("L1 Wavefronts Shared Excessive") if "L1 Wavefronts Shared Excessive" in h else None
tot = 0; tot_i = 0; per = []
for r in rows[hi + 1:]:
    if len(r) <= samp or not r[0].isdigit(): continue
    try: sm = int(r[samp] or 0); ie = int(r[inst] or 0)
    except ValueError: continue
    tot += sm; tot_i += ie
    stalls = {h[i][6:]: int(r[i]) for i in stall_cols if r[i].isdigit() and int(r[i]) > 0}
    ex = int(r[conf]) if conf is not None and r[conf].isdigit() else 0
    per.append((sm, int(r[0]), r[1].strip()[:95], ie, stalls, ex))
print("total samples", tot, "total warp instructions", tot_i)
print("---- by stall samples")
for sm, ln, line, ie, st, ex in sorted(per, key=lambda t: -t[0])[:top]:
    top3 = " ".join(f"{k}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{100.0 * sm / max(tot, 1):5.1f}% L{ln:4d} ie={100.0 * ie / max(tot_i, 1):4.1f}% {line} | {top3}" + (f" | smem-excess-wavefronts={ex}" if ex else ""))
print("---- by instructions executed")
for sm, ln, line, ie, st, ex in sorted(per, key=lambda t: -t[3])[:top]:
    print(f"{100.0 * ie / max(tot_i, 1):5.1f}% L{ln:4d} samples={100.0 * sm / max(tot, 1):4.1f}% {line}")
if len(sys.argv) > 3:  # phase table: "name:lo-hi,name:lo-hi"
    print("---- phases (share of warp instructions, share of stall samples)")
    for spec in sys.argv[3].split(","):
        name, rng = spec.split(":"); lo, hi_ = map(int, rng.split("-"))
        ie = sum(t[3] for t in per if lo <= t[1] <= hi_); sm = sum(t[0] for t in per if lo <= t[1] <= hi_)
        print(f"{name:14s} L{lo}-{hi_}: instr {100.0 * ie / max(tot_i, 1):5.1f}%  samples {100.0 * sm / max(tot, 1):5.1f}%")
