"""Summarise an .ncu-rep: headline metrics + stall breakdown + hottest SASS lines.
usage: python tools/ncu_summary.py report.ncu-rep [kernel_index]"""
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
r = data[kidx]
get = lambda k: r[hdr.index(k)] if k in hdr else None
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]
for k in keys:
    v = get(k)
    if v is not None:
        print(f"{k:82s} {v} {units[hdr.index(k)]}")
print("-- stalls per issue (ratio)")
for i, h in enumerate(hdr):
    m = re.match(r"smsp__average_warps_issue_stalled_(.*)_per_issue_active.ratio", h)
    if m and float(r[i] or 0) > 0.02:
        print(f"   {m.group(1):28s} {float(r[i]):.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
starts = [i for i, x in enumerate(rows) if x and x[0] == "Address"]
if starts:
    s = starts[kidx]
    e = starts[kidx + 1] - 1 if kidx + 1 < len(starts) else len(rows)
    h = rows[s]
    body = [x for x in rows[s + 1:e] if len(x) == len(h)]
    si, so, ex = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
    stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    tot = sum(int(x[si] or 0) for x in body)
    print("-- total samples", tot, " by opcode:")
    agg = {}
    for x in body:
        op = re.sub(r"@!?U?P\d+\s+", "", x[so].strip()).split()[0] if x[so].strip() else "?"
        agg[op] = agg.get(op, 0) + int(x[si] or 0)
    for op, v in sorted(agg.items(), key=lambda kv: -kv[1])[:14]:
        print(f"   {op:22s} {v:8d} {100.0 * v / tot:5.1f}%")
    print("-- stall reasons (all samples)")
    for i, c in stall_cols:
        v = sum(int(x[i] or 0) for x in body)
        if v > 0.01 * tot:
            print(f"   {c:26s} {v:8d} {100.0 * v / tot:5.1f}%")
    print("-- hottest non-DMMA instructions")
    nd = [x for x in body if "DMMA" not in x[so]]
    for x in sorted(nd, key=lambda x: -int(x[si] or 0))[:14]:
        reasons = ",".join(f"{c[6:]}={x[i]}" for i, c in stall_cols if int(x[i] or 0) > 0.25 * int(x[si] or 1))
        print(f"   {x[si]:>6s} {x[ex]:>9s} {x[so].strip()[:70]:70s} {reasons}")
