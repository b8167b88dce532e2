"""Summarise an .ncu-rep (one kernel, --set full, --import-source on) into a small text file:
python tools/ncu_summary.py <rep> <out.txt> "<title / command / algorithmic work note>" """
import csv, io, subprocess, sys
from collections import Counter

rep, out, note = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum"]
lines = [note, ""]
stall = {}
for h, u, v in zip(hdr, units, vals):
    if h in KEEP:
        lines.append(f"{h} = {v} {u}")
    if h.startswith("smsp__pcsamp_warps_issue_stalled") and "not_issued" not in h:
        stall[h.replace("smsp__pcsamp_warps_issue_stalled_", "")] = int(v)
tot = sum(stall.values()) or 1
lines += ["", "warp stall samples: " + ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in sorted(stall.items(), key=lambda x: -x[1])[:8])]
rows = list(csv.reader(io.StringIO(src)))
if len(rows) > 2:
    h2 = rows[1]
    isrc, isamp, iex = h2.index("Source"), h2.index("# Samples"), h2.index("Instructions Executed")
    data, ops = [], Counter()
    for r in rows[2:]:
        try:
            data.append((int(r[isamp]), int(r[iex]), r[isrc]))
        except (ValueError, IndexError):
            continue
    t2 = sum(d[0] for d in data) or 1
    for s, ex, text in data:
        parts = text.split()
        op = (parts[1] if parts and parts[0].startswith("@") and len(parts) > 1 else (parts[0] if parts else "?")).split(".")[0]
        ops[op] += s
    lines += ["", "samples by opcode: " + ", ".join(f"{k} {100 * v / t2:.1f}%" for k, v in ops.most_common(12)), "",
              "top sampled SASS instructions (samples, share, times executed, instruction):"]
    for s, ex, text in sorted(data, key=lambda x: -x[0])[:14]:
        lines.append(f"  {s:8d} {100 * s / t2:5.1f}%  x{ex:<11d} {text[:90]}")
    proof = Counter()
    for _, _, text in data:
        for m in ("UTCHMMA", "LDTM", "UBLKCP", "UTMALDG", "UTCBAR", "SYNCS", "FFMA", "LDS.128", "REDUX", "SHFL"):
            if m in text:
                proof[m] += 1
    lines += ["", "SASS mnemonics present (static count): " + ", ".join(f"{k} x{v}" for k, v in proof.most_common())]
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
