#!/usr/bin/env python
"""Turn an .ncu-rep (from `ncu --set full --import-source on`, see B200_PROFILING.md)
into the text summary committed under profiles/. Usage: summarize.py rep.ncu-rep out.txt [blocks]"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]
STALLS = ["stall_wait", "stall_short_sb", "stall_long_sb", "stall_branch_resolving", "stall_selected", "stall_sleep",
          "stall_barrier", "stall_membar", "stall_lg", "stall_mio", "stall_no_inst", "stall_dispatch", "stall_math"]


def ncu(rep, page):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    blocks = int(sys.argv[3]) if len(sys.argv) > 3 else None
    raw = list(csv.reader(io.StringIO(ncu(rep, "raw"))))
    h, units, v = raw[0], raw[1], raw[2]
    lines = ["# " + rep, "kernel: " + v[h.index("Kernel Name")] if "Kernel Name" in h else ""]
    vals = {}
    for k in KEYS:
        if k in h:
            i = h.index(k)
            vals[k] = v[i]
            lines.append("%-70s %s %s" % (k, v[i], units[i]))
    src = list(csv.reader(io.StringIO(ncu(rep, "source"))))
    hh, rows = src[1], src[2:]
    ix = {n: i for i, n in enumerate(hh)}
    tot_i = sum(int(x[ix["Instructions Executed"]]) for x in rows)
    tot_s = sum(int(x[ix["# Samples"]]) for x in rows)
    lines.append("warp instructions executed (source page): %d   stall samples: %d" % (tot_i, tot_s))
    if blocks:
        lines.append("per 64KB block: %.0f warp instructions, %.1f us of CTA time at this grid" % (
            tot_i / blocks, float(vals.get("gpu__time_duration.sum", 0)) * 1e3 * float(vals.get("launch__grid_size", 1)) / blocks))
    lines.append("stall reasons (share of samples):")
    for k in STALLS:
        if k in ix:
            s = sum(int(x[ix[k]]) for x in rows)
            if s:
                lines.append("  %-26s %5.1f%%" % (k, 100.0 * s / tot_s))
    op = collections.Counter()
    for x in rows:
        t = x[ix["Source"]].split()
        if not t:
            continue
        o = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
        op[o.split(".")[0]] += int(x[ix["Instructions Executed"]])
    lines.append("opcode mix (share of executed warp instructions):")
    for o, c in op.most_common(14):
        lines.append("  %-10s %5.1f%%" % (o, 100.0 * c / tot_i))
    tc = [o for o in op if o.startswith("UTC") or o in ("LDTM", "STTM", "UTMALDG", "UBLKCP", "HMMA")]
    lines.append("tensor/TMA opcodes present: %s (byte/integer kernel: none expected)" % (tc or "none"))
    open(out, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
