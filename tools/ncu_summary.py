#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page + source page) into the numbers quoted in DESIGN.md / profiles/."""
import csv, subprocess, sys
from collections import Counter

def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]

def source(rep, kernel=None):
    """Source page of ONE kernel of the report (the first, or the first whose name contains `kernel`)."""
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    blocks, cur = [], None
    for r in csv.reader(out.splitlines()):
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            blocks.append(cur)
            continue
        if cur is None:
            cur = {"name": "?", "hdr": None, "rows": []}
            blocks.append(cur)
        if cur["hdr"] is None:
            if "Source" in r:
                cur["hdr"] = r
            continue
        if len(r) == len(cur["hdr"]):
            cur["rows"].append(r)
    pick = next((b for b in blocks if kernel and kernel in b["name"]), blocks[0])
    print("== source page of", pick["name"].split("(")[0])
    return pick["hdr"], pick["rows"]

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__maximum_warps_per_active_cycle_pct",
        "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum", "SM_A.TriageCompute.sm__inst_executed_pipe_xu_realtime.avg.pct_of_peak_sustained_elapsed"]

def main():
    rep = sys.argv[1]
    hdr, units, rows = raw(rep)
    for r in rows:
        name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print("== kernel:", name.split("(")[0])
        for i, h in enumerate(hdr):
            if h in KEYS:
                print("  %-80s %s %s" % (h, r[i], units[i]))
    shdr, data = source(rep, sys.argv[3] if len(sys.argv) > 3 else None)
    ix = {h: i for i, h in enumerate(shdr)}
    inst = [int(r[ix["Instructions Executed"]]) for r in data]
    samp = [int(r[ix["# Samples"]]) for r in data]
    tot, ts = sum(inst), max(1, sum(samp))
    c, cs, thr = Counter(), Counter(), Counter()
    for r in data:
        toks = [t for t in r[ix["Source"]].split() if not t.startswith("@")]
        key = toks[0] if toks else "?"
        c[key] += int(r[ix["Instructions Executed"]]); cs[key] += int(r[ix["# Samples"]])
        thr[key] += int(r[ix["Thread Instructions Executed"]])
    print("== SASS opcode mix (warp instructions %d, samples %d)" % (tot, ts))
    for k, v in c.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 28):
        print("  %-26s inst %6.2f%%  samples %6.2f%%  avg active threads %.1f" % (k, 100 * v / tot, 100 * cs[k] / ts, thr[k] / max(v, 1)))
    st = [h for h in shdr if h.startswith("stall_") and "Not Issued" not in h]
    tots = {h: sum(int(r[ix[h]]) for r in data) for h in st}
    print("== stall samples:", {k: v for k, v in sorted(tots.items(), key=lambda kv: -kv[1])[:8]})
    print("== instruction share per 100 SASS lines")
    for b in range(0, len(data), 100):
        if sum(inst[b:b + 100]) * 200 > tot:
            print("  %5d  %5.1f%% inst  %5.1f%% samples  max-exec %d" % (b, 100 * sum(inst[b:b + 100]) / tot, 100 * sum(samp[b:b + 100]) / ts, max(inst[b:b + 100])))

if __name__ == "__main__":
    main()
