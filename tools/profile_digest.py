#!/usr/bin/env python
"""Turn what tools/profile_r2.sh left in gpurun_out/ into the tracked evidence under profiles/:
  r2_launches.csv + r2_launches_summary.txt   the ncu launch list and its per-kernel shares
  r2_top_ncu_raw.txt                          selected ncu --set full metrics, one launch per kernel
  k2_traffic.json                             DRAM bytes and warp-instructions of one scoring-kernel launch
                                              (bench.py reads it for roofline.traffic and roofline.issue)
Usage: python tools/profile_digest.py [bench stage line for the summary header]
"""
import csv, json, os, re, shutil, subprocess, sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

KEYS = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__m_xbar2l1tex_read_sectors_mem_global_op_tma_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "launch__block_size", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__registers_per_thread",
        "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]


def short(name):
    return re.sub(r"\(.*", "", name).replace("void ", "")


def num(v):
    return float(v.replace(",", "")) if v not in ("", "n/a") else 0.0


def to_bytes(v, unit):
    return num(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def to_ms(v, unit):
    return num(v) * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1, "msecond": 1, "second": 1e3, "nsecond": 1e-6}.get(unit, 1)


def launches(note):
    src = os.path.join(OUT, "r2_launches.csv")
    if not os.path.exists(src):
        return
    shutil.copy(src, os.path.join(PROF, "r2_launches.csv"))
    rows = [r for r in csv.reader(open(src)) if len(r) > 14 and r[0].isdigit()]
    agg = OrderedDict()
    for r in rows:
        k = short(r[4])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += to_ms(r[14], r[13])
    total = sum(v[1] for v in agg.values())
    with open(os.path.join(PROF, "r2_launches_summary.txt"), "w") as f:
        f.write("# round 2 launch list: ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 (our kernels only; torch's\n"
                "# tile generator launches thousands of its own before them), python bench.py --tiles 2 --steps 1 --warmup 3 --no-e2e\n"
                "# --no-cpu-baseline --no-cli --no-extra: the first 400 launches = set-up + the first steps over two 1.07 M-cluster tiles.\n"
                "# Per-launch times under ncu are cold-cache and serialised: compare the SHARES with bench.py's stage times\n"
                "# (%s).\n# kernel, launches, total ms, ms per launch, share of the listed kernel time\n" % note)
        for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-48s %4d %10.3f %9.4f %6.1f%%\n" % (k, n, ms, ms / n, 100 * ms / total))
    print("launch list:", len(rows), "launches,", len(agg), "kernels")


def top():
    rep = os.path.join(OUT, "r2_top.ncu-rep")
    if not os.path.exists(rep):
        return
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    best = OrderedDict()            # per kernel (template instances apart): the longest launch captured
    for r in data:
        name = re.sub(r"^void ", "", r[ix["Kernel Name"]]).split("(")[0]
        ms = to_ms(r[ix["gpu__time_duration.sum"]], units[ix["gpu__time_duration.sum"]])
        if name not in best or ms > best[name][0]:
            best[name] = (ms, r)
    with open(os.path.join(PROF, "r2_top_ncu_raw.txt"), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on, one launch each (steady state), tile of 1,070,000 clusters (MiSeq shape)\n"
                "# command: tools/profile_r2.sh + tools/profile_digest.py (report gpurun_out/r2_top.ncu-rep, scratch); per-launch times here are\n"
                "# cold-cache / serialised: compare shares, not absolutes.  k_normalise_score_pack_compact<64> is the instance that does the work on\n"
                "# this tile; <128> finds nothing left for it and returns.\n")
        for name, (ms, r) in best.items():
            f.write("\n== %s\n" % name)
            for k in KEYS:
                if k in ix:
                    f.write("   %-72s %s %s\n" % (k, r[ix[k]], units[ix[k]]))
            rd = to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]])
            wr = to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
            f.write("   dram bytes (read+write) per launch: %.3f Gbyte\n" % ((rd + wr) / 1e9))
    k2 = [(n, v) for n, v in best.items() if "k_normalise_score_pack_compact" in n]
    if k2:
        tot_rd = tot_wr = tot_inst = 0.0
        for n, (ms, r) in k2:       # both instances belong to one pass over the tile
            tot_rd += to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]])
            tot_wr += to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
            tot_inst += num(r[ix["smsp__inst_executed.sum"]])
        old = json.load(open(os.path.join(PROF, "k2_traffic.json")))
        per = dict(old.get("per_kernel", {}))
        for name, (ms, r) in best.items():
            base = name.split("<")[0]
            if base in ("k_decode_demux_seed", "k_polya_ruler"):
                per[base] = to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]]) + \
                            to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
        per["k_normalise_score_pack_compact"] = tot_rd + tot_wr
        bench = {}
        try:
            bench = json.loads(open(os.path.join(OUT, "prof_bench.json")).read().strip().splitlines()[-1])
        except Exception:
            pass
        processed = int(round(bench.get("roofline", {}).get("processed_fraction", 0) * 1070000)) or old.get("processed_clusters")
        json.dump({"kernel": "k_normalise_score_pack_compact", "dram_bytes_per_launch": tot_rd + tot_wr,
                   "clusters_per_launch": 1070000, "processed_clusters": processed,
                   "warp_instructions_per_launch": int(tot_inst),
                   "source": "profiles/r2_top_ncu_raw.txt (ncu --set full of the round-2 build, both instances of the kernel: "
                             "dram__bytes_read.sum + dram__bytes_write.sum, smsp__inst_executed.sum)",
                   "read_write": {"read": tot_rd, "write": tot_wr}, "per_kernel": per},
                  open(os.path.join(PROF, "k2_traffic.json"), "w"), indent=1)
        print("k2: %.3f GB dram, %.3f G warp-instructions" % ((tot_rd + tot_wr) / 1e9, tot_inst / 1e9))


if __name__ == "__main__":
    launches(sys.argv[1] if len(sys.argv) > 1 else "see profiles/r2_bench_builder_runs.jsonl")
    top()
