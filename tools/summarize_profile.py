#!/usr/bin/env python3
"""Turn the ncu artefacts brought back in gpurun_out/ into the tracked summaries under profiles/.

    python tools/summarize_profile.py r01 gpurun_out/launches_r01.csv gpurun_out/prof_expand_r01.ncu-rep

 * profiles/<round>_launches_summary.json   per-kernel launch counts, device time shares, DRAM bytes of one
                                            full step (from the --metrics launch list)
 * profiles/<round>_k_expand_full.md        key counters of the `ncu --set full` capture of k_expand and
                                            the hottest source lines (SASS instruction counts mapped back
                                            through -lineinfo)
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def launches(round_, path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    kn, mn, mv, idc = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    launch = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        k = int(r[idc])
        launch.setdefault(k, {"name": r[kn].split("(")[0].replace("void ", "").replace("ambi::", "").split("<")[0]})
        launch[k][r[mn]] = float(r[mv].replace(",", ""))
    L = list(launch.values())
    idx = [i for i, l in enumerate(L) if l["name"] == "k_bump_gen"]
    step = L[idx[0]:idx[1]] if len(idx) > 1 else L[idx[0]:]
    agg = collections.OrderedDict()
    for l in step:
        d = agg.setdefault(l["name"], {"launches": 0, "time_us": 0.0, "dram_read_MB": 0.0, "dram_write_MB": 0.0, "warp_inst": 0.0})
        d["launches"] += 1
        d["time_us"] += l["gpu__time_duration.sum"] / 1e3
        d["dram_read_MB"] += l.get("dram__bytes_read.sum", 0) / 1e6
        d["dram_write_MB"] += l.get("dram__bytes_write.sum", 0) / 1e6
        d["warp_inst"] += l.get("smsp__inst_executed.sum", 0)
    tot = sum(d["time_us"] for d in agg.values())
    for d in agg.values():
        d["share"] = d["time_us"] / tot
    ex = [l for l in step if l["name"] == "k_expand"]
    tsum = sum(l["gpu__time_duration.sum"] for l in ex)
    out = {"round": round_, "complete_step": len(idx) > 1,
           "note": "one full step (" + os.environ.get("AMBI_PROFILE_BATCH", "32768") + " partitions n=12 / 4 folds: greedy + beam + exact pass) of `AMBI_NO_GRAPH=1 ncu --metrics "
                   "gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,... --clock-control none python bench.py "
                   "--steps 2 --warmup 1 --no-cpu-baseline` (graph replay disabled so that ncu sees the kernels). Per-launch times "
                   "are cold-cache and serialised: compare shares, not absolutes.",
           "step_total_us_under_ncu": tot, "per_kernel": agg,
           "k_expand": {"launches_per_step": len(ex),
                        "dram_bytes_per_launch": sum(l.get("dram__bytes_read.sum", 0) + l.get("dram__bytes_write.sum", 0) for l in ex) / len(ex),
                        "issue_active_pct_time_weighted": sum(l.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0) * l["gpu__time_duration.sum"] for l in ex) / tsum,
                        "threads_per_inst": sum(l.get("smsp__thread_inst_executed_per_inst_executed.ratio", 0) * l.get("smsp__inst_executed.sum", 0) for l in ex) / max(1.0, sum(l.get("smsp__inst_executed.sum", 0) for l in ex)),
                        "per_launch_us": [round(l["gpu__time_duration.sum"] / 1e3, 1) for l in ex]}}
    p = os.path.join(ROOT, "profiles", f"{round_}_launches_summary.json")
    json.dump(out, open(p, "w"), indent=1)
    print("wrote", p, "k_expand share", agg["k_expand"]["share"])


KEYS = ["sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def full(round_, rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    md = [f"# {round_}: `ncu --set full --clock-control none --import-source on -k regex:k_expand` (exact-pass levels of one step)\n",
          "Command: `" + os.environ.get("AMBI_PROFILE_CMD", "AMBI_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:k_expand -s 26 -c 2 python tools/prof_target.py 32768") + "`;",
          "the same command exited 0 without ncu directly before.  Numbers under a profiler are not bench values.\n"]
    stall = [k for k in hdr if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")]
    for r in rows[2:]:
        md.append(f"## launch id {r[0]}: {r[hdr.index('Kernel Name')]}\n")
        md.append("| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in hdr:
                md.append(f"| `{k}` | {r[hdr.index(k)]} | {units[hdr.index(k)]} |")
        md.append("\nWarp stall reasons (warps per issue-active cycle): " +
                  ", ".join("%s=%.2f" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""),
                                         float(r[hdr.index(k)])) for k in stall if float(r[hdr.index(k)]) >= 0.1) + "\n")
    # hottest source lines of the first launch: ncu's own SASS -> CUDA-C mapping (-lineinfo, --import-source on)
    src_csv = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                             capture_output=True, text=True).stdout
    first, cur_file, hdr2, launch = None, None, None, -1
    agg = collections.OrderedDict()
    for r in csv.reader(src_csv.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            first = first or cur_file
            if cur_file == first:
                launch += 1
            continue
        if r[0] == "Line No":
            hdr2 = r; continue
        if hdr2 and len(r) == len(hdr2) and launch == 0:
            try:
                ln = int(r[0])
            except ValueError:
                continue
            a = agg.setdefault((cur_file, ln, r[1].strip()), [0, 0, 0])
            a[0] += int(r[hdr2.index("Instructions Executed")]); a[1] += int(r[hdr2.index("Thread Instructions Executed")])
            a[2] += int(r[hdr2.index("# Samples")])
    tot = sum(v[0] for v in agg.values()) or 1
    tots = sum(v[2] for v in agg.values()) or 1
    md.append(f"## hottest source lines of launch 0 ({tot} warp instructions)\n")
    md.append("| % warp inst | % stall samples | threads / inst | file:line | source |\n|---|---|---|---|---|")
    for (f, l, text), v in sorted(agg.items(), key=lambda t: -t[1][0])[:30]:
        md.append("| %.1f | %.1f | %.1f | %s:%d | `%s` |" % (100 * v[0] / tot, 100 * v[2] / tots, v[1] / max(v[0], 1), f, l,
                                                              text[:100].replace("|", "\\|")))
    p = os.path.join(ROOT, "profiles", f"{round_}_k_expand_full.md")
    open(p, "w").write("\n".join(md) + "\n")
    print("wrote", p)


if __name__ == "__main__":
    rnd = sys.argv[1]
    launches(rnd, sys.argv[2])
    if len(sys.argv) > 3:
        full(rnd, sys.argv[3])
