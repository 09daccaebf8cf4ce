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
        launch.setdefault(k, {"name": r[kn].split("(")[0].replace("ambi::", "")})
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
           "note": "one full step (4096 partitions n=12 / 4 folds: greedy + beam + exact pass) of `AMBI_NO_GRAPH=1 ncu --metrics "
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


KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
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
          "Command: `AMBI_NO_GRAPH=1 ncu --set full ... -k regex:k_expand -s 134 -c 2 python bench.py --steps 2 --warmup 1 --no-cpu-baseline`;",
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
    # hottest source lines of the first launch
    src_csv = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(src_csv.splitlines()))
    shdr = srows[1]
    ie, te, ss = shdr.index("Instructions Executed"), shdr.index("Thread Instructions Executed"), shdr.index("# Samples")
    tbl = []
    for r in srows[2:]:
        if len(r) < 10 or not r[0].startswith("0x"):
            break
        tbl.append(r)
    so = os.path.join(ROOT, "ambigram_b200", "libambigram_b200.so")
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd="/tmp", capture_output=True)
    sass = subprocess.run(["nvdisasm", "-g", "/tmp/stage_a.sm_100a.cubin"], capture_output=True, text=True).stdout.splitlines()
    start = next(i for i, l in enumerate(sass) if l.startswith(".text._ZN4ambi8k_expandENS_9LevelArgsE:"))
    end = next((i for i in range(start + 1, len(sass)) if sass[i].startswith("//---------------------")), len(sass))
    ins_line, cur = [], None
    for ln in sass[start:end]:
        m = re.search(r'stage_a.cu", line (\d+)', ln)
        if m:
            cur = int(m.group(1)); continue
        if re.search(r'ambi_common.cuh", line (\d+)', ln):
            cur = None; continue
        if re.match(r"\s+/\*[0-9a-f]{4}\*/", ln):
            ins_line.append(cur)
    src = open(os.path.join(ROOT, "ambigram_b200", "csrc", "stage_a.cu")).read().splitlines()
    if len(ins_line) == len(tbl):
        agg = collections.defaultdict(lambda: [0, 0, 0])
        for i, r in enumerate(tbl):
            a = agg[ins_line[i]]
            a[0] += int(r[ie]); a[1] += int(r[te]); a[2] += int(r[ss])
        tot = sum(v[0] for v in agg.values()); tots = sum(v[2] for v in agg.values())
        md.append(f"## hottest source lines of launch 0 ({tot} warp instructions, {len(tbl)} SASS instructions = {len(tbl) * 16 // 1024} KB of code)\n")
        md.append("| % warp inst | % stall samples | threads / inst | stage_a.cu line | source |\n|---|---|---|---|---|")
        for l, v in sorted(agg.items(), key=lambda t: -t[1][0])[:25]:
            text = src[l - 1].strip()[:100].replace("|", "\\|") if l else "(helpers in ambi_common.cuh)"
            md.append("| %.1f | %.1f | %.1f | %s | `%s` |" % (100 * v[0] / tot, 100 * v[2] / max(tots, 1), v[1] / max(v[0], 1), l, text))
    else:
        md.append(f"(source mapping skipped: the .so in the tree ({len(ins_line)} SASS instructions) is not the profiled build ({len(tbl)}))")
    p = os.path.join(ROOT, "profiles", f"{round_}_k_expand_full.md")
    open(p, "w").write("\n".join(md) + "\n")
    print("wrote", p)


if __name__ == "__main__":
    rnd = sys.argv[1]
    launches(rnd, sys.argv[2])
    if len(sys.argv) > 3:
        full(rnd, sys.argv[3])
