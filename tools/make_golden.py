#!/usr/bin/env python3
"""Generate tests/golden/cases.json by running the REFERENCE's own code (oracle/_ref/Ambigram_ref,
built from /root/reference by `make -C oracle ref`) in this container.

For every case we record
  * the reference's printed BFB path with HiGHS standing in for cbc (mode highs) -- the whole
    reference pipeline end to end,
  * the optimal objective of the reference's own LP (the matrix its BFB_ILP emitted),
  * (feasible, score2) of the reference's rows on a set of probe assignments -- pins
    oracle/ambi_oracle.c orc_model_eval row for row,
  * the path the reference prints when OUR oracle's Stage-A assignment is injected as the cbc
    answer (mode inject) -- pins Stage B (constructDAG, allTopologicalOrders, getBFB, imperfectFBI,
    expansion, printBFB) independently of solver tie-breaks; also with --reversed.
The GPU box has no /root/reference: tests read only the JSON written here.
"""
import json
import os
import random
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from ambigram_b200 import lh, synth  # noqa: E402
from tests import oracle_lib as orc  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "Ambigram_ref")
BIN = os.path.join(ROOT, "oracle", "_ref", "bin")
PATH_RE = re.compile(r"^\d+[+-](\|{0,2}\d+[+-])*$")


def run_ref(lh_text, juncs_text="", junc_info=False, reversed_=False, inject=None, print_all=False):
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "in.lh"), "w").write(lh_text)
        cmd = [REF, "--op", "bfb", "--in_lh", "in.lh", "--lp_prefix", "p"]
        if juncs_text:
            open(os.path.join(d, "in.juncs"), "w").write(juncs_text)
            cmd += ["--juncdb", "in.juncs"]
        if junc_info:
            cmd += ["--junc_info", "true"]
        if reversed_:
            cmd += ["--reversed"]
        if print_all:
            cmd += ["--all"]
        env = dict(os.environ, PATH=BIN + ":" + os.environ["PATH"])
        if inject is not None:
            with open(os.path.join(d, "inject.txt"), "w") as f:
                for i, v in enumerate(inject):
                    if v:
                        f.write(f"{i} {int(v)}\n")
            env["AMBI_CBC_MODE"] = "inject"
            env["AMBI_CBC_SOL"] = os.path.join(d, "inject.txt")
        r = subprocess.run(cmd, cwd=d, env=env, capture_output=True, text=True, timeout=3600)
        paths = [ln.strip() for ln in r.stdout.splitlines() if PATH_RE.match(ln.strip())]
        sol = open(os.path.join(d, "p.sol")).read() if os.path.exists(os.path.join(d, "p.sol")) else ""
        lp = open(os.path.join(d, "p.lp")).read() if os.path.exists(os.path.join(d, "p.lp")) else ""
        return r.returncode, paths, sol, lp, r.stdout


def parse_sol(sol, nvars):
    obj, x = None, [0] * nvars
    toks = sol.split()
    for i, t in enumerate(toks):
        if t == "value":
            obj = float(toks[i + 1])
        if t.startswith("x") and t[1:].isdigit() and int(t[1:]) < nvars:
            x[int(t[1:])] = int(round(float(toks[i + 1])))
    return obj, x


def lp_eval(lp, x, nvars):
    """(feasible, score2) of the reference's dumped rows on integer x (eps optimal)."""
    lines = lp.splitlines()
    nrows, ncols = map(int, lines[1].split())
    rows = []
    for ln in lines[2:2 + nrows]:
        t = ln.split()
        rows.append((float(t[1]), float(t[2]), [(int(p.split(":")[0]), float(p.split(":")[1])) for p in t[4:]]))
    cols = [ln.split() for ln in lines[2 + nrows:2 + nrows + ncols]]
    ok = True
    for i in range(nvars):
        if x[i] < float(cols[i][1]) or x[i] > float(cols[i][2]):
            ok = False
    neps = ncols - 1 - nvars
    score2 = 0.0
    for k in range(neps):
        lo, hi, r = rows[2 * k]
        val = sum(c * x[i] for i, c in r if i < nvars)
        score2 += 2 * abs(val - lo)
    for (lo, hi, r) in rows[2 * neps + 1:]:
        val = sum(c * x[i] for i, c in r if i < nvars)
        if val < lo - 1e-9 or val > hi + 1e-9:
            ok = False
    return ok, score2


def probes(problem, xopt, rng, k=24):
    n = problem.n
    nv = n * (n + 1)
    out = [list(map(int, xopt)), [0] * nv]
    for _ in range(k):
        x = list(map(int, xopt))
        for _ in range(rng.randint(1, 3)):
            i = rng.randrange(nv)
            x[i] = max(0, x[i] + rng.choice([-1, 1]))
        out.append(x)
    for _ in range(k // 2):
        x = [0] * nv
        for _ in range(rng.randint(1, 4)):
            x[rng.randrange(nv)] = rng.randint(1, 2)
        out.append(x)
    return out


def make_case(name, lh_text, juncs_text="", junc_info=False, want_highs=True, rng=None, unit=1):
    rng = rng or random.Random(1)
    g = lh.parse_lh(lh_text)
    comps = lh.read_components(g, juncs_text) if juncs_text else []
    parts = lh.flatten(g, comps, junc_info, unit)
    assert len(parts) == 1 and parts[0].problem is not None, name
    p = parts[0].problem
    nv = p.n * (p.n + 1)
    case = {"name": name, "inversions": {str(k): list(v) for k, v in parts[0].inversions.items()}, "lh": lh_text, "juncs": juncs_text, "junc_info": junc_info,
            "n": p.n, "seg_cn2": list(p.seg_cn2), "fbi_cn2": list(p.fbi_cn2), "loop_ub": p.effective_loop_ub(),
            "bias": p.bias, "tgs_pairs": list(p.tgs_pairs), "start_id": p.start_id, "unit": unit}
    x, st = orc.solve(p, want_unique=True)
    case["oracle_score2"] = int(st.score2)
    case["oracle_unique"] = int(st.unique)
    case["oracle_x_nz"] = [[int(i), int(v)] for i, v in enumerate(x) if v]
    if want_highs:
        rc, paths, sol, lpt, out = run_ref(lh_text, juncs_text, junc_info)
        obj, xh = parse_sol(sol, nv)
        case["ref_rc"] = rc
        case["ref_highs_objective"] = obj
        case["ref_highs_x_nz"] = [[i, v] for i, v in enumerate(xh) if v]
        case["ref_highs_paths"] = paths
    # inject our Stage-A answer into the reference -> its Stage B output
    for rev in (False, True):
        rc, paths, sol, lpt, out = run_ref(lh_text, juncs_text, junc_info, reversed_=rev, inject=x)
        case["ref_inject_paths_rev" if rev else "ref_inject_paths"] = paths
        case["ref_inject_rc_rev" if rev else "ref_inject_rc"] = rc
    # LP stats + probe evaluations with the reference's own rows
    lines = lpt.splitlines()
    case["lp_rows"], case["lp_cols"] = map(int, lines[1].split())
    pr = probes(p, x, rng)
    # score2 = 2 * unit * sum(eps) of the reference's rows (an integer when the copy numbers are multiples of 1/(2*unit))
    case["probes"] = [{"x_nz": [[i, v] for i, v in enumerate(q) if v], "feasible": bool(f), "score2": int(round(unit * s))}
                      for q in pr for (f, s) in [lp_eval(lpt, q, nv)]]
    return case


def extra_imperfect():
    """cases whose fold-back junctions join two adjacent segments (imperfectFBI, bias > 1); appended to
    an existing cases.json"""
    out = os.path.join(ROOT, "tests", "golden", "cases.json")
    doc = json.load(open(out))
    doc["cases"] = [c for c in doc["cases"] if not c["name"].startswith("imperfect_")]
    rng = random.Random(11)
    for n, folds, k in [(5, 3, 4), (6, 3, 4), (8, 3, 4), (10, 4, 3), (12, 4, 3)]:
        for seed in range(k):
            s = synth.generate(n, folds, 50 + seed, 0.0, imperfect=0.6)
            c = make_case(f"imperfect_n{n}_f{folds}_s{seed}", s.lh_text(), rng=rng)
            doc["cases"].append(c)
            print(c["name"], c["oracle_score2"], c.get("ref_highs_objective"), c["bias"], c["ref_inject_paths"][:1], flush=True)
    json.dump(doc, open(out, "w"), indent=0)
    print("wrote", out, len(doc["cases"]), "cases")


def extra_indel():
    """cases with a deletion / duplication junction inside the partition (getIndelBias before the
    search, indelBFB after it): the reference prints a second path line"""
    out = os.path.join(ROOT, "tests", "golden", "cases.json")
    doc = json.load(open(out))
    doc["cases"] = [c for c in doc["cases"] if not c["name"].startswith("indel_")]
    rng = random.Random(13)
    for n, folds, seed, extra in [(6, 3, 0, "JUNC H:2:+ H:4:+ 15.0 1.0 U B"), (8, 3, 1, "JUNC H:3:+ H:5:+ 15.0 1.0 U B"),
                                  (8, 3, 2, "JUNC H:6:- H:4:- 15.0 1.0 U B"), (10, 4, 3, "JUNC H:5:+ H:4:+ 15.0 1.0 U B"),
                                  (12, 4, 5, "JUNC H:7:+ H:9:+ 15.0 1.0 U B"), (12, 4, 6, "JUNC H:4:+ H:3:+ 15.0 1.0 U B")]:
        s = synth.generate(n, folds, seed, 0.0)
        text = s.lh_text() + extra + "\n"
        c = make_case(f"indel_n{n}_s{seed}", text, rng=rng)
        doc["cases"].append(c)
        print(c["name"], c["oracle_score2"], c.get("ref_highs_objective"), c["seg_cn2"], c["ref_inject_paths"], flush=True)
    json.dump(doc, open(out, "w"), indent=0)
    print("wrote", out, len(doc["cases"]), "cases")


def extra_all():
    """--all: every path the reference prints (our Stage-A answer injected), forward and --reversed"""
    out = os.path.join(ROOT, "tests", "golden", "cases.json")
    doc = json.load(open(out))
    done = 0
    for c in doc["cases"]:
        if len(c["oracle_x_nz"]) > 12:
            continue
        nv = c["n"] * (c["n"] + 1)
        x = [0] * nv
        for i, v in c["oracle_x_nz"]:
            x[i] = v
        for rev in (False, True):
            rc, paths, sol, lpt, so = run_ref(c["lh"], c["juncs"], c["junc_info"], reversed_=rev, inject=x, print_all=True)
            c["ref_inject_paths_all_rev" if rev else "ref_inject_paths_all"] = paths
        done += 1
        print(c["name"], len(c["ref_inject_paths_all"]), len(c["ref_inject_paths_all_rev"]), flush=True)
    json.dump(doc, open(out, "w"), indent=0)
    print("wrote", out, done, "cases with --all output")


def zero_cn(text, hdp=15.0, normal=0.0):
    """same profile with every explicit copy number removed (cn = 0): the reader has to derive it from the depth
    (Graph::calculateHapDepth / calculateCopyNum, Graph.cpp:312-405); depths are cn x `hdp` + `normal` (the share
    of the depth that normal cells contribute when PURITY < 1)"""
    out = []
    for ln in text.splitlines():
        t = ln.split()
        if t and t[0] == "SEG":
            t[2], t[3] = str(hdp * float(t[3]) + normal), "0"
        elif t and t[0] == "JUNC":
            t[3], t[4] = str(hdp * float(t[4]) + normal), "0"
        out.append(" ".join(t))
    return "\n".join(out) + "\n"


def extra_inputs():
    """input shapes beyond the synthetic generator's defaults -> tests/golden/cases_extra.json (CPU parity of the
    oracle and of the host-side flattening against the reference's own rows and printed paths)"""
    out = os.path.join(ROOT, "tests", "golden", "cases_extra.json")
    rng = random.Random(17)
    cases = []

    def add(name, text, **kw):
        c = make_case(name, text, rng=rng, **kw)
        cases.append(c)
        print(c["name"], c["n"], c["seg_cn2"], c["fbi_cn2"], c["oracle_score2"], c.get("ref_highs_objective"), c["ref_inject_paths"][:1], flush=True)

    # copy numbers derived from depth only
    for n, folds, seed in [(6, 3, 0), (8, 3, 1), (12, 4, 2)]:
        add(f"depth_only_n{n}_s{seed}", zero_cn(synth.generate(n, folds, seed, 0.0).lh_text()))
    # ... with an impure sample: haploid depth = 30 * 0.5 / (0.5 * 2 + 0.5 * 2) = 7.5, normal-cell share 30 * 0.5 = 15
    for n, folds, seed in [(6, 3, 3), (10, 4, 4)]:
        t = synth.generate(n, folds, seed, 0.5).lh_text().replace("PURITY 1", "PURITY 0.5")
        add(f"depth_purity_n{n}_s{seed}", zero_cn(t, 7.5, 15.0))
    # copy numbers eight times larger (scores in the hundreds)
    for n, folds, seed in [(6, 3, 5), (10, 4, 6)]:
        lines = []
        for ln in synth.generate(n, folds, seed, 0.5).lh_text().splitlines():
            t = ln.split()
            if t and t[0] == "SEG":
                t[2], t[3] = str(8 * float(t[2])), str(8 * float(t[3]))
            elif t and t[0] == "JUNC":
                t[3], t[4] = str(8 * float(t[3])), str(8 * float(t[4]))
            lines.append(" ".join(t))
        add(f"big_cn_n{n}_s{seed}", "\n".join(lines) + "\n")
    # a fold-back junction with copy number in (0.5, 1): getJuncCN rounds it up to 1 (LGM:4001-4002)
    for n, folds, seed in [(6, 3, 7), (8, 3, 8)]:
        lines, done = [], False
        for ln in synth.generate(n, folds, seed, 0.0).lh_text().splitlines():
            t = ln.split()
            if t and t[0] == "JUNC" and not done and float(t[4]) == 1.0:
                t[3], t[4], done = "10.5", "0.7", True
            lines.append(" ".join(t))
        add(f"junc_cn_0p7_n{n}_s{seed}", "\n".join(lines) + "\n")
    json.dump({"generator": "tools/make_golden.py inputs", "cases": cases}, open(out, "w"), indent=0)
    print("wrote", out, len(cases), "cases")


def two_chrom(n1, f1, s1, n2, f2, s2):
    """two partitions (SOURCE a,b / SINK c,d) in one .lh: the reference searches them one after the other"""
    a, b = synth.generate(n1, f1, s1), synth.generate(n2, f2, s2)
    lines = ["SAMPLE two", "AVG_CHR_SEG_DP 30", "AVG_WHOLE_HOST_DP 30", "AVG_JUNC_DP 30", "PURITY 1", "AVG_TUMOR_PLOIDY 2",
             "PLOIDY 2m1", f"VIRUS_START {n1 + n2 + 1}", f"SOURCE 1,{n1 + 1}", f"SINK {n1},{n1 + n2}"]
    for off, smp, chrom in ((0, a, "chr7"), (n1, b, "chr8")):
        for i, cn in enumerate(smp.seg_cn, 1):
            lines.append(f"SEG H:{off + i}:{chrom}:{1000 * i}:{1000 * i + 999} {15.0 * cn} {float(cn)}")
    for off, smp in ((0, a), (n1, b)):
        for (si, sd, ti, td, cn) in smp.juncs:
            lines.append(f"JUNC H:{off + si}:{'+' if sd > 0 else '-'} H:{off + ti}:{'+' if td > 0 else '-'} {15.0 * cn} {float(cn)} U B")
    return "\n".join(lines) + "\n"


def extra_multi():
    """shapes make_case() does not take (several partitions per file, a one-segment partition, a fold-back
    junction two segments apart): the reference's printed paths with HiGHS as the solver ->
    tests/golden/cases_multi.json.  Paths are only comparable where the optimum is unique; the oracle says where."""
    out = os.path.join(ROOT, "tests", "golden", "cases_multi.json")
    texts = [(f"two_chrom_{a}", two_chrom(*a)) for a in [(6, 3, 0, 6, 3, 1), (8, 3, 1, 5, 3, 2), (5, 3, 3, 10, 4, 4), (4, 2, 5, 7, 3, 6)]]
    texts.append(("one_segment", "SAMPLE one\nAVG_CHR_SEG_DP 30\nAVG_WHOLE_HOST_DP 30\nAVG_JUNC_DP 30\nPURITY 1\nAVG_TUMOR_PLOIDY 2\n"
                  "PLOIDY 2m1\nVIRUS_START 2\nSOURCE 1\nSINK 1\nSEG H:1:chr7:1000:1999 60.0 4.0\nJUNC H:1:+ H:1:- 15.0 1.0 U B\n"))
    texts.append(("foldback_two_apart", synth.generate(8, 3, 1).lh_text().replace("JUNC H:2:- H:2:+", "JUNC H:3:- H:1:+")))
    cases = []
    for name, text in texts:
        rc, paths, sol, lpt, so = run_ref(text)
        parts = lh.flatten(lh.parse_lh(text), [], False)
        uniq = []
        for part in parts:
            x, st = orc.solve(part.problem, want_unique=True)
            uniq.append(int(st.unique))
        cases.append({"name": name, "lh": text, "ref_rc": rc, "ref_highs_paths": paths, "oracle_unique": uniq})
        print(name, rc, paths, uniq, flush=True)
    json.dump({"generator": "tools/make_golden.py multi", "cases": cases}, open(out, "w"), indent=0)
    print("wrote", out, len(cases), "cases")


def extra_fractional():
    """copy numbers that are not half-integral (multiples of 1/8): the reference's LP takes any real right-hand side;
    the oracle's *_units entry points search in units of 1/8 copy (unit = 4) -> tests/golden/cases_fractional.json"""
    out = os.path.join(ROOT, "tests", "golden", "cases_fractional.json")
    rng = random.Random(23)
    cases = []
    for n, folds, seed in [(4, 2, 0), (5, 3, 1), (6, 3, 2), (6, 3, 3), (8, 3, 4), (8, 3, 5), (10, 4, 6), (12, 4, 7)]:
        lines = []
        for ln in synth.generate(n, folds, seed, 0.0).lh_text().splitlines():
            t = ln.split()
            if t and t[0] == "SEG":
                cn = max(0.125, float(t[3]) + rng.randint(-3, 3) / 8.0)
                t[2], t[3] = str(15.0 * cn), str(cn)
            elif t and t[0] == "JUNC":
                cn = float(t[4]) + rng.randint(0, 3) / 8.0          # stays >= 1: no rounding of (0.5, 1) involved
                t[3], t[4] = str(15.0 * cn), str(cn)
            lines.append(" ".join(t))
        c = make_case(f"fractional_n{n}_s{seed}", "\n".join(lines) + "\n", rng=rng, unit=4)
        cases.append(c)
        print(c["name"], c["seg_cn2"], c["fbi_cn2"], c["oracle_score2"], c.get("ref_highs_objective"), c["oracle_unique"], c["ref_inject_paths"][:1], flush=True)
    json.dump({"generator": "tools/make_golden.py fractional", "cases": cases}, open(out, "w"), indent=0)
    print("wrote", out, len(cases), "cases")


REF_FIXED = os.path.join(ROOT, "oracle", "_ref", "Ambigram_ref_fixedsort")


def run_binary(binary, lh_text, inject, reversed_=False, timeout=120):
    """one reference binary with an injected Stage-A assignment; -> (rc or None on timeout, paths, lp text).
    The model dump (p.lp) is written before `cbc` is called, so it exists even when Stage B hangs."""
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "in.lh"), "w").write(lh_text)
        cmd = [binary, "--op", "bfb", "--in_lh", "in.lh", "--lp_prefix", "p"] + (["--reversed"] if reversed_ else [])
        with open(os.path.join(d, "inject.txt"), "w") as f:
            for i, v in enumerate(inject):
                if v:
                    f.write(f"{i} {int(v)}\n")
        env = dict(os.environ, PATH=BIN + ":" + os.environ["PATH"], AMBI_CBC_MODE="inject", AMBI_CBC_SOL=os.path.join(d, "inject.txt"))
        try:
            r = subprocess.run(cmd, cwd=d, env=env, capture_output=True, text=True, timeout=timeout)
            rc, out = r.returncode, r.stdout
        except subprocess.TimeoutExpired as e:
            rc, out = None, (e.stdout.decode() if isinstance(e.stdout, bytes) else (e.stdout or ""))
        paths = [ln.strip() for ln in out.splitlines() if PATH_RE.match(ln.strip())]
        lp = open(os.path.join(d, "p.lp")).read() if os.path.exists(os.path.join(d, "p.lp")) else ""
        return rc, paths, lp


def highs_on_dump(lp_text, time_limit):
    """HiGHS (the cbc stand-in, oracle/cbc_standin.py) on a model the reference dumped -> (objective, x, proven)"""
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "p.lp"), "w").write(lp_text)
        env = dict(os.environ, AMBI_CBC_MODE="highs", AMBI_CBC_TIME=str(time_limit))
        t0 = __import__("time").time()
        subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "cbc_standin.py"), "p.lp", "solve", "solu", "p.sol"],
                       cwd=d, env=env, timeout=time_limit + 600)
        dt = __import__("time").time() - t0
        sol = open(os.path.join(d, "p.sol")).read()
        return sol, dt


def extra_large(which):
    """headline sizes (BASELINE configs 3/4 and the top of config 5): n = 25 / 7 folds and n = 30 / 8 folds, clean and noisy.
    Per case: the optimum of the reference's own dumped LP by HiGHS (value parity of Stage A at these sizes), the
    reference's rows on probe assignments, and the path printed for our Stage-A assignment by (a) the unmodified
    reference (only meaningful for k <= 16) and (b) the reference with compareLoops fixed
    (oracle/_ref/Ambigram_ref_fixedsort, tools/patches/reference_fixed_sort.patch), forward and --reversed
    -> tests/golden/cases_large.json (appended case by case, so an interrupted run keeps what it has)"""
    out = os.path.join(ROOT, "tests", "golden", "cases_large.json")
    doc = json.load(open(out)) if os.path.exists(out) else {"generator": "tools/make_golden.py large", "cases": []}
    have = {c["name"] for c in doc["cases"]}
    rng = random.Random(29)
    spec = [(25, 7, 0.0, 0), (25, 7, 0.0, 1), (25, 7, 0.5, 2), (25, 7, 0.5, 3),
            (30, 8, 0.0, 0), (30, 8, 0.5, 0), (30, 8, 0.5, 1), (30, 8, 0.5, 3),
            (20, 6, 0.5, 4), (20, 6, 0.5, 5), (30, 8, 0.0, 4), (25, 7, 0.5, 6),
            (25, 7, 0.5, 7), (25, 7, 0.0, 5), (30, 8, 0.5, 5), (30, 8, 0.5, 6), (30, 8, 0.0, 2), (28, 8, 0.5, 2),
            (22, 6, 0.5, 1), (20, 6, 0.0, 7)]
    if which:
        spec = [sp for sp in spec if f"n{sp[0]}" in which]
    for n, folds, noise, seed in spec:
        name = f"large_n{n}_f{folds}_noise{noise}_s{seed}"
        if name in have:
            continue
        text = synth.generate(n, folds, seed, noise).lh_text()
        part = lh.flatten(lh.parse_lh(text), [], False)[0]
        p = part.problem
        nv = p.n * (p.n + 1)
        import time
        t0 = time.time()
        x, st = orc.solve(p, want_unique=True)
        t_orc = time.time() - t0
        case = {"name": name, "inversions": {str(k): list(v) for k, v in part.inversions.items()}, "lh": text, "juncs": "", "junc_info": False,
                "n": p.n, "seg_cn2": list(p.seg_cn2), "fbi_cn2": list(p.fbi_cn2), "loop_ub": p.effective_loop_ub(), "bias": p.bias,
                "tgs_pairs": [], "start_id": p.start_id, "unit": 1, "oracle_score2": int(st.score2), "oracle_unique": int(st.unique),
                "oracle_x_nz": [[int(i), int(v)] for i, v in enumerate(x) if v], "oracle_seconds": t_orc}
        k = len(case["oracle_x_nz"])
        lpt = ""
        for rev in (False, True):
            rc, paths, lp1 = run_binary(REF, text, x, rev, timeout=60)
            lpt = lpt or lp1
            case["ref_inject_paths_rev" if rev else "ref_inject_paths"] = paths if rc == 0 else []
            case["ref_inject_rc_rev" if rev else "ref_inject_rc"] = rc
            rc, paths, _ = run_binary(REF_FIXED, text, x, rev, timeout=300)
            case["ref_fixed_paths_rev" if rev else "ref_fixed_paths"] = paths if rc == 0 else []
            case["ref_fixed_rc_rev" if rev else "ref_fixed_rc"] = rc
        lines = lpt.splitlines()
        case["lp_rows"], case["lp_cols"] = map(int, lines[1].split())
        sol, dt = highs_on_dump(lpt, 3000)
        obj, xh = parse_sol(sol, nv)
        case["ref_highs_objective"] = obj
        case["ref_highs_seconds"] = dt
        case["ref_highs_x_nz"] = [[i, v] for i, v in enumerate(xh) if v]
        pr = probes(p, x, rng, k=12)
        case["probes"] = [{"x_nz": [[i, v] for i, v in enumerate(q) if v], "feasible": bool(f), "score2": int(round(s))}
                          for q in pr for (f, s) in [lp_eval(lpt, q, nv)]]
        import fcntl
        with open(out + ".lock", "w") as lk:          # several generator processes may run side by side
            fcntl.flock(lk, fcntl.LOCK_EX)
            doc = json.load(open(out)) if os.path.exists(out) else {"generator": "tools/make_golden.py large", "cases": []}
            if name not in {c["name"] for c in doc["cases"]}:
                doc["cases"].append(case)
            doc["cases"].sort(key=lambda c: c["name"])
            json.dump(doc, open(out, "w"), indent=0)
        print(name, "k", k, "oracle", case["oracle_score2"], f"{t_orc:.1f}s", "highs", obj, f"{dt:.0f}s", "bias", p.bias,
              "ref rc", case["ref_inject_rc"], "fixed rc", case["ref_fixed_rc"], case["ref_fixed_paths"][:1], flush=True)
    print("wrote", out, len(doc["cases"]), "cases")


def extra_trx():
    """translocation modes (PROP I1 / I2 / C1 / C2, LGM:4052-4395, 3839-3939, localhap.cpp:71-88, 295-316): host-only graph
    rewrites around the same search.  The four inputs of the reference's README are not available offline, so these are
    synthetic multi-chromosome profiles built after them; recorded: every path line the unmodified reference prints
    (HiGHS standing in for cbc) -> tests/golden/cases_trx.json.  The tests run the reference's own main() with the two
    seams patched in (oracle/_ref/Ambigram_dropin) on the same inputs."""
    hdr = "SAMPLE trx\nAVG_CHR_SEG_DP 30\nAVG_WHOLE_HOST_DP 30\nAVG_JUNC_DP 30\nPURITY 1\nAVG_TUMOR_PLOIDY 2\nPLOIDY 2m1\n"

    def seg(i, ch, cn):
        return f"SEG H:{i}:{ch}:{1000 * i}:{1000 * i + 999} {15.0 * cn} {float(cn)}"

    def junc(a, da, b, db, cn=1):
        return f"JUNC H:{a}:{da} H:{b}:{db} {15.0 * cn} {float(cn)} U B"

    def lh_text(virus_start, sources, sinks, segs, juncs, prop):
        return (hdr + f"VIRUS_START {virus_start}\nSOURCE {','.join(map(str, sources))}\nSINK {','.join(map(str, sinks))}\n" +
                "\n".join([seg(*x) for x in segs] + [junc(*x) for x in juncs]) + f"\nPROP {prop}\n")

    inputs = []
    for k, (c2, c3, c6) in enumerate([(3, 3, 4), (3, 3, 2), (5, 3, 4)]):            # C2: concatenation after BFB
        inputs.append((f"trx_C2_{k}", lh_text(8, [1, 5], [4, 7],
                       [(1, "chr2", 1), (2, "chr2", c2), (3, "chr2", c3), (4, "chr2", 2), (5, "chr6", 1), (6, "chr6", c6), (7, "chr6", c6)],
                       [(4, "+", 4, "-"), (2, "-", 2, "+"), (7, "+", 7, "-", c6 // 2), (6, "-", 6, "+"), (3, "+", 6, "+")], "C2:chr2:chr6 M:chr2")))
    for k, (c3, c4) in enumerate([(4, 3), (4, 2), (3, 3)]):                         # I2: insertion after BFB
        inputs.append((f"trx_I2_{k}", lh_text(10, [1, 5, 8], [4, 7, 9],
                       [(1, "chr2", 1), (2, "chr2", 3), (3, "chr2", c3), (4, "chr2", c4), (5, "chr6", 1), (6, "chr6", 2), (7, "chr6", 2), (8, "chr13", 1), (9, "chr13", 1)],
                       [(4, "+", 4, "-"), (2, "-", 2, "+"), (4, "+", 4, "-"), (3, "-", 3, "+"), (7, "+", 7, "-"), (3, "+", 5, "+"), (6, "-", 8, "+"), (9, "+", 4, "-")],
                       "I2:chr2:chr6:chr13 M:chr2 S:3")))
    for k, (c2, c4, cv) in enumerate([(4, 2, 4), (4, 2, 2), (3, 2, 3)]):            # I1: insertion before BFB
        inputs.append((f"trx_I1_{k}", lh_text(5, [1, 5], [4, 6],
                       [(1, "chr8", 1), (2, "chr8", c2), (3, "chr8", c2), (4, "chr8", c4), (5, "virus", 0), (6, "virus", cv)],
                       [(4, "+", 4, "-"), (2, "-", 2, "+"), (3, "+", 6, "+", 2), (6, "+", 4, "+", 2)], "I1:chr8:virus:chr8 M:chr8")))
    for k, (c2, c4, cv) in enumerate([(3, 2, 2), (3, 2, 1), (5, 2, 2)]):            # C1: concatenation before BFB
        inputs.append((f"trx_C1_{k}", lh_text(7, [1, 5], [4, 6],
                       [(1, "chr1", 1), (2, "chr1", c2), (3, "chr1", c2), (4, "chr1", c4), (5, "chrV", cv), (6, "chrV", cv)],
                       [(4, "+", 4, "-"), (2, "-", 2, "+"), (6, "+", 2, "+", 2)], "C1:chr1:chrV")))
    trx_re = re.compile(r"^\d+[+-](\|{0,2}\d+[+-])*$")
    cases = []
    for name, text in inputs:
        with tempfile.TemporaryDirectory() as d:
            open(os.path.join(d, "in.lh"), "w").write(text)
            env = dict(os.environ, PATH=BIN + ":" + os.environ["PATH"])
            try:
                r = subprocess.run([REF, "--op", "bfb", "--in_lh", "in.lh", "--lp_prefix", "p"], cwd=d, env=env, capture_output=True, text=True, timeout=300)
            except subprocess.TimeoutExpired:
                print(name, "reference timed out"); continue
            lines = [ln.strip() for ln in r.stdout.splitlines() if trx_re.match(ln.strip())]
            marks = [ln.strip() for ln in r.stdout.splitlines() if ln.startswith(("BFB with translocation", "TRX-BFB mode"))]
            print(name, r.returncode, lines, marks, flush=True)
            if r.returncode == 0 and lines:
                cases.append({"name": name, "lh": text, "ref_rc": r.returncode, "ref_highs_paths": lines, "ref_marks": marks})
    out = os.path.join(ROOT, "tests", "golden", "cases_trx.json")
    json.dump({"generator": "tools/make_golden.py trx", "cases": cases}, open(out, "w"), indent=0)
    print("wrote", out, len(cases), "cases")


def extra_juncs20(seeds):
    """BASELINE config 3 (n = 20 / 6 folds with .juncs and --junc_info): reference-pinned like cases.json
    -> tests/golden/cases_juncs20.json (one process per seed may run side by side)"""
    import fcntl
    out = os.path.join(ROOT, "tests", "golden", "cases_juncs20.json")
    for seed in seeds:
        sm = synth.generate(20, 6, int(seed), 0.0, with_juncs=True)
        case = make_case(f"synth_juncs_n20_s{seed}", sm.lh_text(), sm.juncs_text(), True, rng=random.Random(41 + int(seed)))
        with open(out + ".lock", "w") as lk:
            fcntl.flock(lk, fcntl.LOCK_EX)
            doc = json.load(open(out)) if os.path.exists(out) else {"generator": "tools/make_golden.py juncs20", "cases": []}
            if case["name"] not in {c["name"] for c in doc["cases"]}:
                doc["cases"].append(case)
            doc["cases"].sort(key=lambda c: c["name"])
            json.dump(doc, open(out, "w"), indent=0)
        print(case["name"], "k", len(case["oracle_x_nz"]), "oracle", case["oracle_score2"], "highs", case.get("ref_highs_objective"),
              "tgs", len(case["tgs_pairs"]) // 2, "ref rc", case["ref_inject_rc"], flush=True)


def extra_sc_probe():
    """`--op sc_bfb` on two 4-segment clones: the LP the reference's BFB_ILP_SC (LGM:4754-5093) emits, kept as evidence
    for DESIGN.md section 9 (the joint model's deviation columns alias for every clone after the first)
    -> tests/golden/sc_probe.json"""
    a = synth.generate(4, 2, 0, 0.0).lh_text()
    b = synth.generate(4, 2, 1, 0.0).lh_text()
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "a.lh"), "w").write(a)
        open(os.path.join(d, "b.lh"), "w").write(b)
        env = dict(os.environ, PATH=BIN + ":" + os.environ["PATH"])
        r = subprocess.run([REF, "--op", "sc_bfb", "--in_lh", "a.lh,b.lh", "--lp_prefix", "p"], cwd=d, env=env, capture_output=True, text=True, timeout=600)
        lp = open(os.path.join(d, "p.lp")).read()
    paths = [ln.strip() for ln in r.stdout.splitlines() if PATH_RE.match(ln.strip())]
    out = os.path.join(ROOT, "tests", "golden", "sc_probe.json")
    json.dump({"generator": "tools/make_golden.py sc_probe", "command": "Ambigram --op sc_bfb --in_lh a.lh,b.lh --lp_prefix p",
               "a_lh": a, "b_lh": b, "rc": r.returncode, "paths": paths, "lp": lp}, open(out, "w"), indent=0)
    print("wrote", out, "rc", r.returncode, lp.splitlines()[1], paths)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "sc_probe":
        return extra_sc_probe()
    if len(sys.argv) > 1 and sys.argv[1] == "juncs20":
        return extra_juncs20(sys.argv[2:])
    if len(sys.argv) > 1 and sys.argv[1] == "trx":
        return extra_trx()
    if len(sys.argv) > 1 and sys.argv[1] == "large":
        return extra_large(sys.argv[2:])
    if len(sys.argv) > 1 and sys.argv[1] == "fractional":
        return extra_fractional()
    if len(sys.argv) > 1 and sys.argv[1] == "multi":
        return extra_multi()
    if len(sys.argv) > 1 and sys.argv[1] == "inputs":
        return extra_inputs()
    if len(sys.argv) > 1 and sys.argv[1] == "all":
        return extra_all()
    if len(sys.argv) > 1 and sys.argv[1] == "imperfect":
        return extra_imperfect()
    if len(sys.argv) > 1 and sys.argv[1] == "indel":
        return extra_indel()
    cases = []
    rng = random.Random(7)
    readme = open(os.path.join(ROOT, "tests", "golden", "readme6.lh")).read()
    cases.append(make_case("readme6", readme, rng=rng))
    readme_juncs = "6+ 6- 5- 4- 3- 2- 2+\n2- 2+ 3+ 4+ 5+ 6+ 6-\n6+ 6- 5- 4- 3-\n"
    cases.append(make_case("readme6_juncs", readme, readme_juncs, True, rng=rng))
    spec = [(2, 2, 0.0, 2), (3, 2, 0.5, 3), (4, 2, 0.5, 3), (5, 3, 0.5, 3), (6, 3, 0.0, 3), (6, 3, 0.5, 4),
            (8, 3, 0.5, 4), (10, 4, 0.5, 4), (12, 4, 0.0, 4), (12, 4, 0.5, 4), (14, 4, 0.5, 2), (16, 5, 0.5, 2)]
    for n, folds, noise, k in spec:
        for seed in range(k):
            s = synth.generate(n, folds, seed, noise)
            cases.append(make_case(f"synth_n{n}_f{folds}_noise{noise}_s{seed}", s.lh_text(), rng=rng))
            print(cases[-1]["name"], cases[-1]["oracle_score2"], cases[-1].get("ref_highs_objective"), flush=True)
    for n, folds, seed in [(8, 3, 11), (12, 4, 12), (12, 4, 13)]:
        s = synth.generate(n, folds, seed, 0.0, with_juncs=True)
        cases.append(make_case(f"synth_juncs_n{n}_s{seed}", s.lh_text(), s.juncs_text(), True, rng=rng))
        print(cases[-1]["name"], cases[-1]["oracle_score2"], cases[-1].get("ref_highs_objective"), flush=True)
    # n = 20 : HiGHS is slow here, keep two
    for seed in (2, 3):
        s = synth.generate(20, 6, seed, 0.0)
        cases.append(make_case(f"synth_n20_f6_s{seed}", s.lh_text(), want_highs=(seed == 2), rng=rng))
        print(cases[-1]["name"], cases[-1]["oracle_score2"], cases[-1].get("ref_highs_objective"), flush=True)
    out = os.path.join(ROOT, "tests", "golden", "cases.json")
    json.dump({"generator": "tools/make_golden.py", "cases": cases}, open(out, "w"), indent=0)
    print("wrote", out, len(cases), "cases")


if __name__ == "__main__":
    main()
