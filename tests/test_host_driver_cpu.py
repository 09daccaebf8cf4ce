"""CPU: the C++ `--op bfb` host driver (ambigram_b200/host/ambigram_bfb.cpp) end to end from .lh text to the printed
path, with the device entry points of the C ABI answered by the CPU oracle through a preloaded test shim
(tests/shim/ambi_cpu_shim.c).  Under test: the driver's readers, depth -> copy number, flattening (getJuncCN,
getIndelBias, readComponents), imperfectFBI / expansion / indelBFB / printing -- against what the reference binary
printed for the same Stage-A assignment (tests/golden/*.json, tools/make_golden.py).  The same comparison runs on a
GPU with the real library in tests/test_gpu_cli.py."""
import json
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "ambigram_b200", "ambigram_bfb")
SHIM_SRC = os.path.join(ROOT, "tests", "shim", "ambi_cpu_shim.c")
SHIM = os.path.join(ROOT, "tests", "shim", "_build", "libambi_cpu_shim.so")
PATH_RE = re.compile(r"^\d+[+-](\|{0,2}\d+[+-])*$")


def load(name):
    return json.load(open(os.path.join(ROOT, "tests", "golden", name)))["cases"]


CASES = load("cases.json") + load("cases_extra.json") + load("cases_fractional.json")
MULTI = load("cases_multi.json")


@pytest.fixture(scope="module")
def shim():
    if not os.path.exists(CLI):
        pytest.skip("host driver not built (make -C ambigram_b200/csrc)")
    os.makedirs(os.path.dirname(SHIM), exist_ok=True)
    src = [SHIM_SRC, os.path.join(ROOT, "oracle", "ambi_oracle.c")]
    if not os.path.exists(SHIM) or any(os.path.getmtime(SHIM) < os.path.getmtime(f) for f in src):
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", SHIM] + src + ["-lm"])
    return SHIM


def run_cli(shim, args, cwd):
    env = dict(os.environ, AMBIGRAM_NO_TIME_CSV="1", LD_PRELOAD=shim)
    r = subprocess.run([CLI] + args, cwd=cwd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    return [ln.strip() for ln in r.stdout.splitlines() if PATH_RE.match(ln.strip())]


@pytest.mark.parametrize("reversed_", [False, True])
def test_host_driver_matches_reference_stdout(shim, tmp_path, reversed_):
    key = "ref_inject_paths_rev" if reversed_ else "ref_inject_paths"
    checked = 0
    for c in CASES:
        if len(c["oracle_x_nz"]) > 16 or not c[key]:
            continue                      # k > 16: reference comparator bug
        d = tmp_path / (c["name"] + ("_r" if reversed_ else ""))
        d.mkdir()
        (d / "in.lh").write_text(c["lh"])
        args = ["--op", "bfb", "--in_lh", "in.lh", "--lp_prefix", "p"]
        if c["juncs"]:
            (d / "in.juncs").write_text(c["juncs"])
            args += ["--juncdb", "in.juncs"]
        if c["junc_info"]:
            args += ["--junc_info", "true"]
        if reversed_:
            args += ["--reversed"]
        assert run_cli(shim, args, str(d)) == c[key], c["name"]
        checked += 1
    assert checked >= 80


def test_host_driver_multi_partition_files(shim, tmp_path):
    """several SOURCE/SINK pairs, a one-segment partition, a fold-back two segments apart: one path per partition,
    equal to the reference's wherever the optimum is unique"""
    for c in MULTI:
        d = tmp_path / c["name"].replace(" ", "").replace(",", "_").replace("(", "").replace(")", "")
        d.mkdir()
        (d / "in.lh").write_text(c["lh"])
        got = run_cli(shim, ["--op", "bfb", "--in_lh", "in.lh"], str(d))
        assert len(got) == len(c["ref_highs_paths"]), c["name"]
        for g, ref, uniq in zip(got, c["ref_highs_paths"], c["oracle_unique"]):
            if uniq:
                assert g == ref, c["name"]


def test_host_driver_batch_list(shim, tmp_path):
    names, want = [], []
    for c in CASES:
        if c["juncs"] or c.get("unit", 1) != 1 or len(c["oracle_x_nz"]) > 16 or not c["ref_inject_paths"]:
            continue
        (tmp_path / (c["name"] + ".lh")).write_text(c["lh"])
        names.append(c["name"] + ".lh")
        want += c["ref_inject_paths"]
    (tmp_path / "list.txt").write_text("\n".join(names) + "\n")
    assert run_cli(shim, ["--op", "bfb", "--in_list", "list.txt"], str(tmp_path)) == want


def test_host_driver_print_all(shim, tmp_path):
    """--all: every valid order's path in the reference's print order (the driver asks the library for the finished
    vertex paths: ambi_bfb_all_valid_paths)"""
    checked = 0
    for c in CASES:
        if "ref_inject_paths_all" not in c or c["name"].startswith("indel_") or len(c["oracle_x_nz"]) > 16:
            continue
        for rev, key in ((False, "ref_inject_paths_all"), (True, "ref_inject_paths_all_rev")):
            d = tmp_path / (c["name"] + ("_ar" if rev else "_a"))
            d.mkdir()
            (d / "in.lh").write_text(c["lh"])
            args = ["--op", "bfb", "--in_lh", "in.lh", "--all"] + (["--reversed"] if rev else [])
            if c["juncs"]:
                (d / "in.juncs").write_text(c["juncs"])
                args += ["--juncdb", "in.juncs"]
            if c["junc_info"]:
                args += ["--junc_info", "true"]
            assert run_cli(shim, args, str(d)) == c[key], c["name"]
            checked += 1
    assert checked >= 40
