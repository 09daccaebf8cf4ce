"""CPU: the REFERENCE'S OWN main() (localhap.cpp) with the two seams of INTEGRATION.md patched in
(tools/patches/localhap_seams.patch, built by `make -C oracle dropin` from the reference's unmodified src/ and linked
against libambigram_b200.so) -- stdout path lines equal to the unmodified reference binary's on the committed
fixtures, forward, --reversed, --all and with .juncs.  Without a GPU the device entry points of the library are
answered by the oracle through the preloaded test shim (tests/shim/ambi_cpu_shim.c); tests/test_gpu_cli.py runs the
same binary against the real library on a B200.  The binary is built in this container only (it needs
/root/reference); on a box without it the test skips."""
import json
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "oracle", "_ref", "Ambigram_dropin")
SHIM_SRC = os.path.join(ROOT, "tests", "shim", "ambi_cpu_shim.c")
SHIM = os.path.join(ROOT, "tests", "shim", "_build", "libambi_cpu_shim.so")
PATH_RE = re.compile(r"^\d+[+-](\|{0,2}\d+[+-])*$")


def load(name):
    return json.load(open(os.path.join(ROOT, "tests", "golden", name)))["cases"]


CASES = load("cases.json") + load("cases_extra.json") + load("cases_fractional.json")
MULTI = load("cases_multi.json")


@pytest.fixture(scope="module")
def shim():
    if not os.path.exists(DROPIN):
        if os.path.isdir("/root/reference/src"):
            subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "dropin"])
        else:
            pytest.skip("oracle/_ref/Ambigram_dropin not built (needs /root/reference)")
    os.makedirs(os.path.dirname(SHIM), exist_ok=True)
    src = [SHIM_SRC, os.path.join(ROOT, "oracle", "ambi_oracle.c")]
    if not os.path.exists(SHIM) or any(os.path.getmtime(SHIM) < os.path.getmtime(f) for f in src):
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", SHIM] + src + ["-lm"])
    return SHIM


def run_dropin(shim, case, d, extra=()):
    (d / "in.lh").write_text(case["lh"])
    args = [DROPIN, "--op", "bfb", "--in_lh", "in.lh", "--lp_prefix", "p"]
    if case.get("juncs"):
        (d / "in.juncs").write_text(case["juncs"])
        args += ["--juncdb", "in.juncs"]
    if case.get("junc_info"):
        args += ["--junc_info", "true"]
    env = dict(os.environ, LD_PRELOAD=shim)
    r = subprocess.run(args + list(extra), cwd=str(d), env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (case["name"], r.stderr[-500:])
    return [ln.strip() for ln in r.stdout.splitlines() if PATH_RE.match(ln.strip())]


def usable(c):
    # k > 16: reference comparator bug.  Copy numbers in eighths (cases_fractional.json) are in.
    return len(c["oracle_x_nz"]) <= 16


@pytest.mark.parametrize("reversed_", [False, True])
def test_patched_reference_main_prints_the_reference_paths(shim, tmp_path, reversed_):
    key = "ref_inject_paths_rev" if reversed_ else "ref_inject_paths"
    checked = 0
    for c in CASES:
        if not usable(c) or not c[key]:
            continue
        d = tmp_path / (c["name"] + ("_r" if reversed_ else ""))
        d.mkdir()
        got = run_dropin(shim, c, d, ["--reversed"] if reversed_ else [])
        assert got == c[key], c["name"]
        checked += 1
    assert checked >= 80


def test_patched_reference_main_print_all(shim, tmp_path):
    checked = multi = 0
    for c in CASES:
        if not usable(c) or "ref_inject_paths_all" not in c or c["name"].startswith("indel_"):
            continue
        for rev, key in ((False, "ref_inject_paths_all"), (True, "ref_inject_paths_all_rev")):
            d = tmp_path / (c["name"] + ("_ar" if rev else "_a"))
            d.mkdir()
            got = run_dropin(shim, c, d, ["--all"] + (["--reversed"] if rev else []))
            assert got == c[key], c["name"]
            checked += 1
            multi += len(got) > 1
    assert checked >= 40 and multi >= 10


def test_patched_reference_main_multi_partition_files(shim, tmp_path):
    for c in MULTI:
        d = tmp_path / re.sub(r"[^A-Za-z0-9_]", "_", c["name"])
        d.mkdir()
        got = run_dropin(shim, c, d)
        assert len(got) == len(c["ref_highs_paths"]), c["name"]
        for g, ref, uniq in zip(got, c["ref_highs_paths"], c["oracle_unique"]):
            if uniq:
                assert g == ref, c["name"]


TRX = load("cases_trx.json")


def test_patched_reference_main_translocation_modes(shim, tmp_path):
    """PROP I1 / I2 / C1 / C2: host-only graph rewrites of the reference around the same search
    (insertBeforeBFB, concatBeforeBFB, virusBFB, translocationBFB stay the reference's code).  Synthetic inputs after
    the README's four examples; every path line -- per chromosome and the stitched `||` path -- equal to the unmodified
    reference's.  (HiGHS played cbc there: a tie in Stage A could legitimately print another path; none of the twelve
    inputs has one.)"""
    modes = set()
    for c in TRX:
        d = tmp_path / c["name"]
        d.mkdir()
        assert run_dropin(shim, c, d) == c["ref_highs_paths"], c["name"]
        modes.add(c["name"].split("_")[1])
    assert modes == {"I1", "I2", "C1", "C2"} and len(TRX) >= 12
