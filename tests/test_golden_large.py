"""Headline sizes (BASELINE configs 3/4 and the top of config 5): n = 20 / 25 / 30 fixtures produced by the reference's
own code here (tools/make_golden.py large -> tests/golden/cases_large.json):
  * ref_highs_objective -- optimum of the LP the reference's BFB_ILP emitted, solved exactly by HiGHS (the cbc stand-in);
  * probes -- the reference's rows evaluated on probe assignments;
  * ref_inject_paths[_rev] -- what the unmodified reference prints for our Stage-A assignment (comparable for k <= 16);
  * ref_fixed_paths[_rev]  -- the same from the reference with compareLoops made a strict weak order
    (tools/patches/reference_fixed_sort.patch, oracle/_ref/Ambigram_ref_fixedsort): the parity target for k > 16,
    where the unmodified comparator scrambles the DAG (it hangs or crashes on 11 of these 20 inputs).
CPU part: the oracle against all of it.  GPU part (-m gpu): the CUDA path against the fixtures."""
import json
import os

import numpy as np
import pytest

from ambigram_b200 import capi, lh
from tests import oracle_lib as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = json.load(open(os.path.join(ROOT, "tests", "golden", "cases_large.json")))["cases"]
IDS = [c["name"] for c in CASES]


def partition_of(case):
    return lh.flatten(lh.parse_lh(case["lh"]), [], False)[0]


def dense(nz, nv):
    x = np.zeros(nv, dtype=np.int32)
    for i, v in nz:
        x[i] = v
    return x


def test_fixture_covers_the_headline_sizes():
    ns = [c["n"] for c in CASES]
    assert ns.count(25) >= 4 and ns.count(30) >= 4 and ns.count(20) >= 2
    assert sum(1 for c in CASES if len(c["oracle_x_nz"]) > 16) >= 5          # k > 16 is the common case here
    assert all(c["ref_fixed_rc"] == 0 and c["ref_fixed_paths"] for c in CASES)


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_oracle_value_equals_reference_milp_optimum(case):
    p = partition_of(case).problem
    assert list(p.seg_cn2) == case["seg_cn2"] and list(p.fbi_cn2) == case["fbi_cn2"] and p.bias == case["bias"]
    x = dense(case["oracle_x_nz"], p.n * (p.n + 1))
    ok, s2 = orc.model_eval(p, x)
    assert ok and s2 == case["oracle_score2"]
    assert abs(2 * (case["ref_highs_objective"] + case["bias"]) - case["oracle_score2"]) < 1e-6
    if case["oracle_unique"]:
        assert case["ref_highs_x_nz"] == case["oracle_x_nz"]
    for pr in case["probes"]:
        ok, s2 = orc.model_eval(p, dense(pr["x_nz"], p.n * (p.n + 1)))
        assert ok == pr["feasible"] and s2 == pr["score2"]
    if case["oracle_seconds"] < 4.0:          # the quick ones are searched again here
        x2, st = orc.solve(p, want_unique=False)
        assert st.score2 == case["oracle_score2"] and np.array_equal(x2, x)


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_stage_b_matches_the_reference_for_every_k(case):
    part = partition_of(case)
    p = part.problem
    d = orc.construct_dag(p.n, p.start_id, dense(case["oracle_x_nz"], p.n * (p.n + 1)))
    for rev, kf, ku in ((False, "ref_fixed_paths", "ref_inject_paths"), (True, "ref_fixed_paths_rev", "ref_inject_paths_rev")):
        ok, flipped, order, bkp, _ = orc.first_valid_order(d, rev)
        assert ok
        got = capi.format_path(orc.expand_path(orc.imperfect_fbi(bkp, part.inversions)))
        assert got == case[kf][0]
        if d.k <= 16:
            assert case[ku] and got == case[ku][0]            # the unmodified reference agrees where its sort is sane


@pytest.mark.gpu
def test_gpu_stage_a_and_b_on_the_large_fixtures(ctx):
    parts = [partition_of(c) for c in CASES]
    sols = ctx.solve_batch([pt.problem for pt in parts])
    dags = []
    for c, pt, s in zip(CASES, parts, sols):
        assert s.status == 0 and s.score2 == c["oracle_score2"], c["name"]
        assert abs(2 * (c["ref_highs_objective"] + c["bias"]) - s.score2) < 1e-6
        assert [[int(i), int(v)] for i, v in enumerate(s.element_cn) if v] == c["oracle_x_nz"], c["name"]
        dags.append(capi.construct_dag(pt.problem.n, pt.problem.start_id, s.element_cn))
    for rev, key in ((False, "ref_fixed_paths"), (True, "ref_fixed_paths_rev")):
        res = ctx.first_valid_order_batch(dags, [rev] * len(dags))
        for c, pt, r in zip(CASES, parts, res):
            assert r.found, c["name"]
            assert capi.format_path(capi.expand_path(capi.imperfect_fbi(r.bkp, pt.inversions))) == c[key][0], c["name"]
