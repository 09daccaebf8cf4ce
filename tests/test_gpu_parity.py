"""GPU (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle and the committed
fixtures.  Bit-exact bar: same score2, same element_cn (search-order tie-break), same order and path."""
import json
import os
import random

import numpy as np
import pytest

from ambigram_b200 import capi, lh, synth
from ambigram_b200.capi import Problem
from tests import oracle_lib as orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = json.load(open(os.path.join(ROOT, "tests", "golden", "cases.json")))["cases"]
# + BASELINE config 3 (n = 20 / 6 folds with .juncs), reference-pinned the same way
CASES = CASES + json.load(open(os.path.join(ROOT, "tests", "golden", "cases_juncs20.json")))["cases"]
README_PATH = "1+2+3+4+5+6+|6-5-4-3-2-|2+3+4+|4-3-|3+4+|4-3-2-|2+3+4+5+6+|6-5-4-3-2-1-"


def partition_of(case):
    g = lh.parse_lh(case["lh"])
    comps = lh.read_components(g, case["juncs"]) if case["juncs"] else []
    return lh.flatten(g, comps, case["junc_info"])[0]


def problem_of(case):
    return partition_of(case).problem


def check_against_oracle(p, sol):
    x, st = orc.solve(p, want_unique=False)
    assert sol.status == 0
    assert sol.score2 == st.score2
    assert sol.objective2 == st.score2 - 2 * p.bias
    assert np.array_equal(sol.element_cn, x)
    ok, s2 = orc.model_eval(p, sol.element_cn)
    assert ok and s2 == sol.score2
    assert sol.incumbent_score2 >= sol.score2 and sol.candidates_scored > 0


def test_readme_known_answer(ctx):
    p = lh.problem_from_lh(open(os.path.join(ROOT, "tests", "golden", "readme6.lh")).read())
    sol = ctx.solve(p)
    assert sol.score2 == 2 and [int(i) for i in np.nonzero(sol.element_cn)[0]] == [26, 29, 31, 33]
    dag = capi.construct_dag(p.n, p.start_id, sol.element_cn)
    res = ctx.first_valid_order(dag, False)
    assert res.found and not res.used_flipped and res.order == [0, 1, 2, 3]
    assert capi.format_path(capi.expand_path(res.bkp)) == README_PATH
    rev = ctx.first_valid_order(dag, True)
    assert capi.format_path(capi.expand_path(rev.bkp)) == \
        "6-5-4-3-2-1-|1+2+3+4+5+6+|6-5-4-3-2-|2+3+4+|4-3-|3+4+|4-3-2-|2+3+4+5+6+"


def test_golden_cases_batch(any_ctx):
    ctx = any_ctx
    """every committed fixture (reference-generated) in one device batch"""
    probs = [problem_of(c) for c in CASES]
    sols = ctx.solve_batch(probs)
    dags = []
    for c, p, s in zip(CASES, probs, sols):
        assert s.status == 0 and s.score2 == c["oracle_score2"], c["name"]
        assert [[int(i), int(v)] for i, v in enumerate(s.element_cn) if v] == c["oracle_x_nz"], c["name"]
        if c.get("ref_highs_objective") is not None:
            assert abs(2 * (c["ref_highs_objective"] + c["bias"]) - s.score2) < 1e-5
        dags.append(capi.construct_dag(p.n, p.start_id, s.element_cn))
    for rev, key in ((False, "ref_inject_paths"), (True, "ref_inject_paths_rev")):
        res = ctx.first_valid_order_batch(dags, [rev] * len(dags))
        for c, d, r in zip(CASES, dags, res):
            if d.k > 16 or not c[key]:
                continue
            assert r.found, c["name"]
            bkp = capi.imperfect_fbi(r.bkp, partition_of(c).inversions)      # host step, LGM:3656
            assert capi.format_path(capi.expand_path(bkp)) == c[key][0], c["name"]


@pytest.mark.parametrize("n,folds,noise,count", [(2, 2, 0.5, 12), (3, 2, 0.5, 16), (5, 3, 0.5, 24), (8, 3, 0.5, 24),
                                                 (12, 4, 0.0, 32), (12, 4, 0.5, 32), (16, 5, 0.5, 8), (20, 6, 0.5, 4)])
def test_seeded_synthetic_vs_oracle(any_ctx, n, folds, noise, count):
    ctx = any_ctx
    probs = [lh.problem_from_lh(synth.generate(n, folds, 100 + s, noise).lh_text()) for s in range(count)]
    sols = ctx.solve_batch(probs)
    for p, s in zip(probs, sols):
        check_against_oracle(p, s)
    # single-problem entry point agrees with the batched one
    one = ctx.solve(probs[0])
    assert one.score2 == sols[0].score2 and np.array_equal(one.element_cn, sols[0].element_cn)


def test_random_profiles_ragged_batch(any_ctx):
    ctx = any_ctx
    """ragged batch: n from 1 to 10, arbitrary (non-BFB) profiles, zeros, third-generation rows"""
    rng = random.Random(5)
    probs = []
    for _ in range(200):
        n = rng.randint(1, 10)
        tg = []
        for _ in range(rng.choice([0, 0, 1, 3])):
            a = rng.randrange(n); b = rng.randrange(a, n); tg += [a + 1, b + 1]
        probs.append(Problem(n=n, seg_cn2=[rng.choice([0, 1, 2, 4, 6, 7, 12, 20]) for _ in range(n)],
                             fbi_cn2=[rng.choice([0, 0, 1, 2, 4, 6]) for _ in range(n)], start_id=1,
                             loop_ub=rng.choice([0, 1, 2, 9, 40]), bias=rng.choice([1, 2]), tgs_pairs=tg,
                             junc_info=rng.choice([False, True])))
    sols = ctx.solve_batch(probs)
    for p, s in zip(probs, sols):
        check_against_oracle(p, s)


def test_large_host_batch_compact_upload_in_pieces(ctx):
    """6000 partitions in one ambi_bfb_solve_batch call: the host side flattens them on its worker pool into the compact
    form (with third-generation masks: some partitions have them), copies it in 512-partition pieces and the device
    expands it; results are mapped back on the pool.  200 distinct ragged problems, 30 shuffled copies of each: every
    copy must carry the oracle's answer for its problem."""
    rng = random.Random(77)
    distinct = []
    for _ in range(200):
        n = rng.randint(1, 9)
        tg = []
        for _ in range(rng.choice([0, 0, 1, 2])):
            a = rng.randrange(n); b = rng.randrange(a, n); tg += [a + 1, b + 1]
        distinct.append(Problem(n=n, seg_cn2=[rng.choice([0, 1, 2, 4, 6, 7, 12]) for _ in range(n)],
                                fbi_cn2=[rng.choice([0, 0, 1, 2, 4]) for _ in range(n)], start_id=1,
                                loop_ub=rng.choice([0, 1, 2, 9]), bias=1, tgs_pairs=tg, junc_info=rng.choice([False, True])))
    order = [i for i in range(200) for _ in range(30)]
    rng.shuffle(order)
    sols = ctx.solve_batch([distinct[i] for i in order])
    want = {}
    for i, s in zip(order, sols):
        if i not in want:
            check_against_oracle(distinct[i], s)
            want[i] = (s.score2, tuple(int(v) for v in s.element_cn))
        assert (s.score2, tuple(int(v) for v in s.element_cn)) == want[i]


def test_small_n_vs_bruteforce(any_ctx):
    ctx = any_ctx
    rng = random.Random(9)
    probs = [Problem(n=n, seg_cn2=[rng.randint(0, 12) for _ in range(n)], fbi_cn2=[rng.randint(0, 6) for _ in range(n)],
                     loop_ub=rng.choice([1, 2, 3])) for n in [1, 2, 3] * 20]
    sols = ctx.solve_batch(probs)
    for p, s in zip(probs, sols):
        best, xb, _ = orc.bruteforce(p, lmax=3)
        assert s.score2 == best and np.array_equal(s.element_cn, xb)


def test_results_do_not_depend_on_beam_or_batching():
    probs = [lh.problem_from_lh(synth.generate(12, 4, 300 + s, 0.5).lh_text()) for s in range(24)]
    ref = None
    for beam in (1, 8, 64, 512):
        with capi.Context(device=0, beam_width=beam) as c:
            sols = c.solve_batch(probs)
            got = [(s.score2, tuple(int(v) for v in s.element_cn)) for s in sols]
            if ref is None:
                ref = got
            assert got == ref
            shuffled = list(reversed(probs))
            got2 = [(s.score2, tuple(int(v) for v in s.element_cn)) for s in c.solve_batch(shuffled)]
            assert got2 == list(reversed(ref))


def test_repeated_runs_are_identical(ctx):
    probs = [lh.problem_from_lh(synth.generate(14, 4, 500 + s, 0.5).lh_text()) for s in range(16)]
    a = ctx.solve_batch(probs)
    b = ctx.solve_batch(probs)
    for s, t in zip(a, b):
        assert s.score2 == t.score2 and np.array_equal(s.element_cn, t.element_cn)


def test_stage_b_vs_oracle_batch(ctx):
    probs = [lh.problem_from_lh(synth.generate(n, f, 700 + s, 0.5).lh_text())
             for (n, f) in [(6, 3), (10, 4), (12, 4), (16, 5)] for s in range(12)]
    sols = ctx.solve_batch(probs)
    dags = [capi.construct_dag(p.n, p.start_id, s.element_cn) for p, s in zip(probs, sols)]
    for rev in (False, True):
        res = ctx.first_valid_order_batch(dags, [rev] * len(dags))
        for p, s, d, r in zip(probs, sols, dags, res):
            od = orc.construct_dag(p.n, p.start_id, s.element_cn)
            assert od.nodes() == d.nodes() and od.edges() == d.edges()
            ok, flipped, order, bkp, _ = orc.first_valid_order(od, rev)
            assert r.found == ok
            if ok:
                assert (r.used_flipped, r.order, r.bkp) == (flipped, order, bkp)
                assert capi.expand_path(r.bkp) == orc.expand_path(bkp)


def test_prefix_sharded_solve_combines_to_the_same_optimum():
    """multi-GPU path on one device: every shard explores only the prefixes hashed to it; the arg-min
    of the shard optima in (score, search order) must equal the unsharded answer"""
    probs = [lh.problem_from_lh(synth.generate(n, f, 40 + s, 0.5).lh_text()) for (n, f, s) in
             [(6, 3, 0), (10, 4, 1), (12, 4, 2), (12, 4, 3), (16, 5, 4)]]
    with capi.Context(device=0) as c1:
        want = [c1.solve(p) for p in probs]
    for world in (2, 4):
        shard_sols = []
        for rank in range(world):
            with capi.Context(device=0, rank=rank, world=world) as c:
                shard_sols.append([c.solve_shard(p) for p in probs])
        for i, p in enumerate(probs):
            best = None
            for rank in range(world):
                s = shard_sols[rank][i]
                if best is None or capi.better(p.n, s.score2, s.element_cn, best.score2, best.element_cn):
                    best = s
            assert best.score2 == want[i].score2 and np.array_equal(best.element_cn, want[i].element_cn)


def test_small_buffers_fall_back_to_chunked_depth_first_pass():
    """when a level does not fit the candidate buffers the exact pass is redone depth-first in chunks:
    same answer, bit for bit"""
    probs = [lh.problem_from_lh(synth.generate(n, f, s, 0.5).lh_text()) for (n, f, s) in
             [(12, 4, 1), (16, 5, 0), (16, 5, 3), (20, 6, 1)]]
    for small in (False, True):
        with capi.Context(device=0, node_capacity=4096, small_kernel=small) as c:
            for p in probs:
                check_against_oracle(p, c.solve(p))
            for p, s in zip(probs, c.solve_batch(probs)):
                check_against_oracle(p, s)


def test_overflow_is_reported_loudly(monkeypatch):
    monkeypatch.setenv("AMBI_NO_CHUNKED", "1")
    p = lh.problem_from_lh(synth.generate(16, 5, 0, 0.5).lh_text())
    with capi.Context(device=0, node_capacity=64) as c:
        with pytest.raises(capi.AmbiError) as e:
            c.solve(p)
        assert e.value.status == -5


def test_exhausted_frontier_pool_is_an_error_never_a_silent_answer(monkeypatch):
    """the chunked depth-first pass itself can run out of room (frontier pool exhausted): the partial best it
    holds is not proven optimal, so the call must return AMBI_ERR_FRONTIER_OVERFLOW, never AMBI_OK"""
    monkeypatch.setenv("AMBI_POOL_BYTES", "65536")
    p = lh.problem_from_lh(synth.generate(20, 6, 1, 0.5).lh_text())
    with capi.Context(device=0, node_capacity=4096) as c:
        with pytest.raises(capi.AmbiError) as e:
            c.solve(p)
        assert e.value.status == -5
        with pytest.raises(capi.AmbiError) as e:
            c.solve_batch([p, p])
        assert e.value.status == -5
    monkeypatch.delenv("AMBI_POOL_BYTES")
    with capi.Context(device=0, node_capacity=4096) as c:        # same buffers, default pool: exact answer
        check_against_oracle(p, c.solve(p))


def test_small_kernel_hands_wide_frontiers_to_the_global_path(monkeypatch):
    """a tiny per-warp record capacity: most partitions do not fit the warp-per-partition kernel and are handed on
    to the global-memory path inside the same call; answers are the oracle's either way"""
    monkeypatch.setenv("AMBI_SMALL_C", "32")
    probs = [lh.problem_from_lh(synth.generate(n, f, 60 + s, 0.5).lh_text()) for (n, f) in [(8, 3), (12, 4), (14, 4), (16, 5)] for s in range(6)]
    with capi.Context(device=0, small_kernel=True) as c:
        b = c.upload(probs)
        b.run(); c.sync()
        sols = b.fetch()
        pr = c.profile()
        b.free()
        assert pr["small_capacity"] == 32 and pr["small_solved"] + pr["global_solved"] == len(probs)
        assert pr["global_solved"] > 0 and pr["small_solved"] > 0
        for p, s in zip(probs, sols):
            check_against_oracle(p, s)


def test_time_to_best_and_time_to_proof_per_partition(any_ctx):
    """per-partition device-clock times (BASELINE metric, second half): 0 <= time to best <= time to proof"""
    probs = [lh.problem_from_lh(synth.generate(12, 4, 800 + s, 0.5).lh_text()) for s in range(64)]
    sols = any_ctx.solve_batch(probs)
    for s in sols:
        assert 0.0 <= s.seconds_to_best <= s.seconds and 0.0 < s.seconds < 10.0
    assert len({round(s.seconds_to_best * 1e9) for s in sols}) > 1        # not one number for the whole batch


FRACTIONAL = json.load(open(os.path.join(ROOT, "tests", "golden", "cases_fractional.json")))["cases"]


def test_real_valued_copy_numbers_reference_fixtures(ctx):
    """copy numbers that are not half-integral (multiples of 1/8 copy: unit 4): the reference's LP takes any real
    right-hand side; fixtures solved by HiGHS through the reference binary (tools/make_golden.py fractional)"""
    parts = []
    for c in FRACTIONAL:
        g = lh.parse_lh(c["lh"])
        parts.append(lh.flatten(g, [], False, c["unit"])[0])
    sols = ctx.solve_batch([pt.problem for pt in parts])
    for c, pt, s in zip(FRACTIONAL, parts, sols):
        p = pt.problem
        assert s.status == 0 and s.score2 == c["oracle_score2"], c["name"]
        assert abs(2 * c["unit"] * (c["ref_highs_objective"] + c["bias"]) - s.score2) < 1e-6, c["name"]
        assert s.objective2 == s.score2 - 2 * c["unit"] * c["bias"]
        assert [[int(i), int(v)] for i, v in enumerate(s.element_cn) if v] == c["oracle_x_nz"], c["name"]
        dag = capi.construct_dag(p.n, p.start_id, s.element_cn)
        r = ctx.first_valid_order(dag, False)
        if dag.k <= 16 and c["ref_inject_paths"]:
            assert r.found and capi.format_path(capi.expand_path(capi.imperfect_fbi(r.bkp, pt.inversions))) == c["ref_inject_paths"][0], c["name"]


@pytest.mark.parametrize("unit", [2, 8, 64])
def test_units_vs_oracle_and_scale_invariance(ctx, unit):
    """(i) random profiles on the 1/(2*unit) grid against the oracle's *_units search, mixed with half-copy problems
    in one batch; (ii) a half-copy problem restated in a finer unit has the same assignment and unit x the score"""
    rng = random.Random(100 + unit)
    probs = []
    for s in range(24):
        n = rng.randint(1, 14)
        base = lh.problem_from_lh(synth.generate(max(n, 2), 3, 900 + s, 0.5).lh_text()) if n >= 2 else None
        c2 = [max(0, (2 * unit * v // 2) + rng.randint(-unit, unit)) for v in (base.seg_cn2 if base else [rng.randint(2, 12)])]
        j2 = [max(0, (2 * unit * v // 2) + rng.randint(0, unit // 2)) if v else 0 for v in (base.fbi_cn2 if base else [rng.randint(0, 6)])]
        probs.append(Problem(n=len(c2), seg_cn2=c2, fbi_cn2=j2, loop_ub=rng.choice([2, 9]), unit=unit))
        if base is not None:
            probs.append(base)
    sols = ctx.solve_batch(probs)
    for p, s in zip(probs, sols):
        check_against_oracle(p, s) if p.unit == 1 else None
        x, st = orc.solve(p, want_unique=False)
        assert s.score2 == st.score2 and np.array_equal(s.element_cn, x), (p, s.score2, st.score2)
    halves = [lh.problem_from_lh(synth.generate(12, 4, 950 + s, 0.5).lh_text()) for s in range(8)]
    fine = [Problem(n=p.n, seg_cn2=[unit * v for v in p.seg_cn2], fbi_cn2=[unit * v for v in p.fbi_cn2], loop_ub=p.effective_loop_ub(), unit=unit) for p in halves]
    for a, b in zip(ctx.solve_batch(halves), ctx.solve_batch(fine)):
        assert b.score2 == unit * a.score2 and np.array_equal(a.element_cn, b.element_cn)


def test_program_in_one_cooperative_kernel_gives_the_same_answers(monkeypatch):
    """AMBI_FUSED_MAX: batches of up to that many partitions run their whole launch sequence inside one cooperative
    kernel (k_program), stopping where a frontier outgrows its small grid; same answers as launch by launch"""
    probs = [lh.problem_from_lh(synth.generate(n, f, 70 + s, 0.5).lh_text()) for (n, f) in [(6, 3), (12, 4), (16, 5), (20, 6)] for s in range(3)]
    monkeypatch.setenv("AMBI_FUSED_MAX", "64")
    monkeypatch.setenv("AMBI_FUSED_BIG", "2048")
    with capi.Context(device=0) as c:
        fused = [c.solve(p) for p in probs] + c.solve_batch(probs)
    for p, s in zip(probs + probs, fused):
        check_against_oracle(p, s)
    with capi.Context(device=0, node_capacity=4096) as c:          # overflow inside the cooperative kernel -> chunked pass
        for p in probs[-3:]:
            check_against_oracle(p, c.solve(p))


def test_bad_arguments(ctx):
    with pytest.raises(capi.AmbiError):
        ctx.solve(Problem(n=33, seg_cn2=[2] * 33, fbi_cn2=[0] * 33))


def test_config3_n20_with_juncs_vs_oracle(ctx):
    """BASELINE config 3: n=20 / 6 folds with a .juncs file and --junc_info true (third-generation row)"""
    for seed in (3, 5, 8):
        sm = synth.generate(20, 6, seed, 0.0, with_juncs=True)
        p = lh.problem_from_lh(sm.lh_text(), sm.juncs_text(), True)
        assert p.junc_info and len(p.tgs_pairs) > 0
        check_against_oracle(p, ctx.solve(p))


def test_config4_n30_noisy_vs_oracle(ctx):
    """BASELINE config 4: n=30 / 8 folds, CN noise +-1 w.p. 0.5 (two seeds the oracle finishes in seconds)"""
    for seed in (0, 1):
        p = lh.problem_from_lh(synth.generate(30, 8, seed, 0.5).lh_text())
        check_against_oracle(p, ctx.solve(p))


def test_config5_sweep_properties():
    """BASELINE config 5 at full segment counts (n in 10..25, folds = ceil(n/3.5)): size-independent
    properties instead of the oracle -- the returned assignment satisfies every row of the reference
    model with exactly the reported score, never exceeds the incumbent, and is identical
    under a different beam width and under tiny buffers (chunked depth-first pass)."""
    rng = random.Random(21)
    probs = []
    for s in range(96):
        n = rng.randint(10, 25)
        probs.append(lh.problem_from_lh(synth.generate(n, synth.folds_for(n), 9000 + s, 0.0).lh_text()))
    with capi.Context(device=0) as c:
        a = c.solve_batch(probs)
    for p, s in zip(probs, a):
        ok, s2 = orc.model_eval(p, s.element_cn)
        assert s.status == 0 and ok and s2 == s.score2 and s.score2 <= s.incumbent_score2
        assert s.score2 <= sum(p.seg_cn2) + sum(p.fbi_cn2)
    with capi.Context(device=0, beam_width=64, node_capacity=200000) as c:
        b = c.solve_batch(probs)
    for s, t in zip(a, b):
        assert s.score2 == t.score2 and np.array_equal(s.element_cn, t.element_cn)
    # the small ones also against the oracle
    for p, s in zip(probs, a):
        if p.n <= 16:
            check_against_oracle(p, s)


def test_all_valid_orders_matches_reference_print_all(ctx):
    """--all (printAll): orders, their order, the flip rule and the paths, against the reference's stdout"""
    seen_multi = 0
    for c in CASES:
        if "ref_inject_paths_all" not in c or c["name"].startswith("indel_"):
            continue
        part = partition_of(c)
        p = part.problem
        x = np.zeros(p.n * (p.n + 1), dtype=np.int32)
        for i, v in c["oracle_x_nz"]:
            x[i] = v
        dag = capi.construct_dag(p.n, p.start_id, x)
        for rev, key in ((False, "ref_inject_paths_all"), (True, "ref_inject_paths_all_rev")):
            first, second, paths = ctx.all_valid_orders(dag, rev)
            got = [capi.format_path(capi.expand_path(capi.imperfect_fbi(b, part.inversions))) for b in paths]
            assert got == c[key], c["name"]
            od = orc.construct_dag(p.n, p.start_id, x)
            assert first == orc.all_valid_orders(od, forward=not rev)
            seen_multi += len(first) > 1
    assert seen_multi >= 10


def test_all_valid_paths_device_surgery_matches_reference_print_all(ctx):
    """--all with imperfectFBI and the expansion on the device (ambi_bfb_all_valid_paths): the vertex paths the reference
    prints, incl. the imperfect fold-back fixtures, against its stdout and against the host-side surgery"""
    checked = imperfect = 0
    for c in CASES:
        if "ref_inject_paths_all" not in c or c["name"].startswith("indel_"):
            continue
        part = partition_of(c)
        p = part.problem
        x = np.zeros(p.n * (p.n + 1), dtype=np.int32)
        for i, v in c["oracle_x_nz"]:
            x[i] = v
        dag = capi.construct_dag(p.n, p.start_id, x)
        for rev, key in ((False, "ref_inject_paths_all"), (True, "ref_inject_paths_all_rev")):
            n0, n1, paths = ctx.all_valid_paths(dag, part.inversions, rev)
            assert [capi.format_path(v) for v in paths] == c[key], c["name"]
            first, second, bk = ctx.all_valid_orders(dag, rev)
            assert (n0, n1) == (len(first), len(second))
            assert paths == [capi.expand_path(capi.imperfect_fbi(b, part.inversions)) for b in bk]
            checked += 1
        imperfect += c["name"].startswith("imperfect_")
    assert checked >= 80 and imperfect >= 10


def test_measured_peaks_are_plausible(ctx):
    """on-box roofline denominators: integer issue rate and shared-memory bandwidth"""
    pk = ctx.measure_peaks()
    assert pk["sms"] >= 100
    assert 5.0 < pk["alu_tops"] < 60.0 and 5.0 < pk["alu_fma_tops"] < 60.0
    assert 10.0 < pk["smem_tbs"] < 60.0


def test_profile_funnel_counters(ctx):
    ps = [lh.problem_from_lh(synth.generate(12, 4, s).lh_text()) for s in range(64)]
    b = ctx.upload(ps)
    ctx.set_profile(True)
    b.run(); ctx.sync()
    sols = b.fetch()
    pr = ctx.profile()
    ctx.set_profile(False)
    b.free()
    cand = sum(s.candidates_scored for s in sols)
    # candidates = configurations of the expanded records; the ones whose columns the prefilter refuses outright are
    # dropped from the masks before the enumeration (k_expand phase 1), so fewer are enumerated than counted
    assert cand >= pr["bounds_evaluated"] > 0 and cand + 64 * 12 * 64 >= pr["configs_enumerated"]
    assert pr["configs_enumerated"] >= pr["bounds_evaluated"] >= pr["children_written"] > 0
    assert pr["bound_batches"] * 32 >= pr["bounds_evaluated"]
