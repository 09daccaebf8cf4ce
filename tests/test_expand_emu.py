"""CPU: the product's level-synchronous Stage-A kernels (ambigram_b200/csrc/stage_a.cu: k_expand's body with its
chunked cursor enumeration, mask restriction, bound and child write; the merge; the beam cut; k_materialise; ...) compiled
with g++ against the 32-lane warp emulator (tests/emu/expand_emu.cpp, test infrastructure) and played in the order
StageA::run() launches them -- greedy dive, beam, exact pass -- against the oracle.  The same comparison runs on the
GPU in tests/test_gpu_parity.py; this one pins the kernel logic on a box without a GPU (both instantiation widths,
half-copy and finer units, third-generation rows, batches whose record groups mix partitions)."""
import ctypes as C
import os
import random
import subprocess

import numpy as np
import pytest

from ambigram_b200 import lh, synth
from ambigram_b200.capi import Problem
from tests import oracle_lib as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "emu", "expand_emu.cpp")
LIB = os.path.join(ROOT, "tests", "emu", "_build", "libexpand_emu.so")
DEPS = [SRC, os.path.join(ROOT, "tests", "emu", "simt_emu.h")] + [os.path.join(ROOT, "ambigram_b200", "csrc", f)
                                                                 for f in ("stage_a.cu", "decode_common.cuh", "host_flatten.h", "ambi_common.cuh")]
I32P = C.POINTER(C.c_int32)


@pytest.fixture(scope="module")
def emu():
    if not os.path.exists(LIB) or any(os.path.getmtime(LIB) < os.path.getmtime(f) for f in DEPS):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-w", "-shared", "-fPIC", "-o", LIB, SRC])
    return C.CDLL(LIB)


def solve_batch(emu, problems, cap=20000, beam=16):
    cnt = len(problems)
    n = np.array([p.n for p in problems], dtype=np.int32)
    c2 = np.zeros((cnt, 32), dtype=np.int32); j2 = np.zeros((cnt, 32), dtype=np.int32)
    max_tgs = max(1, max(len(p.tgs_pairs) // 2 for p in problems))
    tgs = np.zeros((cnt, 2 * max_tgs), dtype=np.int32)
    for i, p in enumerate(problems):
        c2[i, :p.n] = p.seg_cn2; j2[i, :p.n] = p.fbi_cn2
        t = [int(v) - p.start_id for v in p.tgs_pairs]
        tgs[i, :len(t)] = t
    lub = np.array([p.effective_loop_ub() for p in problems], dtype=np.int32)
    ush = np.array([int(p.unit).bit_length() - 1 for p in problems], dtype=np.int32)
    ntg = np.array([len(p.tgs_pairs) // 2 for p in problems], dtype=np.int32)
    ji = np.array([int(p.junc_info) for p in problems], dtype=np.int32)
    x = np.zeros((cnt, 32 * 33), dtype=np.int32)
    sc = np.zeros(cnt, dtype=np.int64); inc = np.zeros(cnt, dtype=np.int32); cand = np.zeros(cnt, dtype=np.uint64)
    fun = (C.c_uint64 * 4)()
    ptr = lambda a: a.ctypes.data_as(I32P)
    rc = emu.emu_expand_solve(cnt, ptr(n), ptr(c2), ptr(j2), ptr(lub), ptr(ush), ptr(ntg), ptr(tgs), max_tgs, ptr(ji), cap, beam, ptr(x),
                              sc.ctypes.data_as(C.POINTER(C.c_int64)), ptr(inc), cand.ctypes.data_as(C.POINTER(C.c_uint64)), fun)
    assert rc == 0, rc
    return [(int(sc[i]), x[i, :p.n * (p.n + 1)].copy(), int(inc[i]), int(cand[i])) for i, p in enumerate(problems)], [int(v) for v in fun]


def check(problems, results):
    for p, (score, x, inc, cand) in zip(problems, results):
        xo, st = orc.solve(p, want_unique=False)
        assert score == st.score2 and np.array_equal(x, xo), (p, score, st.score2)
        assert inc >= score and cand > 0


def test_ragged_batch_with_third_generation_rows(emu):
    rng = random.Random(5)
    probs = []
    for _ in range(60):
        n = rng.randint(2, 9)
        tg = []
        for _ in range(rng.choice([0, 0, 1, 3])):
            a = rng.randrange(n); b = rng.randrange(a, n); tg += [a + 1, b + 1]
        probs.append(Problem(n=n, seg_cn2=[rng.choice([0, 1, 2, 4, 6, 7, 12, 20]) for _ in range(n)],
                             fbi_cn2=[rng.choice([0, 0, 1, 2, 4, 6]) for _ in range(n)], start_id=1,
                             loop_ub=rng.choice([0, 1, 2, 9, 40]), bias=1, tgs_pairs=tg, junc_info=rng.choice([False, True])))
    res, fun = solve_batch(emu, probs)
    check(probs, res)
    assert fun[0] >= fun[1] >= fun[2] > 0           # enumerated >= bounds evaluated >= children written


@pytest.mark.parametrize("n,folds,noise,count", [(6, 3, 0.5, 10), (10, 4, 0.5, 6), (12, 4, 0.0, 4), (12, 4, 0.5, 4)])
def test_seeded_synthetic_partitions(emu, n, folds, noise, count):
    probs = [lh.problem_from_lh(synth.generate(n, folds, 100 + s, noise).lh_text()) for s in range(count)]
    res, _ = solve_batch(emu, probs, cap=60000)
    check(probs, res)


def test_wide_instantiation_and_results_independent_of_the_beam(emu):
    """a partition of more than 16 segments selects the 32-column staging; the beam width only changes the incumbent"""
    probs = [lh.problem_from_lh(synth.generate(18, 5, 7, 0.0).lh_text()), lh.problem_from_lh(synth.generate(8, 3, 8, 0.5).lh_text())]
    res, _ = solve_batch(emu, probs, cap=120000)
    check(probs, res)
    probs12 = [lh.problem_from_lh(synth.generate(12, 4, 300 + s, 0.5).lh_text()) for s in range(3)]
    a, _ = solve_batch(emu, probs12, cap=200000, beam=16)
    for w in (4, 64):
        b, _ = solve_batch(emu, probs12, cap=200000, beam=w)
        for (sa, xa, _, _), (sb, xb, _, _) in zip(a, b):
            assert sa == sb and np.array_equal(xa, xb)


def test_copy_numbers_in_finer_units(emu):
    rng = random.Random(104)
    probs = []
    for s in range(10):
        unit = rng.choice([2, 4, 8])
        base = lh.problem_from_lh(synth.generate(rng.randint(3, 9), 3, 900 + s, 0.5).lh_text())
        c2 = [max(0, unit * v + rng.randint(-unit, unit)) for v in base.seg_cn2]
        j2 = [max(0, unit * v + rng.randint(0, unit // 2)) if v else 0 for v in base.fbi_cn2]
        probs.append(Problem(n=base.n, seg_cn2=c2, fbi_cn2=j2, loop_ub=rng.choice([2, 9]), unit=unit))
        probs.append(base)                            # half-copy partitions in the same batch
    res, _ = solve_batch(emu, probs, cap=60000)
    check(probs, res)


def solve_sharded(emu, problems, world, level, cap=200000, beam=16):
    cnt = len(problems)
    n = np.array([p.n for p in problems], dtype=np.int32)
    c2 = np.zeros((cnt, 32), dtype=np.int32); j2 = np.zeros((cnt, 32), dtype=np.int32)
    for i, p in enumerate(problems):
        c2[i, :p.n] = p.seg_cn2; j2[i, :p.n] = p.fbi_cn2
    lub = np.array([p.effective_loop_ub() for p in problems], dtype=np.int32)
    zero = np.zeros(cnt, dtype=np.int32); tgs = np.zeros((cnt, 2), dtype=np.int32)
    x = np.zeros((cnt, 32 * 33), dtype=np.int32)
    sc = np.zeros(cnt, dtype=np.int64); cand = np.zeros(cnt, dtype=np.uint64)
    ptr = lambda a: a.ctypes.data_as(I32P)
    rc = emu.emu_expand_solve_sharded(world, level, cnt, ptr(n), ptr(c2), ptr(j2), ptr(lub), ptr(zero), ptr(zero), ptr(tgs), 1, ptr(zero), cap, beam,
                                      ptr(x), sc.ctypes.data_as(C.POINTER(C.c_int64)), cand.ctypes.data_as(C.POINTER(C.c_uint64)))
    assert rc == 0, rc
    return [(int(sc[i]), x[i, :p.n * (p.n + 1)].copy(), int(cand[i])) for i, p in enumerate(problems)]


@pytest.mark.parametrize("world,level", [(2, 2), (3, 1), (4, 2), (8, 3)])
def test_prefix_shards_combine_to_the_optimum(emu, world, level):
    """north_star's multi-GPU split of one search: each rank keeps the children of row `level` whose
    prefix hashes to it; the arg-min over the ranks in (score, search order) is the unsharded optimum, tie-break included"""
    probs = [lh.problem_from_lh(synth.generate(n, f, 40 + s, noise).lh_text()) for (n, f, s, noise) in
             [(6, 3, 0, 0.5), (9, 3, 1, 0.5), (10, 4, 2, 0.0), (12, 4, 3, 0.5), (12, 4, 4, 0.0)]]
    got = solve_sharded(emu, probs, world, level)
    for p, (score, x, cand) in zip(probs, got):
        xo, st = orc.solve(p, want_unique=False)
        assert score == st.score2 and np.array_equal(x, xo), (p.n, score, st.score2)
