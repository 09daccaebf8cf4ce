"""CPU: the body of the warp-per-partition CUDA kernel (ambigram_b200/csrc/small_core.cuh) compiled with g++ against a
32-lane warp emulator (tests/emu) and checked against the oracle: same score2, same element_cn (search-order
tie-break), frontier-capacity overflow reported rather than answered.  The same source is what nvcc compiles for
sm_100a; the GPU parity tests (-m gpu) run it on the device."""
import random

import numpy as np
import pytest

from ambigram_b200 import lh, synth
from ambigram_b200.capi import Problem
from tests import emu_lib as emu
from tests import oracle_lib as orc


def check(p, capacity=256):
    rc, x, sc, inc, cand, nodes, fun = emu.solve(p, capacity)
    xo, st = orc.solve(p, want_unique=False)
    if rc == 1:
        return False                      # did not fit: the product hands the partition to the global-memory path
    assert rc == 0 and sc == st.score2 and np.array_equal(x, xo), (p, sc, st.score2)
    assert inc >= sc and cand > 0 and nodes > 0
    return True


@pytest.mark.parametrize("n,folds,noise,count", [(2, 2, 0.5, 6), (3, 2, 0.5, 6), (5, 3, 0.5, 8), (8, 3, 0.5, 8), (12, 4, 0.0, 8),
                                                 (12, 4, 0.5, 8), (14, 4, 0.5, 4), (16, 5, 0.0, 4)])
def test_emulated_kernel_vs_oracle(n, folds, noise, count):
    solved = sum(check(lh.problem_from_lh(synth.generate(n, folds, 100 + s, noise).lh_text())) for s in range(count))
    assert solved >= count // 2


def test_emulated_kernel_third_generation_rows_and_ragged_profiles():
    for s in range(6):
        sm = synth.generate(12, 4, 300 + s, 0.0, with_juncs=True)
        check(lh.problem_from_lh(sm.lh_text(), sm.juncs_text(), True), 1024)
    rng = random.Random(5)
    solved = 0
    for _ in range(150):
        n = rng.randint(2, 10)
        tg = []
        for _ in range(rng.choice([0, 0, 1, 3, 6])):
            a = rng.randrange(n); b = rng.randrange(a, n); tg += [a + 1, b + 1]
        p = Problem(n=n, seg_cn2=[rng.choice([0, 1, 2, 4, 6, 7, 12, 20]) for _ in range(n)],
                    fbi_cn2=[rng.choice([0, 0, 1, 2, 4, 6]) for _ in range(n)], start_id=1,
                    loop_ub=rng.choice([0, 1, 2, 9, 40]), bias=rng.choice([1, 2]), tgs_pairs=tg, junc_info=rng.choice([False, True]))
        solved += check(p, 1024)
    assert solved >= 140


def test_emulated_kernel_small_capacity_overflows_or_agrees():
    """the lowered-limit retries (bisection between the root bound and the smallest limit known not to fit) never
    change the answer: whatever capacity, a returned assignment is the oracle's"""
    probs = [lh.problem_from_lh(synth.generate(12, 4, 500 + s, 0.5).lh_text()) for s in range(10)]
    for cap in (32, 64, 4096):
        solved = sum(check(p, cap) for p in probs)
        if cap == 4096:
            assert solved == len(probs)
