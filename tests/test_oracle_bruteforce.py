"""CPU: the oracle's exact search against exhaustive enumeration of the reference model (n <= 3),
including the tie-break (smallest assignment in search order) and the uniqueness flag."""
import random

import numpy as np

from ambigram_b200.capi import Problem
from tests import oracle_lib as orc


def rand_problem(rng, n):
    nt = rng.choice([0, 0, 1, 2, 3])
    tg = []
    for _ in range(nt):
        a = rng.randrange(n)
        b = rng.randrange(a, n)
        tg += [a + 1, b + 1]
    return Problem(n=n, seg_cn2=[rng.randint(0, 14) for _ in range(n)], fbi_cn2=[rng.randint(0, 8) for _ in range(n)],
                   start_id=1, loop_ub=rng.choice([0, 1, 2, 3, 5]), bias=1, tgs_pairs=tg, junc_info=rng.choice([False, True]))


def test_search_equals_bruteforce_small_n():
    rng = random.Random(1)
    for _ in range(300):
        p = rand_problem(rng, rng.choice([1, 2, 2, 3, 3, 3]))
        best, xb, nopt = orc.bruteforce(p, lmax=3)
        x, st = orc.solve(p, want_unique=True)
        assert st.score2 == best
        assert np.array_equal(x, xb), (p, list(x), list(xb))
        assert st.unique == (1 if nopt == 1 else 0)


def test_all_zero_is_feasible_and_scores_sum():
    rng = random.Random(2)
    for n in (1, 2, 5, 9):
        p = rand_problem(rng, n)
        ok, s2 = orc.model_eval(p, np.zeros(n * (n + 1), dtype=np.int32))
        assert ok and s2 == sum(p.seg_cn2) + sum(p.fbi_cn2)


def test_effective_domains():
    """for n >= 2 every loop variable is capped at 2 by the children rows alone."""
    p = Problem(n=3, seg_cn2=[40, 40, 40], fbi_cn2=[20, 20, 20], loop_ub=60)
    x = np.zeros(12, dtype=np.int32)
    x[6 + 2] = 3                      # L[1,3] = 3 violates LGM:4615-4646
    ok, _ = orc.model_eval(p, x)
    assert not ok
    x[6 + 2] = 2
    ok, _ = orc.model_eval(p, x)
    assert ok
