"""Randomised campaign (test infrastructure): the product's Stage-A kernel bodies under the warp emulator (tests/emu/expand_emu.cpp)
against the oracle on random ragged batches -- arbitrary profiles with third-generation rows and finer units, seeded
synthetic partitions of 4..13 segments, beam widths 2 / 16 / 48.   python tools/emu_campaign.py <seed> <rounds of 80 problems>"""
import sys, os, random, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
from tests import test_expand_emu as T
from tests import oracle_lib as orc
from ambigram_b200 import lh, synth
from ambigram_b200.capi import Problem
emu = C.CDLL(T.LIB)
seed0 = int(sys.argv[1]); rounds = int(sys.argv[2])
bad = 0; total = 0
t0 = time.time()
for rd in range(rounds):
    rng = random.Random(seed0 * 1000 + rd)
    probs = []
    for _ in range(80):
        kind = rng.random()
        if kind < 0.5:
            n = rng.randint(2, 11)
            tg = []
            for _ in range(rng.choice([0, 0, 0, 1, 2, 4])):
                a = rng.randrange(n); b = rng.randrange(a, n); tg += [a + 1, b + 1]
            unit = rng.choice([1, 1, 1, 2, 4, 16])
            probs.append(Problem(n=n, seg_cn2=[rng.choice([0, 1, 2, 3, 4, 6, 7, 9, 12, 20, 33]) * (unit if rng.random() < 0.7 else 1) for _ in range(n)],
                                 fbi_cn2=[rng.choice([0, 0, 1, 2, 3, 4, 6]) * (unit if rng.random() < 0.7 else 1) for _ in range(n)], start_id=1,
                                 loop_ub=rng.choice([0, 1, 2, 3, 9, 40]), bias=1, tgs_pairs=tg, junc_info=rng.choice([False, True]), unit=unit))
        else:
            n = rng.randint(4, 13)
            probs.append(lh.problem_from_lh(synth.generate(n, synth.folds_for(n) if rng.random() < 0.5 else 3, rng.randrange(10**6), rng.choice([0.0, 0.5, 1.0])).lh_text()))
    try:
        res, fun = T.solve_batch(emu, probs, cap=400000, beam=rng.choice([2, 16, 16, 48]))
    except AssertionError:
        try:
            res, fun = T.solve_batch(emu, probs, cap=1500000, beam=16)
        except AssertionError:
            print("round", rd, "does not fit the emulated buffers: skipped", flush=True)
            continue
    for p, (score, x, inc, cand) in zip(probs, res):
        xo, st = orc.solve(p, want_unique=False)
        total += 1
        if score != st.score2 or not np.array_equal(x, xo):
            bad += 1
            print("MISMATCH", seed0, rd, p, score, st.score2, flush=True)
    print(f"round {rd}: {total} problems, {bad} mismatches, {time.time() - t0:.0f}s", flush=True)
