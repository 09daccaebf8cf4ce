"""CPU: why `--op sc_bfb` is not built -- executable evidence for DESIGN.md section 9.

tests/golden/sc_probe.json holds the LP that the reference's own BFB_ILP_SC (src/LocalGenomicMap.cpp:4754-5093)
emitted for two 4-segment clones (`tools/make_golden.py sc_probe`: the reference binary built from its unmodified
sources, model dumped by the COIN shim).  The model indexes the deviation epsilons of a clone's segment rows as
`numElements + idx/2` with `idx` the GLOBAL row counter (LGM:4814, 4820, 4858, 4864).  That is right for the first
clone only: the second clone's rows land on columns that the coupling rows |x_i - x_j| use as well (`cnt/2`,
LGM:5026-5070) and its own epsilon columns appear in no row.  The joint model the reference solves for >= 2 clones is
therefore neither "the single-clone model per clone plus coupling" nor anything a restatement could be pinned to
without copying the accident; the single-clone case of sc_bfb is the `bfb` model, which is what the library implements."""
import collections
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROBE = json.load(open(os.path.join(ROOT, "tests", "golden", "sc_probe.json")))

N_SEG, N_CLONES = 4, 2
N_VARS_PER_CLONE = N_SEG * (N_SEG + 1)          # patterns + loops of one clone (variableIdx, localhap.cpp:122-133)
N_ELEMENTS = N_VARS_PER_CLONE * N_CLONES        # LGM:4761
ROWS_PER_CLONE = 64                             # 16 deviation rows (4 per segment) + 48 structural rows at n = 4


def parse(lp_text):
    lines = lp_text.splitlines()
    assert lines[0].startswith("AMBI-LP")
    nrows, ncols = map(int, lines[1].split())
    rows = []
    for ln in lines[2:]:
        if ln.startswith("R "):
            t = ln.split()
            k = int(t[3])
            rows.append({int(a.split(":")[0]): float(a.split(":")[1]) for a in t[4:4 + k]})
    assert len(rows) == nrows
    return rows, ncols


def test_the_dump_is_the_two_clone_model():
    rows, ncols = parse(PROBE["lp"])
    assert PROBE["rc"] == 0 and len(rows) == 168 and ncols == 80
    # clone 1: deviation row pair i uses epsilon column numElements + i, as intended
    for i in range(2 * N_SEG):
        for r in (2 * i, 2 * i + 1):
            eps = [c for c in rows[r] if c >= N_ELEMENTS]
            assert eps == [N_ELEMENTS + i]
        assert rows[2 * i][N_ELEMENTS + i] == 1 and rows[2 * i + 1][N_ELEMENTS + i] == -1


def test_second_clone_deviation_columns_alias_the_coupling_columns():
    rows, ncols = parse(PROBE["lp"])
    use = collections.defaultdict(list)
    for r, ent in enumerate(rows):
        for c in ent:
            use[c].append(r)
    own = range(N_ELEMENTS + 2 * N_SEG, N_ELEMENTS + 4 * N_SEG)          # where clone 2's epsilons were meant to live
    assert all(c not in use for c in own)                                 # ... they appear in no row at all
    # clone 2's deviation rows (global idx 64..79) sit on numElements + idx/2 = 72..79 instead
    for i in range(2 * N_SEG):
        for r in (ROWS_PER_CLONE + 2 * i, ROWS_PER_CLONE + 2 * i + 1):
            eps = [c for c in rows[r] if c >= N_ELEMENTS]
            assert eps == [N_ELEMENTS + (ROWS_PER_CLONE + 2 * i) // 2]
    # ... and the first four of those columns are coupling epsilons as well (rows >= 128 are the |x_i - x_j| pairs)
    shared = [c for c in range(72, 80) if any(r >= 2 * ROWS_PER_CLONE for r in use[c]) and any(r < 2 * ROWS_PER_CLONE for r in use[c])]
    assert shared == [72, 73, 74, 75]
