"""CPU: the C-ABI library loads, exports every symbol include/ambigram_b200.h declares, its host-side
helpers agree with the oracle, and the product path fails loudly without a GPU (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

from ambigram_b200 import capi, lh, synth
from tests import oracle_lib as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    h = open(os.path.join(ROOT, "include", "ambigram_b200.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(ambi_[a-z0-9_]+)\s*\(", h)))


def test_library_exports_every_declared_symbol():
    lib = capi.load_library()
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/ambigram_b200.h but not exported"
    assert sorted(capi.EXPORTS) == syms
    assert lib.ambi_abi_version() == 1


def _has_gpu():
    import torch
    return torch.cuda.is_available()


@pytest.mark.skipif(_has_gpu(), reason="needs a box without a GPU")
def test_no_gpu_fails_loudly():
    with pytest.raises(capi.AmbiError) as e:
        capi.Context(device=0)
    assert e.value.status == -1


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(capi.AmbiError):
        capi.load_library(str(tmp_path / "nope.so"))


def test_construct_dag_matches_oracle():
    for n, folds, seed in [(6, 3, 0), (9, 3, 1), (12, 4, 2), (12, 4, 7), (16, 5, 1)]:
        p = lh.problem_from_lh(synth.generate(n, folds, seed, 0.5).lh_text())
        x, _ = orc.solve(p, want_unique=False)
        d = capi.construct_dag(p.n, p.start_id, x)
        o = orc.construct_dag(p.n, p.start_id, x)
        assert d.nodes() == o.nodes() and d.edges() == o.edges()


def test_dag_node_order_follows_string_keys():
    """std::map<string,int> order (LGM:3279): "l:1,10" sorts before "l:1,2"; loops by decreasing length."""
    n = 12
    x = np.zeros(n * (n + 1), dtype=np.int32)
    h = n * (n + 1) // 2
    for (a, b) in [(0, 11), (0, 1), (0, 9)]:
        x[h + capi.pair_index(n, a, b)] = 1
    for (a, b) in [(0, 9), (0, 1)]:
        x[capi.pair_index(n, a, b)] = 1
    d = capi.construct_dag(n, 1, x)
    assert d.nodes() == [("l", 1, 12, 1), ("l", 1, 10, 1), ("l", 1, 2, 1), ("p", 1, 10, 1), ("p", 1, 2, 1)]


def test_expand_and_format_path():
    bkp = [1, 6, -6, -2, 2, 4]
    assert capi.expand_path(bkp) == orc.expand_path(bkp) == [1, 2, 3, 4, 5, 6, -6, -5, -4, -3, -2, 2, 3, 4]
    assert capi.format_path(capi.expand_path(bkp)) == "1+2+3+4+5+6+|6-5-4-3-2-|2+3+4+"


def test_better_is_score_then_search_order():
    n = 3
    a = np.zeros(12, dtype=np.int32)
    b = np.zeros(12, dtype=np.int32)
    b[capi.pair_index(n, 0, 2)] = 1
    assert capi.better(n, 4, a, 5, b) and not capi.better(n, 5, a, 4, b)
    assert capi.better(n, 4, a, 4, b) and not capi.better(n, 4, b, 4, a)     # zero first at P[0,2]
    c = np.zeros(12, dtype=np.int32)
    c[6 + capi.pair_index(n, 0, 0)] = 1                                      # L[0,0] comes after P[0,*]
    assert capi.better(n, 4, c, 4, b)


def test_synth_generator_is_seeded_and_parses():
    s1, s2 = synth.generate(12, 4, 5), synth.generate(12, 4, 5)
    assert s1.lh_text() == s2.lh_text()
    p = lh.problem_from_lh(s1.lh_text())
    assert p.n == 12 and sum(p.seg_cn2) // 2 == len(s1.string)


def test_imperfect_fbi_matches_oracle_on_golden_cases():
    import json
    cases = json.load(open(os.path.join(ROOT, "tests", "golden", "cases.json")))["cases"]
    seen = 0
    for c in cases:
        g = lh.parse_lh(c["lh"])
        part = lh.flatten(g, lh.read_components(g, c["juncs"]) if c["juncs"] else [], c["junc_info"])[0]
        x = np.zeros(c["n"] * (c["n"] + 1), dtype=np.int32)
        for i, v in c["oracle_x_nz"]:
            x[i] = v
        d = orc.construct_dag(c["n"], c["start_id"], x)
        ok, _, _, bkp, _ = orc.first_valid_order(d, False)
        if not ok:
            continue
        a, b = capi.imperfect_fbi(bkp, part.inversions), orc.imperfect_fbi(bkp, part.inversions)
        assert a == b
        seen += a != list(bkp)
        if d.k <= 16 and c["ref_inject_paths"]:
            assert capi.format_path(capi.expand_path(a)) == c["ref_inject_paths"][0]
    assert seen >= 5        # the imperfect_* fixtures really exercise the rewrite


def test_target_cn_is_the_coverage_of_the_assignment():
    """targetCN (localhap.cpp:222-232): cov_i = sum of P + 2L over the components covering segment i,
    the same vector the model's coverage rows use (LGM:4427-4440)."""
    for n, folds, seed in [(6, 3, 0), (12, 4, 3), (16, 5, 2)]:
        p = lh.problem_from_lh(synth.generate(n, folds, seed, 0.5).lh_text())
        x, _ = orc.solve(p, want_unique=False)
        H = n * (n + 1) // 2
        want, i = [0] * n, 0
        for a in range(n):
            for b in range(a, n):
                for s in range(a, b + 1):
                    want[s] += int(x[i]) + 2 * int(x[H + i])
                i += 1
        assert capi.target_cn(n, x) == want


def test_units_are_marshalled_as_unit_shift():
    """copy numbers in units of 1/(2*unit) copies: the binding passes unit = 2^k as ambi_bfb_problem.unit_shift = k and
    refuses anything that is not a power of two up to 64"""
    pk = capi._Packed([capi.Problem(n=3, seg_cn2=[8, 9, 10], fbi_cn2=[0, 4, 0], unit=4), capi.Problem(n=3, seg_cn2=[8, 9, 10], fbi_cn2=[0, 4, 0])])
    assert pk.arr[0].unit_shift == 2 and pk.arr[1].unit_shift == 0
    for bad in (3, 128, 0):
        with pytest.raises(capi.AmbiError) as e:
            capi._Packed([capi.Problem(n=3, seg_cn2=[8, 9, 10], fbi_cn2=[0, 4, 0], unit=bad)])
        assert e.value.status == -6
