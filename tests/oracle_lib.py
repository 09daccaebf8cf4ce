"""ctypes wrapper of the CPU oracle (oracle/ambi_oracle.c).  TEST INFRASTRUCTURE: imported only by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence, Tuple

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "_build", "libambi_oracle.so")
MAXK = 64
I32P = C.POINTER(C.c_int32)


class Stats(C.Structure):
    _fields_ = [("score2", C.c_int64), ("candidates", C.c_int64), ("nodes", C.c_int64), ("unique", C.c_int32),
                ("status", C.c_int32)]


class Dag(C.Structure):
    _fields_ = [("k", C.c_int32), ("kind", C.c_int8 * MAXK), ("a", C.c_int16 * MAXK), ("b", C.c_int16 * MAXK),
                ("cn", C.c_int16 * MAXK), ("adj", C.c_uint64 * MAXK)]

    def nodes(self):
        return [("l" if self.kind[i] else "p", self.a[i], self.b[i], self.cn[i]) for i in range(self.k)]

    def edges(self):
        return [(i, j) for i in range(self.k) for j in range(self.k) if (self.adj[i] >> j) & 1]


def build(force: bool = False) -> str:
    src = os.path.join(ORACLE_DIR, "ambi_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "_build/libambi_oracle.so"])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_bruteforce.restype = C.c_int64
        _lib.orc_count_valid_orders.restype = C.c_int64
    return _lib


def _arr(v) -> np.ndarray:
    return np.ascontiguousarray(v, dtype=np.int32)


def _tgs_rel(problem) -> Tuple[np.ndarray, int]:
    """absolute (min,max) ids -> 0-based offsets as orc_* expects"""
    t = [int(v) - problem.start_id for v in problem.tgs_pairs]
    return _arr(t if t else [0, 0]), len(t) // 2


def solve(problem, want_unique: bool = True):
    """-> (x[n(n+1)], Stats) for an ambigram_b200.capi.Problem"""
    L = lib()
    c2, j2 = _arr(problem.seg_cn2), _arr(problem.fbi_cn2)
    t, nt = _tgs_rel(problem)
    x = np.zeros(problem.n * (problem.n + 1), dtype=np.int32)
    st = Stats()
    unit = int(getattr(problem, "unit", 1))
    if unit != 1:       # copy numbers in units of 1/(2*unit): Stats.score2 = 2*unit*sum(eps)
        rc = L.orc_solve_units(problem.n, c2.ctypes.data_as(I32P), j2.ctypes.data_as(I32P), unit, problem.effective_loop_ub(), nt,
                               t.ctypes.data_as(I32P), int(problem.junc_info), int(want_unique), 1, 0, -1,
                               x.ctypes.data_as(I32P), C.byref(st))
    else:
        rc = L.orc_solve(problem.n, c2.ctypes.data_as(I32P), j2.ctypes.data_as(I32P), problem.effective_loop_ub(), nt,
                         t.ctypes.data_as(I32P), int(problem.junc_info), int(want_unique), x.ctypes.data_as(I32P), C.byref(st))
    if rc != 0:
        raise RuntimeError(f"orc_solve failed: {rc}")
    return x, st


def solve_shard(problem, world: int, rank: int, shard_level: int = 2):
    """shard-local optimum: -> (x, Stats); Stats.score2 == 2**31-1 when the shard owns no assignment"""
    L = lib()
    c2, j2 = _arr(problem.seg_cn2), _arr(problem.fbi_cn2)
    t, nt = _tgs_rel(problem)
    x = np.zeros(problem.n * (problem.n + 1), dtype=np.int32)
    st = Stats()
    rc = L.orc_solve_shard(problem.n, c2.ctypes.data_as(I32P), j2.ctypes.data_as(I32P), problem.effective_loop_ub(), nt,
                           t.ctypes.data_as(I32P), int(problem.junc_info), 0, world, rank, shard_level,
                           x.ctypes.data_as(I32P), C.byref(st))
    if rc != 0:
        raise RuntimeError(f"orc_solve_shard failed: {rc}")
    return x, st


def model_eval(problem, x) -> Tuple[bool, int]:
    L = lib()
    c2, j2 = _arr(problem.seg_cn2), _arr(problem.fbi_cn2)
    t, nt = _tgs_rel(problem)
    xx = _arr(x)
    s = C.c_int64(0)
    ok = L.orc_model_eval_units(problem.n, c2.ctypes.data_as(I32P), j2.ctypes.data_as(I32P), int(getattr(problem, "unit", 1)),
                                problem.effective_loop_ub(), nt, t.ctypes.data_as(I32P), int(problem.junc_info),
                                xx.ctypes.data_as(I32P), C.byref(s))
    return bool(ok), int(s.value)


def bruteforce(problem, lmax: int = 3):
    L = lib()
    c2, j2 = _arr(problem.seg_cn2), _arr(problem.fbi_cn2)
    t, nt = _tgs_rel(problem)
    x = np.zeros(problem.n * (problem.n + 1), dtype=np.int32)
    nopt = C.c_int64(0)
    best = L.orc_bruteforce(problem.n, c2.ctypes.data_as(I32P), j2.ctypes.data_as(I32P), problem.effective_loop_ub(), nt,
                            t.ctypes.data_as(I32P), int(problem.junc_info), lmax, x.ctypes.data_as(I32P), C.byref(nopt))
    return int(best), x, int(nopt.value)


def construct_dag(n: int, start_id: int, x) -> Dag:
    d = Dag()
    xx = _arr(x)
    rc = lib().orc_construct_dag(n, start_id, xx.ctypes.data_as(I32P), C.byref(d))
    if rc != 0:
        raise RuntimeError("orc_construct_dag failed")
    return d


def first_valid_order(dag: Dag, reversed_: bool = False, naive_cap: Optional[int] = None):
    """-> (found, used_flipped, order, bkp, n_tested)"""
    order = (C.c_int32 * MAXK)()
    bkp = (C.c_int32 * 4096)()
    flipped, blen, tested = C.c_int32(0), C.c_int32(0), C.c_int64(0)
    if naive_cap is None:
        ok = lib().orc_first_valid_order(C.byref(dag), int(reversed_), order, C.byref(flipped), bkp, C.byref(blen), C.byref(tested))
    else:
        ok = lib().orc_first_valid_order_naive(C.byref(dag), int(reversed_), C.c_int64(naive_cap), order, C.byref(flipped), bkp,
                                               C.byref(blen), C.byref(tested))
    return bool(ok), bool(flipped.value), list(order[:dag.k]), list(bkp[:blen.value]), int(tested.value)


def count_valid_orders(dag: Dag, forward: bool = True, cap: int = 0) -> Tuple[int, int]:
    n_orders = C.c_int64(0)
    v = lib().orc_count_valid_orders(C.byref(dag), int(forward), C.c_int64(cap), C.byref(n_orders))
    return int(v), int(n_orders.value)


def expand_path(bkp: Sequence[int]) -> List[int]:
    b = _arr(bkp)
    out = np.zeros(16384, dtype=np.int32)
    ln = lib().orc_expand_path(b.ctypes.data_as(I32P), len(b), out.ctypes.data_as(I32P), len(out))
    return [int(v) for v in out[:ln]]


def max_threads() -> int:
    return int(lib().orc_max_threads())


def solve_batch(problems, threads: int = 0):
    """OpenMP batch solve (no third-generation rows) -> (list of x, scores, candidates)"""
    L = lib()
    cnt = len(problems)
    n = _arr([p.n for p in problems])
    off = np.zeros(cnt, dtype=np.int32)
    xoff = np.zeros(cnt, dtype=np.int64)
    o = xo = 0
    for i, p in enumerate(problems):
        off[i], xoff[i] = o, xo
        o += p.n
        xo += p.n * (p.n + 1)
    c2 = _arr(np.concatenate([np.asarray(p.seg_cn2) for p in problems]))
    j2 = _arr(np.concatenate([np.asarray(p.fbi_cn2) for p in problems]))
    ub = _arr([p.effective_loop_ub() for p in problems])
    x = np.zeros(xo, dtype=np.int32)
    score = np.zeros(cnt, dtype=np.int64)
    cand = np.zeros(cnt, dtype=np.int64)
    I64P = C.POINTER(C.c_int64)
    bad = L.orc_solve_batch(cnt, n.ctypes.data_as(I32P), off.ctypes.data_as(I32P), c2.ctypes.data_as(I32P),
                            j2.ctypes.data_as(I32P), ub.ctypes.data_as(I32P), xoff.ctypes.data_as(I64P),
                            x.ctypes.data_as(I32P), score.ctypes.data_as(I64P), cand.ctypes.data_as(I64P), int(threads))
    if bad:
        raise RuntimeError(f"orc_solve_batch: {bad} problems failed")
    xs = [x[xoff[i]:xoff[i] + problems[i].n * (problems[i].n + 1)] for i in range(cnt)]
    return xs, score, cand


def imperfect_fbi(bkp: Sequence[int], inversions: dict) -> List[int]:
    """imperfectFBI (LGM:3431-3512); inversions: segment id -> (source id, target id)"""
    b = _arr(bkp).copy()
    mx = max([1] + [abs(int(v)) for v in bkp] + list(inversions.keys()))
    src = np.zeros(mx + 2, dtype=np.int32)
    tgt = np.zeros(mx + 2, dtype=np.int32)
    for k, (s_, t_) in inversions.items():
        src[k], tgt[k] = s_, t_
    lib().orc_imperfect_fbi.restype = None
    lib().orc_imperfect_fbi(b.ctypes.data_as(I32P), len(b), src.ctypes.data_as(I32P), tgt.ctypes.data_as(I32P), mx)
    return [int(v) for v in b]


def all_valid_orders(dag: Dag, forward: bool = True, cap: int = 4096) -> List[List[int]]:
    lib().orc_all_valid_orders.restype = C.c_int64
    buf = np.zeros(cap * max(dag.k, 1), dtype=np.int32)
    n = lib().orc_all_valid_orders(C.byref(dag), int(forward), buf.ctypes.data_as(I32P), C.c_int64(cap))
    n = min(int(n), cap)
    return [[int(v) for v in buf[i * dag.k:(i + 1) * dag.k]] for i in range(n)]


def assemble_order(dag: Dag, order: Sequence[int], forward: bool = True):
    o = _arr(order)
    bkp = (C.c_int32 * 4096)()
    blen = C.c_int32(0)
    placed = lib().orc_assemble_order(C.byref(dag), o.ctypes.data_as(I32P), int(forward), bkp, C.byref(blen))
    return int(placed), list(bkp[:blen.value])
