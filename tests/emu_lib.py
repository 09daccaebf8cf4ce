"""ctypes wrapper of tests/emu/small_emu.cpp (TEST INFRASTRUCTURE): the body of the warp-per-partition CUDA kernel
(ambigram_b200/csrc/small_core.cuh) compiled with g++ against a 32-lane warp emulator, for CPU-side checks of
the kernel logic against the oracle.  Imported only by tests/."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "emu", "small_emu.cpp")
LIB = os.path.join(ROOT, "tests", "emu", "_build", "libsmall_emu.so")
DEPS = [SRC, os.path.join(ROOT, "tests", "emu", "simt_emu.h")] + [os.path.join(ROOT, "ambigram_b200", "csrc", f)
                                                                 for f in ("small_core.cuh", "host_flatten.h", "ambi_common.cuh")]
I32P = C.POINTER(C.c_int32)
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB) or any(os.path.getmtime(LIB) < os.path.getmtime(f) for f in DEPS):
            os.makedirs(os.path.dirname(LIB), exist_ok=True)
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-w", "-shared", "-fPIC", "-o", LIB, SRC])
        _lib = C.CDLL(LIB)
    return _lib


def solve(problem, capacity=128, beam_width=16):
    """-> (rc, x, score2, incumbent, candidates, nodes, funnel) ; rc 0 solved, 1 frontier did not fit"""
    n = problem.n
    c2 = np.ascontiguousarray(problem.seg_cn2, dtype=np.int32); j2 = np.ascontiguousarray(problem.fbi_cn2, dtype=np.int32)
    t = [int(v) - problem.start_id for v in problem.tgs_pairs]
    tg = np.ascontiguousarray(t if t else [0, 0], dtype=np.int32)
    x = np.zeros(n * (n + 1), dtype=np.int32)
    sc, inc = C.c_int64(0), C.c_int32(0)
    cand, nodes = C.c_uint64(0), C.c_uint64(0)
    fun = (C.c_uint64 * 4)()
    rc = lib().emu_small_solve(n, c2.ctypes.data_as(I32P), j2.ctypes.data_as(I32P), problem.effective_loop_ub(), len(t) // 2,
                               tg.ctypes.data_as(I32P), int(problem.junc_info), capacity, beam_width, x.ctypes.data_as(I32P),
                               C.byref(sc), C.byref(inc), C.byref(cand), C.byref(nodes), fun)
    return rc, x, int(sc.value), int(inc.value), int(cand.value), int(nodes.value), [int(v) for v in fun]
