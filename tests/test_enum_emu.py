"""CPU: the configuration enumeration of k_expand (ambigram_b200/csrc/decode_common.cuh: enumeration lists, chunked
cursor, column-mask restriction, light test) compiled with g++ (tests/emu/enum_emu.cpp, test infrastructure) and
checked on random records against the flat decode of every configuration number:
  * a cursor started at any number and stepped yields the same configurations and the same light-test verdicts;
  * the configurations that pass the light test under the restricted masks are exactly those that pass under the full
    masks -- restrict_masks() only removes what the light test refuses one by one.
The same code runs on the GPU under tests/test_gpu_parity.py; this test pins its logic where there is none."""
import ctypes as C
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "emu", "enum_emu.cpp")
LIB = os.path.join(ROOT, "tests", "emu", "_build", "libenum_emu.so")
DEPS = [SRC, os.path.join(ROOT, "tests", "emu", "simt_emu.h"), os.path.join(ROOT, "ambigram_b200", "csrc", "decode_common.cuh")]


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB) or any(os.path.getmtime(LIB) < os.path.getmtime(f) for f in DEPS):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-w", "-shared", "-fPIC", "-o", LIB, SRC])
    return C.CDLL(LIB)


@pytest.mark.parametrize("scaled", [0, 1])
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_cursor_and_mask_restriction_against_flat_decode(lib, seed, scaled):
    stats = (C.c_longlong * 5)()
    rc = lib.emu_enum_check(C.c_uint64(seed * 7919 + scaled), 400, scaled, stats)
    assert rc == 0
    checked, passed, removed, slow, big = (int(v) for v in stats)
    # the sample must exercise what it claims to: many steps, passing configurations, records whose masks shrink,
    # records with more class-0 loop multisets than the list holds
    assert checked > 100_000 and passed > 1_000 and removed > 50 and slow > 5
    assert big >= 1              # ... and records with more configurations than 16 bits hold (sampled)
