"""ambigram_b200 -- B200-native drop-in for Ambigram's BFB fold-pattern search.

Only what the hot path needs lives here:
  csrc/      hand-written sm_100a CUDA kernels + the C ABI (libambigram_b200.so)
  capi.py    ctypes binding of include/ambigram_b200.h (what a host program calls)
  lh.py      host-side mirror of the reference's .lh reader and the flattening of a
             Graph into ambi_bfb_problem records (Graph.cpp:109-237, LGM:3989-4050)
  synth.py   seeded synthetic BFB .lh generator
  bfb.py     the `--op bfb` driver mirroring localhap.cpp:49-389 on top of the C ABI

The product path fails loudly when the CUDA library is missing or no GPU is
present; there is no CPU fallback (oracle/ is test infrastructure only).
"""
from .capi import (AmbiError, Context, Problem, Solution, load_library, library_path)  # noqa: F401

__all__ = ["AmbiError", "Context", "Problem", "Solution", "load_library", "library_path"]
