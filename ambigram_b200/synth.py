"""Seeded synthetic BFB profiles and their `.lh` text.

The reference ships no BFB simulator (script/simu.py inserts viral sequence).  This
generator follows the on-disk layout of script/bfb_scripts.py generate_lh
(bfb_scripts.py:576-601): header, SOURCE/SINK, SEG and JUNC lines.
"""
from __future__ import annotations

import math
import random
from dataclasses import dataclass
from typing import List, Optional, Tuple


@dataclass
class SynthSample:
    n: int
    folds: int
    seed: int
    string: List[Tuple[int, int]]            # the simulated BFB haplotype (segment id, +1/-1)
    seg_cn: List[int]                        # per-segment copy number (after noise)
    juncs: List[Tuple[int, int, int, int, int]]   # (src id, src dir, tgt id, tgt dir, cn) in file order
    juncs_file: Optional[List[List[Tuple[int, int]]]] = None   # optional .juncs sub-paths

    def lh_text(self, sample: str = "synth") -> str:
        lines = [f"SAMPLE {sample}", "AVG_CHR_SEG_DP 30", "AVG_WHOLE_HOST_DP 30", "AVG_JUNC_DP 30", "PURITY 1",
                 "AVG_TUMOR_PLOIDY 2", "PLOIDY 2m1", f"VIRUS_START {self.n + 1}", "SOURCE 1", f"SINK {self.n}"]
        for i, cn in enumerate(self.seg_cn, start=1):
            s = 55281001 + (i - 1) * 1000
            lines.append(f"SEG H:{i}:chr7:{s}:{s + 999} {15.0 * cn} {float(cn)}")
        for (si, sd, ti, td, cn) in self.juncs:
            lines.append(f"JUNC H:{si}:{'+' if sd > 0 else '-'} H:{ti}:{'+' if td > 0 else '-'} {15.0 * cn} {float(cn)} U B")
        return "\n".join(lines) + "\n"

    def juncs_text(self) -> str:
        if not self.juncs_file:
            return ""
        return "\n".join(" ".join(f"{i}{'+' if d > 0 else '-'}" for (i, d) in comp) for comp in self.juncs_file) + "\n"


def generate(n: int, folds: int, seed: int, noise: float = 0.0, with_juncs: bool = False,
             imperfect: float = 0.0) -> SynthSample:
    """S = [1+..n+]; each fold appends a prefix of the reverse complement (one breakage-fusion-bridge
    cycle).  Segment CN = occurrences; every adjacent opposite-strand repeat (i,d)->(i,-d) is one copy of a
    fold-back junction.  `noise`: each segment CN +/-1 with that probability (min 1)."""
    rng = random.Random(seed)
    s: List[Tuple[int, int]] = [(i, 1) for i in range(1, n + 1)]
    for _ in range(folds):
        rc = [(i, -d) for (i, d) in reversed(s)]
        k = rng.randint(1, len(s))
        s = s + rc[:k]
    cn = [0] * (n + 1)
    for (i, _d) in s:
        cn[i] += 1
    fb = {}
    for u, v in zip(s, s[1:]):
        if u[0] == v[0] and u[1] != v[1]:
            fb[(u[0], u[1])] = fb.get((u[0], u[1]), 0) + 1
    seg_cn = cn[1:]
    if noise > 0:
        seg_cn = [max(1, x + (rng.choice([-1, 1]) if rng.random() < noise else 0)) for x in seg_cn]
    juncs = []
    for (i, d), v in sorted(fb.items()):
        # `imperfect`: with that probability the fold-back junction is reported between two adjacent
        # segments (what getJuncCN accepts as a fold-back "with error of 2 segments", LGM:4012)
        if imperfect > 0 and rng.random() < imperfect:
            if d > 0 and i < n:
                juncs.append((i, 1, i + 1, -1, v)); continue
            if d < 0 and i > 1:
                juncs.append((i, -1, i - 1, 1, v)); continue
        juncs.append((i, d, i, -d, v))
    jf = None
    if with_juncs:
        jf = []
        for _ in range(rng.randint(2, 4)):
            ln = rng.randint(3, min(len(s), 8))
            st = rng.randint(0, len(s) - ln)
            jf.append(s[st:st + ln])
    return SynthSample(n, folds, seed, s, seg_cn, juncs, jf)


def folds_for(n: int) -> int:
    return int(math.ceil(n / 3.5))
