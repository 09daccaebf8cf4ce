"""Host-side mirror of the reference's `.lh` reader and of the graph -> Stage-A problem flattening.

Follows, without copying, Graph::readGraph (src/Graph.cpp:109-237), calculateCopyNum
(src/Graph.cpp:369-405, only the "given CN > 0 is kept" part), readComponents
(src/LocalGenomicMap.cpp:5096-5156), getJuncCN (LGM:3989-4050), getIndelBias (LGM:3699-3744) and the
bias / component selection of localhap.cpp:141-162.  Pure bookkeeping: no search arithmetic.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

from .capi import Problem


@dataclass
class Junction:
    src: int
    src_dir: str
    tgt: int
    tgt_dir: str
    coverage: float
    cn: float


@dataclass
class LhGraph:
    seg_ids: List[int] = field(default_factory=list)
    seg_chrom: Dict[int, str] = field(default_factory=dict)
    seg_cn: Dict[int, float] = field(default_factory=dict)
    seg_cov: Dict[int, float] = field(default_factory=dict)
    seg_chr_id: Dict[int, int] = field(default_factory=dict)
    juncs: List[Junction] = field(default_factory=list)
    sources: List[int] = field(default_factory=list)
    sinks: List[int] = field(default_factory=list)
    avg_cov_raw: float = -1.0
    purity: float = -1.0
    avg_tumor_ploidy: float = -1.0
    avg_ploidy: float = 0.0
    virus_start: int = 1 << 30
    props: List[str] = field(default_factory=list)


def parse_lh(text: str) -> LhGraph:
    g = LhGraph()
    for raw in text.splitlines():
        line = raw.strip()
        if not line or line.startswith("#"):
            continue
        tok = line.split()
        key = tok[0]
        if key == "AVG_WHOLE_HOST_DP":
            g.avg_cov_raw = float(tok[1])
        elif key == "PURITY":
            g.purity = float(tok[1])
        elif key == "AVG_TUMOR_PLOIDY":
            g.avg_tumor_ploidy = float(tok[1])
        elif key == "AVG_PLOIDY":
            g.avg_ploidy = float(tok[1])
        elif key == "VIRUS_START":
            g.virus_start = int(tok[1])
        elif key == "SOURCE":
            g.sources = [int(v) for v in tok[1].split(",")]
        elif key == "SINK":
            g.sinks = [int(v) for v in tok[1].split(",")]
        elif key == "SEG":
            f = tok[1].split(":")
            sid = int(f[1])
            g.seg_ids.append(sid)
            g.seg_chrom[sid] = f[2]
            g.seg_cov[sid] = max(float(tok[2]), 0.0)
            g.seg_cn[sid] = float(tok[3])
            for i, (s, e) in enumerate(zip(g.sources, g.sinks)):
                if s <= sid <= e:
                    g.seg_chr_id[sid] = i
        elif key == "JUNC":
            cov, cn = float(tok[3]), float(tok[4])
            if cov <= 0 and cn <= 0:
                continue
            a, b = tok[1].split(":"), tok[2].split(":")
            g.juncs.append(Junction(int(a[1]), a[2][0], int(b[1]), b[2][0], cov, cn))
        elif key == "PROP":
            g.props = tok[1:]
    if len(g.sources) != len(g.sinks):
        raise ValueError("SOURCE/SINK count mismatch")
    # calculateHapDepth / calculateCopyNum: only entries without a positive CN are derived from depth
    if any(v <= 0 for v in g.seg_cn.values()) or any(j.cn <= 0 for j in g.juncs):
        # Graph(const char*) starts with mAvgPloidy = 0 (Graph.cpp:38), so calculateHapDepth always takes its
        # "AVG_PLOIDY given" branch (Graph.cpp:332-355): with purity and tumour ploidy present the normal-cell
        # share of the depth, ratio = 1 - pT / (pT + 2(1-p)), is subtracted before dividing by the haploid
        # depth, and a given AVG_PLOIDY is kept only if it is within 0.1 of pT + 2(1-p)
        avg_ploidy, ratio = g.avg_ploidy, 0.0
        if g.avg_tumor_ploidy >= 0 and g.purity >= 0:
            tumour = g.purity * g.avg_tumor_ploidy
            computed = tumour + (1 - g.purity) * 2
            ratio = 1 - tumour / computed
            if abs(avg_ploidy - computed) > 0.1:
                avg_ploidy = computed
        hdp = g.avg_cov_raw * g.purity / avg_ploidy
        for sid in g.seg_ids:
            if g.seg_cn[sid] > 0:
                continue
            if sid >= g.virus_start:
                g.seg_cn[sid] = max(g.seg_cov[sid] / g.avg_cov_raw * 2, 0.0)
            else:
                g.seg_cn[sid] = max((g.seg_cov[sid] - g.avg_cov_raw * ratio) / hdp, 0.0)
        for j in g.juncs:
            if j.cn <= 0:
                j.cn = max((j.coverage - g.avg_cov_raw * ratio) / hdp, 0.0)
    return g


def read_components(g: LhGraph, juncs_text: str) -> List[List[int]]:
    """readComponents (LGM:5096-5156): sub-paths between strand switches become sorted segment lists;
    every switch adds a CN-1 junction or lifts an existing one to CN >= 2."""
    res: List[List[int]] = []
    part = g.seg_chr_id
    for line in juncs_text.splitlines():
        toks = line.split()
        if not toks:
            continue
        segs = [int(t[:-1]) for t in toks]
        sign = [t[-1] for t in toks]
        last = 0
        for i in range(1, len(segs)):
            if part.get(segs[last]) != part.get(segs[i]) or sign[i - 1] != sign[i]:
                if i - last >= 2:
                    res.append(sorted(segs[last:i]))
                s, sd, t, td = segs[i - 1], sign[i - 1], segs[i], sign[i]
                found = None
                comp = {"+": "-", "-": "+"}
                for j in g.juncs:    # Graph::findJunction: same junction or its reverse complement
                    if (j.src, j.src_dir, j.tgt, j.tgt_dir) == (s, sd, t, td) or \
                       (j.src, j.src_dir, j.tgt, j.tgt_dir) == (t, comp[td], s, comp[sd]):
                        found = j
                        break
                if found is None:
                    g.juncs.append(Junction(s, sd, t, td, 30.0, 1.0))
                elif found.cn < 2:
                    found.cn = 2.0
                last = i
        if len(segs) - last >= 2:
            res.append(sorted(segs[last:]))
    uniq = sorted(set(tuple(c) for c in res))
    return [list(c) for c in uniq]


@dataclass
class Partition:
    start_id: int
    end_id: int
    problem: Optional[Problem]
    inversions: Dict[int, Tuple[int, int]]      # segment id -> (source id, target id) of its fold-back junction
    fbi_cn: List[float]
    trivial: bool                                # localhap.cpp:164-170 shortcut (no FBI, no components)


def get_indel_bias(g: LhGraph, start_id: int, end_id: int) -> None:
    """getIndelBias (LGM:3699-3744): same-strand, non-adjacent junctions inside the partition shift the
    CN of the segments they delete / duplicate / insert.  Mutates g.seg_cn like the reference."""
    sv = []
    for j in g.juncs:
        if g.seg_chr_id.get(j.src) != g.seg_chr_id.get(j.tgt):
            continue
        if not (start_id <= j.src <= end_id and start_id <= j.tgt <= end_id):
            continue
        if j.src_dir != j.tgt_dir:
            continue
        if (j.src_dir == "+" and j.tgt - j.src == 1) or (j.src_dir == "-" and j.src - j.tgt == 1):
            continue
        sv.append(j)
    while sv:
        group: List[int] = []
        i = 0
        while i < len(sv):
            s = sv[i].src if sv[i].src_dir == "+" else -sv[i].src
            t = sv[i].tgt if sv[i].tgt_dir == "+" else -sv[i].tgt
            if not group:
                group = [s, t]
            elif t == group[0]:
                group.insert(0, s)
            elif s == -group[0]:
                group.insert(0, -t)
            elif group[-1] == s:
                group.append(t)
            elif group[-1] == -t:
                group.append(-s)
            else:
                i += 1
                continue
            sv.pop(i)
        if len(group) == 2:
            if group[0] < group[1]:
                for k in range(group[0] + 1, group[1]):
                    g.seg_cn[abs(k)] += 1
            else:
                for k in range(group[1], group[0] + 1):
                    g.seg_cn[abs(k)] -= 1
        else:
            for k in group[1:-1]:
                g.seg_cn[abs(k)] -= 1


def flatten(g: LhGraph, components: Optional[List[List[int]]] = None, junc_info: bool = False, unit: int = 1) -> List[Partition]:
    """Per SOURCE/SINK pair: getJuncCN (LGM:3989-4050), bias (localhap.cpp:141-146), getIndelBias,
    component selection (localhap.cpp:158-162) -> the integer arrays of ambi_bfb_problem.
    `unit`: the arrays are 2*unit*copies (copy numbers that are multiples of 1/(2*unit)); the library takes powers of
    two up to 64 (ambi_bfb_problem.unit_shift), the oracle's *_units entry points any unit."""
    components = components or []
    parts: List[Partition] = []
    for s_id, e_id in zip(g.sources, g.sinks):
        n = e_id - s_id + 1
        fbi = {i: 0.0 for i in range(s_id, e_id + 1)}
        inversions: Dict[int, Tuple[int, int]] = {}
        inv_list = []
        for j in g.juncs:
            if not (s_id <= j.src <= e_id and s_id <= j.tgt <= e_id):
                continue
            cn = j.cn
            if 0.5 < cn < 1:
                cn = 1.0
            if j.src_dir != j.tgt_dir and abs(j.src - j.tgt) <= 2:
                inv_list.append(j)
                if j.src not in inversions:
                    inversions[j.src] = (j.src, j.tgt)
                    fbi[j.src] += cn
                elif j.tgt not in inversions:
                    inversions[j.tgt] = (j.src, j.tgt)
                    fbi[j.tgt] += cn
        for j in inv_list:
            inversions.setdefault(j.src, (j.src, j.tgt))
            inversions.setdefault(j.tgt, (j.src, j.tgt))
        bias = 1
        for i in range(s_id, e_id + 1):
            if fbi[i] > 0 and inversions[i][0] != inversions[i][1]:
                bias += int(fbi[i]) % 2
        get_indel_bias(g, s_id, e_id)
        valid = [c for c in components if g.seg_chr_id.get(c[0]) == g.seg_chr_id.get(s_id)]
        fbi_list = [fbi[i] for i in range(s_id, e_id + 1)]
        trivial = abs(sum(fbi_list)) < 1e-6 and not valid
        problem = None
        if not trivial:
            c2 = [2 * unit * g.seg_cn[i] for i in range(s_id, e_id + 1)]
            j2 = [2 * unit * v for v in fbi_list]
            if any(abs(v - round(v)) > 1e-9 for v in c2 + j2):
                raise ValueError("copy numbers are not half-integral (AMBI_ERR_NOT_HALF_INTEGRAL)" if unit == 1 else
                                 f"copy numbers are not multiples of 1/{2 * unit}")
            tgs: List[int] = []
            for c in valid:
                tgs += [min(c[0], c[-1]), max(c[0], c[-1])]
            loop_ub = int(sum(g.seg_cn[i] for i in g.seg_ids))
            problem = Problem(n=n, seg_cn2=[int(round(v)) for v in c2], fbi_cn2=[int(round(v)) for v in j2],
                              start_id=s_id, loop_ub=loop_ub, bias=bias, tgs_pairs=tgs, junc_info=junc_info, unit=unit)
        parts.append(Partition(s_id, e_id, problem, inversions, fbi_list, trivial))
    return parts


def problem_from_lh(text: str, juncs_text: str = "", junc_info: bool = False) -> Problem:
    g = parse_lh(text)
    comps = read_components(g, juncs_text) if juncs_text else []
    parts = flatten(g, comps, junc_info)
    if len(parts) != 1 or parts[0].problem is None:
        raise ValueError("expected exactly one non-trivial partition")
    return parts[0].problem
