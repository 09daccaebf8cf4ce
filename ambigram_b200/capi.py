"""ctypes binding of include/ambigram_b200.h.

Mirrors the reference's seams (localhap.cpp:173-212 and 235-261): a host program
hands over plain integer arrays and receives `elementCN` in the reference's own
variableIdx numbering.  Raises AmbiError when libambigram_b200.so is missing or
no CUDA device is present -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

MAX_SEGMENTS = 32
MAX_COMPONENTS = 64

_HERE = os.path.dirname(os.path.abspath(__file__))


def library_path() -> str:
    return os.path.join(_HERE, "libambigram_b200.so")


class AmbiError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"ambigram_b200 status {status}: {msg}")
        self.status = status


class _Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("rank", C.c_int32), ("world", C.c_int32), ("beam_width", C.c_int32),
                ("node_capacity", C.c_int64), ("shard_level", C.c_int32), ("flags", C.c_uint32)]


class _Problem(C.Structure):
    _fields_ = [("n", C.c_int32), ("start_id", C.c_int32), ("seg_cn2", C.POINTER(C.c_int32)),
                ("fbi_cn2", C.POINTER(C.c_int32)), ("loop_ub", C.c_int32), ("bias", C.c_int32),
                ("n_tgs", C.c_int32), ("tgs_pairs", C.POINTER(C.c_int32)), ("flags", C.c_uint32),
                ("unit_shift", C.c_uint32)]


class _Solution(C.Structure):
    _fields_ = [("element_cn", C.POINTER(C.c_int32)), ("score2", C.c_int64), ("objective2", C.c_int64),
                ("candidates_scored", C.c_uint64), ("nodes_expanded", C.c_uint64), ("status", C.c_int32),
                ("incumbent_score2", C.c_int32), ("seconds", C.c_double), ("seconds_to_best", C.c_double)]


class _Profile(C.Structure):
    _fields_ = [("expand_ms", C.c_double), ("expand_launches", C.c_uint64), ("total_ms", C.c_double),
                ("incumbent_pass_ms", C.c_double), ("record_stride", C.c_uint64), ("configs_enumerated", C.c_uint64),
                ("bounds_evaluated", C.c_uint64), ("children_written", C.c_uint64), ("bound_batches", C.c_uint64), ("exchange_rounds", C.c_uint64),
                ("chunked_runs", C.c_uint64), ("small_ms", C.c_double), ("small_solved", C.c_uint64), ("global_solved", C.c_uint64),
                ("small_capacity", C.c_uint64), ("stage_b_kernel_ms", C.c_double)]


class _Peaks(C.Structure):
    _fields_ = [("alu_tops", C.c_double), ("alu_fma_tops", C.c_double), ("smem_tbs", C.c_double),
                ("sms", C.c_int32), ("reserved", C.c_int32)]


class _Dag(C.Structure):
    _fields_ = [("k", C.c_int32), ("kind", C.c_int8 * MAX_COMPONENTS), ("a", C.c_int16 * MAX_COMPONENTS),
                ("b", C.c_int16 * MAX_COMPONENTS), ("cn", C.c_int16 * MAX_COMPONENTS),
                ("adj", C.c_uint64 * MAX_COMPONENTS)]


class _OrderResult(C.Structure):
    _fields_ = [("found", C.c_int32), ("used_flipped", C.c_int32), ("order", C.c_int32 * MAX_COMPONENTS),
                ("bkp_len", C.c_int32), ("bkp", C.c_int32 * 1024), ("orders_tested", C.c_uint64),
                ("status", C.c_int32), ("reserved", C.c_int32)]


EXPORTS = [
    "ambi_abi_version", "ambi_init", "ambi_destroy", "ambi_last_error", "ambi_status_string",
    "ambi_nccl_unique_id", "ambi_comm_init", "ambi_bfb_solve", "ambi_bfb_solve_batch", "ambi_bfb_solve_shard",
    "ambi_bfb_better", "ambi_bfb_construct_dag", "ambi_bfb_first_valid_order", "ambi_bfb_first_valid_order_batch",
    "ambi_bfb_expand_path", "ambi_bfb_imperfect_fbi", "ambi_bfb_all_valid_orders", "ambi_bfb_all_valid_paths", "ambi_batch_upload", "ambi_batch_run", "ambi_batch_fetch", "ambi_batch_free",
    "ambi_ctx_stream", "ambi_ctx_sync", "ambi_ctx_launch_count", "ambi_ctx_set_profile", "ambi_ctx_get_profile",
    "ambi_measure_peaks", "ambi_bfb_target_cn",
]

_lib = None


def load_library(path: Optional[str] = None):
    """dlopen libambigram_b200.so and declare the prototypes of include/ambigram_b200.h."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or library_path()
    if not os.path.exists(p):
        raise AmbiError(-1, f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(there is no CPU fallback)")
    lib = C.CDLL(p)
    vp = C.c_void_p
    lib.ambi_abi_version.restype = C.c_int
    lib.ambi_init.argtypes = [C.POINTER(vp), C.POINTER(_Config)]
    lib.ambi_destroy.argtypes = [vp]
    lib.ambi_destroy.restype = None
    lib.ambi_last_error.argtypes = [vp]
    lib.ambi_last_error.restype = C.c_char_p
    lib.ambi_status_string.argtypes = [C.c_int]
    lib.ambi_status_string.restype = C.c_char_p
    lib.ambi_nccl_unique_id.argtypes = [C.c_void_p]
    lib.ambi_comm_init.argtypes = [vp, C.c_void_p]
    lib.ambi_bfb_solve.argtypes = [vp, C.POINTER(_Problem), C.POINTER(_Solution)]
    lib.ambi_bfb_solve_shard.argtypes = [vp, C.POINTER(_Problem), C.POINTER(_Solution)]
    lib.ambi_bfb_solve_batch.argtypes = [vp, C.c_int32, C.POINTER(_Problem), C.POINTER(_Solution)]
    lib.ambi_bfb_better.argtypes = [C.c_int32, C.c_int64, C.POINTER(C.c_int32), C.c_int64, C.POINTER(C.c_int32)]
    lib.ambi_bfb_construct_dag.argtypes = [C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(_Dag)]
    lib.ambi_bfb_first_valid_order.argtypes = [vp, C.POINTER(_Dag), C.c_int32, C.POINTER(_OrderResult)]
    lib.ambi_bfb_first_valid_order_batch.argtypes = [vp, C.c_int32, C.POINTER(_Dag), C.POINTER(C.c_int32),
                                                     C.POINTER(_OrderResult)]
    lib.ambi_bfb_all_valid_orders.argtypes = [vp, C.POINTER(_Dag), C.c_int32, C.c_int32, C.POINTER(C.c_int32),
                                              C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                              C.POINTER(C.c_int32), C.POINTER(C.c_uint64)]
    lib.ambi_bfb_all_valid_paths.argtypes = [vp, C.POINTER(_Dag), C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int32, C.c_int32,
                                             C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_int32),
                                             C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_uint64)]
    lib.ambi_bfb_imperfect_fbi.argtypes = [C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                           C.c_int32]
    lib.ambi_bfb_expand_path.argtypes = [C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_int32), C.c_int32]
    lib.ambi_batch_upload.argtypes = [vp, C.c_int32, C.POINTER(_Problem), C.POINTER(vp)]
    lib.ambi_batch_run.argtypes = [vp, vp]
    lib.ambi_batch_fetch.argtypes = [vp, vp, C.POINTER(_Solution)]
    lib.ambi_batch_free.argtypes = [vp, vp]
    lib.ambi_batch_free.restype = None
    lib.ambi_ctx_stream.argtypes = [vp]
    lib.ambi_ctx_stream.restype = vp
    lib.ambi_ctx_sync.argtypes = [vp]
    lib.ambi_ctx_set_profile.argtypes = [vp, C.c_int]
    lib.ambi_ctx_get_profile.argtypes = [vp, C.POINTER(_Profile)]
    lib.ambi_measure_peaks.argtypes = [vp, C.POINTER(_Peaks)]
    lib.ambi_bfb_target_cn.argtypes = [C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    lib.ambi_ctx_launch_count.argtypes = [vp]
    lib.ambi_ctx_launch_count.restype = C.c_uint64
    if path is None:
        _lib = lib
    return lib


def num_vars(n: int) -> int:
    return n * (n + 1)


def pair_index(n: int, a: int, b: int) -> int:
    """variableIdx of pattern (a,b), 0-based offsets (combinations() order, LGM:3254-3264)."""
    return a * n - a * (a - 1) // 2 + (b - a)


@dataclass
class Problem:
    """One partition's Stage-A inputs (what BFB_ILP reads, LGM:4397-4752)."""
    n: int
    seg_cn2: Sequence[int]
    fbi_cn2: Sequence[int]
    start_id: int = 1
    loop_ub: Optional[int] = None
    bias: int = 1
    tgs_pairs: Sequence[int] = ()
    junc_info: bool = False
    unit: int = 1          # seg_cn2 / fbi_cn2 are 2*unit*copies (unit a power of two <= 64 -> ambi_bfb_problem.unit_shift)

    def effective_loop_ub(self) -> int:
        return int(self.loop_ub) if self.loop_ub is not None else int(sum(self.seg_cn2) // (2 * self.unit))


@dataclass
class Solution:
    element_cn: np.ndarray
    score2: int
    objective2: int
    candidates_scored: int
    nodes_expanded: int
    status: int
    incumbent_score2: int
    seconds: float
    seconds_to_best: float

    def components(self, n: int, start_id: int = 1) -> List[tuple]:
        """[(kind, a, b, cn)] with absolute ids, kind 'p' or 'l'."""
        h = n * (n + 1) // 2
        out = []
        for a in range(n):
            for b in range(a, n):
                i = pair_index(n, a, b)
                if self.element_cn[i] > 0:
                    out.append(("p", a + start_id, b + start_id, int(self.element_cn[i])))
                if self.element_cn[h + i] > 0:
                    out.append(("l", a + start_id, b + start_id, int(self.element_cn[h + i])))
        return out


class _Packed:
    """Keeps the numpy buffers behind an array of _Problem alive."""

    def __init__(self, problems: Sequence[Problem]):
        self.count = len(problems)
        self.arr = (_Problem * self.count)()
        self.keep = []
        self.sol = (_Solution * self.count)()
        self.x = []
        for i, p in enumerate(problems):
            unit = int(getattr(p, "unit", 1))
            if unit < 1 or unit > 64 or unit & (unit - 1):
                raise AmbiError(-6, f"problem {i}: unit must be a power of two <= 64 (copy numbers in units of 1/(2*unit) copies)")
            c2 = np.ascontiguousarray(p.seg_cn2, dtype=np.int32)
            j2 = np.ascontiguousarray(p.fbi_cn2, dtype=np.int32)
            tg = np.ascontiguousarray(p.tgs_pairs if len(p.tgs_pairs) else [0, 0], dtype=np.int32)
            if len(c2) != p.n or len(j2) != p.n:
                raise AmbiError(-2, f"problem {i}: arrays must have n={p.n} entries")
            self.keep += [c2, j2, tg]
            q = self.arr[i]
            q.n = p.n
            q.start_id = p.start_id
            q.seg_cn2 = c2.ctypes.data_as(C.POINTER(C.c_int32))
            q.fbi_cn2 = j2.ctypes.data_as(C.POINTER(C.c_int32))
            q.loop_ub = p.effective_loop_ub()
            q.bias = p.bias
            q.n_tgs = len(p.tgs_pairs) // 2
            q.tgs_pairs = tg.ctypes.data_as(C.POINTER(C.c_int32))
            q.flags = 1 if p.junc_info else 0
            q.unit_shift = unit.bit_length() - 1
            x = np.zeros(num_vars(p.n), dtype=np.int32)
            self.x.append(x)
            self.sol[i].element_cn = x.ctypes.data_as(C.POINTER(C.c_int32))

    def solutions(self) -> List[Solution]:
        return [Solution(self.x[i], int(s.score2), int(s.objective2), int(s.candidates_scored),
                         int(s.nodes_expanded), int(s.status), int(s.incumbent_score2), float(s.seconds),
                         float(s.seconds_to_best)) for i, s in enumerate(self.sol)]


@dataclass
class Dag:
    raw: _Dag

    @property
    def k(self) -> int:
        return self.raw.k

    def nodes(self):
        return [("l" if self.raw.kind[i] else "p", self.raw.a[i], self.raw.b[i], self.raw.cn[i]) for i in range(self.k)]

    def edges(self):
        return [(i, j) for i in range(self.k) for j in range(self.k) if (self.raw.adj[i] >> j) & 1]


@dataclass
class OrderResult:
    found: bool
    used_flipped: bool
    order: List[int]
    bkp: List[int]
    orders_tested: int


class ResidentBatch:
    """A batch whose inputs live in HBM (ambi_batch_upload); run() is asynchronous."""

    def __init__(self, ctx: "Context", packed: _Packed):
        self.ctx = ctx
        self.packed = packed
        self.handle = C.c_void_p()
        ctx._check(ctx.lib.ambi_batch_upload(ctx.handle, packed.count, packed.arr, C.byref(self.handle)))

    def run(self):
        self.ctx._check(self.ctx.lib.ambi_batch_run(self.ctx.handle, self.handle))

    def fetch(self) -> List[Solution]:
        self.ctx._check(self.ctx.lib.ambi_batch_fetch(self.ctx.handle, self.handle, self.packed.sol))
        return self.packed.solutions()

    def free(self):
        if self.handle:
            self.ctx.lib.ambi_batch_free(self.ctx.handle, self.handle)
            self.handle = C.c_void_p()


class Context:
    """ambi_ctx: one per process / GPU."""

    def __init__(self, device: int = 0, rank: int = 0, world: int = 1, beam_width: int = 0,
                 node_capacity: int = 0, shard_level: int = 0, lib_path: Optional[str] = None, small_kernel: bool = False):
        self.lib = load_library(lib_path)
        # small_kernel=True (AMBI_CFG_SMALL_KERNEL): partitions of <= 16 segments go through the warp-per-partition kernel first
        cfg = _Config(device, rank, world, beam_width, node_capacity, shard_level, 2 if small_kernel else 0)
        self.handle = C.c_void_p()
        rc = self.lib.ambi_init(C.byref(self.handle), C.byref(cfg))
        if rc != 0:
            msg = self.lib.ambi_last_error(None)
            raise AmbiError(rc, (msg or b"").decode() or self.lib.ambi_status_string(rc).decode())
        self.world, self.rank = world, rank

    def close(self):
        if self.handle:
            self.lib.ambi_destroy(self.handle)
            self.handle = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int):
        if rc != 0:
            raise AmbiError(rc, (self.lib.ambi_last_error(self.handle) or b"").decode()
                            or self.lib.ambi_status_string(rc).decode())

    # ---- Seam A
    def solve(self, problem: Problem) -> Solution:
        pk = _Packed([problem])
        self._check(self.lib.ambi_bfb_solve(self.handle, pk.arr, pk.sol))
        return pk.solutions()[0]

    def solve_shard(self, problem: Problem) -> Solution:
        pk = _Packed([problem])
        self._check(self.lib.ambi_bfb_solve_shard(self.handle, pk.arr, pk.sol))
        return pk.solutions()[0]

    def solve_batch(self, problems: Sequence[Problem]) -> List[Solution]:
        pk = _Packed(problems)
        self._check(self.lib.ambi_bfb_solve_batch(self.handle, pk.count, pk.arr, pk.sol))
        return pk.solutions()

    @staticmethod
    def pack(problems: Sequence[Problem]) -> "_Packed":
        """marshal once (host arrays of ambi_bfb_problem / ambi_bfb_solution) for repeated solve_packed calls"""
        return _Packed(problems)

    def solve_packed(self, pk: "_Packed") -> None:
        """ambi_bfb_solve_batch on pre-marshalled host buffers; results land in pk (pk.solutions())"""
        self._check(self.lib.ambi_bfb_solve_batch(self.handle, pk.count, pk.arr, pk.sol))

    def upload(self, problems: Sequence[Problem]) -> ResidentBatch:
        return ResidentBatch(self, _Packed(problems))

    def sync(self):
        self._check(self.lib.ambi_ctx_sync(self.handle))

    def stream(self) -> int:
        return int(self.lib.ambi_ctx_stream(self.handle) or 0)

    def launch_count(self) -> int:
        return int(self.lib.ambi_ctx_launch_count(self.handle))

    def set_profile(self, on: bool):
        self._check(self.lib.ambi_ctx_set_profile(self.handle, int(on)))

    def profile(self) -> dict:
        p = _Profile()
        self._check(self.lib.ambi_ctx_get_profile(self.handle, C.byref(p)))
        return {"expand_ms": p.expand_ms, "expand_launches": int(p.expand_launches), "total_ms": p.total_ms,
                "incumbent_pass_ms": p.incumbent_pass_ms, "record_stride": int(p.record_stride),
                "configs_enumerated": int(p.configs_enumerated), "bounds_evaluated": int(p.bounds_evaluated),
                "children_written": int(p.children_written), "bound_batches": int(p.bound_batches), "exchange_rounds": int(p.exchange_rounds),
                "chunked_runs": int(p.chunked_runs), "small_ms": p.small_ms, "small_solved": int(p.small_solved),
                "global_solved": int(p.global_solved), "small_capacity": int(p.small_capacity), "stage_b_kernel_ms": p.stage_b_kernel_ms}

    def measure_peaks(self) -> dict:
        """on-box INT32 issue rate (alu pipe; alu + fma pipes) and shared-memory bandwidth"""
        p = _Peaks()
        self._check(self.lib.ambi_measure_peaks(self.handle, C.byref(p)))
        return {"alu_tops": p.alu_tops, "alu_fma_tops": p.alu_fma_tops, "smem_tbs": p.smem_tbs, "sms": int(p.sms)}

    def comm_init(self, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, 128)
        self._check(self.lib.ambi_comm_init(self.handle, buf))

    # ---- Seam B
    def first_valid_order(self, dag: Dag, reversed_: bool = False) -> OrderResult:
        res = _OrderResult()
        self._check(self.lib.ambi_bfb_first_valid_order(self.handle, C.byref(dag.raw), int(reversed_), C.byref(res)))
        return OrderResult(bool(res.found), bool(res.used_flipped), list(res.order[:dag.k]),
                           list(res.bkp[:res.bkp_len]), int(res.orders_tested))

    def all_valid_orders(self, dag: Dag, reversed_: bool = False, max_orders: int = 4096):
        """--all: -> (orders of the initial orientation, orders of the flipped one, their breakpoint paths)"""
        k = max(dag.k, 1)
        orders = np.zeros(max_orders * k, dtype=np.int32)
        bkp = np.zeros(max_orders * 1024, dtype=np.int32)
        blen = np.zeros(max_orders, dtype=np.int32)
        n0, n1, tested = C.c_int32(0), C.c_int32(0), C.c_uint64(0)
        I = C.POINTER(C.c_int32)
        self._check(self.lib.ambi_bfb_all_valid_orders(self.handle, C.byref(dag.raw), int(reversed_), max_orders,
                                                       orders.ctypes.data_as(I), bkp.ctypes.data_as(I), 1024, blen.ctypes.data_as(I),
                                                       C.byref(n0), C.byref(n1), C.byref(tested)))
        rows = [[int(v) for v in orders[r * k:(r + 1) * k]] for r in range(n0.value + n1.value)]
        paths = [[int(v) for v in bkp[r * 1024:r * 1024 + blen[r]]] for r in range(n0.value + n1.value)]
        return rows[:n0.value], rows[n0.value:], paths

    def all_valid_paths(self, dag: Dag, inversions: dict, reversed_: bool = False, max_orders: int = 4096, path_stride: int = 4096):
        """--all with imperfectFBI and the expansion done on the device: -> (number of orders of the initial orientation,
        of the flipped one, the vertex paths in the reference's print order); `inversions` as for imperfect_fbi"""
        max_id = max([0] + [int(k) for k in inversions] + [int(v) for st in inversions.values() for v in st])
        src = np.zeros(max_id + 2, dtype=np.int32); tgt = np.zeros(max_id + 2, dtype=np.int32)
        for k, (s_, t_) in inversions.items():
            src[int(k)], tgt[int(k)] = int(s_), int(t_)
        paths = np.zeros(max_orders * path_stride, dtype=np.int32)
        plen = np.zeros(max_orders, dtype=np.int32)
        n0, n1, tested = C.c_int32(0), C.c_int32(0), C.c_uint64(0)
        I = C.POINTER(C.c_int32)
        self._check(self.lib.ambi_bfb_all_valid_paths(self.handle, C.byref(dag.raw), int(reversed_), src.ctypes.data_as(I), tgt.ctypes.data_as(I),
                                                      max_id, max_orders, None, paths.ctypes.data_as(I), path_stride, plen.ctypes.data_as(I),
                                                      C.byref(n0), C.byref(n1), C.byref(tested)))
        out = [[int(v) for v in paths[r * path_stride:r * path_stride + plen[r]]] for r in range(n0.value + n1.value)]
        return n0.value, n1.value, out

    @staticmethod
    def pack_dags(dags: Sequence[Dag], reversed_: Sequence[bool]):
        """host arrays of ambi_dag / ambi_order_result for first_valid_order_packed"""
        n = len(dags)
        return n, (_Dag * n)(*[d.raw for d in dags]), (C.c_int32 * n)(*[int(r) for r in reversed_]), (_OrderResult * n)()

    def first_valid_order_packed(self, packed) -> None:
        n, arr, rv, res = packed
        self._check(self.lib.ambi_bfb_first_valid_order_batch(self.handle, n, arr, rv, res))

    def first_valid_order_batch(self, dags: Sequence[Dag], reversed_: Sequence[bool]) -> List[OrderResult]:
        n, arr, rv, res = self.pack_dags(dags, reversed_)
        self._check(self.lib.ambi_bfb_first_valid_order_batch(self.handle, n, arr, rv, res))
        return [OrderResult(bool(r.found), bool(r.used_flipped), list(r.order[:dags[i].k]), list(r.bkp[:r.bkp_len]),
                            int(r.orders_tested)) for i, r in enumerate(res)]


def nccl_unique_id(lib_path: Optional[str] = None) -> bytes:
    lib = load_library(lib_path)
    buf = C.create_string_buffer(128)
    rc = lib.ambi_nccl_unique_id(buf)
    if rc != 0:
        raise AmbiError(rc, "ambi_nccl_unique_id failed")
    return buf.raw


def construct_dag(n: int, start_id: int, element_cn: Sequence[int], lib_path: Optional[str] = None) -> Dag:
    """constructDAG (LGM:3276-3378) -- host side of Seam B, no GPU needed."""
    lib = load_library(lib_path)
    x = np.ascontiguousarray(element_cn, dtype=np.int32)
    d = _Dag()
    rc = lib.ambi_bfb_construct_dag(n, start_id, x.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(d))
    if rc != 0:
        raise AmbiError(rc, "ambi_bfb_construct_dag failed")
    return Dag(d)


def expand_path(bkp: Sequence[int], lib_path: Optional[str] = None) -> List[int]:
    lib = load_library(lib_path)
    b = np.ascontiguousarray(bkp, dtype=np.int32)
    out = np.zeros(8192, dtype=np.int32)
    ln = lib.ambi_bfb_expand_path(b.ctypes.data_as(C.POINTER(C.c_int32)), len(b), out.ctypes.data_as(C.POINTER(C.c_int32)), len(out))
    return [int(v) for v in out[:ln]]


def imperfect_fbi(bkp: Sequence[int], inversions: dict, lib_path: Optional[str] = None) -> List[int]:
    """imperfectFBI (LGM:3431-3512) on a breakpoint path; inversions: segment id -> (source id, target id)"""
    lib = load_library(lib_path)
    b = np.ascontiguousarray(bkp, dtype=np.int32).copy()
    mx = max([1] + [abs(int(v)) for v in bkp] + [int(k) for k in inversions.keys()])
    src = np.zeros(mx + 2, dtype=np.int32)
    tgt = np.zeros(mx + 2, dtype=np.int32)
    for k, (s_, t_) in inversions.items():
        src[int(k)], tgt[int(k)] = s_, t_
    rc = lib.ambi_bfb_imperfect_fbi(b.ctypes.data_as(C.POINTER(C.c_int32)), len(b), src.ctypes.data_as(C.POINTER(C.c_int32)),
                                    tgt.ctypes.data_as(C.POINTER(C.c_int32)), mx)
    if rc != 0:
        raise AmbiError(rc, "ambi_bfb_imperfect_fbi failed")
    return [int(v) for v in b]


def better(n: int, score_a: int, x_a, score_b: int, x_b, lib_path: Optional[str] = None) -> bool:
    lib = load_library(lib_path)
    xa = np.ascontiguousarray(x_a, dtype=np.int32)
    xb = np.ascontiguousarray(x_b, dtype=np.int32)
    return bool(lib.ambi_bfb_better(n, score_a, xa.ctypes.data_as(C.POINTER(C.c_int32)), score_b,
                                    xb.ctypes.data_as(C.POINTER(C.c_int32))))


def target_cn(n: int, element_cn) -> List[int]:
    """targetCN of localhap.cpp:222-232 (host helper)"""
    lib = load_library()
    x = (C.c_int32 * (n * (n + 1)))(*[int(v) for v in element_cn])
    out = (C.c_int32 * n)()
    rc = lib.ambi_bfb_target_cn(n, x, out)
    if rc:
        raise AmbiError(rc, "ambi_bfb_target_cn")
    return list(out)


def format_path(path: Sequence[int]) -> str:
    """printBFB (LGM:3411-3429) for a single-chromosome path of signed ids."""
    out = []
    for i, v in enumerate(path):
        out.append(f"{abs(v)}{'+' if v > 0 else '-'}")
        if i + 1 < len(path) and (v > 0) != (path[i + 1] > 0):
            out.append("|")
    return "".join(out)
