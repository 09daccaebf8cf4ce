#!/usr/bin/env python3
"""bench.py -- BFB candidates scored/sec on B200 (BASELINE.json metric), one JSON line.

A "step" is one pass of the hot path (Stage-A fold-pattern search: incumbent pass + exact pass)
over one batch of synthetic BFB partitions.  Workload at N=1 = BASELINE.json configs[1]:
synthetic BFB .lh, 12 segments / 4 folds (seeded, ambigram_b200/synth.py), --batch (default 32768) of them per step.

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, C ABI)
    python bench.py --impl reference --gpus N ...            # CPU arm: the oracle port on host cores

value   = candidates scored per second, whole job, inputs resident in HBM (device-event timed)
e2e     = the same through ambi_bfb_solve_batch with HOST buffers (H2D + D2H inside the timed region)
roofline / cpu_baseline / clocks: see the task contract and DESIGN.md "Measurement".
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "bfb_candidates_scored_per_sec"
UNIT = "candidates/s"
N_SEG, N_FOLDS = 12, 4


def make_problems(count: int, seed0: int, n: int = N_SEG, folds: int = N_FOLDS):
    from ambigram_b200 import lh, synth
    return [lh.problem_from_lh(synth.generate(n, folds, seed0 + s).lh_text()) for s in range(count)]


def host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def profiled_traffic():
    """dram bytes per k_expand launch from the committed ncu capture of this round (profiles/)"""
    for name in ("r02_launches_summary.json", "r01_launches_summary.json"):
        try:
            return float(json.load(open(os.path.join(ROOT, "profiles", name)))["k_expand"]["dram_bytes_per_launch"])
        except Exception:
            continue
    return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.splitlines():
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_baseline(problems, budget_s: float = 12.0, threads: int = 0):
    """oracle port (oracle/ambi_oracle.c, OpenMP over problems) on a bounded sample of the workload:
    the step's partitions, solved repeatedly until about budget_s of CPU wall time has been spent"""
    from tests import oracle_lib as orc
    th = threads or host_threads()
    orc.solve_batch(problems[:64], th)          # warm-up (page in, spin up the thread team)
    tot_c, tot_p, t0 = 0, 0, time.perf_counter()
    while True:
        _, _, cand = orc.solve_batch(problems, th)
        tot_c += int(cand.sum()); tot_p += len(problems)
        dt = time.perf_counter() - t0
        if dt >= budget_s:
            break
    return {"value": float(tot_c / dt), "unit": UNIT, "cores": int(th), "kind": "port",
            "sample": f"{tot_p} partitions (the step's {len(problems)} n={N_SEG}/{N_FOLDS}-fold partitions, "
                      f"{tot_p // len(problems)} passes), oracle/ambi_oracle.c orc_solve_batch, {dt:.1f} s wall",
            "problems_per_s": tot_p / dt, "reference_binary": reference_binary_timing()}


def reference_binary_timing(count: int = 4, n: int = N_SEG, folds: int = N_FOLDS, budget_s: float = 40.0):
    """The reference's own binary end to end (oracle/_ref/Ambigram_ref = its unmodified sources compiled with
    the COIN header shim; HiGHS via scipy plays the external `cbc`, oracle/cbc_standin.py) on the first
    `count` .lh files of the workload, one process per file and single-threaded like the original.
    None when the build did not travel to this box."""
    import tempfile
    from ambigram_b200 import synth
    ref = os.path.join(ROOT, "oracle", "_ref", "Ambigram_ref")
    cbc_dir = os.path.join(ROOT, "oracle", "_ref", "bin")
    if not (os.path.exists(ref) and os.path.exists(os.path.join(cbc_dir, "cbc"))):
        return None
    try:
        import scipy  # noqa: F401  (the cbc stand-in needs it)
    except Exception:
        return None
    times, t_start = [], time.perf_counter()
    with tempfile.TemporaryDirectory() as d:
        env = dict(os.environ, PATH=cbc_dir + ":" + os.environ.get("PATH", ""), OMP_NUM_THREADS="1")
        for s in range(count):
            fn = os.path.join(d, f"s{s}.lh")
            open(fn, "w").write(synth.generate(n, folds, s).lh_text())
            t0 = time.perf_counter()
            try:
                r = subprocess.run([ref, "--op", "bfb", "--in_lh", f"s{s}.lh", "--lp_prefix", f"s{s}"], cwd=d, env=env,   # the reference opens "./<prefix>.sol"
                                   capture_output=True, text=True, timeout=120)
            except Exception:
                continue
            if r.returncode == 0:
                times.append(time.perf_counter() - t0)
            if time.perf_counter() - t_start > budget_s:
                break
    if not times:
        return None
    return {"seconds_per_lh": statistics.mean(times), "lh_per_s": 1.0 / statistics.mean(times), "files": len(times), "cores": 1,
            "what": f"oracle/_ref/Ambigram_ref --op bfb on the first {len(times)} n={n}/{folds}-fold .lh of the workload, "
                    "HiGHS (scipy.optimize.milp) standing in for cbc, process start included"}


def workload_problems(name: str, count: int, rank: int, world: int):
    """the other BASELINE configs (parity-test cases; not the default bench line)"""
    import random
    from ambigram_b200 import lh, synth
    if name == "sweep":          # config 5: n ~ U{10..25}, folds = ceil(n/3.5), seeds 0..count-1, sample-sharded
        rng = random.Random(12345)
        ns = [rng.randint(10, 25) for _ in range(count)]
        # sample-sharded; longest partitions first, dealt round-robin, so that every rank gets the same mix of
        # lengths (search cost grows ~1.6x per segment: dealing by index would leave one rank with the stragglers)
        order = sorted(range(count), key=lambda s: (-ns[s], s))
        mine = order[rank::world]
        return [lh.problem_from_lh(synth.generate(ns[s], synth.folds_for(ns[s]), s).lh_text()) for s in mine], "sample-sharded"
    if name == "n30":            # config 4: n=30 / 8 folds, CN noise +-1 w.p. 0.5, prefix-sharded
        return [lh.problem_from_lh(synth.generate(30, 8, s, 0.5).lh_text()) for s in range(count)], "prefix-sharded"
    if name == "n20j":           # config 3: n=20 / 6 folds with .juncs + --junc_info true, prefix-sharded
        out = []
        for s in range(count):
            sm = synth.generate(20, 6, s, 0.0, with_juncs=True)
            out.append(lh.problem_from_lh(sm.lh_text(), sm.juncs_text(), True))
        return out, "prefix-sharded"
    raise SystemExit(f"unknown workload {name}")


def prefix_sharded_record(local: int, rank: int, world: int, dist):
    """north_star's multi-GPU mode for ONE partition at a time (BASELINE configs 3 and 4): every rank searches the
    prefixes hashed to it, incumbents and pruning limits are exchanged over NCCL, the shard optima are all-gathered
    (ambi_bfb_solve with a communicator).  Strong scaling: the same partitions at every N; timed on the host around
    the collective C-ABI calls, max over ranks."""
    import torch
    from ambigram_b200 import capi
    ctx = capi.Context(device=local, rank=rank, world=world)
    if world > 1:
        uid = [capi.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0])
    out = {}
    for name, count in (("n30", 8), ("n20j", 16)):
        probs, _ = workload_problems(name, count, 0, 1)
        for p in probs[:2]:
            ctx.solve(p)                                   # warm-up (graph capture, pools)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        per_lh, cands, t_best = [], 0, []
        t00 = time.perf_counter()
        for p in probs:
            t0 = time.perf_counter()
            s = ctx.solve(p)
            per_lh.append(time.perf_counter() - t0)
            cands += int(s.candidates_scored)
            t_best.append(float(s.seconds_to_best))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t00
        t = torch.tensor([dt], dtype=torch.float64, device=f"cuda:{local}")
        ranks = [torch.zeros_like(t) for _ in range(world)]
        if world > 1:
            dist.all_gather(ranks, t)
        else:
            ranks = [t]
        prof = ctx.profile()
        out[name] = {"partitions": len(probs), "ms_total": 1e3 * max(float(r[0]) for r in ranks), "ms_per_lh": 1e3 * max(float(r[0]) for r in ranks) / len(probs),
                     "per_rank_ms": [round(1e3 * float(r[0]), 3) for r in ranks], "candidates": cands,
                     "candidates_per_s": cands / max(float(r[0]) for r in ranks), "slowest_lh_ms": 1e3 * max(per_lh),
                     "time_to_best_ms_median": 1e3 * statistics.median(t_best), "exchange_rounds": prof["exchange_rounds"], "chunked_runs": prof["chunked_runs"]}
    out["what"] = ("BASELINE config 4 (n=30 / 8 folds, CN noise, seeds 0..7) and config 3 (n=20 / 6 folds with .juncs, seeds 0..15), one "
                   "partition at a time through ambi_bfb_solve; prefix-sharded x%d over NCCL; strong scaling; host wall clock, max over ranks" % world)
    ctx.close()
    return out


def sweep_record(ctx, local: int, rank: int, world: int, dist, count: int = 2000):
    """BASELINE config 5 (batch of samples with 10..25 segments, folds = ceil(n/3.5)), sample-sharded: a fixed set of
    `count` samples dealt over the ranks (strong scaling), one ambi_bfb_solve_batch per rank with host buffers"""
    import torch
    problems, _ = workload_problems("sweep", count, rank, world)
    pk = ctx.pack(problems)
    ctx.solve_packed(pk)                                   # warm-up
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ctx.solve_packed(pk)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    sols = pk.solutions()
    t = torch.tensor([dt], dtype=torch.float64, device=f"cuda:{local}")
    c = torch.tensor([float(sum(s.candidates_scored for s in sols)), float(len(problems))], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    secs = float(t[0])
    return {"samples": int(c[1]), "seconds": secs, "partitions_per_s": float(c[1]) / secs, "candidates_per_s": float(c[0]) / secs,
            "what": "BASELINE config 5 at %d samples (n ~ U{10..25}, seeds 0..%d), sample-sharded x%d, strong scaling, host buffers, "
                    "host wall clock, max over ranks" % (count, count - 1, world)}


def run_workload(args):
    import torch
    import torch.distributed as dist
    from ambigram_b200 import capi
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    count = args.batch if args.batch > 0 else {"sweep": 10000, "n30": 8, "n20j": 64}[args.workload]
    problems, how = workload_problems(args.workload, count, rank, world)
    sharded = how == "prefix-sharded"
    ctx = capi.Context(device=local, rank=rank if sharded else 0, world=world if sharded else 1)
    if sharded and world > 1:
        uid = [capi.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0])

    def step():
        if sharded:
            return [ctx.solve(p) for p in problems]
        pk = ctx.pack(problems)
        ctx.solve_packed(pk)
        return pk.solutions()

    for _ in range(max(1, min(args.warmup, 2))):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sols = step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    cand = float(sum(s.candidates_scored for s in sols))
    t = torch.tensor([dt], dtype=torch.float64, device=f"cuda:{local}")
    c = torch.tensor([cand if not sharded or rank == 0 else 0.0, float(len(problems)) if not sharded or rank == 0 else 0.0],
                     dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    if rank == 0:
        secs = float(t[0]) / args.steps
        print(json.dumps({"metric": METRIC, "value": float(c[0]) / secs, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": 1e3 * secs, "higher_is_better": True, "scaling": "strong",
                          "vs_baseline": None, "dtype": "int32", "data": "synthetic",
                          "config": {"workload": args.workload, "partitions_per_step": int(c[1]), "parallelism": f"{how} x{world}"},
                          "partitions_per_s": float(c[1]) / secs, "time_per_lh_ms": 1e3 * secs / max(float(c[1]), 1),
                          "timing": "host wall clock around the C-ABI calls (host buffers), max over ranks"}))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def arm_config(batch: int, world: int):
    """`config` of the bench line -- the same object for our arm and for --impl reference"""
    return {"workload": f"synthetic BFB .lh, {N_SEG} segments / {N_FOLDS} folds (BASELINE configs[1]), "
                        f"{batch} partitions per step per GPU, seeds rank*batch..",
            "batch_per_gpu": batch, "l2": "256 MB buffer written between timed steps (GPU arm)",
            "parallelism": f"sample-sharded x{world}"}


def run_reference(args):
    """--impl reference: the reference's CPU path for this metric = the oracle port (Cbc, the reference's
    solver, is not in the image and does not count candidates; DESIGN.md 'Reference arm')."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from tests import oracle_lib as orc
    problems = make_problems(args.batch, 0)
    th = host_threads()         # torchrun exports OMP_NUM_THREADS=1; the arm uses every host core it is allowed
    # bounded sample per step
    probe = problems[:min(64, len(problems))]
    t0 = time.perf_counter(); orc.solve_batch(probe, th); dt = time.perf_counter() - t0
    per_step = int(min(len(problems), max(len(probe), len(probe) * 8.0 / max(dt, 1e-6))))
    sample = problems[:per_step]
    for _ in range(args.warmup):
        orc.solve_batch(sample, th)
    tot_c, tot_t = 0, 0.0
    for _ in range(args.steps):
        t0 = time.perf_counter()
        _, _, cand = orc.solve_batch(sample, th)
        tot_t += time.perf_counter() - t0
        tot_c += int(cand.sum())
    value = tot_c / tot_t
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": arm_config(args.batch, max(1, args.gpus)),         # the config of our own arm, word for word
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": th, "kind": "port",
                             "sample": f"{per_step} of the step's {args.batch} partitions per step (rank 0's), oracle/ambi_oracle.c (OpenMP)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "problems_per_s": per_step * args.steps / tot_t, "reference_binary": reference_binary_timing()}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--batch", type=int, default=0, help="partitions per step per GPU (default 32768 for n12)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="n12", help="n12 (default, BASELINE configs[1]) | n20j | n30 | sweep")
    args = ap.parse_args()
    if args.workload == "n12" and args.batch <= 0:
        args.batch = 32768
    if args.impl == "reference":
        return run_reference(args)
    if args.workload != "n12":
        return run_workload(args)

    import torch
    import torch.distributed as dist
    from ambigram_b200 import capi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    # weak scaling: every rank owns its own --batch partitions (sample-sharded, no data-path collective)
    problems = make_problems(args.batch, rank * args.batch)
    ctx = capi.Context(device=local, rank=0, world=1)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=local)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local}")   # > 126 MB L2

    batch = ctx.upload(problems)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident steps -------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        batch.run(); ctx.sync()
    sols = batch.fetch()
    cand_per_step = sum(s.candidates_scored for s in sols)
    nodes_per_step = sum(s.nodes_expanded for s in sols)
    launches0 = ctx.launch_count()
    sampler = ClockSampler(local) if rank == 0 and not os.environ.get("AMBI_BENCH_NO_SAMPLER") else None
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for k in range(args.steps):
        flush.zero_()                       # L2 flush between timed iterations (untimed)
        torch.cuda.synchronize()
        ev[k][0].record(stream)
        batch.run()
        ev[k][1].record(stream)
    barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    launches = ctx.launch_count() - launches0
    sols = batch.fetch()
    tb = sorted(s.seconds_to_best for s in sols); tp = sorted(s.seconds for s in sols)
    pct = lambda v, q: v[min(len(v) - 1, int(q * len(v)))]
    time_to_best = {"unit": "ms since the search of the batch was enqueued (device %globaltimer, per .lh)",
                    "best_p50": 1e3 * pct(tb, 0.5), "best_p99": 1e3 * pct(tb, 0.99), "best_max": 1e3 * tb[-1],
                    "proof_p50": 1e3 * pct(tp, 0.5), "proof_p99": 1e3 * pct(tp, 0.99), "proof_max": 1e3 * tp[-1]}
    t_best = statistics.mean(tb)

    # ---- Stage B on the same batch (first valid topological order per partition) and single-.lh latency
    dags = [capi.construct_dag(p.n, p.start_id, s.element_cn) for p, s in zip(problems, sols)]
    dpk = ctx.pack_dags(dags, [False] * len(dags))
    ctx.first_valid_order_packed(dpk)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        ctx.first_valid_order_packed(dpk)
    stage_b_s = (time.perf_counter() - t0) / 5
    sb_kernel_s = ctx.profile()["stage_b_kernel_ms"] * 1e-3
    sb_place = sum(int(r.orders_tested) for r in dpk[3])
    # algorithmic work of a placement: about k * len integer operations, len = breakpoint-path length
    sb_ops = sum(int(r.orders_tested) * max(1, d.k) * max(2, int(r.bkp_len)) for d, r in zip(dags, dpk[3]))
    stage_b = {"dags": len(dags), "seconds": stage_b_s, "partitions_per_s": len(dags) / stage_b_s,
               "placements_per_s": sb_place / stage_b_s, "kernel_ms": 1e3 * sb_kernel_s,
               "kernel_partitions_per_s": len(dags) / max(sb_kernel_s, 1e-9), "kernel_int_tops": sb_ops / max(sb_kernel_s, 1e-9) / 1e12,
               "found": sum(1 for r in dpk[3] if r.found), "note": "host wall clock around ambi_bfb_first_valid_order_batch (H2D + kernel + D2H)"}
    one = problems[0]
    for _ in range(3):
        ctx.solve(one)
    t0 = time.perf_counter()
    single_best = []
    for _ in range(20):
        single_best.append(ctx.solve(one).seconds_to_best)
    single_ms = (time.perf_counter() - t0) / 20 * 1e3
    time_to_best["single_lh_best_ms"] = 1e3 * statistics.median(single_best)
    batch.free()
    batch = ctx.upload(problems)

    # ---- on-box integer / shared-memory peaks (not in MEASURED_PEAKS.json) ----
    peaks = ctx.measure_peaks() if rank == 0 else None

    # ---- roofline pass: per-launch duration of the dominant kernel (k_expand) ----
    ctx.set_profile(True)
    batch.run(); ctx.sync(); batch.fetch()
    prof = ctx.profile()
    ctx.set_profile(False)

    # ---- end to end through the C ABI with host buffers --------------------------
    pk = ctx.pack(problems)                 # host buffers of the C ABI, marshalled once (not part of the path)
    for _ in range(2):
        ctx.solve_packed(pk)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.solve_packed(pk)                # H2D of the problems, search, D2H of the results, every step
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_sols = pk.solutions()
    clocks = sampler.stop() if sampler else None
    e2e_cand = sum(s.candidates_scored for s in e2e_sols)

    # ---- north_star's multi-GPU split (one partition at a time, prefix-sharded over NCCL): configs 4 and 3
    sharded = None
    if not os.environ.get("AMBI_BENCH_NO_SHARDED"):
        sharded = prefix_sharded_record(local, rank, world, dist)
    sweep = None if os.environ.get("AMBI_BENCH_NO_SWEEP") else sweep_record(ctx, local, rank, world, dist)

    if os.environ.get("AMBI_BENCH_DEBUG"):
        print(f"[rank {rank}] dev_ms/step {dev_ms / args.steps:.3f} e2e_ms/step {1e3 * e2e_s / args.steps:.3f} cand/step {cand_per_step}", file=sys.stderr)
    # max over ranks / sum over ranks
    t = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device=f"cuda:{local}")
    c = torch.tensor([float(cand_per_step), float(e2e_cand), float(launches)], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    dev_ms_max, e2e_s_max = float(t[0]), float(t[1])
    cand_all, e2e_cand_all, launches_all = float(c[0]), float(c[1]), int(c[2])

    if rank == 0:
        hbm_peak, peak_src, sm_max = measured_peaks()
        ms_per_step = dev_ms_max / args.steps
        value = cand_all / (ms_per_step * 1e-3)
        e2e_value = e2e_cand_all * args.steps / e2e_s_max
        # algorithmic work of the dominant kernel: per candidate 8n+16 INT32 ops;
        # per record materialised in a frontier buffer (read or written) 4n+16 bytes
        n = N_SEG
        exp_s = prof["expand_ms"] * 1e-3
        recs = 2 * nodes_per_step            # every expanded record was written once and read once
        alg_bytes = recs * (4 * n + 16)
        alg_ops = cand_per_step * (8 * n + 16)
        int_peak = max(peaks["alu_tops"], peaks["alu_fma_tops"]) * 1e12   # measured on this box: the better of add+logic and add+mad chains
        roof = {"bound": "hbm", "achieved": alg_bytes / exp_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": alg_bytes / exp_s / 1e9 / hbm_peak, "traffic": profiled_traffic(),
                "algorithmic_bytes_per_launch": alg_bytes / max(prof["expand_launches"], 1), "peak_source": peak_src,
                "kernel": "k_expand", "launches": prof["expand_launches"],
                "avg_launch_us": 1e3 * prof["expand_ms"] / max(prof["expand_launches"], 1),
                "share_of_step": prof["expand_ms"] / prof["total_ms"],
                "note": "k_expand is INT32-issue bound, not HBM bound: see roofline_int32"}
        roof_int = {"bound": "int32", "achieved": alg_ops / exp_s / 1e12, "peak": int_peak / 1e12, "unit": "Top/s",
                    "frac": alg_ops / exp_s / int_peak, "ops_per_candidate": 8 * n + 16,
                    "peak_source": "measured on this box by ambi_measure_peaks (best of add+logic and add+mad register chains)",
                    "peak_add_logic": peaks["alu_tops"], "peak_add_mad": peaks["alu_fma_tops"], "smem_peak_tbs": peaks["smem_tbs"],
                    "nominal": 148 * 128 * sm_max * 1e6 / 1e12,
                    "funnel": {k: prof[k] for k in ("configs_enumerated", "bounds_evaluated", "children_written", "bound_batches")},
                    "hardware_counters": "profiles/ (ncu): alu pipe and issue-slot utilisation of k_expand"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "int32", "data": "synthetic",
                "config": arm_config(args.batch, world),
                "e2e": {"value": e2e_value, "unit": UNIT,
                        # per partition: compact problem (4 header words + one 2c|2J word per segment) + active flag in;
                        # score, incumbent, row codes, two 64-bit counters, four time stamps, active flag out
                        "h2d_bytes_per_step": args.batch * (4 * (4 + N_SEG) + 1) * world,
                        "d2h_bytes_per_step": args.batch * (4 + 4 + 4 * N_SEG + 8 + 8 + 32 + 1) * world},
                "gpu_launches": launches_all,
                "candidates_per_step": cand_all, "records_expanded_per_step": nodes_per_step, "problems_per_s": args.batch * world / (ms_per_step * 1e-3),
                "time_to_best_ms_per_lh": 1e3 * t_best, "time_to_best": time_to_best, "single_lh_solve_ms": single_ms, "stage_b": stage_b,
                "prefix_sharded": sharded, "sweep": sweep,
                "roofline": roof, "roofline_int32": roof_int, "clocks": clocks}
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(problems)
        print(json.dumps(line))
    batch.free()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
