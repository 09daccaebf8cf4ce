"""CPU, world_size 2 over gloo: the host logic of the N>1 paths.

 * prefix sharding: each rank owns the subtrees whose prefix hashes to it (oracle emulation of
   k_expand's shard filter, identical hash); the all-gathered shard optima are combined with the
   library's own rule (ambi_bfb_better) and must equal the unsharded optimum;
 * sample sharding (what bench.py does): ranks draw disjoint seed ranges and the job total is the
   sum over ranks.
"""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ambigram_b200 import capi, lh, synth
    from tests import oracle_lib as orc
    import bench

    ok = True
    # ---- prefix-sharded single partition
    for (n, f, seed) in [(6, 3, 1), (10, 4, 2), (12, 4, 3), (14, 4, 4)]:
        p = lh.problem_from_lh(synth.generate(n, f, seed, 0.5).lh_text())
        x, st = orc.solve_shard(p, world, rank, 2)
        nv = n * (n + 1)
        mine = torch.zeros(nv + 1, dtype=torch.int64)
        mine[0] = int(st.score2)
        mine[1:] = torch.from_numpy(x.astype(np.int64))
        gathered = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        best = None
        for g in gathered:
            s, xs = int(g[0]), g[1:].numpy().astype(np.int32)
            if best is None or capi.better(n, s, xs, best[0], best[1]):
                best = (s, xs)
        xg, sg = orc.solve(p, want_unique=False)
        ok &= best[0] == sg.score2 and np.array_equal(best[1], xg)
    # ---- sample-sharded batch (bench.py's weak-scaling layout)
    batch = 6
    probs = bench.make_problems(batch, rank * batch)
    _, score, cand = orc.solve_batch(probs, 1)
    tot = torch.tensor([float(cand.sum()), float(score.sum())], dtype=torch.float64)
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    allp = bench.make_problems(batch * world, 0)
    _, score_all, cand_all = orc.solve_batch(allp, 1)
    ok &= int(tot[0]) == int(cand_all.sum()) and int(tot[1]) == int(score_all.sum())
    open(os.path.join(out_dir, f"rank{rank}.ok"), "w").write("1" if ok else "0")
    dist.destroy_process_group()


def test_world2_gloo_shard_and_combine(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(tmp_path / f"rank{r}.ok").read() == "1"
