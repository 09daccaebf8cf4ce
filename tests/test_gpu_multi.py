"""GPU (-m gpu), needs >= 2 devices: the NCCL path of ambi_bfb_solve (prefix-sharded search, one
process per GPU, all-gather of the shard optima over NVLink) against the oracle."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker_small_buffers(rank, world, port, out_dir):
    """candidate buffers far too small: the shards continue depth-first and exchange pruning limits over NCCL"""
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ambigram_b200 import capi, lh, synth
    from tests import oracle_lib as orc
    torch.cuda.set_device(rank)
    # rank 1 gets roomy buffers for the first problems, so that only one rank overflows there
    ok = True
    for caps in [(4096, 4096), (4096, 1 << 20)]:
        ctx = capi.Context(device=rank, rank=rank, world=world, node_capacity=caps[rank % 2])
        uid = [capi.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0])
        for (n, f, seed) in [(12, 4, 1), (16, 5, 0), (16, 5, 3), (20, 6, 1)]:
            p = lh.problem_from_lh(synth.generate(n, f, seed, 0.5).lh_text())
            s = ctx.solve(p)
            x, st = orc.solve(p, want_unique=False)
            ok &= s.status == 0 and s.score2 == st.score2 and np.array_equal(s.element_cn, x)
        pr = ctx.profile()
        if caps[rank % 2] == 4096:
            ok &= pr["chunked_runs"] > 0
        ok &= pr["exchange_rounds"] > 0
        ctx.close()
    open(os.path.join(out_dir, f"rank{rank}.ok"), "w").write("1" if ok else "0")
    dist.destroy_process_group()


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ambigram_b200 import capi, lh, synth
    from tests import oracle_lib as orc
    torch.cuda.set_device(rank)
    ctx = capi.Context(device=rank, rank=rank, world=world)
    uid = [capi.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(uid[0])
    ok = True
    for (n, f, seed) in [(6, 3, 1), (12, 4, 3), (16, 5, 4), (20, 6, 1), (1, 1, 0)]:
        p = lh.problem_from_lh(synth.generate(n, f, seed, 0.5).lh_text())
        s = ctx.solve(p)
        x, st = orc.solve(p, want_unique=False)
        ok &= s.status == 0 and s.score2 == st.score2 and np.array_equal(s.element_cn, x)
    # BASELINE config 3: n=20 / 6 folds with .juncs and --junc_info true, prefix-sharded
    sm = synth.generate(20, 6, 3, 0.0, with_juncs=True)
    p = lh.problem_from_lh(sm.lh_text(), sm.juncs_text(), True)
    s = ctx.solve(p)
    x, st = orc.solve(p, want_unique=False)
    ok &= s.status == 0 and s.score2 == st.score2 and np.array_equal(s.element_cn, x)
    open(os.path.join(out_dir, f"rank{rank}.ok"), "w").write("1" if ok else "0")
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_nccl_prefix_sharded_solve(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(tmp_path / f"rank{r}.ok").read() == "1"


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_nccl_bound_exchange_when_shards_go_depth_first(tmp_path):
    world = 2
    mp.spawn(_worker_small_buffers, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(tmp_path / f"rank{r}.ok").read() == "1"
