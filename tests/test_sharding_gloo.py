"""Multi-GPU host logic on CPU: world_size-2 gloo processes shard the job list the way bench.py / the farm
do (job i -> rank i mod N, no data-path collective) and agree on the max-over-ranks timing reduction."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    n_jobs = 10
    rng = np.random.default_rng(0)
    imgs = [rng.integers(0, 256, (40 + i, 60 + i, 3), dtype=np.uint8) for i in range(n_jobs)]
    mine = list(range(rank, n_jobs, world))                       # round-robin shard
    pix = 0
    digest = 0
    for i in mine:
        code, step, out = O.run_chain(imgs[i], resize="16,16")
        assert code == 0
        pix += out.shape[0] * out.shape[1]
        digest += int(out.astype(np.int64).sum()) * (i + 1)
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)      # stand-in for this rank's elapsed time
    dist.barrier()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    tot = torch.tensor([pix, digest], dtype=torch.int64)
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    q.put((rank, mine, float(t.item()), int(tot[0]), int(tot[1])))
    dist.destroy_process_group()


def test_round_robin_sharding_world_size_2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs: p.join(timeout=60)
    assert res[0][1] == [0, 2, 4, 6, 8] and res[1][1] == [1, 3, 5, 7, 9]         # disjoint, complete
    assert res[0][2] == res[1][2] == 2.0                                        # max over ranks
    assert res[0][3] == res[1][3] == 10 * 256                                   # every job processed exactly once
    # checksum-of-checksums equals the single-process result
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    rng = np.random.default_rng(0)
    imgs = [rng.integers(0, 256, (40 + i, 60 + i, 3), dtype=np.uint8) for i in range(10)]
    want = sum(int(O.run_chain(im, resize="16,16")[2].astype(np.int64).sum()) * (i + 1) for i, im in enumerate(imgs))
    assert res[0][4] == want
