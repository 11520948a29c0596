"""Multi-GPU host logic on CPU: world_size-2 gloo processes shard a mixed-size job list with the PRODUCT's own assignment
(imp_gpu_farm_assign of libimp_gpu.so: round-robin and size-aware; bench.py's lpt_assign / job_order for the strong-scaled
cfg5 leg), run their shares through the product planner + per-pixel headers (the hostsim twin), and agree on the
max-over-ranks timing reduction and on a checksum of checksums that must equal the oracle's single-process result."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SHAPES = [(40 + 13 * i, 60 + 7 * ((i * 5) % 9), 3 if i % 3 else 4) for i in range(12)]


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _images():
    rng = np.random.default_rng(0)
    return [rng.integers(0, 256, s, dtype=np.uint8) for s in SHAPES]


def _plans(L, api):
    return [L.plan(s[1], s[0], s[2], api.Config(max_w=0, max_h=0), resize="16,16") for s in SHAPES]


def _worker(rank, world, port, q, policy):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ngx_http_imgproc_b200 as M
    from ngx_http_imgproc_b200 import api
    from conftest import HostSim
    L = M.library()                                               # no imp_gpu_init: validation and sharding need no device
    plans = _plans(L, api)
    owner = api.farm_assign(L, plans, world, policy)              # the library's own sharding
    mine = [i for i, g in enumerate(owner) if g == rank]
    load = sum(plans[i].algorithmic_bytes for i in mine)
    sim = HostSim()
    imgs = _images()
    pix = digest = 0
    for i in mine:
        code, step, out = sim.run(imgs[i], api.Config(max_w=0, max_h=0), resize="16,16")
        assert code == 0
        pix += out.shape[0] * out.shape[1]
        digest += int(out.astype(np.int64).sum()) * (i + 1)
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)      # stand-in for this rank's elapsed time
    dist.barrier()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    tot = torch.tensor([pix, digest], dtype=torch.int64)
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    q.put((rank, mine, float(t.item()), int(tot[0]), int(tot[1]), int(load)))
    dist.destroy_process_group()


@pytest.mark.parametrize("policy", [0, 1], ids=["round_robin", "size_aware"])
def test_farm_sharding_world_size_2(policy):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, policy)) for r in range(2)]
    for p in procs: p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs: p.join(timeout=60)
    n = len(SHAPES)
    assert sorted(res[0][1] + res[1][1]) == list(range(n))                       # disjoint, complete
    if policy == 0:
        assert res[0][1] == list(range(0, n, 2)) and res[1][1] == list(range(1, n, 2))
    else:
        total = res[0][5] + res[1][5]
        assert abs(res[0][5] - res[1][5]) <= 0.15 * total                       # largest-first evens the bytes out
    assert res[0][2] == res[1][2] == 2.0                                        # max over ranks
    assert res[0][3] == res[1][3] == n * 256                                    # every job processed exactly once
    # checksum-of-checksums equals the oracle's single-process result
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    want = sum(int(O.run_chain(im, resize="16,16", cfg=O.OracleConfig(max_w=0, max_h=0))[2].astype(np.int64).sum()) * (i + 1) for i, im in enumerate(_images()))
    assert res[0][4] == want


def test_bench_sharding_helpers():
    """bench.py's strong-scaled cfg5 leg: job j reads source j mod 512 and the two assignments cover every job once."""
    sys.path.insert(0, ROOT)
    import bench
    wl = bench.workload("cfg5", 64)
    order = bench.job_order(wl)
    assert len(order) == 65536 // 64 and order[:5] == [0, 1, 2, 3, 4] and order[512:515] == [0, 1, 2]
    costs = [wl["jobs"][s][0][0] * wl["jobs"][s][0][1] * wl["jobs"][s][0][2] for s in order]
    for n in (2, 4, 8):
        owner = bench.lpt_assign(costs, n)
        loads = [sum(c for c, g in zip(costs, owner) if g == k) for k in range(n)]
        assert len(owner) == len(costs) and set(owner) == set(range(n))
        assert max(loads) - min(loads) <= max(costs)                             # LPT: within one job of each other
        rr = [sum(c for j, c in enumerate(costs) if j % n == k) for k in range(n)]
        assert max(loads) <= max(rr)
