"""world_size-2 gloo test (CPU) of the only multi-rank logic the path has: contiguous env
sharding and the episode-return all-gather."""
import os
import socket

import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from pcc_rl_b200 import distributed as D
    r, lr, w = D.init_from_env(backend="gloo")
    lo, hi = D.shard_range(1001, r, w)
    rets = torch.arange(lo, hi, dtype=torch.float64)[: (3 if r == 0 else 5)]
    st = D.gather_episode_returns(rets)
    mx = D.max_over_ranks(float(r + 1), torch.device("cpu"))
    sm = D.sum_over_ranks(float(hi - lo), torch.device("cpu"))
    q.put((r, lo, hi, st, mx, sm))
    torch.distributed.destroy_process_group()


def test_gloo_world2_shard_and_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (r0, lo0, hi0, st0, mx0, sm0), (r1, lo1, hi1, st1, mx1, sm1) = out
    assert (lo0, hi0, lo1, hi1) == (0, 500, 500, 1001)
    want = [0.0, 1.0, 2.0] + [500.0, 501.0, 502.0, 503.0, 504.0]
    for st in (st0, st1):
        assert st["count"] == 8 and abs(st["mean"] - sum(want) / 8) < 1e-12
        assert st["per_rank"][0] == (3, 1.0) and st["per_rank"][1] == (5, 502.0)
    assert mx0 == mx1 == 2.0 and sm0 == sm1 == 1001.0
