"""CPU suite, part 3: the N>1 path.  Block pairs are independent; ranks take contiguous,
cell-balanced ranges of the reference-ordered job list (yb_plan_split) and the host concatenates results
in job order.  Here two gloo ranks each align their range -- with the CPU oracle standing in for the
device, this is a test of the partition/gather logic, not of the kernels -- and rank 0 checks that the
gathered scripts equal the unsharded run, with no data-path collective other than the final gather."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from multiz_b200 import plan_split
        from oracle.oracle_py import Oracle
        from tools.synth import SynthBatch
        rng = np.random.default_rng(7)
        n = 60
        Ks, Ls = rng.integers(1, 6, n), rng.integers(1, 3, n)
        Ms = np.where(rng.random(n) < 0.2, rng.integers(300, 900, n), rng.integers(5, 120, n))   # ragged sizes
        sb = SynthBatch(99, Ks, Ls, Ms, R=12)                      # every rank builds the same job list
        cuts = plan_split(sb.cells_per_pair(), world)
        lo, hi = int(cuts[rank]), int(cuts[rank + 1])
        orc = Oracle(70)
        mine = []
        for i in range(lo, hi):
            A, B, LB, RB = sb.problem(i)
            o = orc.yama(A, B, LB, RB, want_tback=False)
            mine.append((i, o["cdi"].tolist(), o["script"].tobytes()))
        gathered = [None] * world
        dist.gather_object(mine, gathered if rank == 0 else None, dst=0)
        if rank == 0:
            flat = [x for part in gathered for x in part]          # rank order == reference order
            assert [x[0] for x in flat] == list(range(n))
            cells = sb.cells_per_pair()
            shares = [int(cells[cuts[r]:cuts[r + 1]].sum()) for r in range(world)]
            for i, cdi, script in flat:
                A, B, LB, RB = sb.problem(i)
                o = orc.yama(A, B, LB, RB, want_tback=False)
                assert cdi == o["cdi"].tolist() and script == o["script"].tobytes()
            q.put(("ok", shares))
    except Exception as e:  # pragma: no cover
        if rank == 0:
            q.put(("fail", repr(e)))
        raise
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_matches_unsharded():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    status, shares = q.get(timeout=10)
    assert status == "ok", shares
    assert max(shares) <= 0.65 * sum(shares)        # cell-balanced although pair sizes are ragged
