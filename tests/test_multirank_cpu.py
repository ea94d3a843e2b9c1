"""CPU tests (gloo, world_size 2) of the multi-rank plumbing: the partition the library uses
for target slices / pair-tile ranges covers the work exactly once, and the host-side helpers
bench.py uses (unique-id broadcast, max-over-ranks timing) work under torch.distributed."""
import ctypes as C
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import unsflow_b200 as uf
    L = uf.lib()
    res = []
    for variant in (1, 2):
        b, e = C.c_int64(), C.c_int64()
        assert L.unsflow_partition(n, world, rank, variant, C.byref(b), C.byref(e)) == 0
        res += [b.value, e.value]
    t = torch.tensor(res, dtype=torch.int64)
    gathered = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)
    # unique-id style broadcast (bench.py) and max-over-ranks timing
    uid = torch.arange(128, dtype=torch.uint8) if rank == 0 else torch.zeros(128, dtype=torch.uint8)
    dist.broadcast(uid, 0)
    tm = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    if rank == 0:
        out.put(([g.tolist() for g in gathered], uid.tolist(), tm.item()))
    dist.barrier()
    dist.destroy_process_group()


def test_partition_covers_work_exactly_once_world2():
    world, n = 2, 1_000_001
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    parts, uid, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # direct: target slices tile [0, n) with no gap/overlap
    assert parts[0][0] == 0 and parts[0][1] == parts[1][0] and parts[1][1] == n
    # symmetric: tile ranges tile [0, W)
    W = _ntiles(n, 2)
    assert parts[0][2] == 0 and parts[0][3] == parts[1][2] and parts[1][3] == W
    assert abs((parts[0][3] - parts[0][2]) - (parts[1][3] - parts[1][2])) <= 1    # balanced
    assert uid == list(range(128)) and tmax == 11.0


def _ntiles(n, world=1):
    nb = (n + 2047) // 2048                   # 2048-wide pair tiles (sym_rows_per_thread, induce.cu)
    return 32 * (nb * (nb + 1) // 2)          # work units: 1/32 of a pair tile (kSymSub)


def test_partition_many_ranks_single_process():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import unsflow_b200 as uf
    L = uf.lib()
    for n in (1, 1023, 1024, 1025, 99_999, 1_000_000):
        for world in (1, 2, 4, 8):
            for variant, total in ((1, n), (2, _ntiles(n, world))):
                prev = 0
                for r in range(world):
                    b, e = C.c_int64(), C.c_int64()
                    assert L.unsflow_partition(n, world, r, variant, C.byref(b), C.byref(e)) == 0
                    assert b.value == prev and e.value >= b.value
                    prev = e.value
                assert prev == total
