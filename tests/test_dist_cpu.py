"""Host-side logic of the multi-GPU path, exercised with world_size 2 over gloo on the CPU: the partition map tiles the
rows, and the exchange plans two ranks derive independently are mirror images (what one sends the other receives)."""
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


def _worker(rank, world, port, out):
    import ctypes as C
    sys.path.insert(0, ROOT)
    from meshlessmultigridpoisson_b200 import capi

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L = capi.load()
    ok = True
    for n in (10, 4000000, 62500, 7):
        b = capi.partition_bounds(n, world)
        mine = torch.tensor([int(b[rank]), int(b[rank + 1])])
        allb = [torch.zeros(2, dtype=torch.long) for _ in range(world)]
        dist.all_gather(allb, mine)
        ok &= allb[0][0].item() == 0 and allb[-1][1].item() == n
        ok &= all(allb[r][1].item() == allb[r + 1][0].item() for r in range(world - 1))            # blocks tile [0, n)
        sizes = [int(t[1] - t[0]) for t in allb]
        ok &= max(sizes) - min(sizes) <= 1
        # every rank reads a band of half-width bw around its own block (banded operator in the reference order)
        bw = max(1, n // 7)
        need = np.array([[max(0, b[r] - bw), min(n, b[r + 1] + bw)] for r in range(world)], np.int32).ravel()
        ns, nr = C.c_int(), C.c_int()
        sends, recvs = np.zeros(3 * world, np.int32), np.zeros(3 * world, np.int32)
        rc = L.mmg_debug_exchange_plan(rank, world, need, np.ascontiguousarray(b, np.int32), C.byref(ns), sends, C.byref(nr), recvs)
        ok &= rc == 0
        msg = torch.zeros(2, 3 * world, dtype=torch.long)
        msg[0, : 3 * ns.value] = torch.from_numpy(sends[: 3 * ns.value].astype(np.int64))
        msg[1, : 3 * nr.value] = torch.from_numpy(recvs[: 3 * nr.value].astype(np.int64))
        allm = [torch.zeros_like(msg) for _ in range(world)]
        dist.all_gather(allm, msg)
        for r in range(world):                      # what r sends to me == what I receive from r
            theirs = [tuple(allm[r][0, 3 * i: 3 * i + 3].tolist()) for i in range(world) if allm[r][0, 3 * i + 2] > 0]
            to_me = [(o, c) for (p, o, c) in theirs if p == rank]
            from_r = [(int(recvs[3 * i + 1]), int(recvs[3 * i + 2])) for i in range(nr.value) if recvs[3 * i] == r]
            ok &= to_me == from_r
            for (o, c) in from_r:                   # and it lies inside r's block, outside mine
                ok &= b[r] <= o and o + c <= b[r + 1]
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_partition_and_exchange_plans_are_consistent_across_two_ranks():
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out[0] and out[1]
