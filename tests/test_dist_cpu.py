"""world_size-2 gloo tests of the multi-GPU host logic (mode A: independent meshes sharded over ranks, no data-path
collective; only barrier / max-over-ranks / result gather go through torch.distributed)."""
import os
import socket

import torch
import torch.multiprocessing as mp

from dual_dmp_b200 import dist as D


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_items, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    assert D.init(backend="gloo")
    mine = D.shard_items(n_items, rank, world)
    # every rank "fits" its meshes independently (stand-in: a deterministic per-mesh number) and times itself
    results = {i: float(i * i) for i in mine}
    t = D.max_over_ranks(10.0 + rank)
    D.barrier()
    everything = D.gather_results(results)
    q.put((rank, mine, t, everything))
    torch.distributed.destroy_process_group()


def test_shard_items_cover_exactly_once():
    for n, w in ((64, 1), (64, 2), (64, 8), (5, 4), (0, 2)):
        seen = sorted(i for r in range(w) for i in D.shard_items(n, r, w))
        assert seen == list(range(n))
        sizes = [len(D.shard_items(n, r, w)) for r in range(w)]
        assert max(sizes) - min(sizes) <= 1


def test_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port, world, n_items = _free_port(), 2, 7
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_items, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0][1] == [0, 2, 4, 6] and got[1][1] == [1, 3, 5]
    assert got[0][2] == got[1][2] == 11.0                      # max over ranks
    merged = {}
    for part in got[0][3]:
        merged.update(part)
    assert merged == {i: float(i * i) for i in range(n_items)} and got[0][3] == got[1][3]


def test_single_process_is_a_noop():
    os.environ.pop("WORLD_SIZE", None)
    assert D.init() is False
    assert D.max_over_ranks(3.5) == 3.5 and D.gather_results("x") == ["x"]
    D.barrier()
