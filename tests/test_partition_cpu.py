"""Partitioned mode (SURVEY.md §8e mode B): the host-side index logic, simulated without GPUs — every rank's local
CSR + halo exchange lists must reproduce the global aggregation exactly — and the all-gather autograd op under gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from dual_dmp_b200 import synth
from dual_dmp_b200.partition import PartitionPlan, split_bounds
from dual_dmp_b200.util.mesh import Mesh


def _graphs():
    vs, faces = synth.icosphere(7)
    m = Mesh(vs=vs, faces=faces)
    e = torch.from_numpy(m.edges.T.astype(np.int64))
    yield torch.cat([e, e[[1, 0]]], dim=1), len(vs), m.vs
    yield torch.from_numpy(m.f_edges), len(faces), m.fc
    vs, faces = synth.open_patch(6, 0.1)
    m = Mesh(vs=vs, faces=faces)
    yield torch.from_numpy(m.f_edges), len(faces), m.fc


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_local_parts_reproduce_global_aggregation(world):
    rng = np.random.RandomState(0)
    for ei, n, coords in _graphs():
        plan = PartitionPlan(ei, n, coords, world)
        assert plan.bounds[0] == 0 and plan.bounds[-1] == n
        H = rng.randn(n, 5).astype(np.float32)                         # rows in Morton order
        ref = np.zeros_like(H)
        np.add.at(ref, plan.row, plan.w[:, None] * H[plan.col])
        parts = [plan.local(r) for r in range(world)]
        covered = 0
        for r, p in enumerate(parts):
            # simulate the all-to-all: what rank r receives = the senders' packed rows, in source-rank order
            recv = []
            for q, pq in enumerate(parts):
                off = int(pq["send_counts"][:r].sum())
                cnt = int(pq["send_counts"][r])
                rows = pq["send_idx"][off:off + cnt].astype(np.int64) + pq["lo"]
                recv.append(H[rows])
                assert cnt == int(p["recv_counts"][q])
            H_ext = np.concatenate([H[p["lo"]:p["hi"]]] + recv, axis=0)
            assert H_ext.shape[0] == p["n_own"] + p["n_halo"]
            assert np.array_equal(H_ext[p["n_own"]:], H[p["halo"]])     # halo rows land where the local CSR expects them
            rows_local = np.repeat(np.arange(p["n_own"]), np.diff(p["rowptr"]))
            out = np.zeros((p["n_own"], 5), dtype=np.float32)
            np.add.at(out, rows_local, p["w"][:, None] * H_ext[p["col"]])
            assert np.allclose(out, ref[p["lo"]:p["hi"]], rtol=0, atol=1e-6)
            covered += p["n_own"]
            assert p["send_counts"][r] == 0 and p["recv_counts"][r] == 0
        assert covered == n
        if 1 < world <= 3:
            halo = sum(p["n_halo"] for p in parts)
            assert halo < 0.6 * n                                      # contiguous Morton ranges: boundary << interior


def test_split_bounds():
    assert split_bounds(10, 3).tolist() == [0, 3, 6, 10]
    assert split_bounds(8, 8).tolist() == list(range(9))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gather_worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from types import SimpleNamespace
    from dual_dmp_b200.partition import PartitionedGraph, _AllGatherRows
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 11
    perm = torch.from_numpy(np.random.RandomState(3).permutation(n))
    bounds = split_bounds(n, world)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    pg = SimpleNamespace(plan=SimpleNamespace(bounds=bounds), world=world, group=None, perm_all=perm,
                         own_ids=perm[lo:hi])
    pg.all_gather_equal = lambda mine: PartitionedGraph.all_gather_equal(pg, mine)     # the product's transport
    full_ref = torch.arange(n * 3, dtype=torch.float32).view(n, 3)            # caller numbering
    own = full_ref[perm[lo:hi]].clone().requires_grad_(True)
    full = _AllGatherRows.apply(own, pg)
    w = torch.arange(n, dtype=torch.float32).view(n, 1) + 1
    (full * w).sum().backward()
    ok = torch.equal(full, full_ref) and torch.equal(own.grad, w[perm[lo:hi]].expand(-1, 3))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_all_gather_rows_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]
