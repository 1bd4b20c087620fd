"""CPU, world_size 2, gloo: host logic of the independent-ensemble mode (member dealing, execution on the
owning rank only, rank-0 gather in member order).  The member runner is a stub -- the CUDA driver is
exercised by the GPU tests."""
import os
import socket

import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from fenics_b200 import ensemble
from fenics_b200.drivers import jbp


def test_sweep_matches_reference_loops():
    s = jbp.sweep()
    # 4 We values x 3 Ma values, Re = re_row[2] = 25, lambda_D = lambda_row[1] = 0 (fenepmp_jbp.py:291-297)
    assert len(s) == 12
    assert {m["Re"] for m in s} == {25.0} and {m["lambda_d"] for m in s} == {0.0}
    assert [m["We"] for m in s[:4]] == [0.001, 0.1, 0.5, 1.0] and [m["Ma"] for m in s[::4]] == [0.01, 0.05, 0.1]
    assert [(m["j"], m["jjj"]) for m in s[:5]] == [(1, 0), (2, 0), (3, 0), (4, 0), (1, 1)]


def test_assign_round_robin_partitions_members():
    members = list(range(12))
    shares = [ensemble.assign(members, r, 8) for r in range(8)]
    assert sorted(i for sh in shares for i, _ in sh) == members
    assert [len(sh) for sh in shares] == [2, 2, 2, 2, 1, 1, 1, 1]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ran = []

    def run_member(i, m):
        ran.append(i)
        return dict(index=i, We=m["We"], Ma=m["Ma"], rank=rank, chi=m["We"] * 10 + m["Ma"])

    res = ensemble.run_ensemble(jbp.sweep(), run_member)
    q.put((rank, ran, res))
    dist.barrier()
    dist.destroy_process_group()


def test_run_ensemble_world2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    out = dict()
    for _ in range(world):
        rank, ran, res = q.get(timeout=120)
        out[rank] = (ran, res)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert out[0][0] == list(range(0, 12, 2)) and out[1][0] == list(range(1, 12, 2))  # each member ran once
    assert out[1][1] is None
    res = out[0][1]
    assert [r["index"] for r in res] == list(range(12))
    assert [r["rank"] for r in res] == [i % 2 for i in range(12)]
    sw = jbp.sweep()
    assert all(abs(r["chi"] - (m["We"] * 10 + m["Ma"])) < 1e-15 for r, m in zip(res, sw))
