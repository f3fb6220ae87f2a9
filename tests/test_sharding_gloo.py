"""CPU, world_size 2 over gloo: the N>1 host logic.  Every rank owns the groups whose
home_gpu is its rank (no data-path collective); the union of the per-rank results must equal a
single-engine run of all groups, and the max-over-ranks timing reduction works."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gigapaxos_b200 import shard
from gigapaxos_b200.paxos_manager import HashChainApp, PaxosManager
from helpers import Engine, abi, make_config, oracle_library

NODES = [100, 101, 102]
NAMES = [f"NoopPaxosApp{i}" for i in range(60)]


def run_shard(names, rounds=5):
    lib = oracle_library()
    eng = Engine(lib, make_config(lib, max_groups=128, max_batch_recs=4096, max_batch_payload=1 << 20))
    pm = PaxosManager(eng, [HashChainApp() for _ in NODES], NODES)
    pm.createPaxosInstanceBatch({n: None for n in names}, NODES)
    for r in range(rounds):
        for n in names:
            pm.propose(n, f"{n}:{r}".encode())
        pm.run_round()
    return {n: pm.apps[0].state[n].hex() for n in names}, pm.num_decisions


def worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard.shard_names(NAMES, world)[rank]
    states, nd = run_shard(mine)
    t = torch.tensor([float(nd), float(rank + 1)], dtype=torch.float64)
    tot = t.clone()
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    mx = t.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)  # the bench's max-over-ranks reduction
    dist.barrier()
    q.put((rank, states, nd, tot.tolist(), mx.tolist()))
    dist.destroy_process_group()


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_home_gpu_is_the_reference_hash():
    for n in NAMES[:10]:
        h = abi.java_string_hash(n)
        assert shard.home_gpu(n, 8) == abs(h) % 8
    assert shard.replica_gpus("NoopPaxosApp1", 3, 8, packed=True) == [shard.home_gpu("NoopPaxosApp1", 8)] * 3
    g = shard.replica_gpus("NoopPaxosApp1", 3, 8, packed=False)
    assert g == [(g[0] + j) % 8 for j in range(3)]
    parts = shard.shard_names(NAMES, 4)
    assert sorted(sum(parts.values(), [])) == sorted(NAMES)


def test_two_rank_sharded_run_equals_single_engine():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    single, nd_single = run_shard(NAMES)
    merged = {}
    for rank, states, nd, tot, mx in res:
        assert set(states) == set(shard.shard_names(NAMES, world)[rank])
        merged.update(states)
        assert tot[0] == nd_single and mx[1] == world
    assert merged == single
