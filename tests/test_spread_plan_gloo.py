"""CPU, world_size 2 and 3 over gloo: the host arithmetic behind spread placement in libgpx (gpx_spread_plan_node,
include/gpx.h) -- every rank plans its own bucket arena from the SAME capacity matrix, and what rank s plans to send
rank d for a packet type must be exactly what rank d plans to receive from s (the sizes ncclSend / ncclRecv are called
with, and the offsets the peer-memory transport stores to).  No GPU: the planner is pure host code.
"""
import os

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from test_spread_gloo import free_port


def worker(rank, world, port, q, R):
    try:
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import gigapaxos_b200
        from gigapaxos_b200.spread import coordinator_of, members_of, spread_caps, spread_config, spread_plan
        lib = gigapaxos_b200.load_library()
        G, node_ids = 500, [100 + i for i in range(world)]
        coord = np.zeros(G, dtype=np.int64)
        member_of = np.zeros((G, world), dtype=bool)
        for g in range(G):
            nm = f"NoopPaxosApp{g}"
            mem = [node_ids[m] for m in members_of(nm, world, R)]
            coord[g] = coordinator_of(nm, mem) - 100
            member_of[g, [m - 100 for m in mem]] = True
        cap = spread_caps(coord, member_of, slack=1)
        cfg = spread_config(node_ids, cap, blob_per_rec=48, max_reqs=G)
        p = spread_plan(lib, cfg, rank)
        mine = {"send": [[int(p.send_bytes[k][d]) for d in range(world)] for k in range(3)],
                "recv": [[int(p.recv_bytes[k][s]) for s in range(world)] for k in range(3)],
                "recv_off": [[int(p.recv_off[k][s]) for s in range(world)] for k in range(3)],
                "vbase": [int(p.vbase[s]) for s in range(world)], "vtotal": int(p.vtotal), "arena": int(p.arena_bytes)}
        allp = [None] * world
        dist.all_gather_object(allp, mine)
        for k in range(3):
            for d in range(world):
                assert allp[rank]["send"][k][d] == allp[d]["recv"][k][rank], (k, rank, d)
                if d != rank and allp[rank]["send"][k][d]:
                    assert allp[d]["recv_off"][k][rank] + allp[d]["recv"][k][rank] <= allp[d]["arena"]
        # ACCEPT buckets carry records + blob, the other two records only; the loop-back bucket exists iff I coordinate
        for d in range(world):
            c = int(cap[rank, d])
            assert mine["send"][0][d] == (64 + c * 48 + c * 48 if c else 0)
            assert mine["send"][2][d] == (64 + c * 32 if c else 0)
            assert mine["send"][1][d] == (64 + int(cap[d, rank]) * 32 if cap[d, rank] else 0)
        # the receive side's virtual index space: 256-aligned, in source order, covering every capacity
        v = 0
        for s in range(world):
            assert mine["vbase"][s] == v and v % 256 == 0
            v += (int(cap[s, rank]) + 255) // 256 * 256
        assert mine["vtotal"] == v
        # every rank planned from the same matrix
        caps = [None] * world
        dist.all_gather_object(caps, cap.tolist())
        assert all(c == caps[0] for c in caps)
        dist.barrier()
        q.put((rank, "ok"))
        dist.destroy_process_group()
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, "fail: " + traceback.format_exc()))
        raise


@pytest.mark.parametrize("world,R", [(3, 3), (2, 2)])
def test_spread_plans_agree_over_gloo(world, R):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=worker, args=(r, world, port, q, R)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, status in res:
        assert status == "ok", f"rank {rank}: {status}"
