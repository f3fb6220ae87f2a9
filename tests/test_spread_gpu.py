"""GPU parity of spread placement (replicas of a group in different engines, records exchanged between them).

N single-lane engines -- here all on one GPU, exchanging through device copies (`LocalExchange`); the NCCL
exchange of the multi-GPU run moves the same buffers -- against ONE oracle engine that hosts all N nodes as
lanes and is driven phase by phase (propose, all ACCEPTs, all replies, all DECISIONs).  Per node: request
status, the EXEC sequence, every state row and the non-VOID log images + logged payload bytes must be identical.
"""
import numpy as np
import pytest

from helpers import Engine, abi, canon, make_config, make_requests
from gigapaxos_b200.spread import LocalExchange, SpreadCluster, SpreadNode, coordinator_of, members_of

pytestmark = pytest.mark.gpu

NODE0 = 100


def build(oracle_lib, cuda_lib, N, G, R=3, **cfg):
    import torch
    node_ids = [NODE0 + i for i in range(N)]
    names = [f"NoopPaxosApp{i}" for i in range(G)]
    descs = np.zeros(G, dtype=abi.group_desc_dtype)
    descs["gid"] = np.arange(G)
    descs["n_members"] = R
    descs["init_mode"] = abi.INIT_BATCH
    coord = np.zeros(G, dtype=np.int64)
    member_of = np.zeros((G, N), dtype=bool)
    for g, nm in enumerate(names):
        mem = [node_ids[m] for m in members_of(nm, N, R)]
        descs["name_hash"][g] = abi.java_string_hash(nm)
        descs["members"][g, :R] = mem
        coord[g] = coordinator_of(nm, mem) - NODE0
        member_of[g, [m - NODE0 for m in mem]] = True
    eo = Engine(oracle_lib, make_config(oracle_lib, max_groups=G, n_lanes=N, lane_node=node_ids, max_group_size=R,
                                        max_batch_recs=4 * G, max_batch_payload=1 << 22, **cfg))
    eo.create_groups(descs)
    dev = torch.device("cuda", 0)
    nodes = []
    for i in range(N):
        nd = SpreadNode(cuda_lib, i, node_ids, dev, max_groups=G, max_batch=4 * G, max_payload=1 << 22,
                        max_group_size=R, **cfg)
        nd.engine.create_groups(descs[member_of[:, i]])
        nodes.append(nd)
    return eo, SpreadCluster(nodes, LocalExchange(N), N), descs, coord, member_of


def oracle_round(eo, reqs, pay, N):
    """the same round on the oracle: all nodes are lanes of one engine, phase by phase"""
    acc, blob, status = eo.propose(reqs, pay)
    replies, x1 = eo.handle_accepts(acc, blob, extra_cap=4 * len(acc) + 16)
    dec = eo.handle_accept_replies(replies)
    ex, x2 = eo.handle_decisions(dec, extra_cap=4 * len(acc) + 16)
    return status, ex, np.concatenate([x1, x2])


def by_gid(recs):
    r = recs[(recs["flags"] & abi.F_VOID) == 0]
    return r[np.argsort(r["gid"], kind="stable")]


@pytest.mark.parametrize("N,G,P", [(4, 600, 1), (3, 200, 40), (5, 333, 17), (8, 900, 5), (4, 600, 17),
                                   (5, 100, 17)])
def test_spread_round_parity(oracle_lib, cuda_lib, N, G, P):
    import torch
    eo, cluster, descs, coord, member_of = build(oracle_lib, cuda_lib, N, G, checkpoint_interval=3)
    dev = cluster.nodes[0].device
    rng = np.random.default_rng(5)
    for r in range(5):
        # some groups get several requests (one batched slot), some none
        counts = rng.choice([0, 1, 1, 1, 2, 3], size=G)
        gids = np.repeat(np.arange(G), counts)
        lens = rng.integers(1, P + 1, size=len(gids))
        reqs, pay = make_requests(gids, payload_len=lens, seed=31, round_no=r)
        reqs["flags"] = (coord[gids].astype(np.uint32) << 8)  # entry lane (oracle) = the coordinator's lane
        reqs["entry_node"] = NODE0 + coord[gids]
        so, xo, extra_o = oracle_round(eo, reqs, pay, N)
        # the same requests, split by coordinator node; node engines have a single lane -> entry lane 0
        batches, index = {}, {}
        for i in range(N):
            sel = np.nonzero(coord[gids] == i)[0]
            index[i] = sel
            if len(sel) == 0:
                continue
            rq = reqs[sel].copy()
            rq["flags"] &= ~np.uint32(0xF00)
            # re-pack this node's payloads (16-byte aligned like make_requests)
            stride = ((rq["payload_len"] + 15) // 16) * 16
            offs = np.concatenate([[0], np.cumsum(stride)[:-1]]).astype(np.uint32)
            buf = np.zeros(int(stride.sum()), dtype=np.uint8)
            for k in range(len(rq)):
                o, ln = int(reqs["payload_off"][sel[k]]), int(rq["payload_len"][k])
                buf[offs[k]: offs[k] + ln] = pay[o: o + ln]
            rq["payload_off"] = offs
            batches[i] = (torch.from_numpy(rq.view(np.uint8).copy()).to(dev), torch.from_numpy(buf).to(dev), len(rq))
        res = cluster.round(batches, extra_cap=4096)
        torch.cuda.synchronize()
        assert len(extra_o) == 0
        for i in range(N):
            s = res[i]
            if len(index[i]):
                assert np.array_equal(s["status"].cpu().numpy()[: len(index[i])], so[index[i]]), f"status node {i}"
            assert s["n_extra"] == 0
            ex = s["exec"].cpu().numpy().view(abi.exec_dtype)[: s["n_exec"]]
            lanes_o = (xo["flags"] >> 12) & 0xF
            want = by_gid(xo[lanes_o == i])
            got = by_gid(ex)
            assert len(want) == len(got) > 0, (i, len(want), len(got), {k: v for k, v in s.items() if k.startswith("n")},
                                               cluster.nodes[i].engine.counters())
            for f in ("gid", "slot", "req_id"):
                assert np.array_equal(want[f], got[f]), f"node {i} exec {f}"
            # flags: STOP/CKPT + nreq equal; the lane nibble is the lane inside the engine (0 on the node)
            assert np.array_equal(want["flags"] & ~np.uint32(0xF000), got["flags"] & ~np.uint32(0xF000))
    # state: every row of every node
    for i, nd in enumerate(cluster.nodes):
        g = np.nonzero(member_of[:, i])[0]
        ro, rg = eo.dump_rows(g, i), nd.engine.dump_rows(g, 0)
        for f in ro.dtype.names:
            if f != "lane":
                assert np.array_equal(ro[f], rg[f]), f"node {i} row field {f}"
    # logs: non-VOID images and logged payload bytes, segment by segment
    for i, nd in enumerate(cluster.nodes):
        so_ = [s for s in abi.parse_log(eo.log_read(i)) if int(s[0]["n_valid"]) > 0 or len(canon(s[1])) > 0]
        sg_ = [s for s in abi.parse_log(nd.engine.log_read(0)) if len(canon(s[1])) > 0]
        so_ = [s for s in so_ if len(canon(s[1])) > 0]
        assert len(so_) == len(sg_) > 0
        for (ho, io_, po, _), (hg, ig, pg, _) in zip(so_, sg_):
            assert int(ho["type"]) == int(hg["type"])
            co, cg = canon(io_), canon(ig)
            assert len(co) == len(cg)
            for f in co.dtype.names:
                if f not in ("payload_off", "dst_mask"):
                    assert np.array_equal(co[f], cg[f]), f"node {i} log image {f}"
            if int(ho["rec_bytes"]) == 48:
                for a, b in zip(co, cg):
                    ao, bo, ln = int(a["payload_off"]), int(b["payload_off"]), int(a["payload_len"])
                    x, y = po[ao: ao + ln].copy(), pg[bo: bo + ln].copy()
                    nreq = int(a["nreq"])
                    if nreq > 1:  # batched blob: the per-request flags carry the ENTRY LANE, a per-engine notion
                        for t in (x, y):
                            t[: 16 * nreq].view(abi.batch_ent_dtype)["flags"] &= ~np.uint32(0xF00)
                    if not np.array_equal(x, y):
                        pgb = pg.tobytes()
                        where = pgb.find(x.tobytes())
                        raise AssertionError(("log payload", i, int(a["gid"]), int(coord[int(a["gid"])]), nreq, ln, ao, bo,
                                              x.tolist(), y.tolist(), where, int(hg["n_slots"]), int(hg["payload_bytes"]),
                                              [(int(q["gid"]), int(q["payload_off"]), int(q["payload_len"])) for q in cg][:12]))
    # counters: the nodes together did what the oracle's lanes did
    co = eo.counters()
    tot = {k: 0 for k in co}
    for nd in cluster.nodes:
        for k, v in nd.engine.counters().items():
            tot[k] += v
    for k in ("accepts_handled", "accepts_acked", "accepts_logged", "replies_handled", "decisions_made",
              "decisions_handled", "executed", "checkpoints_due", "proposals", "requests_batched"):
        assert co[k] == tot[k], k


# ---- one process per GPU over NCCL (needs >= 3 GPUs: skipped on single-GPU boxes) ------------------------------
def _nccl_worker(rank, world, port, q, G, rounds):
    import os
    import traceback
    import torch
    import torch.distributed as dist
    try:
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        os.environ["NCCL_DEBUG"] = "WARN"
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        import gigapaxos_b200
        from gigapaxos_b200.spread import DistExchange
        from helpers import oracle_library
        from test_spread_gloo import by_gid as by_gid2, make_groups, node_batch, workload
        R = 3
        node_ids, descs, coord, member_of = make_groups(world, G, R)
        lib = gigapaxos_b200.load_library()
        nd = SpreadNode(lib, rank, node_ids, dev, max_groups=G, max_batch=8 * G, max_payload=1 << 22, max_group_size=R,
                        checkpoint_interval=3)
        nd.engine.create_groups(descs[member_of[:, rank]])
        cluster = SpreadCluster([nd], DistExchange(), world)
        olib = oracle_library()
        ref = Engine(olib, make_config(olib, max_groups=G, n_lanes=world, lane_node=node_ids, max_group_size=R,
                                       max_batch_recs=8 * G, max_batch_payload=1 << 22, checkpoint_interval=3))
        ref.create_groups(descs)
        n_exec = 0
        for r in range(rounds):
            gids, reqs, pay = workload(G, coord, r)
            acc, blob, so = ref.propose(reqs, pay)
            rep, _ = ref.handle_accepts(acc, blob)
            dec = ref.handle_accept_replies(rep)
            xo, _ = ref.handle_decisions(dec)
            sel = np.nonzero(coord[gids] == rank)[0]
            batches = {}
            if len(sel):
                a, b, n = node_batch(reqs, pay, sel)
                batches[rank] = (a.to(dev), b.to(dev), n)
            s = cluster.round(batches)[rank]
            torch.cuda.synchronize()
            if len(sel):
                assert np.array_equal(s["status"].cpu().numpy()[: len(sel)], so[sel])
            got = by_gid2(s["exec"].cpu().numpy().view(abi.exec_dtype)[: s["n_exec"]])
            want = by_gid2(xo[((xo["flags"] >> 12) & 0xF) == rank])
            assert len(got) == len(want)
            for f in ("gid", "slot", "req_id"):
                assert np.array_equal(got[f], want[f]), f
            n_exec += len(got)
        g = np.nonzero(member_of[:, rank])[0]
        ro, rg = ref.dump_rows(g, rank), nd.engine.dump_rows(g, 0)
        for f in ro.dtype.names:
            if f != "lane":
                assert np.array_equal(ro[f], rg[f]), f
        dist.barrier()
        q.put((rank, "ok", n_exec))
        dist.destroy_process_group()
    except Exception:  # pragma: no cover
        q.put((rank, "fail: " + traceback.format_exc(), 0))
        raise


def test_spread_over_nccl():
    import torch
    import torch.multiprocessing as mp
    from test_spread_gloo import free_port
    world = min(torch.cuda.device_count(), 4)
    if world < 3:
        pytest.skip("spread placement over NCCL needs >= 3 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, q, 500, 4)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, status, n_exec in res:
        assert status == "ok", f"rank {rank}: {status}"
        assert n_exec > 0
