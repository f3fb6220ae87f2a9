"""CPU, world_size 3 over gloo: a view change between NODES -- phase 1 as packets.

Every rank is one node of a spread placement (a single-lane engine hosting one member of every group; the CPU oracle
behind the C ABI here, the CUDA engine on a GPU box).  The candidate's PREPAREs travel as the JSON of PreparePacket, every
acceptor answers from its own engine (gpx_handle_prepares) and ships its reply as the JSON of PrepareReplyPacket with the
accepted pvalues' request bodies read from ITS log ring (gigapaxos_b200/packets_json.py); the candidate turns the packets
back into reply records + a payload arena and elects all groups with ONE gpx_handle_prepare_replies call.  The result
must be what a single-process engine hosting all three nodes as lanes computes from the same history."""
import os

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from test_spread_gloo import free_port

NODES = [100, 101, 102]
G, R = 24, 3


def history(seed=5):
    """the same ACCEPT history on every rank: two slots per group under ballot (0, coordinator), reaching a majority /
    a minority of the acceptors; the second slot of some groups is a batched one.  -> [(accept records, blob, reach mask)]"""
    from helpers import Engine, abi, group_descs, make_config, make_requests, oracle_library
    lib = oracle_library()
    eng = Engine(lib, make_config(lib, max_groups=G, max_batch_recs=4096, max_batch_payload=1 << 20))
    eng.create_groups(group_descs(G, members=tuple(NODES)))
    gids = np.arange(G, dtype=np.uint32)
    rows0 = eng.dump_rows(gids, 0)
    coord = np.array([NODES.index(int(x)) for x in rows0["acc_bcoord"]])
    out = []
    rng = np.random.default_rng(seed)
    for k, reach in enumerate((0b011, 0b101, 0b100)):
        per = np.where(rng.random(G) < 0.3, 2, 1) if k == 1 else np.ones(G, dtype=int)
        g2 = np.repeat(gids, per)
        reqs, pay = make_requests(g2, payload_len=6 + k, seed=seed, round_no=k)
        reqs["flags"] = coord[g2] << 8
        reqs["entry_node"] = np.array(NODES)[coord[g2]]
        acc, blob, st = eng.propose(reqs, pay)
        assert len(acc) == G
        out.append((acc.copy(), blob.copy(), reach))
    return eng, out, coord


def worker(rank, world, port, q):
    try:
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from gigapaxos_b200 import abi, packets_json as pj
        from helpers import Engine, group_descs, make_config, oracle_library
        lib = oracle_library()
        me = NODES[rank]
        eng = Engine(lib, make_config(lib, n_lanes=1, lane_node=[me], max_group_size=R, max_groups=G, max_batch_recs=4096,
                                      max_batch_payload=1 << 20))
        eng.create_groups(group_descs(G, members=tuple(NODES)))
        _, hist, coord = history()
        for acc, blob, reach in hist:  # the ACCEPTs that reached this node
            if reach & (1 << rank):
                a = acc.copy()
                a["dst_mask"] = 1
                eng.handle_accepts(a, blob)
        gids = np.arange(G, dtype=np.uint32)
        names = {int(g): (f"NoopPaxosApp{int(g)}", 0) for g in gids}
        cand = 2  # node 102 runs for coordinator of every group
        # ---- PREPARE: candidate -> all (PreparePacket JSON)
        if rank == cand:
            cur = eng.dump_rows(gids, 0)
            preps = [pj.prepare_packet_json(names[int(g)][0], 0, int(cur[i]["acc_bnum"]) + 1, me, int(cur[i]["acc_slot"]))
                     for i, g in enumerate(gids)]
        else:
            preps = None
        box = [preps]
        dist.broadcast_object_list(box, src=cand)
        preps = [pj.parse_packet(p) for p in box[0]]
        prep = np.zeros(G, dtype=abi.decision_dtype)
        prep["gid"] = gids
        prep["slot"] = [p["first_undecided_slot"] for p in preps]
        prep["bnum"], prep["bcoord"] = [p["bnum"] for p in preps], [p["bcoord"] for p in preps]
        prep["flags"], prep["dst_mask"] = abi.F_PREPARE, 1
        # ---- PREPARE_REPLY: every acceptor -> candidate (PrepareReplyPacket JSON, bodies from its own ring)
        replies = eng.handle_prepares(prep)
        read_body = lambda lane, frame_ref, n: bytes(eng.log_read(0, frame_ref * 16, n)) if n else b""
        mine = pj.prepare_replies_to_packets(replies, names, [me], read_body)
        assert len(mine) == G
        allp = [None] * world if rank == cand else None
        dist.gather_object(mine, allp, dst=cand)
        if rank == cand:
            recs, arena, first = [], bytearray(), []
            for i, g in enumerate(gids):
                first.append(sum(len(r) for r in recs))
                for src in range(world):  # in the order the packets "arrived": node 100, 101, 102
                    r, a = pj.prepare_reply_to_records(allp[src][i], int(g), NODES, me)
                    for rec in r:
                        for k in range(int(rec["n_accepted"])):
                            rec["accepted"][k]["frame_ref"] += len(arena) // 16
                    recs.append(r)
                    arena += a
            n_of = [sum(len(r) for r in recs[i * world:(i + 1) * world]) for i in range(G)]
            recs = np.concatenate(recs)
            els = np.zeros(G, dtype=abi.election_dtype)
            els["gid"], els["lane"], els["bnum"], els["bcoord"], els["slot"] = gids, 0, prep["bnum"], me, prep["slot"]
            els["first_reply"], els["n_replies"] = np.concatenate([[0], np.cumsum(n_of)[:-1]]), n_of
            out = eng.handle_prepare_replies(els, recs)
            # ---- the same history in ONE engine hosting the three nodes as lanes
            ref, hist2, _ = history()
            for acc, blob, reach in hist2:
                a = acc.copy()
                a["dst_mask"] = reach
                ref.handle_accepts(a, blob)
            prep3 = prep.copy()
            prep3["dst_mask"] = 0b111
            rep3 = ref.handle_prepares(prep3)
            els3 = els.copy()
            els3["lane"], els3["first_reply"], els3["n_replies"] = cand, np.arange(G) * 3, 3
            want = ref.handle_prepare_replies(els3, rep3)
            assert np.all(want["verdict"] == abi.EL_MAJORITY) and int(want["n_plan"].max()) >= 2
            for f in ("verdict", "next_slot", "n_plan", "flags", "node_slots"):
                assert np.array_equal(out[f], want[f]), f
            n_batched = 0
            for a, b in zip(out, want):
                for k in range(int(b["n_plan"])):
                    pa, pb = a["plan"][k], b["plan"][k]
                    assert pa["slot"] == pb["slot"] and pa["kind"] == pb["kind"]
                    if int(pb["kind"]) == abi.CO_PVALUE:
                        for f in ("slot", "bnum", "bcoord", "req_id", "payload_len", "flags"):
                            assert pa["pv"][f] == pb["pv"][f], f
                        # the body that came over the wire is the body the reporting acceptor holds
                        n = int(pb["pv"]["payload_len"])
                        got = bytes(arena[int(pa["pv"]["frame_ref"]) * 16: int(pa["pv"]["frame_ref"]) * 16 + n])
                        orig = bytes(ref.log_read(int(pb["src_reply"]) % 3, int(pb["pv"]["frame_ref"]) * 16, n))
                        nreq = int(pb["pv"]["flags"]) >> 16
                        if nreq <= 1:
                            assert got == orig
                        else:
                            n_batched += 1
                            assert got[16 * nreq:] == orig[16 * nreq:]
            assert n_batched > 0
            row = eng.dump_rows(gids, 0)
            assert np.all(row["coord_exists"] == 1) and np.all(row["coord_active"] == 1) and np.all(row["coord_bcoord"] == me)
            assert np.array_equal(row["next_proposal_slot"], want["next_slot"])
        dist.barrier()
        q.put((rank, "ok"))
        dist.destroy_process_group()
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, "fail: " + traceback.format_exc()))
        raise


def test_view_change_between_nodes_over_gloo():
    world = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res


# ---- catching up between nodes: SYNC_DECISIONS -> the donor's journal (gpx_log_find + gpx_log_gather) -> DECISIONs back ---
def sync_history(seed=8, rounds=11):
    """11 slots decided by nodes 100 and 101 while node 102 is cut off: -> (3-lane reference engine after them,
    [(accept records, blob, decision records)] per round)"""
    from helpers import Engine, abi, group_descs, make_config, make_requests, oracle_library
    lib = oracle_library()
    ref = Engine(lib, make_config(lib, max_groups=G, max_batch_recs=4096, max_batch_payload=1 << 20, checkpoint_interval=100))
    ref.create_groups(group_descs(G, members=tuple(NODES)))
    gids = np.arange(G, dtype=np.uint32)
    rows0 = ref.dump_rows(gids, 0)
    coord = np.array([NODES.index(int(x)) for x in rows0["acc_bcoord"]])
    ok = coord != 2  # groups node 102 coordinates cannot progress without it: leave them alone
    g2 = gids[ok]
    out = []
    for k in range(rounds):
        per = np.where(np.arange(len(g2)) % 5 == k % 5, 2, 1)
        gg = np.repeat(g2, per)
        reqs, pay = make_requests(gg, payload_len=5 + k % 6, seed=seed, round_no=k)
        reqs["flags"] = coord[gg] << 8
        reqs["entry_node"] = np.array(NODES)[coord[gg]]
        acc, blob, st = ref.propose(reqs, pay)
        acc["dst_mask"] = 0b011
        rep, _ = ref.handle_accepts(acc, blob)
        dec = ref.handle_accept_replies(rep)
        assert len(dec) == len(g2)
        dec["dst_mask"] = 0b011
        ref.handle_decisions(dec)
        out.append((acc.copy(), blob.copy(), dec.copy()))
    return ref, out, g2


def sync_worker(rank, world, port, q):
    try:
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from gigapaxos_b200 import abi, packets_json as pj
        from helpers import Engine, group_descs, make_config, oracle_library
        lib = oracle_library()
        me = NODES[rank]
        eng = Engine(lib, make_config(lib, n_lanes=1, lane_node=[me], max_group_size=R, max_groups=G, max_batch_recs=4096,
                                      max_batch_payload=1 << 20, checkpoint_interval=100))
        eng.create_groups(group_descs(G, members=tuple(NODES)))
        ref, hist, g2 = sync_history()
        executed = 0
        if rank != 2:  # nodes 100 and 101 accept and commit; 102 hears nothing
            for acc, blob, dec in hist:
                a, d = acc.copy(), dec.copy()
                a["dst_mask"], d["dst_mask"] = 1, 1
                eng.handle_accepts(a, blob)
                ex, extra = eng.handle_decisions(d)
                executed += int(((ex["flags"] & abi.F_VOID) == 0).sum()) + len(extra)
            assert executed == len(hist) * len(g2)
        names = {int(g): f"NoopPaxosApp{int(g)}" for g in g2}
        donor, lagging = 0, 2
        # ---- SYNC_DECISIONS: 102 -> 100, one packet per group (it knows of no decision: MISS = [its slot])
        if rank == lagging:
            rows = eng.dump_rows(g2, 0)
            reqs = [pj.sync_decisions_json(names[int(g)], 0, me, int(rows[i]["acc_slot"]) - 1, [int(rows[i]["acc_slot"])])
                    for i, g in enumerate(g2)]
        else:
            reqs = None
        box = [reqs]
        dist.broadcast_object_list(box, src=lagging)
        # ---- the donor looks its journal up (gpx_log_find + gpx_log_gather) and answers with DECISIONs
        if rank == donor:
            rows = eng.dump_rows(g2, 0)
            assert pj.parse_packet(box[0][0])["kind"] == "SYNC_DECISIONS"
            answer = [pj.serve_sync_request(eng, 0, int(g), box[0][i], int(rows[i]["acc_slot"]) - 1) for i, g in enumerate(g2)]
            assert all(len(a) == len(hist) for a in answer)
        else:
            answer = None
        box = [answer]
        dist.broadcast_object_list(box, src=donor)
        if rank == lagging:  # replay, W slots at a time: accept (the value), then commit
            W = int(eng.cfg.window)
            got_exec = 0
            for i, g in enumerate(g2):
                acc, blob, dec = pj.decisions_to_records(box[0][i], int(g), 0)
                for s in range(0, len(acc), W):
                    a = acc[s: s + W].copy()
                    lo = int(a["payload_off"].min())
                    hi = int((a["payload_off"] + (a["payload_len"] + 15) // 16 * 16).max())
                    a["payload_off"] -= lo
                    eng.handle_accepts(a, blob[lo:hi])
                    ex, extra = eng.handle_decisions(dec[s: s + W])
                    got_exec += int(((ex["flags"] & abi.F_VOID) == 0).sum()) + len(extra)
            assert got_exec == len(hist) * len(g2)
            mine, theirs = eng.dump_rows(g2, 0), ref.dump_rows(g2, 0)
            for f in ("acc_slot", "acc_bnum", "acc_bcoord"):
                assert np.array_equal(mine[f], theirs[f]), f
            # what was executed here is what the other nodes executed: the same request ids slot by slot
            w = np.zeros(len(g2), dtype=abi.log_want_dtype)
            w["gid"], w["min_slot"], w["n_slots"] = np.sort(g2), int(theirs["acc_slot"].min()) - len(hist), len(hist)
            a, b = eng.log_find(0, w), ref.log_find(0, w)
            assert np.array_equal(a["decision"]["req_id"], b["decision"]["req_id"])
            assert ((b["decision"]["flags"] & abi.F_VOID) == 0).sum() == len(hist) * len(g2)
        dist.barrier()
        q.put((rank, "ok"))
        dist.destroy_process_group()
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, "fail: " + traceback.format_exc()))
        raise


def test_sync_between_nodes_over_gloo():
    world = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=sync_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
