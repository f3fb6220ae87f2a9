"""The stringified inter-replica packets and the cross-group packet batcher: ACCEPT_REPLY singletons,
BATCHED_ACCEPT (digest mode), BATCHED_PAXOS_PACKET.  The expected JSON objects are written out by hand from the
reference's toJSONObjectImpl methods (key names: paxospackets/PaxosPacket.java:45-200), not produced by the module."""
import hashlib
import json

import numpy as np
import pytest

from gigapaxos_b200 import abi, journal, packets_json as pj


def test_accept_reply_singleton_matches_the_reference_keys():
    # AcceptReplyPacket.toJSONObjectImpl :196-206 over PaxosPacket.toJSONObject :478-494
    b = pj.accept_reply_json("paxos0", 3, acceptor=101, bnum=7, bcoord=102, slot_number=55, max_checkpointed_slot=40,
                             request_id=-123456789012)
    assert json.loads(b) == {"type": 90, "PT": 8, "ID": "paxos0", "V": 3, "SNDR": 101, "B": "7:102", "S": 55, "CP_S": 40,
                             "QID": -123456789012}
    u = pj.accept_reply_json("paxos0", 3, 101, 7, 102, 55, 40, 9, undigest_request=True)
    assert json.loads(u)["NACK"] is True and "NACK" not in json.loads(b)
    v = pj.parse_packet(u)
    assert v == {"kind": "ACCEPT_REPLY", "paxos_id": "paxos0", "version": 3, "acceptor": 101, "bnum": 7, "bcoord": 102,
                 "slot_number": 55, "max_checkpointed_slot": 40, "request_id": 9, "undigest_request": True}


def test_nacks_among_engine_replies_become_singletons():
    r = np.zeros(4, dtype=abi.reply_dtype)
    r["gid"], r["slot"], r["bnum"], r["bcoord"], r["max_cp"], r["req_id"] = [0, 0, 1, 1], [5, 5, 9, 9], [2, 4, 1, 1], \
        [100, 102, 100, 100], [0, 3, 0, 0], [11, 11, 12, 12]
    # lane 0 acks; lane 1 answers with its higher ballot 4:102; group 1: one ack, one VOID hole
    r["who"] = [0 | (0 << 8), 1 | (0 << 8) | (abi.F_NACK << 16), 0, 1 | (abi.F_VOID << 16)]
    out = pj.accept_replies_to_packets(r, {0: ("g0", 0), 1: ("g1", 2)}, node_of_lane=[100, 101, 102])
    assert len(out) == 1
    assert json.loads(out[0]) == {"type": 90, "PT": 8, "ID": "g0", "V": 0, "SNDR": 101, "B": "4:102", "S": 5, "CP_S": 3,
                                  "QID": 11}


def test_batched_accept_json_and_merge_rule():
    d1, d2, d3 = (hashlib.md5(x).digest() for x in (b"a", b"b", b"c"))
    b = pj.BatchedAccept("name", 1, 3, 100, median_cp=10, group=[102, 100, 101])
    b.add_accept(7, d1, 1001, median_cp=10).add_accept(-2, d2, 1002, median_cp=12)
    j = json.loads(b.to_json())
    # BatchedAccept.toJSONObjectImpl :99-146: TreeMap order = natural signed order, digest = new String(bytes, ISO-8859-1)
    assert j == {"type": 90, "PT": 36, "ID": "name", "V": 1, "B": "3:100", "GC_S": 12, "GROUP": [100, 101, 102],
                 "S_DIGS": [[-2, d2.decode("iso-8859-1")], [7, d1.decode("iso-8859-1")]],
                 "S_QIDS": [[-2, 1002], [7, 1001]]}
    assert b.to_json().isascii()  # digest bytes >= 0x80 travel as \u00XX escapes: charset-neutral
    other = pj.BatchedAccept("name", 1, 3, 100, median_cp=11, group=[100, 101, 102]).add_accept(7, d3, 1003, 11)
    assert b.add_batched_accept(other) and b.median_cp == 12  # 11 is not ahead of 12
    assert b.slot_digests[7] == d3 and b.slot_request_ids[7] == 1003  # putAll: the later one wins
    wrap = pj.BatchedAccept("name", 1, 3, 100, median_cp=-(1 << 31) + 5, group=[100]).add_accept(8, d1, 1, -(1 << 31) + 5)
    big = pj.BatchedAccept("name", 1, 3, 100, median_cp=(1 << 31) - 5, group=[100])
    big.add_batched_accept(wrap)
    assert big.median_cp == -(1 << 31) + 5  # wrap-aware: a - b > 0 in int arithmetic (BatchedAccept.java:202)
    with pytest.raises(RuntimeError):
        b.add_batched_accept(pj.BatchedAccept("name", 1, 4, 100, 0, [100]))
    back = pj.BatchedAccept.from_json(b.to_json())
    assert (back.slot_digests, back.slot_request_ids, back.group, back.median_cp) == \
        (b.slot_digests, b.slot_request_ids, b.group, b.median_cp)
    # a frame written by a Java node holds the digest characters as raw ISO-8859-1 bytes
    raw = json.dumps(json.loads(b.to_json()), ensure_ascii=False).encode("iso-8859-1")
    assert not raw.isascii()
    assert pj.BatchedAccept.from_json(raw).slot_digests == b.slot_digests


def test_batched_accept_to_engine_records_and_back():
    from gigapaxos_b200.digests import DigestedAccept
    accs = []
    for gid, slot, req in [(0, 1, 50), (0, 2, 51), (1, 1, 60)]:
        r = np.zeros((), dtype=abi.accept_dtype)
        r["gid"], r["slot"], r["bnum"], r["bcoord"], r["median_cp"], r["req_id"], r["nreq"] = gid, slot, 0, 100, slot - 1, req, 1
        accs.append(DigestedAccept(r, hashlib.md5(b"%d" % req).digest(), 4))
    bs = pj.batch_digested_accepts(accs, {0: ("g0", 0), 1: ("g1", 0)}, {0: [100, 101, 102], 1: [101, 102, 103]})
    assert [(b.paxos_id, b.slots()) for b in bs] == [("g0", [1, 2]), ("g1", [1])]
    assert bs[0].median_cp == 1
    back = bs[0].to_digested_accepts(gid=0, sender_lane=0, value_len=4)
    assert [int(a.rec["slot"]) for a in back] == [1, 2] and [int(a.rec["req_id"]) for a in back] == [50, 51]
    assert [a.digest for a in back] == [accs[0].digest, accs[1].digest]


def test_packet_batcher_groups_by_recipient_set_in_first_seen_order():
    dec = journal.decision_json("g0", 0, 5, 1, 100, 3, 77, 100, 0)
    prep = journal.prepare_json("g1", 0, 2, 101, 4)
    nack = pj.accept_reply_json("g2", 0, 101, 9, 102, 1, 0, 5)
    byteified = b"\x00\x00\x00\x5a" + b"rest-of-a-byteified-accept"
    tasks = [([101, 102], [dec]), ([100], [nack]), ([102, 101], [prep]), ([100], [byteified]), ([], [dec]), ([103], [])]
    # more than MIN_PP_BATCH_SIZE tasks: grouped (PaxosPacketBatcher.process :270-277)
    out = pj.batch_messaging_tasks(tasks)
    assert [r for r, _ in out] == [[101, 102], [100]]
    (_, w0), (_, w1) = out
    assert len(w0) == 1
    j = json.loads(w0[0])
    assert j["type"] == 90 and j["PT"] == 37 and j["V"] == -1 and "ID" not in j  # BatchedPaxosPacket(null): no paxosID
    assert j["PP"] == [json.loads(dec), json.loads(prep)]
    assert w1[0] == byteified and json.loads(w1[1])["PP"] == [json.loads(nack)]  # Byteable packets are not wrapped
    assert [p["PT"] for p in pj.unbatch(w0[0])] == [6, 2] and pj.unbatch(nack) == [json.loads(nack)]
    kinds = [p["kind"] for p in pj.parse_packet(w0[0])["packets"]]
    assert kinds == ["DECISION", "PREPARE"]
    # at or below MIN_PP_BATCH_SIZE tasks: sent as they are
    out = pj.batch_messaging_tasks(tasks[:3])
    assert out == [([101, 102], [dec]), ([100], [nack]), ([102, 101], [prep])]
    assert pj.batch_messaging_tasks(tasks, batch_across_groups=False)[0] == ([101, 102], [dec])


def test_journal_frames_are_charset_neutral():
    # a paxosID with a non-ASCII character: the frame is pure ASCII, and a Java-style raw ISO-8859-1 frame parses too
    b = journal.decision_json("café", 0, 1, 0, 100, 0, 9, 100, 0)
    assert b.isascii() and journal.parse_packet(b)["paxos_id"] == "café"
    raw = json.dumps(json.loads(b), ensure_ascii=False).encode("iso-8859-1")
    assert journal.parse_packet(raw)["paxos_id"] == "café"


# ---- PREPARE / PREPARE_REPLY (phase 1 between nodes; the input format of gpx_handle_prepare_replies) ------------------------
def test_prepare_reply_packets_round_trip_and_elect(oracle_lib):
    """acceptors' reply records -> the PREPARE_REPLY JSON a Java coordinator reads (PrepareReplyPacket.toJSONObjectImpl
    :131-143) -> back to records + a payload arena -> gpx_handle_prepare_replies: the same election result as from the
    original records, and the request bodies survive (incl. a batched slot)"""
    import json
    import numpy as np
    from gigapaxos_b200 import abi, packets_json as pj
    from helpers import Engine, group_descs, make_config, make_requests
    NODES = [100, 101, 102]
    G = 6
    eng = Engine(oracle_lib, make_config(oracle_lib, max_groups=G, max_batch_recs=4096, max_batch_payload=1 << 20))
    eng.create_groups(group_descs(G))
    gids = np.arange(G, dtype=np.uint32)
    for r in range(2):
        reqs, pay = make_requests(gids, payload_len=4, seed=3, round_no=r)
        eng.round(reqs, pay)
    rows0 = eng.dump_rows(gids, 0)
    coord = np.array([NODES.index(int(x)) for x in rows0["acc_bcoord"]])
    # two more slots reach only some acceptors; the second is a batched slot (two requests of the group in one call)
    for k, (reach, per) in enumerate(((0b011, 1), (0b110, 2))):
        g2 = np.repeat(gids, per)
        reqs, pay = make_requests(g2, payload_len=5 + 3 * k, seed=4, round_no=k)
        reqs["flags"] = coord[g2] << 8
        reqs["entry_node"] = np.array(NODES)[coord[g2]]
        acc, blob, st = eng.propose(reqs, pay)
        assert len(acc) == G and (per == 1 or np.all(acc["nreq"] == 2))
        acc["dst_mask"] = reach
        eng.handle_accepts(acc, blob)
    cand = 2
    cur = eng.dump_rows(gids, cand)
    prep = np.zeros(G, dtype=abi.decision_dtype)
    prep["gid"], prep["slot"], prep["bnum"], prep["bcoord"] = gids, cur["acc_slot"], cur["acc_bnum"] + 1, NODES[cand]
    prep["flags"], prep["dst_mask"] = abi.F_PREPARE, 0b111
    replies = eng.handle_prepares(prep)
    names = {int(g): (f"NoopPaxosApp{int(g)}", 0) for g in gids}
    read_body = lambda lane, frame_ref, n: bytes(eng.log_read(lane, frame_ref * 16, n)) if n else b""
    pkts = pj.prepare_replies_to_packets(replies, names, NODES, read_body)
    assert len(pkts) == 3 * G
    j = json.loads(pkts[1].decode("ascii"))
    assert set(j) == {"type", "PT", "ID", "V", "ACCPTR", "B", "ACC_MAP", "PREPLY_MIN", "MAX_S", "MIN_S", "TOT_S", "CT"}
    assert j["type"] == 90 and j["PT"] == 7 and j["TOT_S"] == len(j["ACC_MAP"])
    some = [a for p in pkts for a in json.loads(p.decode("ascii"))["ACC_MAP"]]
    assert some and all({"type", "PT", "ID", "V", "B", "GC_S", "S", "QID", "QV", "ET", "E"} <= set(a) for a in some)
    assert any("BATCH" in a and len(a["BATCH"]) == 1 for a in some)  # the batched slot: one request latched along
    assert pj.parse_packet(pkts[0])["kind"] == "PREPARE_REPLY"
    # back to records: one arena per packet, concatenated; frame_ref rebased
    recs, arena = [], bytearray()
    for k, p in enumerate(pkts):
        r, a = pj.prepare_reply_to_records(p, int(replies[k]["gid"]), NODES, NODES[cand])
        for rec in r:
            for i in range(int(rec["n_accepted"])):
                rec["accepted"][i]["frame_ref"] += len(arena) // 16
        recs.append(r)
        arena += a
    recs = np.concatenate(recs)
    assert len(recs) == len(replies)
    for f in ("gid", "first_slot", "bnum", "bcoord", "n_accepted"):
        assert np.array_equal(recs[f], replies[f]), f
    assert np.array_equal(recs["who"] & 0xFFFF, replies["who"] & 0xFFFF)  # acceptor and preparer indices
    for k, (a, b) in enumerate(zip(recs, replies)):
        for i in range(int(b["n_accepted"])):
            pa, pb = a["accepted"][i], b["accepted"][i]
            for f in ("slot", "bnum", "bcoord", "req_id", "payload_len", "flags"):
                assert pa[f] == pb[f], f
            n, nreq = int(pb["payload_len"]), int(pb["flags"]) >> 16
            got = bytes(arena[int(pa["frame_ref"]) * 16: int(pa["frame_ref"]) * 16 + n])
            orig = read_body(k % 3, int(pb["frame_ref"]), n)  # the body in the log ring of the acceptor that replied
            if nreq <= 1:
                assert got == orig
            else:  # (the entry-lane bits of a batched slot's table travel as "E", not in the blob)
                ea = np.frombuffer(got[: 16 * nreq], dtype=abi.batch_ent_dtype)
                eb = np.frombuffer(orig[: 16 * nreq], dtype=abi.batch_ent_dtype)
                assert np.array_equal(ea["req_id"], eb["req_id"]) and np.array_equal(ea["len"], eb["len"])
                assert got[16 * nreq:] == orig[16 * nreq:]
    els = np.zeros(G, dtype=abi.election_dtype)
    els["gid"], els["lane"], els["bnum"], els["bcoord"], els["slot"] = gids, cand, prep["bnum"], NODES[cand], prep["slot"]
    els["first_reply"], els["n_replies"] = np.arange(G) * 3, 3
    twin = Engine(oracle_lib, make_config(oracle_lib, max_groups=G, max_batch_recs=4096, max_batch_payload=1 << 20))
    twin.create_groups(group_descs(G))
    o1 = eng.handle_prepare_replies(els, replies)
    o2 = twin.handle_prepare_replies(els, recs)
    assert np.all(o1["verdict"] == abi.EL_MAJORITY) and int(o1["n_plan"].max()) >= 1
    for f in ("verdict", "next_slot", "n_plan", "flags", "node_slots"):
        assert np.array_equal(o1[f], o2[f]), f
    for a, b in zip(o1, o2):
        for i in range(int(a["n_plan"])):
            assert a["plan"][i]["slot"] == b["plan"][i]["slot"] and a["plan"][i]["kind"] == b["plan"][i]["kind"]
            assert a["plan"][i]["pv"]["req_id"] == b["plan"][i]["pv"]["req_id"]


def test_prepare_reply_longer_than_the_window_is_continued():
    from gigapaxos_b200 import abi, packets_json as pj
    import numpy as np
    acc = []
    for s in range(11):
        pv = np.zeros((), dtype=abi.accepted_pvalue_dtype)
        pv["slot"], pv["bnum"], pv["bcoord"], pv["req_id"], pv["flags"] = 20 + s, 1, 100, 900 + s, 1 << 16
        acc.append(pj.accepted_pvalue_obj("p", 0, pv, b"v%d" % s, entry_replica=100))
    pkt = pj.prepare_reply_json("p", 0, 101, 2, 102, 19, acc)
    d = pj.parse_packet(pkt)
    assert (d["first_slot"], d["min_slot"], d["max_slot"], d["total_count"]) == (20, 20, 30, 11)
    recs, arena = pj.prepare_reply_to_records(pkt, 5, [100, 101, 102], 102)
    assert len(recs) == 2 and [int(r["n_accepted"]) for r in recs] == [8, 3]
    assert abi.who_flags(int(recs[0]["who"])) & abi.F_MORE and not abi.who_flags(int(recs[1]["who"])) & abi.F_MORE
    assert abi.who_acc(int(recs[0]["who"])) == 1 and abi.who_dst(int(recs[0]["who"])) == 2
    assert [int(r["first_slot"]) for r in recs] == [19, 19] and len(arena) == 11 * 16
    assert pj.parse_packet(pj.prepare_packet_json("p", 0, 2, 102, 20)) == {
        "kind": "PREPARE", "paxos_id": "p", "version": 0, "bnum": 2, "bcoord": 102, "first_undecided_slot": 20}


def test_sync_decisions_packet():
    from gigapaxos_b200 import packets_json as pj
    import json
    p = pj.sync_decisions_json("paxos0", 2, 101, 17, [12, 13, 15])
    assert json.loads(p) == {"type": 90, "PT": 32, "ID": "paxos0", "V": 2, "SNDR": 101, "MAX_S": 17, "MISS": [12, 13, 15]}
    assert pj.parse_packet(p) == {"kind": "SYNC_DECISIONS", "paxos_id": "paxos0", "version": 2, "node": 101,
                                  "max_decision_slot": 17, "missing": [12, 13, 15]}
    assert "MISS" not in json.loads(pj.sync_decisions_json("p", 0, 1, 3, []))  # SyncDecisionsPacket.toJSONObjectImpl :84-86
