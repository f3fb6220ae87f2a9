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
