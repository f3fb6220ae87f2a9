"""CPU: the product's wire/journal codec (gpx_wire.cpp) against the independent layout
restatement (oracle/wire_oracle.py), the committed golden vectors and hand-derived bytes of
the reference's own codec round-trip test (BatchedAcceptReply.java:220-240)."""
import json
import os
import sys

import numpy as np
import pytest

from helpers import ROOT, abi

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import wire_oracle as wo  # noqa: E402

from gigapaxos_b200 import wire  # noqa: E402

GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "wire_vectors.json")))

REQ1 = wire.Request("NoopPaxosApp0", 0, 0x0123456789ABCDEF, b"a", entry_replica=100, entry_time=1700000000000)
REQ2 = wire.Request("pid1", 3, 42, b"STOP_REQUEST", stop=True, entry_replica=101, entry_time=1,
                    client=("127.0.0.1", 2000), listen=("10.0.0.7", 50000), should_return=True, forward_count=2,
                    response=b"ok")
B1 = wire.Request("g", 0, 7, b"xyz", entry_replica=102, entry_time=5)
B2 = wire.Request("g", 0, 8, b"", entry_replica=102, entry_time=6)
REQ3 = wire.Request("g", 0, 6, b"first", entry_replica=102, entry_time=4, batched=(B1, B2))


def test_reference_main_batched_accept_reply_bytes_by_hand():
    """AcceptReplyPacket(23, Ballot(0,234), 1, -1) + (.., 2, -1), paxosID "pid1" version 0
    (BatchedAcceptReply.java:224-231), bytes written out by hand from the layouts."""
    hand = ("0000005a" "00000022" "00000000" "04" + b"pid1".hex() +  # PaxosPacket header, type 34
            "00000017" "00000000" "000000ea" "00000001" "ffffffff" "0000000000000000" "00"  # AcceptReplyPacket part
            "00000002" "00000001" "0000000000000000" "00000002" "0000000000000000")  # 2 x (slot, requestID)
    got = wire.encode_batched_accept_reply("pid1", 0, 23, 0, 234, 1, -1, 0, [1, 2], [0, 0])
    assert got.hex() == hand == GOLD["batched_accept_reply_ref_main"]["hex"]
    d = wire.decode_batched_accept_reply(got)
    assert d["acceptor"] == 23 and (d["bnum"], d["bcoord"]) == (0, 234) and d["slots"] == [1, 2] and d["max_cp"] == -1
    assert d["paxos_id"] == "pid1"


def test_request_and_accept_goldens():
    assert wire.encode_request(REQ1).hex() == GOLD["request_simple"]["hex"]
    assert wire.encode_request(REQ2).hex() == GOLD["request_full"]["hex"]
    assert wire.encode_request(REQ3).hex() == GOLD["request_batched"]["hex"]
    a = wire.encode_accept(REQ1, 5, 0, 101, False, 3, 101)
    assert a.hex() == GOLD["accept_simple"]["hex"]
    assert a[3] == 90 and a[7] == 3  # SQLPaxosLogger.toBytes asserts bytes[3] == PAXOS_PACKET (:1090)
    # fixed part: 13+len(id) header, then SIZEOF_REQUEST_FIXED (:779-798) = 55 incl. the 4 length ints
    assert len(wire.encode_request(REQ1)) == 13 + len("NoopPaxosApp0") + 55 + 1
    assert len(a) == len(wire.encode_request(REQ1)) + 4 + 14 + 4  # SIZEOF_PROPOSAL + SIZEOF_PVALUE + SIZEOF_ACCEPT
    ab = wire.encode_accept(REQ3, -2, 7, 100, True, -1, 100)
    assert ab.hex() == GOLD["accept_batched"]["hex"]
    v = wire.decode_accept(ab)
    assert (v["slot"], v["bnum"], v["bcoord"], v["median_cp"], v["sender"], v["recovery"]) == (-2, 7, 100, -1, 100, 1)
    assert v["request_id"] == 6 and v["n_batched"] == 2 and v["paxos_id"] == "g" and v["value_len"] == 5
    with pytest.raises(ValueError):
        wire.decode_accept(ab[:-1])


def test_batched_commit_and_journal_goldens():
    bc = wire.encode_batched_commit("NoopPaxosApp12", 1, 3, 100, 8, [11, 9, 10], [100, 101, 102])
    assert bc.hex() == GOLD["batched_commit"]["hex"]
    d = wire.decode_batched_commit(bc)
    assert d["slots"] == [9, 10, 11] and d["group"] == [100, 101, 102] and d["median_cp"] == 8
    bar = wire.encode_batched_accept_reply("NoopPaxosApp12", 1, 102, 3, 100, 9, 8, 77, [9, 11, 10], [77, -5, 1 << 40])
    assert bar.hex() == GOLD["batched_accept_reply"]["hex"]
    assert wire.decode_batched_accept_reply(bar)["req_ids"] == [77, 1 << 40, -5]
    j = wire.journal_frame(wire.encode_accept(REQ1, 5, 0, 101, False, 3, 101))
    assert j.hex() == GOLD["journal_frame"]["hex"]
    assert int.from_bytes(j[:4], "big") == len(j) - 4


def test_codec_matches_layout_restatement_randomised():
    rng = np.random.default_rng(3)
    for _ in range(200):
        pid = "".join(chr(int(c)) for c in rng.integers(48, 123, size=int(rng.integers(1, 40))))
        val = bytes(rng.integers(0, 256, size=int(rng.integers(0, 300))).astype(np.uint8))
        rid = int(rng.integers(-(1 << 62), 1 << 62))
        ver = int(rng.integers(0, 1000))
        er = int(rng.integers(-1, 1 << 20))
        et = int(rng.integers(0, 1 << 50))
        stop = bool(rng.integers(0, 2))
        r = wire.Request(pid, ver, rid, val, stop=stop, entry_replica=er, entry_time=et)
        kw = dict(paxos_id=pid, version=ver, request_id=rid, stop=stop, value=val, entry_replica=er, entry_time=et)
        assert wire.encode_request(r) == wo.request(**kw)
        slot, bn, bc, mc = (int(x) for x in rng.integers(-(1 << 31), 1 << 31, size=4))
        assert wire.encode_accept(r, slot, bn, bc, False, mc, bc) == wo.accept(kw, slot, bn, bc, False, mc, bc)
        n = int(rng.integers(1, 20))
        slots = rng.integers(-50, 50, size=n)
        rids = rng.integers(-(1 << 62), 1 << 62, size=n)
        m = {}
        for s, q in zip(slots, rids):
            m[int(s)] = int(q)
        assert wire.encode_batched_accept_reply(pid, ver, 5, bn, bc, int(slots[0]), mc, rid, slots, rids) == \
            wo.batched_accept_reply(pid, ver, 5, bn, bc, int(slots[0]), mc, rid, m)
        assert wire.encode_batched_commit(pid, ver, bn, bc, mc, slots, [3, 1, 2]) == \
            wo.batched_commit(pid, ver, bn, bc, mc, [int(s) for s in slots], [3, 1, 2])


def test_fuse_commits_median_is_wraparound_max():
    """BatchedCommit.addCommit (:113-121): same (paxosID, ballot) fuse, medianCP = wrap-aware max."""
    d = np.zeros(6, dtype=abi.decision_dtype)
    d["gid"] = [1, 1, 1, 2, 2, 3]
    d["slot"] = [4, 5, 6, 9, 10, 1]
    d["bnum"] = [0, 0, 1, 0, 0, 0]
    d["bcoord"] = 100
    d["median_cp"] = [2, 3, 1, 2**31 - 1, -(2**31) + 5, 0]
    rs, rm = wire.fuse_commits(d)
    assert rs.tolist() == [0, 2, 3, 5]
    assert rm.tolist() == [3, 1, -(2**31) + 5, 0]


def test_reference_main_request_packet_with_25_batched():
    """RequestPacket.main (paxospackets/RequestPacket.java:1531-1563): a STOP request "asd999" with 25 STOP requests
    "asd0".."asd24" latched to it survives toBytes -> fromBytes.  Here: the product codec encodes it, the layout
    restatement produces the same bytes, and the ACCEPT that carries it decodes to the same header fields."""
    subs = [wire.Request("pid", 0, 1000 + i, f"asd{i}".encode(), stop=True, entry_replica=100, entry_time=i)
            for i in range(25)]
    req = wire.Request("pid", 0, 999, b"asd999", stop=True, entry_replica=100, entry_time=77, batched=tuple(subs))
    enc = wire.encode_request(req)
    kw = lambda r: dict(paxos_id=r.paxos_id, version=r.version, request_id=r.request_id, stop=r.stop, value=r.value,
                        entry_replica=r.entry_replica, entry_time=r.entry_time)
    assert enc == wo.request(**kw(req), batched=[kw(s) for s in subs])
    acc = wire.encode_accept(req, 12, 3, 100, False, 9, 100)
    v = wire.decode_accept(acc)
    assert v["n_batched"] == 25 and v["request_id"] == 999 and v["stop"] == 1 and v["value_len"] == 6
    assert (v["slot"], v["bnum"], v["bcoord"], v["median_cp"]) == (12, 3, 100, 9)
    # every latched request is length-prefixed inside the parent (RequestPacket.toBytes :935-945)
    assert enc.count(b"asd") == 26


# ---- property tests (hypothesis): encode -> decode is the identity on every field the decoder returns ---------------
from hypothesis import given, settings, strategies as st  # noqa: E402

i32 = st.integers(-(1 << 31), (1 << 31) - 1)
i64 = st.integers(-(1 << 63), (1 << 63) - 1)
pid = st.text(alphabet=st.characters(min_codepoint=33, max_codepoint=126), min_size=1, max_size=60)


@settings(max_examples=150, deadline=None)
@given(pid, st.integers(0, 1 << 20), i64, st.binary(max_size=400), st.booleans(), i32, i32, i32, i32, st.integers(0, 40))
def test_accept_roundtrip_property(paxos_id, version, rid, value, stop, slot, bnum, bcoord, median, nb):
    subs = tuple(wire.Request(paxos_id, version, rid ^ (k + 1), value[:k], stop=False, entry_replica=7, entry_time=k)
                 for k in range(nb))
    r = wire.Request(paxos_id, version, rid, value, stop=stop, entry_replica=bcoord, entry_time=12345, batched=subs)
    b = wire.encode_accept(r, slot, bnum, bcoord, False, median, bcoord)
    v = wire.decode_accept(b)
    assert (v["paxos_id"], v["version"], v["request_id"], v["slot"], v["bnum"], v["bcoord"], v["median_cp"],
            v["sender"], v["n_batched"], v["value_len"], bool(v["stop"])) == (
        paxos_id, version, rid, slot, bnum, bcoord, median, bcoord, nb, len(value), stop)
    kw = dict(paxos_id=paxos_id, version=version, request_id=rid, stop=stop, value=value, entry_replica=bcoord,
              entry_time=12345,
              batched=[dict(paxos_id=s.paxos_id, version=s.version, request_id=s.request_id, stop=s.stop, value=s.value,
                            entry_replica=s.entry_replica, entry_time=s.entry_time) for s in subs])
    assert b == wo.accept(kw, slot, bnum, bcoord, False, median, bcoord)
    j = wire.journal_frame(b)
    assert int.from_bytes(j[:4], "big", signed=True) == len(b) and j[4:] == b


@settings(max_examples=150, deadline=None)
@given(pid, st.integers(0, 1000), i32, i32, i32, st.lists(i32, min_size=1, max_size=64, unique=True),
       st.lists(i32, min_size=1, max_size=16, unique=True))
def test_batched_commit_roundtrip_property(paxos_id, version, bnum, bcoord, median, slots, group):
    b = wire.encode_batched_commit(paxos_id, version, bnum, bcoord, median, slots, group)
    d = wire.decode_batched_commit(b)
    assert sorted(d["slots"]) == sorted(slots) and d["group"] == group
    assert (d["bnum"], d["bcoord"], d["median_cp"]) == (bnum, bcoord, median)
    assert b == wo.batched_commit(paxos_id, version, bnum, bcoord, median, slots, group)
