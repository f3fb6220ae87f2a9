"""Generates tests/golden/wire_vectors.json from oracle/wire_oracle.py (layout restatement).
Run from the repo root: python oracle/make_wire_golden.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import wire_oracle as w  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

req1 = dict(paxos_id="NoopPaxosApp0", version=0, request_id=0x0123456789ABCDEF, stop=False, value=b"a",
            entry_replica=100, entry_time=1700000000000)
req2 = dict(paxos_id="pid1", version=3, request_id=42, stop=True, value=b"STOP_REQUEST", entry_replica=101,
            entry_time=1, client=("127.0.0.1", 2000), listen=("10.0.0.7", 50000), should_return=True,
            forward_count=2, response=b"ok")
b1 = dict(paxos_id="g", version=0, request_id=7, stop=False, value=b"xyz", entry_replica=102, entry_time=5)
b2 = dict(paxos_id="g", version=0, request_id=8, stop=False, value=b"", entry_replica=102, entry_time=6)
req3 = dict(paxos_id="g", version=0, request_id=6, stop=False, value=b"first", entry_replica=102, entry_time=4,
            batched=(b1, b2))

vectors = {
    "request_simple": {"args": {k: (v.decode() if isinstance(v, bytes) else v) for k, v in req1.items()},
                       "hex": w.request(**req1).hex()},
    "request_full": {"hex": w.request(**req2).hex()},
    "request_batched": {"hex": w.request(**req3).hex()},
    "accept_simple": {"slot": 5, "bnum": 0, "bcoord": 101, "recovery": 0, "median_cp": 3, "sender": 101,
                      "hex": w.accept(req1, 5, 0, 101, False, 3, 101).hex()},
    "accept_batched": {"slot": -2, "bnum": 7, "bcoord": 100, "recovery": 1, "median_cp": -1, "sender": 100,
                       "hex": w.accept(req3, -2, 7, 100, True, -1, 100).hex()},
    # BatchedAcceptReply.main BatchedAcceptReply.java:220-240: acceptor 23, ballot 0:234, slots 1,2, "pid1", maxCP -1
    "batched_accept_reply_ref_main": {"hex": w.batched_accept_reply("pid1", 0, 23, 0, 234, 1, -1, 0, {1: 0, 2: 0}).hex()},
    "batched_accept_reply": {"hex": w.batched_accept_reply("NoopPaxosApp12", 1, 102, 3, 100, 9, 8, 77,
                                                           {9: 77, 11: -5, 10: 1 << 40}).hex()},
    "batched_commit": {"hex": w.batched_commit("NoopPaxosApp12", 1, 3, 100, 8, [11, 9, 10], [100, 101, 102]).hex()},
    "journal_frame": {"hex": w.journal_frame(w.accept(req1, 5, 0, 101, False, 3, 101)).hex()},
}
os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
json.dump(vectors, open(os.path.join(ROOT, "tests", "golden", "wire_vectors.json"), "w"), indent=1, sort_keys=True)
print("wrote", len(vectors), "vectors")
