"""wire_oracle.py -- TEST INFRASTRUCTURE: independent restatement (python struct) of the reference's
big-endian byte codecs, used to check gigapaxos_b200/csrc/gpx_wire.cpp and to generate
tests/golden/wire_vectors.json (oracle/make_wire_golden.py).

Layouts (paths under /root/reference/src/edu/umass/cs/gigapaxos/paxospackets/):
  PaxosPacket.toBytes           PaxosPacket.java:459-476
  RequestPacket.toBytes         RequestPacket.java:819-949 (fixed part :779-798)
  AcceptPacket.toBytes          AcceptPacket.java:95-138 (SIZEOF_PROPOSAL 4, SIZEOF_PVALUE 4+4+1+4+1, SIZEOF_ACCEPT 4)
  AcceptReplyPacket.toBytes     AcceptReplyPacket.java:174-184 (SIZEOF_ACCEPTREPLY 4+8+4+4+8+1)
  BatchedAcceptReply.toBytes    BatchedAcceptReply.java:120-173
  BatchedCommit.toBytes         BatchedCommit.java:184-252
  journal frame                 ../SQLPaxosLogger.java:1000-1003
The reference JVM cannot run in this image, so expected bytes are derived from these layouts by
hand, not produced by the reference ("parity pinned by layout reading").
"""
import struct

PAXOS_PACKET, REQUEST, ACCEPT, BATCHED_ACCEPT_REPLY, BATCHED_COMMIT = 90, 1, 3, 34, 35


def header(ptype: int, version: int, paxos_id: str) -> bytes:
    pid = paxos_id.encode("iso-8859-1")
    return struct.pack(">iiiB", PAXOS_PACKET, ptype, version, len(pid)) + pid


def request(paxos_id, version, request_id, stop, value: bytes, entry_replica, entry_time, client=None, listen=None,
            should_return=False, forward_count=0, broadcasted=False, digest=None, response=b"", batched=(),
            ptype=REQUEST) -> bytes:
    def addr(a):
        if a is None:
            return bytes(4) + struct.pack(">H", 0)
        ip, port = a
        return bytes(int(x) for x in ip.split(".")) + struct.pack(">H", port)
    b = header(ptype, version, paxos_id)
    b += struct.pack(">qB", request_id, 1 if stop else 0)
    b += addr(client) + addr(listen)
    b += struct.pack(">iqBi", entry_replica, entry_time, 1 if should_return else 0, forward_count)
    b += struct.pack(">B", 1 if broadcasted else 0)
    b += struct.pack(">i", len(digest) if digest else 0) + (digest or b"")
    b += struct.pack(">i", len(value)) + value
    b += struct.pack(">i", len(response)) + response
    b += struct.pack(">i", len(batched))
    for r in batched:
        rb = request(**r)
        b += struct.pack(">i", len(rb)) + rb
    return b


def accept(req_kwargs: dict, slot, bnum, bcoord, recovery, median_cp, sender) -> bytes:
    kw = dict(req_kwargs)
    kw["ptype"] = ACCEPT
    return request(**kw) + struct.pack(">iiiBiBi", slot, bnum, bcoord, 1 if recovery else 0, median_cp, 0, sender)


def batched_accept_reply(paxos_id, version, acceptor, bnum, bcoord, slot_number, max_cp, request_id, slots: dict) -> bytes:
    b = header(BATCHED_ACCEPT_REPLY, version, paxos_id)
    b += struct.pack(">iiiiiqB", acceptor, bnum, bcoord, slot_number, max_cp, request_id, 0)
    b += struct.pack(">i", len(slots))
    for s in sorted(slots):
        b += struct.pack(">iq", s, slots[s])
    return b


def batched_commit(paxos_id, version, bnum, bcoord, median_cp, slots, group) -> bytes:
    b = header(BATCHED_COMMIT, version, paxos_id)
    b += struct.pack(">iii", bnum, bcoord, median_cp)
    ss = sorted(set(slots))
    b += struct.pack(">i", len(ss)) + b"".join(struct.pack(">i", s) for s in ss)
    b += struct.pack(">i", len(group)) + b"".join(struct.pack(">i", g) for g in group)
    return b


def journal_frame(packet: bytes) -> bytes:
    return struct.pack(">i", len(packet)) + packet
