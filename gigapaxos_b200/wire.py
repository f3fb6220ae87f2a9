"""ctypes bindings of include/gpx_wire.h: the reference's big-endian wire / journal byte codecs
(paxospackets/RequestPacket.java:819-1024, AcceptPacket.java:95-138, BatchedAcceptReply.java:103-173,
BatchedCommit.java:184-252, SQLPaxosLogger.java:1000-1003)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import abi, load_library

PT_REQUEST, PT_ACCEPT, PT_BATCHED_ACCEPT_REPLY, PT_BATCHED_COMMIT = 1, 3, 34, 35


class WireRequest(C.Structure):
    pass


WireRequest._fields_ = [
    ("paxos_id", C.c_char_p), ("paxos_id_len", C.c_uint32), ("version", C.c_int32), ("request_id", C.c_int64),
    ("stop", C.c_uint8), ("client_ip", C.c_uint8 * 4), ("client_port", C.c_uint16), ("listen_ip", C.c_uint8 * 4),
    ("listen_port", C.c_uint16), ("entry_replica", C.c_int32), ("entry_time", C.c_int64),
    ("should_return_request_value", C.c_uint8), ("forward_count", C.c_int32), ("broadcasted", C.c_uint8),
    ("digest", C.c_char_p), ("digest_len", C.c_uint32), ("value", C.c_char_p), ("value_len", C.c_uint32),
    ("response", C.c_char_p), ("response_len", C.c_uint32), ("n_batched", C.c_uint32),
    ("batched", C.POINTER(WireRequest)),
]


class AcceptView(C.Structure):
    _fields_ = [("packet_type", C.c_int32), ("version", C.c_int32), ("paxos_id", C.c_void_p),
                ("paxos_id_len", C.c_uint32), ("request_id", C.c_int64), ("stop", C.c_uint8),
                ("entry_replica", C.c_int32), ("entry_time", C.c_int64), ("value", C.c_void_p),
                ("value_len", C.c_uint32), ("n_batched", C.c_uint32), ("request_bytes", C.c_size_t),
                ("slot", C.c_int32), ("bnum", C.c_int32), ("bcoord", C.c_int32), ("median_cp", C.c_int32),
                ("sender", C.c_int32), ("recovery", C.c_uint8)]


@dataclass
class Request:
    """RequestPacket fields (paxospackets/RequestPacket.java:55-)."""
    paxos_id: str
    version: int
    request_id: int
    value: bytes
    stop: bool = False
    entry_replica: int = -1
    entry_time: int = 0
    client: Optional[tuple] = None  # (ip, port)
    listen: Optional[tuple] = None
    should_return: bool = False
    forward_count: int = 0
    broadcasted: bool = False
    digest: Optional[bytes] = None
    response: bytes = b""
    batched: Sequence["Request"] = field(default_factory=tuple)

    def _c(self, keep: list) -> WireRequest:
        w = WireRequest()
        pid = self.paxos_id.encode("iso-8859-1")
        keep.append(pid)
        w.paxos_id, w.paxos_id_len, w.version, w.request_id = pid, len(pid), self.version, self.request_id
        w.stop = 1 if self.stop else 0
        for name, a in (("client", self.client), ("listen", self.listen)):
            if a is not None:
                ip = [int(x) for x in a[0].split(".")]
                for i in range(4):
                    getattr(w, name + "_ip")[i] = ip[i]
                setattr(w, name + "_port", a[1])
        w.entry_replica, w.entry_time = self.entry_replica, self.entry_time
        w.should_return_request_value = 1 if self.should_return else 0
        w.forward_count, w.broadcasted = self.forward_count, 1 if self.broadcasted else 0
        if self.digest:
            keep.append(self.digest)
            w.digest, w.digest_len = self.digest, len(self.digest)
        keep.append(self.value)
        w.value, w.value_len = self.value, len(self.value)
        keep.append(self.response)
        w.response, w.response_len = self.response, len(self.response)
        if self.batched:
            arr = (WireRequest * len(self.batched))(*[b._c(keep) for b in self.batched])
            keep.append(arr)
            w.n_batched, w.batched = len(self.batched), C.cast(arr, C.POINTER(WireRequest))
        return w


def _lib():
    L = load_library().lib
    L.gpx_wire_encode_request.restype = C.c_size_t
    L.gpx_wire_encode_accept.restype = C.c_size_t
    L.gpx_wire_request_size.restype = C.c_size_t
    L.gpx_wire_encode_batched_accept_reply.restype = C.c_size_t
    L.gpx_wire_encode_batched_commit.restype = C.c_size_t
    L.gpx_wire_journal_frame.restype = C.c_size_t
    return L


def encode_request(r: Request, packet_type: int = PT_REQUEST) -> bytes:
    L, keep = _lib(), []
    w = r._c(keep)
    n = L.gpx_wire_request_size(C.byref(w))
    buf = C.create_string_buffer(n)
    got = L.gpx_wire_encode_request(C.byref(w), C.c_int32(packet_type), buf, C.c_size_t(n))
    assert got == n
    return buf.raw


def encode_accept(r: Request, slot, bnum, bcoord, recovery, median_cp, sender) -> bytes:
    L, keep = _lib(), []
    w = r._c(keep)
    n = L.gpx_wire_request_size(C.byref(w)) + 22
    buf = C.create_string_buffer(n)
    got = L.gpx_wire_encode_accept(C.byref(w), C.c_int32(slot), C.c_int32(bnum), C.c_int32(bcoord),
                                   C.c_uint8(1 if recovery else 0), C.c_int32(median_cp), C.c_int32(sender), buf,
                                   C.c_size_t(n))
    assert got == n
    return buf.raw


def decode_accept(b: bytes) -> dict:
    L = _lib()
    v = AcceptView()
    rc = L.gpx_wire_decode_accept(b, C.c_size_t(len(b)), C.byref(v))
    if rc != 0:
        raise ValueError("malformed ACCEPT")
    out = {k: getattr(v, k) for k in ("packet_type", "version", "request_id", "stop", "entry_replica", "entry_time",
                                      "value_len", "n_batched", "request_bytes", "slot", "bnum", "bcoord",
                                      "median_cp", "sender", "recovery")}
    out["paxos_id"] = b[13: 13 + v.paxos_id_len].decode("iso-8859-1")
    # requestValue: after the header and the 39 fixed bytes comes {int digestLen, digest}{int valueLen, value}
    off = 13 + v.paxos_id_len + 39
    dl = int.from_bytes(b[off: off + 4], "big")
    off += 4 + dl
    out["value"] = b[off + 4: off + 4 + v.value_len]
    return out


def encode_batched_accept_reply(paxos_id, version, acceptor, bnum, bcoord, slot_number, max_cp, request_id,
                                slots, req_ids) -> bytes:
    L = _lib()
    pid = paxos_id.encode("iso-8859-1")
    s = np.ascontiguousarray(slots, dtype=np.int32)
    q = np.ascontiguousarray(req_ids, dtype=np.int64)
    cap = 13 + len(pid) + 29 + 4 + 12 * len(s)
    buf = C.create_string_buffer(cap)
    n = L.gpx_wire_encode_batched_accept_reply(pid, C.c_uint32(len(pid)), C.c_int32(version), C.c_int32(acceptor),
                                               C.c_int32(bnum), C.c_int32(bcoord), C.c_int32(slot_number),
                                               C.c_int32(max_cp), C.c_int64(request_id), C.c_uint32(len(s)),
                                               s.ctypes.data_as(C.c_void_p), q.ctypes.data_as(C.c_void_p), buf,
                                               C.c_size_t(cap))
    assert n > 0
    return buf.raw[:n]


def decode_batched_accept_reply(b: bytes) -> dict:
    L = _lib()
    ver, acc, bn, bc, sn, mc, n = (C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32(),
                                   C.c_uint32())
    pid = C.create_string_buffer(128)
    pl = C.c_uint32()
    cap = max(1, len(b) // 12)
    slots = np.zeros(cap, np.int32)
    rids = np.zeros(cap, np.int64)
    rc = L.gpx_wire_decode_batched_accept_reply(b, C.c_size_t(len(b)), C.byref(ver), pid, C.byref(pl), C.byref(acc),
                                                C.byref(bn), C.byref(bc), C.byref(sn), C.byref(mc), C.byref(n),
                                                slots.ctypes.data_as(C.c_void_p), rids.ctypes.data_as(C.c_void_p),
                                                C.c_uint32(cap))
    if rc != 0:
        raise ValueError("malformed BATCHED_ACCEPT_REPLY")
    return dict(paxos_id=pid.raw[: pl.value].decode("iso-8859-1"), version=ver.value, acceptor=acc.value,
                bnum=bn.value, bcoord=bc.value, slot_number=sn.value, max_cp=mc.value,
                slots=slots[: n.value].tolist(), req_ids=rids[: n.value].tolist())


def encode_batched_commit(paxos_id, version, bnum, bcoord, median_cp, slots, group) -> bytes:
    L = _lib()
    pid = paxos_id.encode("iso-8859-1")
    s = np.ascontiguousarray(slots, dtype=np.int32)
    g = np.ascontiguousarray(group, dtype=np.int32)
    cap = 13 + len(pid) + 12 + 4 * (len(s) + len(g) + 2)
    buf = C.create_string_buffer(cap)
    n = L.gpx_wire_encode_batched_commit(pid, C.c_uint32(len(pid)), C.c_int32(version), C.c_int32(bnum),
                                         C.c_int32(bcoord), C.c_int32(median_cp), C.c_uint32(len(s)),
                                         s.ctypes.data_as(C.c_void_p), C.c_uint32(len(g)),
                                         g.ctypes.data_as(C.c_void_p), buf, C.c_size_t(cap))
    assert n > 0
    return buf.raw[:n]


def decode_batched_commit(b: bytes) -> dict:
    L = _lib()
    ver, bn, bc, mc, ns, ng = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32(), C.c_uint32(), C.c_uint32()
    pid = C.create_string_buffer(128)
    pl = C.c_uint32()
    cap = max(1, len(b) // 4)
    slots = np.zeros(cap, np.int32)
    group = np.zeros(cap, np.int32)
    rc = L.gpx_wire_decode_batched_commit(b, C.c_size_t(len(b)), C.byref(ver), pid, C.byref(pl), C.byref(bn),
                                          C.byref(bc), C.byref(mc), C.byref(ns), slots.ctypes.data_as(C.c_void_p),
                                          C.c_uint32(cap), C.byref(ng), group.ctypes.data_as(C.c_void_p),
                                          C.c_uint32(cap))
    if rc != 0:
        raise ValueError("malformed BATCHED_COMMIT")
    return dict(paxos_id=pid.raw[: pl.value].decode("iso-8859-1"), version=ver.value, bnum=bn.value, bcoord=bc.value,
                median_cp=mc.value, slots=slots[: ns.value].tolist(), group=group[: ng.value].tolist())


def journal_frame(packet: bytes) -> bytes:
    L = _lib()
    buf = C.create_string_buffer(len(packet) + 4)
    n = L.gpx_wire_journal_frame(packet, C.c_size_t(len(packet)), buf, C.c_size_t(len(packet) + 4))
    assert n == len(packet) + 4
    return buf.raw


def fuse_commits(decs: np.ndarray):
    """PaxosPacketBatcher.fuseBatchedCommits: (run_start, run_median_cp) of same-(gid,ballot) runs."""
    L = _lib()
    L.gpx_wire_fuse_commits.restype = C.c_uint32
    d = np.ascontiguousarray(decs, dtype=abi.decision_dtype)
    rs = np.zeros(max(len(d), 1), np.uint32)
    rm = np.zeros(max(len(d), 1), np.int32)
    k = L.gpx_wire_fuse_commits(C.c_uint32(len(d)), d.ctypes.data_as(C.c_void_p), rs.ctypes.data_as(C.c_void_p),
                                rm.ctypes.data_as(C.c_void_p))
    return rs[:k].copy(), rm[:k].copy()
