"""Digest mode (DIGEST_REQUESTS, PaxosConfig.java:788; off by default): host-side mirror.

The entry replica broadcasts the request body to all replicas; the coordinator's ACCEPT then carries only
(requestID, MD5 of the request value) -- AcceptPacket.digest, paxospackets/AcceptPacket.java:162-170 -- and every
acceptor joins the two by requestID and checks the digest before handing the full ACCEPT to the acceptor logic:
paxosutil/PendingDigests.java:82-145 (`match` when the ACCEPT arrives, `release` when the request body arrives,
whichever comes second completes the pair).  The digests of a whole batch come from the engine's MD5 kernel
(`gpx_digest_requests`, RequestPacket.getDigest :1414-1430).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import abi


@dataclass
class DigestedAccept:
    """an ACCEPT without its request body: the 48-byte record (payload_len = 0) + the 16-byte digest"""
    rec: np.void          # abi.accept_dtype scalar; payload_off/payload_len are meaningless
    digest: bytes
    value_len: int        # length of the value the digest stands for (for the blob that undigest rebuilds)


def digest_accepts(engine: abi.Engine, accepts: np.ndarray, blob: np.ndarray) -> List[DigestedAccept]:
    """AcceptPacket.digest for a batch of single-request ACCEPTs: MD5 over each request value, on the device"""
    if len(accepts) == 0:
        return []
    if np.any(accepts["nreq"] > 1):
        raise ValueError("digest mode and request batching are alternatives (RequestPacket.batched stays undigested)")
    reqs = np.zeros(len(accepts), dtype=abi.request_dtype)
    reqs["gid"], reqs["req_id"] = accepts["gid"], accepts["req_id"]
    reqs["payload_off"], reqs["payload_len"] = accepts["payload_off"], accepts["payload_len"]
    dig = engine.digest_requests(reqs, blob)
    out = []
    for a, d in zip(accepts, dig):
        r = a.copy()
        ln = int(r["payload_len"])
        r["payload_off"], r["payload_len"] = 0, 0
        out.append(DigestedAccept(r, bytes(d), ln))
    return out


class PendingDigests:
    """paxosutil/PendingDigests.java: requests and digested ACCEPTs waiting for each other, keyed by requestID"""

    def __init__(self, md5=None):
        import hashlib
        self.requests: Dict[int, Tuple[int, bytes]] = {}       # requestID -> (gid, value): RequestAndCallback
        self.accepts: Dict[int, DigestedAccept] = {}
        self.anomalies = 0
        self._md5 = md5 or (lambda v: hashlib.md5(v).digest())

    def enqueue(self, gid: int, request_id: int, value: bytes):
        """the broadcast request body arrived (PaxosManager keeps it in `outstanding`)"""
        self.requests[request_id] = (gid, bytes(value))

    def _undigest(self, acc: DigestedAccept, value: bytes):
        """AcceptPacket.undigest :172-178"""
        r = acc.rec.copy()
        r["payload_len"] = len(value)
        return r, value

    def match(self, acc: DigestedAccept):
        """PendingDigests.match :82-101: the ACCEPT arrived; returns (accept record, value) if the request is here and
        its digest agrees, else parks the ACCEPT and returns None"""
        rid = int(acc.rec["req_id"])
        rc = self.requests.get(rid)
        if rc is None:
            self.accepts[rid] = acc
            return None
        gid, value = rc
        if gid == int(acc.rec["gid"]):
            if self._md5(value) == acc.digest:
                return self._undigest(acc, value)
            self.anomalies += 1  # logAnomaly: mismatched digests for matching requestIDs
        return None

    def release(self, gid: int, request_id: int, value: bytes, remove: bool = True):
        """PendingDigests.release :106-135: the request body arrived; returns the parked ACCEPT completed with it"""
        acc = self.accepts.get(request_id)
        if acc is not None and int(acc.rec["gid"]) == gid:
            if self._md5(bytes(value)) == acc.digest:
                if remove:
                    del self.accepts[request_id]
                return self._undigest(acc, bytes(value))
            self.anomalies += 1
        return None


def assemble(pairs) -> Tuple[np.ndarray, np.ndarray]:
    """(accept record, value) pairs -> an ACCEPT batch with its blob (16-byte aligned values), grouped by gid"""
    pairs = sorted(pairs, key=lambda p: int(p[0]["gid"]))
    recs = np.zeros(len(pairs), dtype=abi.accept_dtype)
    blob = bytearray()
    for k, (r, v) in enumerate(pairs):
        recs[k] = r
        recs[k]["payload_off"], recs[k]["payload_len"] = len(blob), len(v)
        blob += v + bytes(-len(v) % 16)
    return recs, np.frombuffer(bytes(blob), dtype=np.uint8)
