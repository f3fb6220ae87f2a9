"""Journal persistence: drain the engine's HBM log rings into files byte-compatible with the
reference's journal, and read them back.

Mirrors SQLPaxosLogger.Journaler (gigapaxos/SQLPaxosLogger.java:685-848): files
`<logdir>paxos_journal.<node>/log.<node>.<hex millis>`, contents `{int32 BE length}{packet
bytes}*` (journal() :1000-1003, appendToLogFile :814-826), a new file once the current one
exceeds MAX_LOG_FILE_SIZE (rollLogFile :789-812, checked after a batch as in journal() :1041).

ACCEPTs are written with the reference's byte codec (AcceptPacket.toBytes, what
SQLPaxosLogger.toBytes :1084-1096 journals for ACCEPT packets).  The reference journals
DECISIONs as JSON strings (JSON codecs are out of scope); this writer therefore
produces the journal of the reference's DONT_LOG_DECISIONS mode (:977-978) and keeps the
decision images in a side file `<journal file>.decisions` (raw 32-byte gpx_decision_rec images; logged
PREPAREs -- the promised ballots, GPX_F_PREPARE -- share the form and the file, the reference journals
those as JSON as well).
"""
from __future__ import annotations

import os
import struct
import time
from typing import Callable, Dict, Iterator, List, Tuple

import numpy as np

from . import abi, wire
from .abi import Engine

MAX_LOG_FILE_SIZE = 64 * 1024 * 1024  # PC.MAX_LOG_FILE_SIZE PaxosConfig.java:314


class Journaler:
    """SQLPaxosLogger.Journaler: append-only length-prefixed frames, rolled by size."""

    SUBDIR, PREFIX, POSTPREFIX = "paxos_journal.", "log.", "."

    def __init__(self, logdir: str, node_id, max_log_file_size: int = MAX_LOG_FILE_SIZE):
        self.node_id = node_id
        self.dir = os.path.join(logdir, f"{self.SUBDIR}{node_id}")
        os.makedirs(self.dir, exist_ok=True)
        self.max = max_log_file_size
        self.files: List[str] = []
        self._seq = 0
        self.fos = None
        self.dec = None
        self.cur_size = 0
        self._create()

    def _name(self) -> str:  # generateLogfileName :729-734 (USE_HEX_TIMESTAMP)
        ts = int(time.time() * 1000) + self._seq
        self._seq += 1
        return os.path.join(self.dir, f"{self.PREFIX}{self.node_id}{self.POSTPREFIX}{ts:x}")

    def _create(self):
        if self.fos:
            self.fos.flush()
            self.fos.close()
            self.dec.close()
        self.cur = self._name()
        self.fos = open(self.cur, "wb")
        self.dec = open(self.cur + ".decisions", "wb")
        self.files.append(self.cur)
        self.cur_size = 0

    def append(self, packet: bytes):
        frame = wire.journal_frame(packet)  # {int32 BE len}{bytes}
        self.fos.write(frame)
        self.cur_size += len(frame)

    def append_decisions(self, images: np.ndarray):
        self.dec.write(images.tobytes())

    def end_batch(self):
        self.fos.flush()  # FLUSH=true, SYNC=false (PaxosConfig.java:720,725)
        self.dec.flush()
        if self.cur_size > self.max:
            self._create()

    def close(self):
        if self.fos:
            self.fos.flush()
            self.fos.close()
            self.dec.close()
            self.fos = None


RequestLookup = Callable[[int, int], Tuple[str, int, int, float]]
"""(gid, req_id) -> (paxosID, version, entryReplica, entryTime): the RequestPacket fields the engine
does not carry (PaxosManager keeps them in its outstanding table)."""


class LogDrainer:
    """Drains one lane's log ring into a Journaler, incrementally (the BatchedLogger's role,
    AbstractPaxosLogger.java:691-716, with the device having done the append already)."""

    def __init__(self, engine: Engine, lane: int, journaler: Journaler, lookup: RequestLookup):
        self.engine, self.lane, self.j, self.lookup = engine, lane, journaler, lookup
        self.offset = 0
        self.accepts_written = 0
        self.decisions_written = 0

    def drain(self) -> int:
        head = self.engine.log_head(self.lane)
        if head <= self.offset:
            return 0
        buf = self.engine.log_read(self.lane, self.offset, head - self.offset)
        n = 0
        for hdr, imgs, payload, _ in abi.parse_log(buf):
            if int(hdr["rec_bytes"]) == 48:
                for a in imgs:
                    if a["flags"] & abi.F_VOID:
                        continue
                    self.j.append(self._accept_bytes(a, payload))
                    n += 1
                    self.accepts_written += 1
            else:  # DECISION images and logged PREPAREs (GPX_F_PREPARE: the promised ballot) share the 32-byte form
                logged = imgs[(imgs["flags"] & abi.F_VOID) == 0]
                if len(logged):
                    self.j.append_decisions(logged)
                    self.decisions_written += len(logged)
        self.j.end_batch()
        self.offset = head
        return n

    def _accept_bytes(self, a, payload: np.ndarray) -> bytes:
        off, ln, nreq = int(a["payload_off"]), int(a["payload_len"]), int(a["nreq"])
        blob = payload[off: off + ln].tobytes()
        pid, ver, entry, etime = self.lookup(int(a["gid"]), int(a["req_id"]))
        if nreq <= 1:
            req = wire.Request(pid, ver, int(a["req_id"]), blob, stop=bool(a["flags"] & abi.F_STOP),
                               entry_replica=entry, entry_time=int(etime))
        else:  # RequestPacket.batched: [nreq x {req_id,len,flags}][values]
            ents = np.frombuffer(blob[: 16 * nreq], dtype=abi.batch_ent_dtype)
            vals, p = [], 16 * nreq
            for e in ents:
                vals.append(blob[p: p + int(e["len"])])
                p += int(e["len"])
            subs = []
            for e, v in zip(ents[1:], vals[1:]):
                spid, sver, sentry, setime = self.lookup(int(a["gid"]), int(e["req_id"]))
                subs.append(wire.Request(spid, sver, int(e["req_id"]), v, stop=bool(int(e["flags"]) & abi.F_STOP),
                                         entry_replica=sentry, entry_time=int(setime)))
            req = wire.Request(pid, ver, int(ents[0]["req_id"]), vals[0], stop=bool(int(ents[0]["flags"]) & abi.F_STOP),
                               entry_replica=entry, entry_time=int(etime), batched=tuple(subs))
        return wire.encode_accept(req, int(a["slot"]), int(a["bnum"]), int(a["bcoord"]), False, int(a["median_cp"]),
                                  int(a["sender"]))


def read_journal(path: str) -> Iterator[bytes]:
    """The packets of one journal file (the reader side of {int32 BE len}{bytes}*)."""
    with open(path, "rb") as f:
        data = f.read()
    p = 0
    while p + 4 <= len(data):
        (ln,) = struct.unpack(">i", data[p: p + 4])
        if ln < 0 or p + 4 + ln > len(data):
            break  # torn tail: the reference tolerates a partially written last frame
        yield data[p + 4: p + 4 + ln]
        p += 4 + ln


def replay_accepts(files: List[str]) -> Dict[str, Dict[int, dict]]:
    """Recovery read path (the ACCEPT half of PaxosManager's roll forward, PaxosManager.java:1852-2055):
    per paxosID, the highest-ballot logged ACCEPT of every slot."""
    out: Dict[str, Dict[int, dict]] = {}
    for path in files:
        for pkt in read_journal(path):
            v = wire.decode_accept(pkt)
            cur = out.setdefault(v["paxos_id"], {}).get(v["slot"])
            if cur is None or (v["bnum"], v["bcoord"]) >= (cur["bnum"], cur["bcoord"]):
                out[v["paxos_id"]][v["slot"]] = v
    return out
