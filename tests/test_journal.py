"""Journal persistence: the log ring drained into Journaler-compatible files
({int32 BE len}{packet}*, SQLPaxosLogger.java:1000-1003: ACCEPTs byteified, DECISIONs as the JSON string of the meta
decision, :1084-1098) and read back the way SQLPaxosLogger's reader walks a file (:685-848)."""
import os

import numpy as np
import pytest

from gigapaxos_b200 import journal, wire
from gigapaxos_b200.paxos_manager import NoopPaxosApp, PaxosManager
from helpers import Engine, abi, make_config

NODES = [100, 101, 102]


def run(lib, tmp_path):
    eng = Engine(lib, make_config(lib, max_groups=32, max_batch_recs=1024, max_batch_payload=1 << 18))
    pm = PaxosManager(eng, [NoopPaxosApp() for _ in NODES], NODES)
    names = [f"NoopPaxosApp{i}" for i in range(6)]
    pm.createPaxosInstanceBatch({n: None for n in names}, NODES)
    meta = {}
    sent = {}

    def lookup(gid, rid):
        return meta[rid]

    drainers = [journal.LogDrainer(eng, l, journal.Journaler(str(tmp_path) + "/", NODES[l], max_log_file_size=2000),
                                   lookup) for l in range(3)]
    rng = np.random.default_rng(4)
    for r in range(6):
        for n in names:
            for _ in range(int(rng.integers(1, 4))):
                val = bytes(rng.integers(97, 123, size=int(rng.integers(1, 30))).astype(np.uint8))
                rid = pm.propose(n, val, entry_node=NODES[r % 3])
                req = pm.outstanding[rid]
                meta[rid] = (n, 0, req.entry_replica, int(req.entry_time * 1000))
                sent[rid] = (n, val)
        pm.run_round()
        for d in drainers:
            d.drain()
    for d in drainers:
        d.j.close()
    return pm, drainers, sent


def check(pm, drainers, sent):
    for l, d in enumerate(drainers):
        assert len(d.j.files) > 1  # rolled at MAX_LOG_FILE_SIZE (rollLogFile :789-812)
        assert all(os.path.basename(f).startswith(f"log.{NODES[l]}.") for f in d.j.files)
        assert os.path.basename(os.path.dirname(d.j.files[0])) == f"paxos_journal.{NODES[l]}"
        seen = {}
        slots = {}
        dslots = {}
        for f in d.j.files:
            assert not os.path.exists(f + ".decisions")  # one journal, no side file
            for pkt in journal.read_journal(f):
                v = journal.parse_packet(pkt)
                if v["kind"] == "ACCEPT":
                    assert pkt[3] == 90 and pkt[7] == 3  # PAXOS_PACKET / ACCEPT (SQLPaxosLogger.toBytes :1090)
                    slots.setdefault(v["paxos_id"], []).append(v["slot"])
                    seen[v["request_id"]] = (v["paxos_id"], v["value"])
                else:
                    assert v["kind"] == "DECISION" and pkt[:1] == b"{"  # the JSON string of the meta decision
                    assert v["median_cp"] == -1  # PValuePacket.getMetaDecision :212-218
                    dslots.setdefault(v["paxos_id"], []).append((v["slot"], v["request_id"]))
        for name, ss in dslots.items():  # every slot's decision is journaled once, in slot order, for the logged request
            assert [x[0] for x in ss] == list(range(1, len(ss) + 1))
            assert all(seen[rid][0] == name for _, rid in ss)
        assert journal.replay_decisions(d.j.files).keys() == dslots.keys()
        # every first request of every decided slot is in every replica's journal, byte-exact
        for rid, (name, val) in seen.items():
            assert sent[rid] == (name, val)
        for name, ss in slots.items():
            assert ss == list(range(1, len(ss) + 1))  # one logged ACCEPT per slot, in slot order
        assert d.accepts_written == sum(len(v) for v in slots.values()) == pm.num_decisions
        assert d.decisions_written == pm.num_decisions
        rec = journal.replay_accepts(d.j.files)
        assert set(rec) == set(slots) and all(len(rec[n]) == len(slots[n]) for n in rec)


def test_journal_roundtrip_cpu(oracle_lib, tmp_path):
    check(*run(oracle_lib, tmp_path))


@pytest.mark.gpu
def test_journal_roundtrip_gpu(cuda_lib, oracle_lib, tmp_path):
    pm, dr, sent = run(cuda_lib, tmp_path / "gpu")
    check(pm, dr, sent)
    pm2, dr2, _ = run(oracle_lib, tmp_path / "cpu")
    # same packets in the same order on both (entry times differ: compare with entryTime masked)
    for a, b in zip(dr, dr2):
        pa = [p for f in a.j.files for p in journal.read_journal(f)]
        pb = [p for f in b.j.files for p in journal.read_journal(f)]
        assert len(pa) == len(pb)
        for x, y in zip(pa, pb):
            vx, vy = journal.parse_packet(x), journal.parse_packet(y)
            assert vx["kind"] == vy["kind"]
            keys = (("paxos_id", "request_id", "slot", "bnum", "bcoord", "median_cp", "sender", "value", "n_batched")
                    if vx["kind"] == "ACCEPT" else ("paxos_id", "request_id", "slot", "bnum", "bcoord", "median_cp", "stop"))
            for k in keys:
                assert vx[k] == vy[k]
