"""Digest path: MD5 of requestValue.  The digest arithmetic lives in the JDK
(java.security.MessageDigest), not under /root/reference, and no reference test pins digest bytes:
parity is pinned by the RFC 1321 A.5 vectors and hashlib ("parity unpinned by the reference")."""
import hashlib

import numpy as np
import pytest

from helpers import Engine, abi, make_config

RFC1321 = {b"": "d41d8cd98f00b204e9800998ecf8427e", b"a": "0cc175b9c0f1b6a831c399e269772661",
           b"abc": "900150983cd24fb0d6963f7d28e17f72", b"message digest": "f96b697d7cb7938d525a2f31aaf161d0",
           b"abcdefghijklmnopqrstuvwxyz": "c3fcd3d76192e4007dfb496cca67e13b",
           b"ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789": "d174ab98d277d9f5a5611c2c9f419d9f",
           b"1234567890" * 8: "57edf4a22be3c955ac49da2e2107b67a"}


def batch(values):
    n = len(values)
    reqs = np.zeros(n, dtype=abi.request_dtype)
    chunks, off = [], 0
    for i, v in enumerate(values):
        reqs[i]["payload_off"], reqs[i]["payload_len"] = off, len(v)
        pad = (-len(v)) % 16 if len(v) else 16
        chunks.append(np.frombuffer(v + bytes(pad), dtype=np.uint8))
        off += len(v) + pad
    return reqs, np.concatenate(chunks)


def check(lib):
    e = Engine(lib, make_config(lib, max_groups=4, max_batch_recs=4096, max_batch_payload=1 << 20))
    vals = list(RFC1321)
    rng = np.random.default_rng(2)
    vals += [bytes(rng.integers(0, 256, size=int(l)).astype(np.uint8)) for l in
             list(range(50, 70)) + [119, 120, 121, 127, 128, 129, 1000, 1024]]
    reqs, pay = batch(vals)
    d = e.digest_requests(reqs, pay)
    for v, dig in zip(vals, d):
        assert bytes(dig).hex() == hashlib.md5(v).hexdigest()
        if v in RFC1321:
            assert bytes(dig).hex() == RFC1321[v]
    return d


def test_digest_oracle(oracle_lib):
    check(oracle_lib)


@pytest.mark.gpu
def test_digest_gpu(cuda_lib, oracle_lib):
    assert np.array_equal(check(cuda_lib), check(oracle_lib))


# ---- digest mode end to end: AcceptPacket.digest / undigest + PendingDigests -------------------------
def drive_digest_mode(lib, digest: bool):
    from gigapaxos_b200.digests import PendingDigests, assemble, digest_accepts
    from helpers import group_descs, make_requests
    G, L = 40, 3
    e = Engine(lib, make_config(lib, max_groups=G, max_batch_recs=4096, max_batch_payload=1 << 20, batching_enabled=0))
    e.create_groups(group_descs(G))
    pend = [PendingDigests() for _ in range(L)]  # one per acceptor node
    rng = np.random.default_rng(7)
    executed = 0
    for r in range(5):
        reqs, pay = make_requests(np.arange(G), payload_len=rng.integers(1, 90, size=G), seed=3, round_no=r)
        acc, blob, st = e.propose(reqs, pay)
        assert np.all(st > 0)
        if not digest:
            rep, _ = e.handle_accepts(acc, blob)
        else:
            bodies = {int(q["req_id"]): (int(q["gid"]), bytes(pay[int(q["payload_off"]): int(q["payload_off"]) + int(q["payload_len"])]))
                      for q in reqs}
            dacc = digest_accepts(e, acc, blob)  # MD5 on the device: the ACCEPTs travel without their bodies
            assert all(int(d.rec["payload_len"]) == 0 and len(d.digest) == 16 for d in dacc)
            rep = np.zeros(len(acc) * L, dtype=abi.reply_dtype)
            rep["who"] = abi.who(0xFF, 0xFF, abi.F_VOID)
            for l in range(L):
                pairs = []
                order = rng.permutation(len(dacc))
                for k in order:  # the broadcast body and the digested ACCEPT race each other
                    d = dacc[k]
                    rid = int(d.rec["req_id"])
                    gid, value = bodies[rid]
                    if rng.random() < 0.5:
                        pend[l].enqueue(gid, rid, value)
                        got = pend[l].match(d)
                    else:
                        assert pend[l].match(d) is None
                        pend[l].enqueue(gid, rid, value)
                        got = pend[l].release(gid, rid, value)
                    assert got is not None
                    pairs.append(got)
                a2, b2 = assemble(pairs)
                a2["dst_mask"] = 1 << l
                rl, _ = e.handle_accepts(a2, b2)
                rl = rl.reshape(len(a2), L)[:, l]
                # put lane l's replies where the undigested run has them (accept index * L + lane)
                pos = {int(x["req_id"]): i for i, x in enumerate(acc)}
                for x in rl:
                    rep[pos[int(x["req_id"])] * L + l] = x
            assert not any(p.accepts for p in pend) and all(p.anomalies == 0 for p in pend)
        dec = e.handle_accept_replies(rep)
        ex, extra = e.handle_decisions(dec)
        executed += int(((ex["flags"] & abi.F_VOID) == 0).sum())
    assert executed == 5 * G * L
    return e


def check_digest_mode(lib):
    from gigapaxos_b200.digests import DigestedAccept, PendingDigests
    a, b = drive_digest_mode(lib, True), drive_digest_mode(lib, False)
    gids = np.arange(40)
    for l in range(3):
        ra, rb = a.dump_rows(gids, l), b.dump_rows(gids, l)
        for f in ra.dtype.names:
            assert np.array_equal(ra[f], rb[f]), (l, f)
    # a body that does not hash to the ACCEPT's digest is an anomaly, not an accept (PendingDigests.logAnomaly)
    p = PendingDigests()
    rec = np.zeros(1, dtype=abi.accept_dtype)[0]
    rec["gid"], rec["req_id"] = 3, 77
    p.enqueue(3, 77, b"the body")
    assert p.match(DigestedAccept(rec, hashlib.md5(b"another body").digest(), 12)) is None and p.anomalies == 1
    assert p.match(DigestedAccept(rec, hashlib.md5(b"the body").digest(), 8)) is not None


def test_digest_mode_cpu(oracle_lib):
    check_digest_mode(oracle_lib)


@pytest.mark.gpu
def test_digest_mode_gpu(cuda_lib):
    check_digest_mode(cuda_lib)
