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
