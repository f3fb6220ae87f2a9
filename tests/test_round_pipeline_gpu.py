"""GPU parity of the pipelined round API (gpx_round_submit / gpx_round_wait), full and compact output.

The oracle runs the same rounds one after the other (gpxo_round); the device has up to GPX_PIPE_DEPTH of
them in flight.  Full mode is compared bit-exact record by record.  In compact mode the device reports the
plain in-order executions as one 8-byte summary per request and everything else through the extra queue, so
the comparison is on the canonical set of executions (lane, gid, slot, req_id, flags, nreq) of each round,
plus state and log equality at the end.
"""
import numpy as np
import pytest

from helpers import Engine, abi, exec_by_lane, group_descs, make_config, make_requests
from test_round_parity_gpu import both, compare_logs, compare_state

pytestmark = pytest.mark.gpu

FMASK = abi.F_STOP | abi.F_CKPT  # flags that matter to the application


def exec_tuples(recs):
    """canonical executions of full EXEC records"""
    out = []
    for r in recs:
        fl = int(r["flags"])
        if fl & abi.F_VOID:
            continue
        out.append(((fl >> 12) & 0xF, int(r["gid"]), int(r["slot"]), int(r["req_id"]), fl & FMASK, fl >> 16))
    return sorted(out)


def sum_tuples(sums, reqs, n_lanes):
    out = []
    for s, r in zip(sums, reqs):
        for l in range(n_lanes):
            if (int(s["lane_mask"]) >> l) & 1:
                out.append((l, int(r["gid"]), int(s["slot"]), int(r["req_id"]), int(s["flags"]) & FMASK, int(s["nreq"])))
    return out


def test_pipelined_full_mode_is_gpx_round(oracle_lib, cuda_lib):
    G, rounds = 1500, 9
    eo, eg = both(oracle_lib, cuda_lib, max_groups=G, max_batch_recs=4096, max_batch_payload=1 << 20,
                  checkpoint_interval=4)
    d = group_descs(G)
    eo.create_groups(d)
    eg.create_groups(d)
    gids = np.arange(G)
    batches = [make_requests(gids, payload_len=1 + 13 * (r % 3), seed=11, round_no=r, entry_lane=r % 3)
               for r in range(rounds)]
    tickets, done = [], []
    for reqs, pay in batches:
        if len(tickets) - len(done) == abi.PIPE_DEPTH:
            done.append(eg.round_wait(tickets[len(done)]))
        tickets.append(eg.round_submit(reqs, pay))
    assert tickets == list(range(rounds))
    while len(done) < rounds:
        done.append(eg.round_wait(tickets[len(done)]))
    for (reqs, pay), res in zip(batches, done):
        so, xo, ex_o = eo.round(reqs, pay)
        assert np.array_equal(so, res["status"])
        assert res["n_extra"] == len(ex_o) == 0
        for a, b in zip(exec_by_lane(xo, 3), exec_by_lane(res["exec"], 3)):
            assert len(a) == G and np.array_equal(a, b)
    compare_state(eo, eg, gids, 3)
    compare_logs(eo, eg, 3)
    co, cg = eo.counters(), eg.counters()
    co.pop("kernel_launches"), cg.pop("kernel_launches")
    assert co == cg and co["checkpoints_due"] > 0


@pytest.mark.parametrize("batching", [1, 0])
def test_compact_mode_executions(oracle_lib, cuda_lib, batching):
    """single-request groups take the in-order path (summaries); multi-request runs, STOPs and requests of
    stopped groups take the general path (extra queue / status codes)"""
    G, rounds = 600, 7
    eo, eg = both(oracle_lib, cuda_lib, max_groups=G, max_batch_recs=8192, max_batch_payload=1 << 20,
                  checkpoint_interval=3, batching_enabled=batching)
    d = group_descs(G)
    eo.create_groups(d)
    eg.create_groups(d)
    rng = np.random.default_rng(17)
    n_sum = n_ext = 0
    pending = []
    for r in range(rounds):
        counts = rng.choice([1, 1, 1, 1, 2, 3], size=G)
        counts[rng.random(G) < 0.05] = 0
        gids = np.repeat(np.arange(G), counts)
        lens = rng.integers(1, 50, size=len(gids))
        stop = (rng.random(len(gids)) < 0.01) if r >= 3 else None
        reqs, pay = make_requests(gids, payload_len=lens, seed=23, round_no=r, entry_lane=int(rng.integers(0, 3)),
                                  stop_mask=stop)
        so, xo, ex_o = eo.round(reqs, pay, extra_cap=3 * len(reqs) + 64)
        t = eg.round_submit(reqs, pay, compact=True, extra_cap=4 * len(reqs) + 64)
        pending.append((t, reqs, so, xo, ex_o))
        if len(pending) == 3 or r == rounds - 1:
            for (t, reqs, so, xo, ex_o) in pending:
                res = eg.round_wait(t)
                sums = res["sum"]
                assert res["n_extra"] == len(res["extra"])
                # status: the summary's slot field is the request's status
                assert np.array_equal(sums["slot"], so)
                want = sorted(exec_tuples(xo) + exec_tuples(ex_o))
                got = sorted(sum_tuples(sums, reqs, 3) + exec_tuples(res["extra"]))
                assert got == want
                n_sum += int((sums["lane_mask"] != 0).sum())
                n_ext += len(res["extra"])
            pending = []
    assert n_sum > G and n_ext > 0  # both report paths were exercised
    compare_state(eo, eg, np.arange(G), 3)
    compare_logs(eo, eg, 3)
    co, cg = eo.counters(), eg.counters()
    co.pop("kernel_launches"), cg.pop("kernel_launches")
    assert co == cg


def test_pipeline_depth_and_order_are_enforced(cuda_lib):
    from helpers import Engine, make_config
    eg = Engine(cuda_lib, make_config(cuda_lib, max_groups=64, max_batch_recs=256))
    eg.create_groups(group_descs(64))
    reqs, pay = make_requests(np.arange(64))
    ts = [eg.round_submit(reqs, pay) for _ in range(abi.PIPE_DEPTH)]
    with pytest.raises(abi.GpxError):
        eg.round_submit(reqs, pay)  # GPX_PIPE_DEPTH rounds in flight
    with pytest.raises(abi.GpxError):
        eg.L.check(eg.L.fn("round_wait")(eg.handle, abi.C.c_uint64(ts[1]), None, None))  # out of order
    first = None
    for k, t in enumerate(ts):
        st = eg.round_wait(t)["status"]
        first = int(st[0]) if first is None else first
        assert np.all(st == first + k)  # rounds run in submission order
    assert eg.dump_rows(np.arange(64), 0)["acc_slot"].tolist() == [first + abi.PIPE_DEPTH] * 64


def test_packed_requests(oracle_lib, cuda_lib):
    """GPX_ROUND_PACKED_REQS: 16-byte requests, payloads back to back; the device expands them (k_unpack) to the
    records the oracle is given explicitly"""
    G, rounds = 700, 6
    eo, eg = both(oracle_lib, cuda_lib, max_groups=G, max_batch_recs=8192, max_batch_payload=1 << 20,
                  checkpoint_interval=5)
    d = group_descs(G)
    eo.create_groups(d)
    eg.create_groups(d)
    rng = np.random.default_rng(29)
    nodes = [100, 101, 102]
    for r in range(rounds):
        counts = rng.choice([0, 1, 1, 1, 2, 5], size=G)
        gids = np.repeat(np.arange(G), counts).astype(np.uint32)
        n = len(gids)  # > 1024 so the scan crosses unpack blocks
        lens = rng.integers(0, 70, size=n).astype(np.uint32)
        lane = int(rng.integers(0, 3))
        packed = np.zeros(n, dtype=abi.request_packed_dtype)
        packed["gid"], packed["payload_len"] = gids, lens
        packed["flags"] = (lane << 8) | np.where(rng.random(n) < (0.01 if r > 2 else 0), abi.F_STOP, 0)
        packed["req_id"] = rng.integers(1, 1 << 62, size=n)
        pay = rng.integers(48, 123, size=int(lens.sum()), dtype=np.uint8)
        full = np.zeros(n, dtype=abi.request_dtype)
        full["gid"], full["flags"], full["req_id"], full["payload_len"] = gids, packed["flags"], packed["req_id"], lens
        full["payload_off"] = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint32)
        full["entry_node"] = nodes[lane]
        full["client"] = np.arange(n, dtype=np.uint32)
        so, xo, ex_o = eo.round(full, pay, extra_cap=3 * n + 64)
        res = eg.round_wait(eg.round_submit(packed, pay, packed=True, extra_cap=3 * n + 64))
        assert np.array_equal(so, res["status"])
        assert sorted(exec_tuples(xo) + exec_tuples(ex_o)) == sorted(exec_tuples(res["exec"]) + exec_tuples(res["extra"]))
    compare_state(eo, eg, np.arange(G), 3)
    compare_logs(eo, eg, 3)


def test_round_device_compact_parity(oracle_lib, cuda_lib):
    """gpx_round_device_compact (k_propose + k_build_blobs + k_act on device buffers) == gpx_propose followed by
    gpx_handle_accepts_fused on the oracle: status, EXEC per lane, rows, counters -- with batched slots of 1..40 requests"""
    import ctypes as C
    import torch
    from gigapaxos_b200.abi import DevRoundBufs
    G = 700
    eo, eg = (Engine(lib, make_config(lib, max_groups=G, max_batch_recs=1 << 15, max_batch_payload=1 << 21))
              for lib in (oracle_lib, cuda_lib))
    d = group_descs(G)
    eo.create_groups(d)
    eg.create_groups(d)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    rng = np.random.default_rng(21)
    for r in range(5):
        counts = rng.choice([0, 1, 1, 2, 7, 40], size=G)
        gids = np.repeat(np.arange(G), counts)
        lens = rng.integers(1, 50, size=len(gids))
        reqs, pay = make_requests(gids, payload_len=lens, seed=13, round_no=r, entry_lane=r % 3)
        acc, blob, so = eo.propose(reqs, pay)
        rep, dec, xo, extra_o = eo.handle_accepts_fused(acc, blob)
        n = len(reqs)
        d_reqs = torch.from_numpy(reqs.view(np.uint8).copy()).to(dev)
        d_pay = torch.from_numpy(np.concatenate([pay, np.zeros(16, np.uint8)])).to(dev)
        d_status = torch.zeros(n, dtype=torch.int32, device=dev)
        d_exec = torch.zeros(n * 3 * 24, dtype=torch.uint8, device=dev)
        bufs = DevRoundBufs(d_reqs.data_ptr(), d_pay.data_ptr(), len(pay), n, d_status.data_ptr(), d_exec.data_ptr())
        torch.cuda.synchronize()
        cuda_lib.check(cuda_lib.fn("round_device_compact")(eg.handle, C.byref(bufs), C.c_void_p(stream.cuda_stream)))
        torch.cuda.synchronize()
        assert np.array_equal(d_status.cpu().numpy(), so)
        xg = d_exec.cpu().numpy().view(abi.exec_dtype)[: len(acc) * 3]
        for a, b in zip(exec_by_lane(xo, 3), exec_by_lane(xg, 3)):
            assert len(a) == len(b) == int((counts > 0).sum())
            for f in a.dtype.names:
                if f != "payload_off":
                    assert np.array_equal(a[f], b[f]), f
    for l in range(3):
        ro, rg = eo.dump_rows(np.arange(G), l), eg.dump_rows(np.arange(G), l)
        for f in ro.dtype.names:
            assert np.array_equal(ro[f], rg[f]), f
    co, cg = eo.counters(), eg.counters()
    co.pop("kernel_launches"), cg.pop("kernel_launches")
    assert co == cg
