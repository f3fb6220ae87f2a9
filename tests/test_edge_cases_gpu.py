"""GPU parity on the edges: empty and ragged batches, the RequestBatcher cut rules, back-pressure on a small
window, single-replica groups, unknown / destroyed groups, capacity errors."""
import numpy as np
import pytest

from helpers import abi, exec_by_lane, group_descs, make_requests
from test_round_parity_gpu import both, compare_logs, compare_state

pytestmark = pytest.mark.gpu
MODES = ["round", "round_phases"]


def same_round(eo, eg, mode, reqs, pay, n_lanes=3, extra_cap=4096):
    so, xo, ex_o = getattr(eo, mode)(reqs, pay, extra_cap=extra_cap)
    sg, xg, ex_g = getattr(eg, mode)(reqs, pay, extra_cap=extra_cap)
    assert np.array_equal(so, sg)
    key = lambda r: (int(r["gid"]), int(r["slot"]), int(r["flags"]) >> 12 & 0xF)
    for a, b in zip(exec_by_lane(np.concatenate([xo, ex_o]), n_lanes), exec_by_lane(np.concatenate([xg, ex_g]), n_lanes)):
        a, b = sorted(a.tolist(), key=lambda t: (t[0], t[1])), sorted(b.tolist(), key=lambda t: (t[0], t[1]))
        assert len(a) == len(b)
        for x, y in zip(a, b):
            assert x[:3] == y[:3] and (x[4] & ~abi.F_EXTRA) == (y[4] & ~abi.F_EXTRA)  # gid, slot, req_id, flags
    return so


@pytest.mark.parametrize("mode", MODES)
def test_empty_and_zero_length(oracle_lib, cuda_lib, mode):
    G = 40
    eo, eg = both(oracle_lib, cuda_lib, max_groups=G, max_batch_recs=256, max_batch_payload=1 << 16)
    d = group_descs(G)
    eo.create_groups(d)
    eg.create_groups(d)
    empty = np.zeros(0, dtype=abi.request_dtype)
    for e in (eo, eg):  # an empty batch is a no-op
        s, x, ex = getattr(e, mode)(empty, np.zeros(0, np.uint8))
        assert len(s) == len(x) == len(ex) == 0
    assert eo.counters()["proposals"] == eg.counters()["proposals"] == 0
    # zero-length request values, alone and inside batched slots
    gids = np.repeat(np.arange(G), np.arange(G) % 3 + 1)
    lens = np.where(np.arange(len(gids)) % 2 == 0, 0, 7)
    reqs, pay = make_requests(gids, payload_len=lens, seed=4)
    so = same_round(eo, eg, mode, reqs, pay)
    assert (so > 0).sum() == G
    # all payloads empty: payload arena of 0 bytes
    reqs, pay = make_requests(np.arange(G), payload_len=0, seed=5, round_no=1)
    assert pay.size == 0
    same_round(eo, eg, mode, reqs, pay)
    compare_state(eo, eg, np.arange(G), 3)
    compare_logs(eo, eg, 3)


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("window", [1, 2, 8])
def test_batch_cut_rules_and_backpressure(oracle_lib, cuda_lib, mode, window):
    """RequestBatcher.java:198-219: a batch is cut by MAX_BATCH_SIZE and by the byte limit (lengthEstimate = len +
    SIZE_ESTIMATE); every cut is a further slot; more than W slots of a group in flight -> GPX_RS_BACKPRESSURE"""
    G = 24
    kw = dict(max_groups=G, max_batch_recs=2048, max_batch_payload=1 << 20, window=window, max_batch_size=5,
              max_batch_bytes=3000, request_size_estimate=100)
    eo, eg = both(oracle_lib, cuda_lib, **kw)
    d = group_descs(G)
    eo.create_groups(d)
    eg.create_groups(d)
    rng = np.random.default_rng(8)
    seen_bp = False
    for r in range(5):
        counts = rng.integers(0, 30, size=G)
        gids = np.repeat(np.arange(G), counts)
        lens = rng.choice([1, 10, 400, 900], size=len(gids))
        stop = (rng.random(len(gids)) < 0.01) if r >= 2 else None  # a STOP inside a run: later batches are refused
        reqs, pay = make_requests(gids, payload_len=lens, seed=12, round_no=r, entry_lane=r % 3, stop_mask=stop)
        so = same_round(eo, eg, mode, reqs, pay, extra_cap=8192)
        seen_bp = seen_bp or bool((so == abi.RS_BACKPRESSURE).any())
        compare_state(eo, eg, np.arange(G), 3)
    assert seen_bp == (window < 8) or window == 8
    co, cg = eo.counters(), eg.counters()
    co.pop("kernel_launches"), cg.pop("kernel_launches")
    assert co == cg
    compare_logs(eo, eg, 3)


@pytest.mark.parametrize("mode", MODES)
def test_single_replica_unknown_and_destroyed_groups(oracle_lib, cuda_lib, mode):
    G = 30
    kw = dict(max_groups=G + 10, n_lanes=1, lane_node=[100], max_group_size=1, max_batch_recs=512,
              max_batch_payload=1 << 16)
    eo, eg = both(oracle_lib, cuda_lib, **kw)
    d = group_descs(G, members=(100,))
    eo.create_groups(d)
    eg.create_groups(d)
    for e in (eo, eg):
        e.destroy_groups(np.array([3, 4], dtype=np.uint32))
    gids = np.concatenate([np.arange(G), [G + 2, G + 5]]).astype(np.uint32)  # two never-created gids at the end
    for r in range(3):
        reqs, pay = make_requests(gids, payload_len=3, seed=2, round_no=r)
        so = same_round(eo, eg, mode, reqs, pay, n_lanes=1)
        assert np.all(so[[3, 4, G, G + 1]] == abi.RS_DROPPED) and (so > 0).sum() == G - 2  # majority of one
    live = np.array([g for g in range(G) if g not in (3, 4)])
    compare_state(eo, eg, live, 1)
    for e in (eo, eg):
        assert np.all(e.dump_rows(np.array([3, 4], dtype=np.uint32), 0)["state"] == abi.ST_FREE)
    compare_logs(eo, eg, 1)


def test_capacity_errors_are_api_errors(oracle_lib, cuda_lib):
    eo, eg = both(oracle_lib, cuda_lib, max_groups=8, max_batch_recs=16, max_batch_payload=64)
    d = group_descs(8)
    eo.create_groups(d)
    eg.create_groups(d)
    reqs, pay = make_requests(np.repeat(np.arange(8), 3), payload_len=1)  # 24 > max_batch_recs
    big, bigpay = make_requests(np.arange(8), payload_len=32)            # 256 B > max_batch_payload
    with pytest.raises(abi.GpxError) as ei:  # the engine's scratch is sized by the configuration
        eg.round(reqs, pay)
    assert ei.value.code == abi.GPX_ERANGE
    with pytest.raises(abi.GpxError):
        eg.round(big, bigpay)
    with pytest.raises(abi.GpxError):
        eg.round_submit(reqs, pay)
    # nothing happened
    assert eg.counters()["proposals"] == 0
    compare_state(eo, eg, np.arange(8), 3)


@pytest.mark.parametrize("window", [2, 8])
def test_cut_batches_through_the_pipelined_compact_api(oracle_lib, cuda_lib, window):
    """the same cut / back-pressure / STOP schedules through gpx_round_submit with compact summaries and packed
    requests: multi-batch runs take the general path and report through the extra queue"""
    from test_round_pipeline_gpu import exec_tuples, sum_tuples
    G = 24
    kw = dict(max_groups=G, max_batch_recs=2048, max_batch_payload=1 << 20, window=window, max_batch_size=5,
              max_batch_bytes=3000, request_size_estimate=100)
    eo, eg = both(oracle_lib, cuda_lib, **kw)
    d = group_descs(G)
    eo.create_groups(d)
    eg.create_groups(d)
    rng = np.random.default_rng(18)
    nodes = [100, 101, 102]
    for r in range(5):
        counts = rng.integers(0, 25, size=G)
        gids = np.repeat(np.arange(G), counts).astype(np.uint32)
        n = len(gids)
        lens = rng.choice([1, 10, 400, 900], size=n).astype(np.uint32)
        lane = r % 3
        packed = np.zeros(n, dtype=abi.request_packed_dtype)
        packed["gid"], packed["payload_len"] = gids, lens
        packed["flags"] = (lane << 8) | np.where(rng.random(n) < (0.01 if r >= 2 else 0), abi.F_STOP, 0)
        packed["req_id"] = rng.integers(1, 1 << 62, size=n)
        pay = rng.integers(48, 123, size=int(lens.sum()), dtype=np.uint8)
        full = np.zeros(n, dtype=abi.request_dtype)
        full["gid"], full["flags"], full["req_id"], full["payload_len"] = gids, packed["flags"], packed["req_id"], lens
        full["payload_off"] = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint32)
        full["entry_node"], full["client"] = nodes[lane], np.arange(n, dtype=np.uint32)
        so, xo, ex_o = eo.round(full, pay, extra_cap=8192)
        res = eg.round_wait(eg.round_submit(packed, pay, compact=True, packed=True, extra_cap=8192))
        assert np.array_equal(res["sum"]["slot"], so)
        want = sorted(exec_tuples(xo) + exec_tuples(ex_o))
        got = sorted(sum_tuples(res["sum"], full, 3) + exec_tuples(res["extra"]))
        assert got == want
        compare_state(eo, eg, np.arange(G), 3)
    compare_logs(eo, eg, 3)
