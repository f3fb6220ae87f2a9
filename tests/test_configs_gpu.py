"""GPU parity at the shapes of BASELINE.json's configs 4 and 5 (scaled to what the oracle finishes in seconds).

config 4: 5 replicas, request sizes uniform in 1..1024 B, reconfiguration churn -- every round some groups
          receive a STOP; stopped groups are destroyed and re-created at the next epoch (version + 1) with fresh
          state (PaxosManager.kill :2162 + createPaxosInstance :632; the version drop rule :441 lives where names
          map to gids).
config 5: accept-batch sweep, 1..1024 requests of one group in one batch -> ONE slot carrying the whole batch
          (RequestBatcher.java:198-219, MAX_BATCH_SIZE 2000).
"""
import numpy as np
import pytest

from helpers import abi, exec_by_lane, group_descs, make_requests
from test_round_parity_gpu import both, compare_logs, compare_state

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["round", "round_phases"])
def test_config4_five_replicas_mixed_sizes_churn(oracle_lib, cuda_lib, mode):
    G, R = 300, 5
    nodes = [100, 101, 102, 103, 104]
    kw = dict(max_groups=G, n_lanes=R, lane_node=nodes, max_group_size=R, max_batch_recs=4096,
              max_batch_payload=1 << 21, log_ring_bytes=1 << 26)
    eo, eg = both(oracle_lib, cuda_lib, **kw)
    version = np.zeros(G, dtype=np.int32)
    d = group_descs(G, members=nodes)
    eo.create_groups(d)
    eg.create_groups(d)
    rng = np.random.default_rng(4)
    stopped_total = 0
    for r in range(7):
        gids = np.arange(G)
        lens = rng.integers(1, 1025, size=G)
        stop = rng.random(G) < 0.03
        reqs, pay = make_requests(gids, payload_len=lens, seed=2, round_no=r, entry_lane=r % R, stop_mask=stop)
        so, xo, extra_o = getattr(eo, mode)(reqs, pay)
        sg, xg, extra_g = getattr(eg, mode)(reqs, pay)
        assert np.array_equal(so, sg)
        assert len(extra_o) == len(extra_g) == 0
        for a, b in zip(exec_by_lane(xo, R), exec_by_lane(xg, R)):
            assert len(a) == G and np.array_equal(a, b)
        compare_state(eo, eg, gids, R)
        # churn: every group whose STOP executed is killed and re-created at the next epoch
        lane0 = exec_by_lane(xg, R)[0]
        dead = lane0["gid"][(lane0["flags"] & abi.F_STOP) != 0]
        assert set(dead.tolist()) == set(gids[stop].tolist())
        if len(dead):
            stopped_total += len(dead)
            rows = eg.dump_rows(dead, 0)
            assert np.all(rows["state"] == abi.ST_STOPPED)
            version[dead] += 1
            nd = group_descs(G, members=nodes)[dead]
            nd["version"] = version[dead]
            for e in (eo, eg):
                e.destroy_groups(dead)
                e.create_groups(nd)
            compare_state(eo, eg, dead, R)
    assert stopped_total > 0
    co, cg = eo.counters(), eg.counters()
    co.pop("kernel_launches"), cg.pop("kernel_launches")
    assert co == cg and co["stops_executed"] == R * stopped_total
    compare_logs(eo, eg, R)


@pytest.mark.parametrize("mode", ["round", "round_phases"])
def test_config5_accept_batch_sweep(oracle_lib, cuda_lib, mode):
    G = 64
    eo, eg = both(oracle_lib, cuda_lib, max_groups=G, max_batch_recs=8192, max_batch_payload=1 << 20)
    d = group_descs(G)
    eo.create_groups(d)
    eg.create_groups(d)
    for r, b in enumerate([1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024]):
        active = np.arange(r % 7, G, 17)  # a small active subset, the rest idle
        gids = np.repeat(active, b)
        reqs, pay = make_requests(gids, payload_len=1, seed=6, round_no=r)
        so, xo, _ = getattr(eo, mode)(reqs, pay)
        sg, xg, _ = getattr(eg, mode)(reqs, pay)
        assert np.array_equal(so, sg)
        # one slot per active group, the rest of the batch latched into it
        assert int((so > 0).sum()) == len(active) and int((so == abi.RS_BATCHED).sum()) == len(gids) - len(active)
        for a, c in zip(exec_by_lane(xo, 3), exec_by_lane(xg, 3)):
            assert len(a) == len(active)
            for f in a.dtype.names:
                if f != "payload_off":
                    assert np.array_equal(a[f], c[f]), f
            assert np.all((a["flags"] >> 16) == b)
    compare_state(eo, eg, np.arange(G), 3)
    compare_logs(eo, eg, 3)
    co, cg = eo.counters(), eg.counters()
    co.pop("kernel_launches"), cg.pop("kernel_launches")
    assert co == cg
