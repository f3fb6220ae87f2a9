"""GPU parity: the fused round (propose -> accept -> tally -> commit) against the oracle.

Integer/byte work: everything is compared bit-exact.  Streams whose inter-group order is
decided by block scheduling on the device (ACCEPT / DECISION compaction) are compared after
a stable sort by gid, which preserves the per-group order that the protocol defines.
"""
import numpy as np
import pytest

from helpers import Engine, abi, canon, exec_by_lane, group_descs, make_config, make_requests

pytestmark = pytest.mark.gpu


def both(oracle_lib, cuda_lib, **kw):
    return Engine(oracle_lib, make_config(oracle_lib, **kw)), Engine(cuda_lib, make_config(cuda_lib, **kw))


def compare_state(eo, eg, gids, n_lanes):
    for l in range(n_lanes):
        ro, rg = eo.dump_rows(gids, l), eg.dump_rows(gids, l)
        for f in ro.dtype.names:
            assert np.array_equal(ro[f], rg[f]), f"lane {l} field {f}"


def compare_logs(eo, eg, n_lanes):
    for l in range(n_lanes):
        so, sg = abi.parse_log(eo.log_read(l)), abi.parse_log(eg.log_read(l))
        assert len(so) == len(sg)
        for (ho, io_, po, _), (hg, ig, pg, _) in zip(so, sg):
            # n_valid is an upper bound on the image slots in use (the phase pipeline leaves VOID holes where the
            # batcher refused a batch, the oracle compacts them): the images themselves are compared below
            for f in ("type", "lane", "payload_bytes", "seq", "rec_bytes"):
                assert int(ho[f]) == int(hg[f]), f"lane {l} seg hdr {f}"
            co, cg = canon(io_), canon(ig)
            assert len(co) == len(cg)
            cmpf = [f for f in co.dtype.names if f != "payload_off"]
            for f in cmpf:
                assert np.array_equal(co[f], cg[f]), f"lane {l} log image field {f}"
            if int(ho["rec_bytes"]) == 48:  # logged payload bytes
                for a, b in zip(co, cg):
                    ao, bo = int(a["payload_off"]), int(b["payload_off"])
                    ln = int(a["payload_len"])
                    assert np.array_equal(po[ao: ao + ln], pg[bo: bo + ln])


@pytest.mark.parametrize("mode", ["round", "round_phases"])
@pytest.mark.parametrize("G,P,rounds,init", [(1000, 1, 6, abi.INIT_BATCH), (777, 64, 5, abi.INIT_DEFAULT),
                                             (300, 200, 4, abi.INIT_BATCH)])
def test_round_parity(oracle_lib, cuda_lib, G, P, rounds, init, mode):
    eo, eg = both(oracle_lib, cuda_lib, max_groups=G + 5, max_batch_recs=4096, max_batch_payload=1 << 20)
    assert eg.L.has("round_phases") and eo.L.has("round_phases")
    d = group_descs(G, init_mode=init)
    eo.create_groups(d)
    eg.create_groups(d)
    gids = np.arange(G)
    compare_state(eo, eg, gids, 3)
    for r in range(rounds):
        reqs, pay = make_requests(gids, payload_len=P, seed=3, round_no=r, entry_lane=r % 3)
        so, xo, eo_x = getattr(eo, mode)(reqs, pay)
        sg, xg, eg_x = getattr(eg, mode)(reqs, pay)
        assert np.array_equal(so, sg)
        assert len(eo_x) == len(eg_x) == 0
        for a, b in zip(exec_by_lane(xo, 3), exec_by_lane(xg, 3)):
            assert len(a) == G
            assert np.array_equal(a, b)
        compare_state(eo, eg, gids, 3)
    co, cg = eo.counters(), eg.counters()
    cg.pop("kernel_launches"), co.pop("kernel_launches")
    assert co == cg
    assert co["decisions_made"] == G * rounds and co["executed"] == 3 * G * rounds
    compare_logs(eo, eg, 3)
    # RSM invariant (TESTPaxosApp.java:190): slots consecutive from 1 on every replica
    rows = eg.dump_rows(gids, 0)
    assert np.all(rows["acc_slot"] == rounds + 1)


@pytest.mark.parametrize("mode", ["round", "round_phases"])
def test_batched_requests_parity(oracle_lib, cuda_lib, mode):
    """RequestBatcher: several requests of one group in a call share one slot."""
    G = 50
    eo, eg = both(oracle_lib, cuda_lib, max_groups=G, max_batch_recs=4096, max_batch_payload=1 << 20)
    d = group_descs(G)
    eo.create_groups(d)
    eg.create_groups(d)
    rng = np.random.default_rng(5)
    for r in range(4):
        counts = rng.integers(1, 6, size=G)
        gids = np.repeat(np.arange(G), counts)
        lens = rng.integers(1, 40, size=len(gids))
        reqs, pay = make_requests(gids, payload_len=lens, seed=9, round_no=r)
        so, xo, _ = getattr(eo, mode)(reqs, pay)
        sg, xg, _ = getattr(eg, mode)(reqs, pay)
        assert np.array_equal(so, sg)
        for a, b in zip(exec_by_lane(xo, 3), exec_by_lane(xg, 3)):
            assert len(a) == G
            fa = [f for f in a.dtype.names if f != "payload_off"]
            for f in fa:
                assert np.array_equal(a[f], b[f]), f
            assert np.array_equal((a["flags"] >> 16), counts)
    compare_state(eo, eg, np.arange(G), 3)
    compare_logs(eo, eg, 3)
