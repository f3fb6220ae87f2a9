"""The kernels beside the round that were written after this round's GPU minutes were spent, against the oracle through
the C ABI -- PaxosManager's per-instance sweeps as launches:

  k_prepare_tally (gpx_handle_prepare_replies)   phase 1b: random batches of elections (every verdict, GPX_F_MORE
      continuation records, slots across the int wrap, R = 1..5, nodes of a spread placement) -- the 896-byte result
      records byte for byte and every row of every lane afterwards; 20,000 elections in one launch; whole view changes
      and the mass fail-over of the host mirror with the tally inside the engine;
  k_pause_groups / k_select_groups / k_clear_flags (gpx_pause_groups, gpx_select_groups, gpx_clear_group_flags)   the
      deactivation sweep and the slow-path list;
  k_log_dir / k_log_scan / k_log_hits / k_log_gather (gpx_log_find, gpx_log_gather)   the journal's index as a scan of the
      log ring, incl. a ring that wrapped; the mirror's catch-up and a lagging preparer's logged accepts through it.

The same kernel sources also run on the host (tests/emu: test_phase1b.py, test_pause_batch.py, test_log_find.py).  The
file sorts last on purpose: of these kernels only k_prepare_tally has met a B200 (its two R = 3 cases, with the round's
last GPU seconds: profiles/r2l_phase1b_gpu_first_run.txt); for the rest the driver's round-end pass is the first run, and
nothing else depends on them."""
import numpy as np
import pytest

from helpers import abi
from p1b_cases import assert_same_out, assert_same_rows, make_engine, preconditions, random_elections

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("R,seed,wrap,lane_nodes", [(3, 21, False, None), (3, 22, True, None), (5, 23, False, None),
                                                    (5, 24, True, None), (1, 25, False, None), (4, 26, False, None),
                                                    (2, 27, False, None),
                                                    (3, 28, False, [101]), (5, 29, True, [103])])  # nodes of a spread placement (one lane per engine)
def test_phase1b_kernel_equals_oracle(cuda_lib, oracle_lib, R, seed, wrap, lane_nodes):
    G = 300
    rng = np.random.default_rng(seed)
    eg, eo = make_engine(cuda_lib, R, G, lane_nodes), make_engine(oracle_lib, R, G, lane_nodes)
    verdicts = set()
    for rnd in range(4):
        st = rng.bit_generator.state
        preconditions(eg, R, G, rng)
        rng.bit_generator.state = st
        preconditions(eo, R, G, rng)
        els, reps = random_elections(R, G, rng, wrap, lane_nodes)
        got, want = eg.handle_prepare_replies(els, reps), eo.handle_prepare_replies(els, reps)
        assert_same_out(got, want)
        verdicts |= set(int(v) for v in want["verdict"])
        assert_same_rows(eg, eo, R, G, rnd)
    if R >= 3:
        assert verdicts == {abi.EL_WAITING, abi.EL_MAJORITY, abi.EL_PREEMPTED, abi.EL_DROPPED, abi.EL_OVERFLOW}
    assert eg.counters()["kernel_launches"] > 0


@pytest.mark.parametrize("R,seed", [(3, 51), (5, 52)])
def test_select_groups_kernel_equals_oracle(cuda_lib, oracle_lib, R, seed):
    from test_pause_batch import busy_engine
    G = 120
    eg, eo = busy_engine(cuda_lib, G, seed, 1, R), busy_engine(oracle_lib, G, seed, 1, R)
    some = 0
    for lane in range(R):
        for mask, value in ((abi.GF_NOT_CAUGHT_UP, 0), (abi.GF_NOT_CAUGHT_UP, abi.GF_NOT_CAUGHT_UP), (0, 0),
                            (abi.GF_NEEDS_SYNC, abi.GF_NEEDS_SYNC), (abi.GF_OVERFLOW | abi.GF_NEEDS_SYNC, 0)):
            got, want = eg.select_groups(lane, mask, value), eo.select_groups(lane, mask, value)
            assert np.array_equal(got, want), (lane, mask, value)
            some += len(want)
    assert some > G
    # the sweep through the mirror: idle groups out in one batch, back on demand
    from gigapaxos_b200.paxos_manager import HashChainApp, PaxosManager
    from helpers import Engine, make_config

    def drive(lib):
        eng = Engine(lib, make_config(lib, max_groups=64, max_batch_recs=4096, max_batch_payload=1 << 20))
        pm = PaxosManager(eng, [HashChainApp() for _ in range(3)], [100, 101, 102])
        names = [f"TESTPaxosApp{i}" for i in range(24)]
        pm.createPaxosInstanceBatch({n: None for n in names}, [100, 101, 102])
        for r in range(2):
            for n in names:
                pm.propose(n, f"{n}:{r}".encode())
            pm.run_round()
        pm.propose(names[5], b"waiting")
        res = pm.syncAndDeactivate()
        pm.run_round()
        for n in names:
            assert pm.propose(n, b"again") is not None
        pm.run_round()
        return res, pm.apps[0].state
    assert drive(cuda_lib) == drive(oracle_lib)


def test_missing_decisions_kernel_equals_oracle(cuda_lib, oracle_lib):
    """k_missing_decisions: a lane with a slot without a commit, a commit without its accept, a stopped acceptor, groups that
    are caught up, gids that do not exist -- the SYNC_DECISIONS fields byte for byte"""
    from test_pause_batch import engine_with_holes
    (eg, sel_g), (eo, sel_o) = engine_with_holes(cuda_lib), engine_with_holes(oracle_lib)
    assert np.array_equal(sel_g, sel_o)
    gids = np.arange(26, dtype=np.uint32)
    some = 0
    for lane in range(3):
        for size_limit, gap in ((400, 400), (1, 400), (400, 3)):
            got, want = eg.missing_decisions(lane, gids, size_limit, gap), eo.missing_decisions(lane, gids, size_limit, gap)
            assert got.tobytes() == want.tobytes(), (lane, size_limit, gap)
            some += int((want["n_missing"] > 1).sum())
    assert some > 0


# ---- the deactivation sweep (k_pause_groups behind gpx_pause_groups; also written after the GPU minutes were spent) -------
@pytest.mark.parametrize("R,seed", [(3, 31), (5, 32), (1, 33)])
def test_pause_groups_kernel_equals_oracle(cuda_lib, oracle_lib, R, seed):
    from test_pause_batch import busy_engine
    G = 120
    eg, eo = busy_engine(cuda_lib, G, seed, 1, R), busy_engine(oracle_lib, G, seed, 1, R)
    assert_same_rows(eg, eo, R, G, "before")
    gids = np.random.default_rng(seed).permutation(G + 3).astype(np.uint32)
    rows_g, ok_g = eg.pause_groups(gids)
    rows_o, ok_o = eo.pause_groups(gids)
    assert np.array_equal(ok_g, ok_o) and 0 < ok_o.sum() < len(gids)
    for f in rows_o.dtype.names:
        assert np.array_equal(rows_g[f], rows_o[f]), f
    assert_same_rows(eg, eo, R, G, "after the sweep")
    eg.load_rows(rows_g[ok_g].reshape(-1))  # unpause
    eo.load_rows(rows_o[ok_o].reshape(-1))
    assert_same_rows(eg, eo, R, G, "after unpausing")
    # the engine goes on deciding in the unpaused groups
    from helpers import make_requests
    back = np.sort(gids[ok_o])
    reqs, pay = make_requests(back, payload_len=9, seed=seed, round_no=11)
    sg, xg, _ = eg.round(reqs, pay)
    so, xo, _ = eo.round(reqs, pay)
    assert np.array_equal(sg, so) and np.all(so > 0)
    assert_same_rows(eg, eo, R, G, "after a round")


def test_mass_failover_one_launch(cuda_lib, oracle_lib):
    """a node is lost: every group elects in ONE call (20,000 elections, R = 3, each with a majority of replies carrying
    accepted pvalues); result records and all rows equal the oracle's"""
    R, G = 3, 20000
    rng = np.random.default_rng(5)
    eg, eo = make_engine(cuda_lib, R, G), make_engine(oracle_lib, R, G)
    els = np.zeros(G, dtype=abi.election_dtype)
    els["gid"] = rng.permutation(G)
    els["lane"] = rng.integers(0, R, size=G)
    els["bnum"], els["bcoord"] = 1, np.array([100, 101, 102])[els["lane"]]
    els["slot"] = rng.integers(0, 50, size=G)
    els["first_reply"], els["n_replies"] = np.arange(G) * 2, 2
    reps = np.zeros(2 * G, dtype=abi.prepare_reply_dtype)
    reps["gid"] = np.repeat(els["gid"], 2)
    reps["bnum"], reps["bcoord"] = 1, np.repeat(els["bcoord"], 2)
    reps["first_slot"] = np.repeat(els["slot"], 2) - 1
    for j in range(2):
        acc = (els["lane"] + j) % R
        reps["who"][j::2] = acc | (els["lane"] << 8)
        na = rng.integers(0, 4, size=G)
        reps["n_accepted"][j::2] = na
        for a in range(3):
            sel = np.nonzero(na > a)[0]
            pv = reps["accepted"][j::2][:, a]
            pv["slot"][sel] = els["slot"][sel] + a + j
            pv["bnum"][sel], pv["bcoord"][sel] = 0, 100 + j
            pv["req_id"][sel] = rng.integers(1, 1 << 40, size=len(sel))
            pv["payload_len"][sel], pv["flags"][sel] = 8, 1 << 16
            v = reps["accepted"][j::2]
            v[:, a] = pv
            reps["accepted"][j::2] = v
    got, want = eg.handle_prepare_replies(els, reps), eo.handle_prepare_replies(els, reps)
    assert np.all(want["verdict"] == abi.EL_MAJORITY) and int(want["n_plan"].max()) >= 3
    assert_same_out(got, want)
    assert_same_rows(eg, eo, R, G)


# ---- the journal's index as a scan (k_log_dir / k_log_scan / k_log_hits behind gpx_log_find; same late arrival) -----------
def _same_hits(eg, eo, lane, hg, ho):
    """decision images byte for byte; accept images field by field except payload_off (the position of a batched slot's
    constructed blob inside its segment depends on block scheduling) -- the blob CONTENT is compared"""
    assert hg["decision"].tobytes() == ho["decision"].tobytes()
    for f in abi.accept_dtype.names:
        if f != "payload_off":
            assert np.array_equal(hg["accept"][f], ho["accept"][f]), f
    found = np.argwhere((ho["accept"]["flags"] & abi.F_VOID) == 0)
    for i, k in found[:5]:
        n = int(ho[i, k]["accept"]["payload_len"])
        if n:
            assert bytes(eg.log_read(lane, int(hg[i, k]["blob_pos"]), n)) == bytes(eo.log_read(lane, int(ho[i, k]["blob_pos"]), n))
    if len(found):  # all bodies in one copy on each side (gpx_log_gather)
        sel = (found[:, 0], found[:, 1])
        lens = ho["accept"]["payload_len"][sel]
        assert eg.log_gather(lane, hg["blob_pos"][sel], lens) == eo.log_gather(lane, ho["blob_pos"][sel], lens)
    return len(found)


@pytest.mark.parametrize("seed", [41, 42])
def test_log_find_kernels_equal_oracle(cuda_lib, oracle_lib, seed):
    from test_log_find import logged_engine, random_wants
    G = 40
    (eg, hg_), (eo, ho_) = logged_engine(cuda_lib, G, seed), logged_engine(oracle_lib, G, seed)
    rng = np.random.default_rng(seed)
    found = 0
    for lane in range(3):
        assert len(hg_[lane]) == len(ho_[lane])  # the same calls logged on both sides
        for i in (0, len(ho_[lane]) // 2):  # from the start and from the boundary of a call in the middle
            wants = random_wants(G, rng)
            found += _same_hits(eg, eo, lane, eg.log_find(lane, wants, hg_[lane][i]), eo.log_find(lane, wants, ho_[lane][i]))
    assert found > 50


def test_log_find_on_a_ring_that_wrapped(cuda_lib, oracle_lib):
    """a 128 KiB ring (no back-pressure: the oldest bytes are overwritten), the scan starts at the oldest call boundary that
    is still intact: stale laps and skipped tails are not taken for segments, what is found is what the oracle finds in
    the same calls"""
    from test_log_find import logged_engine, random_wants
    G, ring = 40, 1 << 17
    kw = dict(rounds=70, log_ring_bytes=ring, max_batch_payload=1 << 14)
    (eg, hg_), (eo, ho_) = logged_engine(cuda_lib, G, 43, **kw), logged_engine(oracle_lib, G, 43, **kw)
    rng = np.random.default_rng(43)
    found = 0
    for lane in range(3):
        assert len(hg_[lane]) == len(ho_[lane]) and hg_[lane][-1] > 2 * ring  # wrapped at least twice
        i = next(j for j, p in enumerate(hg_[lane]) if hg_[lane][-1] - p <= ring)
        assert 0 < i < len(hg_[lane]) - 1
        wants = random_wants(G, rng, max_slot=72)
        found += _same_hits(eg, eo, lane, eg.log_find(lane, wants, hg_[lane][i]), eo.log_find(lane, wants, ho_[lane][i]))
        with pytest.raises(abi.GpxError):  # one lap too far back
            eg.log_find(lane, wants, hg_[lane][0])
    assert found > 20


def test_view_changes_with_the_tally_in_the_engine(cuda_lib, oracle_lib):
    from test_paxos_manager import _same_end_state, drive_auto_election, drive_lagging_election, drive_view_change
    _same_end_state(drive_view_change(cuda_lib, p1b=True), drive_view_change(oracle_lib, p1b=True))
    _same_end_state(drive_auto_election(cuda_lib, p1b=True), drive_auto_election(oracle_lib, p1b=True))
    _same_end_state(drive_lagging_election(cuda_lib, p1b=True), drive_lagging_election(oracle_lib, p1b=True))


def test_mass_failover_through_the_mirror(cuda_lib, oracle_lib):
    """PaxosManager.runForCoordinators: one gpx_handle_prepares + one gpx_handle_prepare_replies call for all groups of the
    lost node, then plan entry j of every elected group per round"""
    from test_paxos_manager import _same_end_state
    from test_phase1b import drive_mass_failover
    _same_end_state(drive_mass_failover(cuda_lib, batched=True), drive_mass_failover(oracle_lib, batched=True))


def test_pause_batch_through_the_mirror(cuda_lib, oracle_lib):
    from gigapaxos_b200.paxos_manager import HashChainApp, PaxosManager
    from helpers import Engine, make_config

    def drive(lib):
        eng = Engine(lib, make_config(lib, max_groups=64, max_batch_recs=4096, max_batch_payload=1 << 20))
        pm = PaxosManager(eng, [HashChainApp() for _ in range(3)], [100, 101, 102])
        names = [f"TESTPaxosApp{i}" for i in range(20)]
        pm.createPaxosInstanceBatch({n: None for n in names}, [100, 101, 102])
        for r in range(3):
            for n in names:
                pm.propose(n, f"{n}:{r}".encode())
            pm.run_round()
        pm.propose(names[3], b"queued")
        done = pm.pauseBatch(names[:12])
        table = {n: list(pm.paused[n]) for n in done}
        pm.run_round()
        for n in names:
            assert pm.propose(n, f"{n}:later".encode()) is not None
        pm.run_round()
        assert not pm.paused and all(a.state == pm.apps[0].state for a in pm.apps)
        return pm, table
    (g, tg), (o, to) = drive(cuda_lib), drive(oracle_lib)
    assert tg == to and len(to) == 11  # the same HotRestoreInfo strings in the pause table
    assert g.apps[0].state == o.apps[0].state and g.num_decisions == o.num_decisions


def test_sync_and_lagging_election_with_the_scan(cuda_lib, oracle_lib):
    """the host mirror's catch-up (syncDecisions) and a lagging preparer's logged accepts, both looked up by gpx_log_find"""
    from test_paxos_manager import _same_end_state, drive_lagging_election, drive_sync
    g, o = drive_sync(cuda_lib, p1b=True), drive_sync(oracle_lib, p1b=True)
    assert g.apps[2].state == o.apps[2].state and g.num_decisions == o.num_decisions
    _same_end_state(drive_lagging_election(cuda_lib, p1b=True), drive_lagging_election(oracle_lib, p1b=True))


def test_slow_path_list_end_to_end(cuda_lib, oracle_lib):
    """a replica that missed more decisions than its window holds is flagged, gpx_select_groups names it, the mirror catches
    it up (the donor's journal looked up by gpx_log_find) and gpx_clear_group_flags takes it off the list"""
    from test_paxos_manager import _same_end_state
    from test_pause_batch import drive_flagged_sync
    _same_end_state(drive_flagged_sync(cuda_lib), drive_flagged_sync(oracle_lib))
