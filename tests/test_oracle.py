"""CPU tests of the oracle: the reference's own self-checking tests (transliterated inside
oracle/gpx_oracle.cpp::gpxo_selftest), the one literal-valued reference test, and the RSM
invariants the reference's TESTPaxosApp asserts (testing/TESTPaxosApp.java:179-213)."""
import ctypes as C

import numpy as np
import pytest

from helpers import Engine, abi, exec_by_lane, group_descs, make_config, make_requests


def test_reference_selftests(oracle_lib):
    """PaxosAcceptor.java:749-776, PaxosCoordinatorState.java:1179-1214, WaitforUtility.java:147-163,
    PaxosPacketBatcher.java:556-567, HotRestoreInfo.java:159-175, RFC 1321 A.5."""
    f = oracle_lib.fn("selftest")
    f.argtypes = [C.c_uint64]
    for seed in (1, 2, 12345, 987654321):
        rc = f(seed)
        assert rc == 0, f"selftest check {rc}: {oracle_lib.last_error()}"


def test_hot_restore_info_literal(oracle_lib):
    """The only literal-valued JUnit test near the path: HotRestoreInfoTest.testToStringAndBack."""
    out = C.create_string_buffer(256)
    s = b"paxos0|2|[1,4,67]|5|3:4|3|45:67|34|[1,3,5]"  # Util.arrayOfIntToString (utils/Util.java:241-248): no blanks
    assert oracle_lib.fn("hri_roundtrip")(s, out, C.c_size_t(256)) == 0
    assert out.value == s
    s2 = b"name|0|[100,101,102]|1|0:101|-1|null|-1|null"
    assert oracle_lib.fn("hri_roundtrip")(s2, out, C.c_size_t(256)) == 0
    assert out.value == s2
    # Util.stringToIntArray :184-193 strips brackets and white space: a blank-separated array parses to the same
    assert oracle_lib.fn("hri_roundtrip")(b"paxos0|2|[1, 4, 67]|5|3:4|3|45:67|34|[1, 3, 5]", out, C.c_size_t(256)) == 0
    assert out.value == s


def test_java_helpers(oracle_lib):
    h = oracle_lib.fn("java_string_hash")
    h.restype = C.c_int32
    h.argtypes = [C.c_char_p, C.c_size_t]
    for s in ["", "a", "paxos0", "NoopPaxosApp123", "TESTPaxosApp0", "x" * 100]:
        assert h(s.encode(), len(s)) == abi.java_string_hash(s)
    assert abi.java_string_hash("paxos0") == -995235643
    rr = oracle_lib.fn("round_robin_coordinator")
    rr.restype = C.c_int32
    m = (C.c_int32 * 3)(100, 101, 102)
    for s in ["NoopPaxosApp0", "NoopPaxosApp1", "paxos0"]:
        hh = abi.java_string_hash(s)
        assert rr(C.c_int32(hh), m, C.c_int32(3), C.c_int32(0)) == [100, 101, 102][abs(hh) % 3]
    lcs = oracle_lib.fn("last_checkpoint_slot")
    lcs.restype = C.c_int32
    assert lcs(C.c_int32(805), C.c_int32(400)) == 800
    assert lcs(C.c_int32(399), C.c_int32(400)) == 0


def test_create_modes(oracle_lib):
    """Batch creation (createHRI: gc -1, nodeSlots 0) vs default creation (gc 0, nodeSlots -1)."""
    e = Engine(oracle_lib, make_config(oracle_lib, max_groups=8))
    d = group_descs(4, init_mode=abi.INIT_BATCH)
    d2 = group_descs(4, init_mode=abi.INIT_DEFAULT, gid0=4, prefix="Other")
    e.create_groups(d)
    e.create_groups(d2)
    for gid in range(8):
        name = (f"NoopPaxosApp{gid}" if gid < 4 else f"Other{gid - 4}")
        coord = [100, 101, 102][abs(abi.java_string_hash(name)) % 3]
        for lane in range(3):
            r = e.dump_rows([gid], lane)[0]
            assert r["acc_slot"] == 1 and r["acc_bnum"] == 0 and r["acc_bcoord"] == coord
            assert r["acc_gc_slot"] == (-1 if gid < 4 else 0)
            assert r["state"] == abi.ST_ACTIVE_1
            is_coord = [100, 101, 102][lane] == coord
            assert bool(r["coord_exists"]) == is_coord
            if is_coord:
                assert r["next_proposal_slot"] == 1 and r["coord_active"] == 1
                assert list(r["node_slots"][:3]) == ([0, 0, 0] if gid < 4 else [-1, -1, -1])


@pytest.mark.parametrize("W", [0, 4, 8])
def test_rsm_invariant_rounds(oracle_lib, W):
    """Every replica executes every slot in order, gap-free, with identical requests
    (TESTPaxosApp.java:190 seqnum == slot; RSMInvariant :382-396)."""
    G = 200
    e = Engine(oracle_lib, make_config(oracle_lib, max_groups=G, window=W))
    e.create_groups(group_descs(G))
    seqs = [dict() for _ in range(3)]
    for r in range(12):
        reqs, pay = make_requests(np.arange(G), payload_len=3, seed=2, round_no=r, entry_lane=r % 3)
        st, ex, extra = e.round(reqs, pay)
        assert np.all(st == r + 1) and len(extra) == 0
        for l, xs in enumerate(exec_by_lane(ex, 3)):
            assert len(xs) == G
            assert np.array_equal(xs["slot"], np.full(G, r + 1))
            assert np.array_equal(xs["req_id"], reqs["req_id"])
            for x in xs:
                seqs[l].setdefault(int(x["gid"]), []).append((int(x["slot"]), int(x["req_id"])))
    assert seqs[0] == seqs[1] == seqs[2]
    c = e.counters()
    assert c["decisions_made"] == 12 * G and c["executed"] == 36 * G and c["accepts_nacked"] == 0
    # medianCP / gc: after round r every acceptor's gc has caught up to slot-2 or better
    rows = e.dump_rows(np.arange(G), 1)
    assert np.all(rows["acc_slot"] == 13) and np.all(rows["acc_gc_slot"] >= 10)


def test_window_mode_equals_unbounded(oracle_lib):
    """With in-flight depth <= W the bounded-window rules never fire: W=0 (Java TreeMaps) and W=8 agree."""
    G = 64
    outs = []
    for W in (0, 8):
        e = Engine(oracle_lib, make_config(oracle_lib, max_groups=G, window=W))
        e.create_groups(group_descs(G))
        rng = np.random.default_rng(7)
        acc = []
        for r in range(8):
            cnt = rng.integers(1, 4, size=G)
            reqs, pay = make_requests(np.repeat(np.arange(G), cnt), payload_len=rng.integers(1, 30), seed=4,
                                      round_no=r)
            st, ex, extra = e.round(reqs, pay)
            acc.append((st.copy(), [x.copy() for x in exec_by_lane(ex, 3)]))
        outs.append((acc, e.dump_rows(np.arange(G), 0), e.counters()))
    for (s0, x0), (s1, x1) in zip(outs[0][0], outs[1][0]):
        assert np.array_equal(s0, s1)
        for a, b in zip(x0, x1):
            assert np.array_equal(a, b)
    assert np.array_equal(outs[0][1], outs[1][1]) and outs[0][2] == outs[1][2]


def test_stop_and_refusal(oracle_lib):
    """STOP: executed() flips the acceptor to STOPPED (PaxosAcceptor.java:462-474); proposals after an
    outstanding STOP are refused (PaxosCoordinatorState.java:235-239); stopped groups drop packets."""
    G = 4
    e = Engine(oracle_lib, make_config(oracle_lib, max_groups=G))
    e.create_groups(group_descs(G))
    reqs, pay = make_requests([0, 0, 1, 2, 3], payload_len=2, stop_mask=[1, 0, 0, 0, 0])
    e.cfg.batching_enabled = 0
    e2 = Engine(oracle_lib, make_config(oracle_lib, max_groups=G, batching_enabled=0))
    e2.create_groups(group_descs(G))
    st, ex, _ = e2.round(reqs, pay)
    assert st[0] == 1 and st[1] == abi.RS_REFUSED_STOP and list(st[2:]) == [1, 1, 1]
    xs = exec_by_lane(ex, 3)
    for l in range(3):
        stops = xs[l][xs[l]["gid"] == 0]
        assert len(stops) == 1 and stops[0]["flags"] & abi.F_STOP and stops[0]["flags"] & abi.F_CKPT
        assert e2.dump_rows([0], l)[0]["state"] == abi.ST_STOPPED
    reqs2, pay2 = make_requests([0, 1], payload_len=2, round_no=1)
    st2, ex2, _ = e2.round(reqs2, pay2)
    assert st2[0] == abi.RS_DROPPED and st2[1] == 2
    c = e2.counters()
    assert c["stops_executed"] == 3 and c["requests_rejected"] == 2


def test_checkpoint_flag(oracle_lib):
    """PISM.shouldCheckpoint :2037-2041: slot % CPI == 0."""
    e = Engine(oracle_lib, make_config(oracle_lib, max_groups=2, checkpoint_interval=3))
    e.create_groups(group_descs(2))
    flags = []
    for r in range(7):
        reqs, pay = make_requests([0, 1], payload_len=1, round_no=r)
        _, ex, _ = e.round(reqs, pay)
        flags.append(bool(exec_by_lane(ex, 3)[0][0]["flags"] & abi.F_CKPT))
    assert flags == [False, False, True, False, False, True, False]
