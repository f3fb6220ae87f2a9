"""CPU: adversarial schedules against the oracle alone -- Paxos safety (agreement, gap-free
in-order execution, TESTPaxosApp.java:179-213) must hold under NACKs, preemption, duplicates,
reordering, lost multicasts, commits overtaking accepts, STOPs and coordinator changes; and the
bounded-window engine rules (W=8) must agree with the unbounded Java maps (W=0) whenever no
overflow was flagged."""
import numpy as np
import pytest

from fuzz import Fuzzer


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
@pytest.mark.parametrize("R,fused", [(3, 0.0), (5, 0.0), (3, 0.5), (5, 1.0)])
def test_safety_under_adversarial_schedules(oracle_lib, seed, R, fused):
    f = Fuzzer([oracle_lib], G=48, R=R, W=8, seed=seed)
    decided = f.run(steps=50, fused_prob=fused)
    assert decided > 100
    c = f.engines[0].counters()
    assert c["accepts_nacked"] > 0 and c["preempted"] >= 0
    assert fused == 1.0 or c["placeholders"] > 0
    f.close()


def test_unbounded_and_windowed_oracle_agree(oracle_lib):
    """Same schedule through W=0 (TreeMap semantics) and W=8; identical unless overflow fired."""
    class Two(Fuzzer):
        pass
    f = Fuzzer([oracle_lib, oracle_lib], G=40, R=3, W=8, seed=11)
    # re-create engine 0 unbounded
    from helpers import Engine, group_descs, make_config
    f.engines[0].close()
    f.engines[0] = Engine(oracle_lib, make_config(oracle_lib, max_groups=40, window=0, max_batch_recs=1 << 14,
                                                  max_batch_payload=1 << 20))
    f.engines[0].create_groups(group_descs(40))
    try:
        f.run(steps=40, rival=False, view_changes=False, stop_prob=0.0)
    except AssertionError as e:
        # only acceptable divergence: the bounded engine dropped records beyond the window
        c = f.engines[1].counters()
        assert c["window_overflow"] > 0, str(e)
    f.close()


@pytest.mark.parametrize("seed,R,fn", [(31, 3, "round"), (32, 5, "round"), (33, 3, "round_phases")])
def test_fused_rounds_inside_adversarial_schedules(oracle_lib, seed, R, fn):
    """whole rounds (gpxo_round) interleaved with lossy / duplicated / reordered phase steps, rival coordinators
    and view changes: safety holds, and the windowed and the unbounded oracle agree"""
    f = Fuzzer([oracle_lib], G=48, R=R, W=8, seed=seed)
    assert f.run(steps=60, fused_prob=0.3, round_prob=0.5, round_fn=fn) > 100
    f.close()


@pytest.mark.parametrize("seed,R", [(61, 3), (62, 5)])
def test_prepares_inside_adversarial_schedules(oracle_lib, seed, R):
    """phase 1a (PISM.handlePrepare) interleaved with everything else: ballots only grow, NACKs carry no pvalues,
    accepted pvalues come back in slot order, safety still holds"""
    f = Fuzzer([oracle_lib], G=32, R=R, W=8, seed=seed)
    assert f.run(steps=60, fused_prob=0.3, round_prob=0.3, prepares=True) > 50
    f.close()


def test_handle_prepare_semantics(oracle_lib):
    """PaxosAcceptor.handlePrepare :239-275 case by case"""
    from helpers import Engine, abi, group_descs, make_config, make_requests
    e = Engine(oracle_lib, make_config(oracle_lib, max_groups=4, journaling_enabled=0))
    e.create_groups(group_descs(4))
    reqs, pay = make_requests(np.repeat(np.arange(4), 1), payload_len=5)
    acc, blob, _ = e.propose(reqs, pay)
    acc["dst_mask"] = 0b011  # lane 2 misses the ACCEPTs: nothing gets decided at lane 2
    e.handle_accepts(acc, blob)
    row = e.dump_rows(np.array([1], dtype=np.uint32), 0)[0]
    coord = int(row["acc_bcoord"])
    p = np.zeros(3, dtype=abi.decision_dtype)
    p["gid"], p["flags"], p["dst_mask"] = 1, abi.F_PREPARE, 0b111
    p["slot"] = [1, 1, 3]            # firstUndecidedSlot
    p["bnum"] = [5, 2, 5]            # raise, stale (NACK), equal
    p["bcoord"] = [101, 102, 101]
    out = e.handle_prepares(p).reshape(3, 3)
    first = out[0]
    assert np.all(first["bnum"] == 5) and np.all(first["bcoord"] == 101)
    fl = abi.who_flags(first["who"])
    assert np.all(fl & abi.F_LOGGED) and not np.any(fl & abi.F_NACK)  # the promise is logged before the reply
    assert first["n_accepted"].tolist() == [1, 1, 0] and first["accepted"]["slot"][0, 0] == 1
    assert first["accepted"]["bcoord"][0, 0] == coord and first["accepted"]["payload_len"][0, 0] == 5
    stale = out[1]
    assert np.all(abi.who_flags(stale["who"]) & abi.F_NACK) and np.all(stale["n_accepted"] == 0)
    assert np.all(stale["bnum"] == 5)  # the NACK carries the higher ballot
    eq = out[2]
    assert not np.any(abi.who_flags(eq["who"]) & (abi.F_NACK | abi.F_LOGGED)) and np.all(eq["n_accepted"] == 0)
    assert np.all(eq["first_slot"] == 2)  # max(gcSlot, firstUndecidedSlot - 1)
    segs = abi.parse_log(e.log_read(0))
    assert int(segs[-1][0]["type"]) == abi.F_PREPARE
    img = segs[-1][1]
    assert [int(x) & abi.F_VOID for x in img["flags"]] == [0, abi.F_VOID, abi.F_VOID]  # only the raise is logged
