"""Phase 1b at the would-be coordinator (gpx_handle_prepare_replies; PISM.handlePrepareReply :1017-1068, PCS :264-587).

CPU: the oracle's entry point (a C++ restatement of the Java classes) against the host-language twin in
gigapaxos_b200/paxos_manager.py (tally_prepare_replies / combine_carryover + gpx_patch), a second, independently
structured restatement -- verdict, nodeSlotNumbers, plan and the rows both leave behind, over random elections."""
import numpy as np
import pytest

from gigapaxos_b200.paxos_manager import NoopPaxosApp, PaxosManager
from helpers import abi
from p1b_cases import NODES5, assert_same_out, dump_all, make_engine, preconditions, random_elections


def twin_election(pm: PaxosManager, R: int, el, reps: np.ndarray):
    """the host-language twin on one election: -> election_out record (and the engine of `pm` patched as it patches)"""
    out = np.zeros(1, dtype=abi.election_out_dtype)[0]
    out["gid"], out["verdict"] = el["gid"], abi.EL_DROPPED
    out["node_slots"][:] = -1
    gid, lane = int(el["gid"]), int(el["lane"])
    eng = pm.engine
    if lane >= eng.n_lanes:
        return out
    row = eng.dump_rows(np.array([gid], dtype=np.uint32), lane)[0]
    if int(row["state"]) not in (abi.ST_ACTIVE_1, abi.ST_ACTIVE_2):
        return out
    # logical replies: a record and its GPX_F_MORE continuations (the twin takes the overflow as `logged`)
    logical, logged, first_rec = [], {}, []
    k, end = int(el["first_reply"]), int(el["first_reply"]) + int(el["n_replies"])
    while k < end:
        head = reps[k].copy()
        fl = abi.who_flags(int(head["who"]))
        extra = []
        first_rec.append(k)
        k += 1
        while (fl & abi.F_MORE) and not (fl & abi.F_VOID) and k < end:
            nxt = reps[k]
            extra += [(pv.copy(), k) for pv in nxt["accepted"][: int(nxt["n_accepted"])]]
            fl = abi.who_flags(int(nxt["who"]))
            k += 1
        logical.append(head)
        if extra:
            logged[len(logical) - 1] = extra
    my = (int(el["bnum"]), int(el["bcoord"]))
    # the twin reports the index of the logical reply; the engine the index of the RECORD: translate
    src_of = {}
    lg = {}
    for l, ex in logged.items():
        lg[l] = [pv for pv, _ in ex]
        for pv, rec in ex:
            src_of[(l, int(pv["slot"]))] = rec
    verdict, ns, carry = PaxosManager.tally_prepare_replies(logical, R, my, lg)
    out["node_slots"][:R] = ns
    out["verdict"] = {"waiting": abi.EL_WAITING, "majority": abi.EL_MAJORITY, "preempted": abi.EL_PREEMPTED,
                      "overflow": abi.EL_OVERFLOW}[verdict]
    if verdict != "majority":
        return out
    comb = PaxosManager.combine_carryover(carry, ns, int(el["slot"]))
    if comb is None:
        out["verdict"] = abi.EL_OVERFLOW
        return out
    plan, next_slot, flags = comb
    out["next_slot"], out["n_plan"], out["flags"] = next_slot, len(plan), flags
    for j, (sl, kind, pv, l) in enumerate(plan):
        c = out["plan"][j]
        c["slot"], c["kind"] = sl, kind
        if kind == abi.CO_PVALUE:
            c["pv"] = pv
            c["src_reply"] = src_of.get((l, int(pv["slot"])), first_rec[l])
    # ... and installs through gpx_patch exactly as PaxosManager._phase1b_host does
    res = pm._phase1b_host(gid, lane, R, my, int(el["slot"]), logical, lg)
    assert res is not None
    return out


@pytest.mark.parametrize("R,seed,wrap", [(3, 1, False), (3, 2, True), (5, 3, False), (5, 4, True), (1, 5, False),
                                         (4, 6, False)])
def test_oracle_phase1b_equals_the_host_twin(oracle_lib, R, seed, wrap):
    G = 160
    rng = np.random.default_rng(seed)
    ea, eb = make_engine(oracle_lib, R, G), make_engine(oracle_lib, R, G)
    pm = PaxosManager(eb, [NoopPaxosApp() for _ in range(R)], NODES5[:R])
    verdicts = set()
    for rnd in range(4):
        st = rng.bit_generator.state
        preconditions(ea, R, G, rng)
        rng.bit_generator.state = st
        preconditions(eb, R, G, rng)
        els, reps = random_elections(R, G, rng, wrap)
        got = ea.handle_prepare_replies(els, reps)
        want = np.array([twin_election(pm, R, el, reps) for el in els], dtype=abi.election_out_dtype)
        assert_same_out(got, want)
        verdicts |= set(int(v) for v in got["verdict"])
        for ra, rb in zip(dump_all(ea, R, G), dump_all(eb, R, G)):
            for f in ra.dtype.names:
                assert np.array_equal(ra[f], rb[f]), f
    if R >= 3:
        assert verdicts == {abi.EL_WAITING, abi.EL_MAJORITY, abi.EL_PREEMPTED, abi.EL_DROPPED, abi.EL_OVERFLOW}


def test_phase1b_api_misuse(oracle_lib):
    eng = make_engine(oracle_lib, 3, 8)
    els = np.zeros(2, dtype=abi.election_dtype)
    els["gid"] = [1, 1]
    with pytest.raises(abi.GpxError):  # one election per group per call
        eng.handle_prepare_replies(els, np.zeros(0, dtype=abi.prepare_reply_dtype))
    els["gid"] = [1, 2]
    els["n_replies"] = [0, 3]
    with pytest.raises(abi.GpxError):  # replies out of range
        eng.handle_prepare_replies(els, np.zeros(2, dtype=abi.prepare_reply_dtype))
    assert len(eng.handle_prepare_replies(els[:0], np.zeros(0, dtype=abi.prepare_reply_dtype))) == 0


def test_elected_coordinator_proposes_the_plan(oracle_lib):
    """end to end on one group: two ACCEPTs reach a minority / a majority, the coordinator dies, the next node is elected
    by gpx_handle_prepare_replies and re-proposes the plan; the decided values are the carried-over ones, in slot order"""
    from test_paxos_manager import drive_view_change
    pm = drive_view_change(oracle_lib, p1b=True)
    assert pm.device_phase1b
