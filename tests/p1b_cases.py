"""Random elections for gpx_handle_prepare_replies (phase 1b): engine preconditions + election records + PREPARE_REPLY
record streams.  Shared by the CPU test (oracle entry point vs the host-language twin) and the GPU test (CUDA kernel vs
oracle)."""
from __future__ import annotations

import numpy as np

from helpers import Engine, abi, group_descs, make_config

NODES5 = [100, 101, 102, 103, 104]


def make_engine(lib, R: int, G: int, lane_nodes=None):
    """lane_nodes: the nodes this engine hosts (default: all R members, lane l = member l).  A single-lane engine that hosts
    one member of every group is a node of the spread placement: the other acceptors' replies come in as records."""
    nodes = NODES5[:R]
    lane_nodes = nodes if lane_nodes is None else list(lane_nodes)
    eng = Engine(lib, make_config(lib, n_lanes=len(lane_nodes), max_groups=G, max_group_size=R, max_batch_recs=4096,
                                  max_batch_payload=1 << 20, lane_node=lane_nodes))
    eng.create_groups(group_descs(G, members=tuple(nodes)))
    return eng


def preconditions(eng, R: int, G: int, rng) -> None:
    """coordinators of assorted ballots on assorted lanes, a few stopped acceptors, a few destroyed groups"""
    pts = []
    L = eng.n_lanes
    for gid in range(G):
        for l in range(L):
            k = int(rng.integers(0, 6))
            if k == 0:
                pts.append((gid, l, abi.PATCH_RESIGN_COORD, 0, 0, 0, 0))
            elif k == 1:
                pts.append((gid, l, abi.PATCH_INSTALL_COORD, int(rng.integers(0, 9)), int(eng.cfg.lane_node[l]), int(rng.integers(0, 30)),
                            int(rng.integers(0, 2))))
        if rng.integers(0, 40) == 0:
            pts.append((gid, int(rng.integers(0, L)), abi.PATCH_SET_STATE, abi.ST_STOPPED, 0, 0, 0))
    p = np.zeros(len(pts), dtype=abi.patch_dtype)
    for i, t in enumerate(pts):
        p[i]["gid"], p[i]["lane"], p[i]["op"], p[i]["a"], p[i]["b"], p[i]["c"], p[i]["d"] = t
    if len(p):
        eng.patch(p)
    dead = [g for g in range(G) if rng.integers(0, 60) == 0]
    if dead:
        eng.destroy_groups(dead)


def random_elections(R: int, G: int, rng, wrap: bool = False, lane_nodes=None):
    """-> (election records, reply records).  One election per group (a random subset of the groups, shuffled)."""
    lane_nodes = NODES5[:R] if lane_nodes is None else list(lane_nodes)
    L = len(lane_nodes)
    gids = [g for g in rng.permutation(G) if rng.integers(0, 10) < 8]
    els = np.zeros(len(gids), dtype=abi.election_dtype)
    recs = []
    base = 0x7FFFFFF0 if wrap else 0  # slots straddling the int wrap
    jint = lambda v: ((int(v) + (1 << 31)) % (1 << 32)) - (1 << 31)
    for i, gid in enumerate(gids):
        lane = L if rng.integers(0, 16) == 0 else int(rng.integers(0, L))  # L = not a lane: dropped
        my = (int(rng.integers(1, 8)), lane_nodes[lane % L])
        fus = int(rng.integers(0, 20))
        els[i]["gid"], els[i]["lane"], els[i]["bnum"], els[i]["bcoord"] = gid, lane, my[0], my[1]
        els[i]["slot"] = jint(base + fus)
        els[i]["first_reply"] = len(recs)
        style = int(rng.integers(0, 10))
        n_logical = int(rng.integers(0, R + 3))
        order = list(rng.permutation(R)) + [int(rng.integers(0, R + 2)) for _ in range(3)]
        for k in range(n_logical):
            idx = int(order[k % len(order)])
            r = np.zeros(1, dtype=abi.prepare_reply_dtype)[0]
            r["gid"] = gid
            c = int(rng.integers(0, 20))
            ballot = my
            if c == 0:
                ballot = (my[0] + int(rng.integers(0, 2)), my[1] + 1)  # higher: preempts
            elif c == 1:
                ballot = (my[0] - 1, my[1] + int(rng.integers(-1, 2)))  # lower: ignored
            elif c == 2:
                ballot = (my[0], my[1] - 1)
            r["bnum"], r["bcoord"] = ballot
            gc = fus - 1 + int(rng.integers(0, 4)) * int(rng.integers(0, 2))
            r["first_slot"] = jint(base + gc)
            flags = abi.F_VOID if rng.integers(0, 25) == 0 else 0
            # accepted pvalues: distinct slots around fus, ballots below mine
            if style < 2:
                n_acc = 0
            elif style < 8:
                n_acc = int(rng.integers(0, 3 + style))
            else:
                n_acc = int(rng.integers(0, 28))  # continuation records; sometimes beyond GPX_MAX_CARRY in total
            span = 10 if style < 8 else (40 if style == 9 else 14)
            slots = sorted(rng.choice(np.arange(fus - 2, fus + span), size=min(n_acc, span + 2), replace=False).tolist())
            pvs = []
            for sl in slots:
                pv = np.zeros(1, dtype=abi.accepted_pvalue_dtype)[0]
                pv["slot"] = jint(base + sl)
                pv["bnum"], pv["bcoord"] = int(rng.integers(0, my[0] + 1)), NODES5[int(rng.integers(0, R))]
                if pv["bnum"] == my[0]:
                    pv["bcoord"] = my[1] - 1 - int(rng.integers(0, 2))
                pv["frame_ref"] = int(rng.integers(0, 1 << 20))
                pv["req_id"] = int(rng.integers(1, 1 << 40))
                pv["payload_len"] = int(rng.integers(0, 200))
                pv["flags"] = (2 if rng.integers(0, 9) == 0 else 0) | (int(rng.integers(1, 4)) << 16)
                pvs.append(pv)
            chunks = [pvs[j: j + abi.GPX_MAX_WINDOW] for j in range(0, len(pvs), abi.GPX_MAX_WINDOW)] or [[]]
            for ci, ch in enumerate(chunks):
                q = r.copy()
                more = abi.F_MORE if ci + 1 < len(chunks) else 0
                q["who"] = abi.who(idx if idx < R or rng.integers(0, 2) else 0xFF, NODES5.index(my[1]), flags | more)
                q["n_accepted"] = len(ch)
                for j, pv in enumerate(ch):
                    q["accepted"][j] = pv
                recs.append(q)
        els[i]["n_replies"] = len(recs) - int(els[i]["first_reply"])
    reps = np.array(recs, dtype=abi.prepare_reply_dtype) if recs else np.zeros(0, dtype=abi.prepare_reply_dtype)
    return els, reps


def dump_all(eng, R: int, G: int):
    gids = np.arange(G, dtype=np.uint32)
    return [eng.dump_rows(gids, l) for l in range(eng.n_lanes)]


def assert_same_out(a: np.ndarray, b: np.ndarray):
    """two arrays of election_out records, field by field (plan entries beyond n_plan and the pvalue of a non-PVALUE
    entry are zero on both sides)"""
    assert len(a) == len(b)
    for f in ("gid", "verdict", "next_slot", "n_plan", "flags", "node_slots"):
        assert np.array_equal(a[f], b[f]), (f, a[f], b[f])
    assert a.tobytes() == b.tobytes()


def assert_same_rows(eng_a, eng_b, R: int, G: int, ctx=None):
    """every field of every row of every lane; rows of destroyed groups (state FREE) only agree on being FREE"""
    for ra, rb in zip(dump_all(eng_a, R, G), dump_all(eng_b, R, G)):
        assert np.array_equal(ra["state"], rb["state"]), (ctx, "state")
        live = rb["state"] != abi.ST_FREE
        for f in ra.dtype.names:
            assert np.array_equal(ra[f][live], rb[f][live]), (ctx, f)
