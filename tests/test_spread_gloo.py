"""CPU, world_size 3 over gloo: the host logic of SPREAD placement (gigapaxos_b200/spread.py).

Every rank is one node.  The node behind SpreadCluster is a test double that answers the device-resident phase
calls with the CPU oracle on host tensors (the product's nodes are CUDA engines; this file only checks the
orchestration: bucketing contract, count exchange, point-to-point bucket transfer, chunk bookkeeping, the order in
which reply buckets are tallied).  Each rank compares what its node executed and its final state rows with a
single-process oracle run that hosts all nodes as lanes of one engine.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gigapaxos_b200.spread import (CTL_N_ACCEPTS, CTL_N_DECISIONS, CTL_N_EXTRA, K_ACCEPT, K_DECISION, K_REPLY,
                                   DistExchange, SpreadCluster, coordinator_of, members_of)
from helpers import Engine, abi, make_config, make_requests, oracle_library

NODE0 = 100
CPU = torch.device("cpu")


def make_groups(N, G, R):
    node_ids = [NODE0 + i for i in range(N)]
    descs = np.zeros(G, dtype=abi.group_desc_dtype)
    descs["gid"] = np.arange(G)
    descs["n_members"] = R
    descs["init_mode"] = abi.INIT_BATCH
    coord = np.zeros(G, dtype=np.int64)
    member_of = np.zeros((G, N), dtype=bool)
    for g in range(G):
        nm = f"NoopPaxosApp{g}"
        mem = [node_ids[m] for m in members_of(nm, N, R)]
        descs["name_hash"][g] = abi.java_string_hash(nm)
        descs["members"][g, :R] = mem
        coord[g] = coordinator_of(nm, mem) - NODE0
        member_of[g, [m - NODE0 for m in mem]] = True
    return node_ids, descs, coord, member_of


class OracleNode:
    """SpreadNode's interface on host tensors, backed by the oracle's host-buffer phase calls"""

    def __init__(self, index, node_ids, descs, R, G):
        lib = oracle_library()
        self.index, self.node_ids, self.device = index, list(node_ids), CPU
        self.engine = Engine(lib, make_config(lib, max_groups=G, n_lanes=1, lane_node=[node_ids[index]],
                                              max_group_size=R, max_batch_recs=8 * G, max_batch_payload=1 << 22,
                                              checkpoint_interval=3))
        mine = descs[[node_ids[index] in d["members"][: d["n_members"]] for d in descs]]
        self.engine.create_groups(mine)
        self.members = {int(d["gid"]): [int(x) for x in d["members"][: d["n_members"]]] for d in mine}
        n = len(node_ids)
        self.ctl = torch.zeros(8, dtype=torch.int32)
        self.cnt = torch.zeros((n, 2), dtype=torch.int32)
        self.dropped = torch.zeros(1, dtype=torch.int32)
        self.blob_space = np.zeros(0, dtype=np.uint8)

    def _ctl(self):
        return self.ctl.numpy()

    def propose(self, reqs, payload, n, status, accepts):
        acc, blob, st = self.engine.propose(reqs.numpy().view(abi.request_dtype)[:n], payload.numpy())
        accepts.numpy().view(abi.accept_dtype)[: len(acc)] = acc
        status.numpy()[:n] = st
        self._ctl()[CTL_N_ACCEPTS] = len(acc)
        self.blob_space = blob  # [payload arena | batched blobs], what payload_off refers to

    def route(self, kind, recs, n_ptr, n_max, payload, out_recs, cap, out_blob, blob_cap):
        dt = {K_ACCEPT: abi.accept_dtype, K_DECISION: abi.decision_dtype, K_REPLY: abi.reply_dtype}[kind]
        n = int(self._ctl()[(n_ptr - self.ctl.data_ptr()) // 4]) if n_ptr else n_max
        r = recs.numpy().view(dt)[:n]
        N = len(self.node_ids)
        out = out_recs.numpy().view(dt)
        cnt = np.zeros((N, 2), dtype=np.int32)
        ob = out_blob.numpy() if out_blob is not None else None
        for x in r:
            mem = self.members.get(int(x["gid"]))
            if kind == K_REPLY:
                if abi.who_flags(int(x["who"])) & abi.F_VOID:
                    continue
                dests = [mem[abi.who_dst(int(x["who"]))]] if mem else []
            else:
                if int(x["flags"]) & abi.F_VOID:
                    continue
                dests = mem or []
            for node in dests:
                d = self.node_ids.index(node)
                y = x.copy()
                if kind == K_ACCEPT:
                    ln, u = int(x["payload_len"]), (int(x["payload_len"]) + 15) // 16
                    boff = int(cnt[d, 1]) * 16
                    ob[d * blob_cap + boff: d * blob_cap + boff + 16 * u] = 0
                    o = int(x["payload_off"])
                    ob[d * blob_cap + boff: d * blob_cap + boff + ln] = self.blob_space[o: o + ln]
                    y["payload_off"] = boff
                    cnt[d, 1] += u
                out[d * cap + int(cnt[d, 0])] = y
                cnt[d, 0] += 1
        self.cnt = torch.from_numpy(cnt)
        return self.cnt

    def _ingest(self, r):
        r["dst_mask"] = [1 if int(g) in self.members else 0 for g in r["gid"]]

    def accepts(self, recs, n, blob, rec_end, blob_base, replies, extra, extra_cap):
        r = recs.numpy().view(abi.accept_dtype)[:n]
        self._ingest(r)
        lo = 0
        for end, base in zip(rec_end, blob_base):
            r["payload_off"][lo:end] += np.uint32(base)
            lo = end
        rep, ex = self.engine.handle_accepts(r, blob.numpy(), extra_cap=extra_cap)
        assert len(ex) == 0
        replies.numpy().view(abi.reply_dtype)[:n] = rep

    def replies(self, recs, byte_off, n, decisions):
        r = recs.numpy()[byte_off: byte_off + 32 * n].view(abi.reply_dtype)
        dec = self.engine.handle_accept_replies(r)
        k = int(self._ctl()[CTL_N_DECISIONS])
        decisions.numpy().view(abi.decision_dtype)[k: k + len(dec)] = dec
        self._ctl()[CTL_N_DECISIONS] = k + len(dec)

    def decisions(self, recs, n, exec_out, extra, extra_cap):
        r = recs.numpy().view(abi.decision_dtype)[:n]
        self._ingest(r)
        ex, xx = self.engine.handle_decisions(r, extra_cap=extra_cap)
        assert len(xx) == 0
        exec_out.numpy().view(abi.exec_dtype)[:n] = ex
        self._ctl()[CTL_N_EXTRA] = 0


def workload(G, coord, r):
    rng = np.random.default_rng(100 + r)
    counts = rng.choice([0, 1, 1, 2, 3], size=G)
    gids = np.repeat(np.arange(G), counts)
    lens = rng.integers(1, 30, size=len(gids))
    reqs, pay = make_requests(gids, payload_len=lens, seed=77, round_no=r)
    reqs["flags"] = coord[gids].astype(np.uint32) << 8
    reqs["entry_node"] = NODE0 + coord[gids]
    return gids, reqs, pay


def node_batch(reqs, pay, sel):
    rq = reqs[sel].copy()
    rq["flags"] &= ~np.uint32(0xF00)
    stride = ((rq["payload_len"] + 15) // 16) * 16
    offs = np.concatenate([[0], np.cumsum(stride)[:-1]]).astype(np.uint32)
    buf = np.zeros(int(stride.sum()), dtype=np.uint8)
    for k in range(len(rq)):
        o, ln = int(reqs["payload_off"][sel[k]]), int(rq["payload_len"][k])
        buf[offs[k]: offs[k] + ln] = pay[o: o + ln]
    rq["payload_off"] = offs
    return torch.from_numpy(rq.view(np.uint8).copy()), torch.from_numpy(buf), len(rq)


def by_gid(recs):
    r = recs[(recs["flags"] & abi.F_VOID) == 0]
    return r[np.argsort(r["gid"], kind="stable")]


def worker(rank, world, port, q, G, R, rounds):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        node_ids, descs, coord, member_of = make_groups(world, G, R)
        node = OracleNode(rank, node_ids, descs, R, G)
        cluster = SpreadCluster([node], DistExchange(), world)
        # the reference run: all nodes as lanes of ONE oracle engine, phase by phase
        lib = oracle_library()
        ref = Engine(lib, make_config(lib, max_groups=G, n_lanes=world, lane_node=node_ids, max_group_size=R,
                                      max_batch_recs=8 * G, max_batch_payload=1 << 22, checkpoint_interval=3))
        ref.create_groups(descs)
        n_exec = 0
        for r in range(rounds):
            gids, reqs, pay = workload(G, coord, r)
            acc, blob, so = ref.propose(reqs, pay)
            rep, _ = ref.handle_accepts(acc, blob)
            dec = ref.handle_accept_replies(rep)
            xo, _ = ref.handle_decisions(dec)
            sel = np.nonzero(coord[gids] == rank)[0]
            batches = {rank: node_batch(reqs, pay, sel)} if len(sel) else {}
            s = cluster.round(batches)[rank]
            if len(sel):
                assert np.array_equal(s["status"].numpy()[: len(sel)], so[sel])
            got = by_gid(s["exec"].numpy().view(abi.exec_dtype)[: s["n_exec"]])
            want = by_gid(xo[((xo["flags"] >> 12) & 0xF) == rank])
            assert len(got) == len(want)
            for f in ("gid", "slot", "req_id"):
                assert np.array_equal(got[f], want[f]), f
            assert np.array_equal(got["flags"] & ~np.uint32(0xF000), want["flags"] & ~np.uint32(0xF000))
            n_exec += len(got)
        g = np.nonzero(member_of[:, rank])[0]
        ro, rg = ref.dump_rows(g, rank), node.engine.dump_rows(g, 0)
        for f in ro.dtype.names:
            if f != "lane":
                assert np.array_equal(ro[f], rg[f]), f
        dist.barrier()
        q.put((rank, "ok", n_exec))
        dist.destroy_process_group()
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "fail: " + traceback.format_exc(), 0))
        raise e


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,R", [(3, 3), (2, 2)])
def test_spread_cluster_over_gloo(world, R):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    G, rounds = 40, 4
    procs = [ctx.Process(target=worker, args=(r, world, port, q, G, R, rounds)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, status, n_exec in res:
        assert status == "ok", f"rank {rank}: {status}"
        assert n_exec > 0
    assert all(p.exitcode == 0 for p in procs)
