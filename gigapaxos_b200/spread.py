"""Spread placement: the replicas of a group live on DIFFERENT GPUs.

One engine per GPU hosts ONE node (one lane).  A group with members {m0 < m1 < m2} has its coordinator
(PaxosInstanceStateMachine.roundRobinCoordinator :2251-2256) on one of them; requests enter at the
coordinator's GPU.  A round is the reference's message flow with one exchange per inter-replica packet
type:

    propose (k_propose)                         coordinator's engine
      -> route ACCEPTs by member node (k_route) -> all-to-all -> handleAccept (k_ingest + k_accept) everywhere
      -> route ACCEPT_REPLYs to the coordinator -> all-to-all -> handleAcceptReply (k_tally), one bucket of
                                                   one acceptor at a time, in node order
      -> route DECISIONs by member node         -> all-to-all -> handleBatchedCommit / execute (k_commit)

Records stay in HBM from the request batch to the EXEC records; the all-to-all moves device buffers (NCCL over
NVLink, `NcclExchange`, one process per GPU) or -- for several engines inside one process, e.g. all on one GPU in
the single-GPU test-suite -- plain device copies (`LocalExchange`).  The host only reads the per-destination
counts (one small device->host copy per exchange) to size the transfers.

torch is used for device memory, streams and torch.distributed; every kernel is the engine's own.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import abi
from .abi import Engine, Library, java_string_hash

K_ACCEPT, K_DECISION, K_REPLY = abi.F_ACCEPT, abi.F_DECISION, 0
REC_BYTES = {K_ACCEPT: 48, K_DECISION: 32, K_REPLY: 32}
CTL_N_ACCEPTS, CTL_N_DECISIONS, CTL_N_EXTRA = 0, 1, 2  # uint32 indices into gpx_dev_ctl


def members_of(name: str, n_nodes: int, n_replicas: int) -> List[int]:
    """node indices of a group's replicas: (home + j) mod n_nodes, sorted (PISM ctor sorts members :205)"""
    h = java_string_hash(name)
    a = -h if h < 0 else h
    if a >= (1 << 31):
        a = 0
    home = a % n_nodes
    return sorted((home + j) % n_nodes for j in range(n_replicas))


def coordinator_of(name: str, member_nodes: Sequence[int]) -> int:
    """PISM.roundRobinCoordinator(0): members[|hash(paxosID)| % R] (node id)"""
    h = java_string_hash(name)
    a = -h if h < 0 else h  # Math.abs; abs(Integer.MIN_VALUE) stays negative in Java, the engine takes |.| of the index
    return member_nodes[(a % (1 << 31) if a < (1 << 31) else (1 << 31)) % len(member_nodes)]


class LocalExchange:
    """all nodes are engines of this process: the all-to-all is a set of device copies"""

    def __init__(self, n_nodes: int):
        self.n = n_nodes
        self.local = list(range(n_nodes))

    def counts(self, per_node: List[torch.Tensor]) -> List[np.ndarray]:
        """per_node[i] = int32 [n_nodes, k] (what node i sends to each destination); returns for every local node
        the [n_nodes, k] array of what it RECEIVES from each source (one host sync)"""
        host = torch.stack(per_node).cpu().numpy()  # [src, dst, k]
        return [host[:, j, :].copy() for j in range(self.n)], [host[i] for i in range(self.n)]

    def all_to_all(self, send: List[List[torch.Tensor]], recv: List[List[torch.Tensor]]):
        for j in range(self.n):
            for i in range(self.n):
                if recv[j][i].numel():
                    recv[j][i].copy_(send[i][j], non_blocking=True)


class DistExchange:
    """one process per node (torch.distributed; backend nccl over NVLink on GPUs, gloo in the CPU tests):
    node index == rank.  Counts travel by all_gather, the buckets by one batch of point-to-point
    sends/receives straight out of / into the bucket views (no packing copy)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.n = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.local = [self.rank]

    def counts(self, per_node):
        mine = per_node[0].contiguous()  # [n, k] what I send to each destination
        allc = [torch.empty_like(mine) for _ in range(self.n)]
        self.dist.all_gather(allc, mine, group=self.group)
        host = torch.stack(allc).cpu().numpy()  # [src, dst, k]
        return [host[:, self.rank, :].copy()], [host[self.rank]]

    def all_to_all(self, send, recv):
        dist, me = self.dist, self.rank
        ops = []
        for p in range(self.n):
            if p == me:
                continue
            if recv[0][p].numel():
                ops.append(dist.P2POp(dist.irecv, recv[0][p], p, self.group))
            if send[0][p].numel():
                ops.append(dist.P2POp(dist.isend, send[0][p], p, self.group))
        if recv[0][me].numel():
            recv[0][me].copy_(send[0][me], non_blocking=True)  # loopback
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()


NcclExchange = DistExchange


def _on(dev: torch.device):
    """device context (no-op for the CPU nodes of the test-suite)"""
    import contextlib
    return torch.cuda.device(dev) if dev.type == "cuda" else contextlib.nullcontext()


class SpreadNode:
    """one node = one engine with a single lane on one device"""

    def __init__(self, lib: Library, node_index: int, node_ids: Sequence[int], device: torch.device,
                 max_groups: int, max_batch: int, max_payload: int, **cfg_overrides):
        self.index = node_index
        self.node_ids = list(node_ids)
        self.device = device
        cfg = lib.config_defaults()
        cfg.device = device.index or 0
        cfg.max_groups = max_groups
        cfg.n_lanes = 1
        cfg.lane_node[0] = self.node_ids[node_index]
        cfg.max_batch_recs = max_batch
        cfg.max_batch_payload = max_payload
        for k, v in cfg_overrides.items():
            setattr(cfg, k, v)
        self.engine = Engine(lib, cfg)
        self.lib = lib
        n = len(node_ids)
        self.ctl = torch.zeros(8, dtype=torch.int32, device=device)       # gpx_dev_ctl
        self.cnt = torch.zeros((n, 2), dtype=torch.int32, device=device)  # per destination: records, blob units
        self.dropped = torch.zeros(1, dtype=torch.int32, device=device)
        self._rc = torch.zeros(n, dtype=torch.int32, device=device)
        self._ru = torch.zeros(n, dtype=torch.int32, device=device)
        self.dest = (C.c_int32 * n)(*self.node_ids)

    # ---- thin wrappers over the device-resident C ABI -------------------------------------------------
    def _st(self):
        st = torch.cuda.current_stream(self.device).cuda_stream
        assert st != 0, "spread rounds run on an explicit stream (SpreadCluster.round sets it)"
        return C.c_void_p(st)

    def _call(self, name, *args):
        self.lib.check(self.lib.fn(name)(self.engine.handle, *args))

    def propose(self, reqs: torch.Tensor, payload: torch.Tensor, n: int, status: torch.Tensor, accepts: torch.Tensor):
        self._call("propose_device", C.c_void_p(reqs.data_ptr()), C.c_void_p(payload.data_ptr()),
                   C.c_uint64(payload.numel()), C.c_uint32(n), C.c_void_p(status.data_ptr()),
                   C.c_void_p(accepts.data_ptr()), C.c_void_p(self.ctl.data_ptr()), self._st())

    def route(self, kind: int, recs: torch.Tensor, n_ptr: Optional[int], n_max: int, payload: Optional[torch.Tensor],
              out_recs: torch.Tensor, cap: int, out_blob: Optional[torch.Tensor], blob_cap: int):
        self._rc.zero_()  # records per destination
        self._ru.zero_()  # blob bytes / 16 per destination
        self._call("route_device", C.c_uint32(kind), C.c_void_p(recs.data_ptr()),
                   C.c_void_p(n_ptr) if n_ptr else None, C.c_uint32(n_max),
                   C.c_void_p(payload.data_ptr()) if payload is not None and payload.numel() else None,
                   C.c_uint64(payload.numel() if payload is not None else 0), C.c_uint32(len(self.node_ids)), self.dest,
                   C.c_void_p(out_recs.data_ptr()), C.c_uint32(cap), C.c_void_p(self._rc.data_ptr()),
                   C.c_void_p(out_blob.data_ptr()) if out_blob is not None else None, C.c_uint64(blob_cap),
                   C.c_void_p(self._ru.data_ptr()), C.c_void_p(self.dropped.data_ptr()), self._st())
        self.cnt = torch.stack([self._rc, self._ru], dim=1).contiguous()
        return self.cnt

    def accepts(self, recs: torch.Tensor, n: int, blob: torch.Tensor, rec_end, blob_base, replies: torch.Tensor,
                extra: torch.Tensor, extra_cap: int):
        k = len(rec_end)
        self._call("accepts_device", C.c_void_p(recs.data_ptr()), C.c_uint32(n),
                   C.c_void_p(blob.data_ptr()) if blob.numel() else None, C.c_uint64(blob.numel()), C.c_uint32(k),
                   (C.c_uint32 * k)(*rec_end), (C.c_uint64 * k)(*blob_base), C.c_void_p(replies.data_ptr()),
                   C.c_void_p(extra.data_ptr()), C.c_uint32(extra_cap), C.c_void_p(self.ctl.data_ptr()), self._st())

    def replies(self, recs: torch.Tensor, byte_off: int, n: int, decisions: torch.Tensor):
        self._call("replies_device", C.c_void_p(recs.data_ptr() + byte_off), C.c_uint32(n),
                   C.c_void_p(decisions.data_ptr()), C.c_void_p(self.ctl.data_ptr()), self._st())

    def decisions(self, recs: torch.Tensor, n: int, exec_out: torch.Tensor, extra: torch.Tensor, extra_cap: int):
        self._call("decisions_device", C.c_void_p(recs.data_ptr()), C.c_uint32(n), C.c_void_p(exec_out.data_ptr()),
                   C.c_void_p(extra.data_ptr()), C.c_uint32(extra_cap), C.c_void_p(self.ctl.data_ptr()), self._st())


class SpreadCluster:
    """Drives the local nodes of a spread deployment through rounds.  With `LocalExchange` all nodes are local
    (one process), with `NcclExchange` exactly one (this rank's)."""

    def __init__(self, nodes: List[SpreadNode], exchange, n_nodes: int):
        self.nodes = nodes
        self.x = exchange
        self.N = n_nodes
        assert [nd.index for nd in nodes] == list(exchange.local)
        self.timing = False
        self._pool: Dict[tuple, torch.Tensor] = {}  # persistent per-node buffers (grown, never shrunk)
        self.streams = {}
        for nd in nodes:
            if nd.device.type == "cuda" and nd.device not in self.streams:
                self.streams[nd.device] = torch.cuda.Stream(device=nd.device)

    def _buf(self, nd: "SpreadNode", key: str, nbytes: int, zero: bool = False) -> torch.Tensor:
        """a byte buffer of the node that lives across rounds: no allocator traffic on the round's critical path.
        What round() returns (status / exec / extra) points into these buffers and is valid until the next round."""
        k = (nd.index, key)
        t = self._pool.get(k)
        if t is None or t.numel() < nbytes:
            t = torch.empty(max(int(nbytes * 1.5), 256), dtype=torch.uint8, device=nd.device)
            self._pool[k] = t
        v = t[: max(nbytes, 1)]
        if zero:
            v.zero_()
        return v

    def _exchange(self, kind: int, outs: List[torch.Tensor], caps: List[int], blobs=None, blob_caps=None):
        """outs[k]: node k's routed buckets [N][cap] (bytes).  Returns per local node (recv records, per-source
        counts, recv blob, per-source blob bytes)."""
        rb = REC_BYTES[kind]
        recv_cnt, send_cnt = self.x.counts([nd.cnt for nd in self.nodes])
        send_r, recv_r, send_b, recv_b, res = [], [], [], [], []
        for k, nd in enumerate(self.nodes):
            sc, rc = send_cnt[k], recv_cnt[k]  # [N, 2]
            cap = caps[k]
            send_r.append([outs[k][d * cap * rb: d * cap * rb + int(sc[d, 0]) * rb] for d in range(self.N)])
            tot = int(rc[:, 0].sum())
            buf = self._buf(nd, "recv%d" % kind, max(tot, 1) * rb)
            offs = np.concatenate([[0], np.cumsum(rc[:, 0])]).astype(np.int64)
            recv_r.append([buf[offs[s] * rb: offs[s + 1] * rb] for s in range(self.N)])
            bbuf, boffs = None, None
            if blobs is not None:
                bc = blob_caps[k]
                send_b.append([blobs[k][d * bc: d * bc + int(sc[d, 1]) * 16] for d in range(self.N)])
                btot = int(rc[:, 1].sum()) * 16
                bbuf = self._buf(nd, "brecv%d" % kind, max(btot, 16))
                boffs = np.concatenate([[0], np.cumsum(rc[:, 1].astype(np.int64) * 16)])
                recv_b.append([bbuf[boffs[s]: boffs[s + 1]] for s in range(self.N)])
            res.append((buf, rc[:, 0].astype(np.int64), offs, bbuf, boffs))
        self.x.all_to_all(send_r, recv_r)
        if blobs is not None:
            self.x.all_to_all(send_b, recv_b)
        return res

    def round(self, batches: Dict[int, tuple], extra_cap: int = 4096):
        """One round on the cluster's own stream(s): the engine launches on the stream it is handed, torch copies and
        NCCL on torch's current stream -- they must be the same, explicit, stream (the legacy default stream does
        not order against the engine's non-blocking stream)."""
        import contextlib
        with contextlib.ExitStack() as stack:
            for dev, strm in self.streams.items():
                strm.wait_stream(torch.cuda.current_stream(dev))
            for strm in self.streams.values():
                stack.enter_context(torch.cuda.stream(strm))
            out = self._round(batches, extra_cap)
        for dev, strm in self.streams.items():
            torch.cuda.current_stream(dev).wait_stream(strm)
        return out

    def _round(self, batches: Dict[int, tuple], extra_cap: int = 4096):
        """batches[node_index] = (reqs uint8 tensor [n*32], payload uint8 tensor, n) on that node's device, every
        request belonging to a group whose coordinator is that node.  Returns per local node a dict with
        status [n] (int32 tensor), exec (uint8 tensor of 24-B EXEC records, one per received DECISION), n_exec,
        extra / n_extra, and the intermediate counts."""
        N = self.N
        st = {}
        marks = {nd.index: [] for nd in self.nodes}

        def mark(nd, name):  # CUDA events at the phase boundaries (only when self.timing)
            if self.timing and nd.device.type == "cuda":
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(torch.cuda.current_stream(nd.device))
                marks[nd.index].append((name, ev))

        for nd in self.nodes:
            mark(nd, "start")
        # ---- propose + route ACCEPTs -------------------------------------------------------------------
        outs, caps, blobs, bcaps = [], [], [], []
        for nd in self.nodes:
            reqs, payload, n = batches.get(nd.index, (None, None, 0))
            dev = nd.device
            with _on(dev):
                nd.ctl.zero_()
                nd.dropped.zero_()
                s = {"n": n, "status": self._buf(nd, "status", 4 * max(n, 1), zero=True).view(torch.int32),
                     "extra": self._buf(nd, "extra", extra_cap * 24, zero=True)}
                acc = self._buf(nd, "acc", max(n, 1) * 48)
                pbytes = payload.numel() if n else 0
                bcap = ((pbytes + 15) // 16 * 16) + 32 * max(n, 1)  # payload + batched-blob tables, per bucket
                out = self._buf(nd, "a_out", N * max(n, 1) * 48)
                ob = self._buf(nd, "a_blob", N * bcap)
                if n:
                    nd.propose(reqs, payload, n, s["status"], acc)
                    nd.route(K_ACCEPT, acc, nd.ctl.data_ptr() + 4 * CTL_N_ACCEPTS, n, payload, out, max(n, 1), ob, bcap)
                else:
                    nd.cnt = torch.zeros((N, 2), dtype=torch.int32, device=dev)
                s["_keep"] = (acc, out, ob)
                st[nd.index] = s
                outs.append(out), caps.append(max(n, 1)), blobs.append(ob), bcaps.append(bcap)
                mark(nd, "propose+route")
        got = self._exchange(K_ACCEPT, outs, caps, blobs, bcaps)
        for nd in self.nodes:
            mark(nd, "exchange_accepts")
        # ---- handleAccept + route replies ----------------------------------------------------------------
        outs, caps = [], []
        for nd, (abuf, cnts, offs, bbuf, boffs) in zip(self.nodes, got):
            s = st[nd.index]
            na = int(offs[-1])
            s["n_accepts_in"] = na
            with _on(nd.device):
                rep = self._buf(nd, "rep", max(na, 1) * 32)
                out = self._buf(nd, "r_out", N * max(na, 1) * 32)
                if na:
                    nd.accepts(abuf, na, bbuf[: int(boffs[-1])], [int(x) for x in offs[1:]], [int(x) for x in boffs[:-1]],
                               rep, s["extra"], extra_cap)
                    nd.route(K_REPLY, rep, None, na, None, out, max(na, 1), None, 0)
                else:
                    nd.cnt = torch.zeros((N, 2), dtype=torch.int32, device=nd.device)
                s["_keep2"] = (abuf, bbuf, rep, out)
                outs.append(out), caps.append(max(na, 1))
                mark(nd, "accept+route")
        got = self._exchange(K_REPLY, outs, caps)
        for nd in self.nodes:
            mark(nd, "exchange_replies")
        # ---- tally (one acceptor's bucket at a time, node order) + route DECISIONs ----------------------------
        outs, caps = [], []
        for nd, (rbuf, cnts, offs, _, _) in zip(self.nodes, got):
            s = st[nd.index]
            nr = int(offs[-1])
            n = s["n"]
            with _on(nd.device):
                dec = self._buf(nd, "dec", max(n, 1) * 32)
                out = self._buf(nd, "d_out", N * max(n, 1) * 32)
                for src in range(N):
                    if cnts[src]:
                        nd.replies(rbuf, int(offs[src]) * 32, int(cnts[src]), dec)
                if n:
                    nd.route(K_DECISION, dec, nd.ctl.data_ptr() + 4 * CTL_N_DECISIONS, n, None, out, max(n, 1), None, 0)
                else:
                    nd.cnt = torch.zeros((N, 2), dtype=torch.int32, device=nd.device)
                s["n_replies_in"] = nr
                s["_keep3"] = (rbuf, dec, out)
                outs.append(out), caps.append(max(n, 1))
                mark(nd, "tally+route")
        got = self._exchange(K_DECISION, outs, caps)
        for nd in self.nodes:
            mark(nd, "exchange_decisions")
        # ---- commit + execute -------------------------------------------------------------------------------
        for nd, (dbuf, cnts, offs, _, _) in zip(self.nodes, got):
            s = st[nd.index]
            ndec = int(offs[-1])
            with _on(nd.device):
                ex = self._buf(nd, "exec", max(ndec, 1) * 24)
                if ndec:
                    nd.decisions(dbuf, ndec, ex, s["extra"], extra_cap)
                s["exec"], s["n_exec"], s["_keep4"] = ex, ndec, dbuf
                mark(nd, "commit")
        for nd in self.nodes:
            s = st[nd.index]
            s["n_extra"] = int(nd.ctl[CTL_N_EXTRA].item())
            if int(nd.dropped.item()) != 0:
                raise RuntimeError("k_route dropped records: destination not served or buckets too small")
            if marks[nd.index]:
                torch.cuda.synchronize(nd.device)
                m = marks[nd.index]
                s["ms"] = {m[k + 1][0]: m[k][1].elapsed_time(m[k + 1][1]) for k in range(len(m) - 1)}
            for k in ("_keep", "_keep2", "_keep3", "_keep4"):
                s.pop(k, None)
        return st


# ======================================================================================================
# The data plane behind the C ABI (include/gpx.h gpx_spread_*): no Python, no host count reads per round.
# ======================================================================================================
def spread_caps(coord: np.ndarray, member_of: np.ndarray, slots_per_round: int = 1, slack: int = 0) -> np.ndarray:
    """cap[s][d] = groups node s coordinates that node d is a member of (x slots a round may open per group):
    the bucket capacities every node of the spread group must agree on (gpx_spread_config.cap)."""
    n = member_of.shape[1]
    cap = np.zeros((SPREAD_MAX_NODES_PY, SPREAD_MAX_NODES_PY), dtype=np.uint32)
    for s in range(n):
        mine = member_of[coord == s]
        if len(mine):
            cap[s, :n] = mine.sum(axis=0) * slots_per_round
    cap[cap > 0] += slack
    return cap


SPREAD_MAX_NODES_PY = abi.SPREAD_MAX_NODES


def spread_config(node_ids: Sequence[int], cap: np.ndarray, blob_per_rec: int, max_reqs: int,
                  graph: bool = False, p2p: bool = False) -> "abi.SpreadConfig":
    c = abi.SpreadConfig()
    c.n_nodes = len(node_ids)
    for i, x in enumerate(node_ids):
        c.node_ids[i] = int(x)
    for s in range(abi.SPREAD_MAX_NODES):
        for d in range(abi.SPREAD_MAX_NODES):
            c.cap[s][d] = int(cap[s, d])
    c.blob_per_rec = (int(blob_per_rec) + 15) // 16 * 16
    c.max_reqs = int(max_reqs)
    c.flags = (abi.SPREAD_GRAPH if graph else 0) | (abi.SPREAD_P2P if p2p else 0)
    return c


def spread_plan(lib: Library, cfg: "abi.SpreadConfig", rank: int) -> "abi.SpreadPlan":
    p = abi.SpreadPlan()
    lib.check(lib.fn("spread_plan_node")(C.byref(cfg), C.c_uint32(rank), C.byref(p)))
    return p


class Spread:
    """gpx_spread handle: `engines` are this process's nodes (all of them = local mode; exactly one + an NCCL
    unique id = one process per GPU)."""

    def __init__(self, lib: Library, engines: Sequence[Engine], cfg: "abi.SpreadConfig", rank: Optional[int] = None,
                 unique_id: Optional[bytes] = None):
        self.lib, self.cfg, self.engines = lib, cfg, list(engines)
        self._h = C.c_void_p()
        if unique_id is None:
            arr = (C.c_void_p * len(engines))(*[e.handle for e in engines])
            lib.check(lib.fn("spread_create_local")(arr, C.byref(cfg), C.byref(self._h)))
            self.ranks = list(range(len(engines)))
        else:
            assert len(engines) == 1 and rank is not None and len(unique_id) == 128
            buf = C.create_string_buffer(unique_id, 128)
            lib.check(lib.fn("spread_create_nccl")(engines[0].handle, C.byref(cfg), C.c_uint32(rank), buf,
                                                   C.byref(self._h)))
            self.ranks = [rank]
        self.plans = [spread_plan(lib, cfg, r) for r in self.ranks]

    @staticmethod
    def unique_id(lib: Library) -> bytes:
        buf = C.create_string_buffer(128)
        lib.check(lib.fn("spread_unique_id")(buf))
        return buf.raw

    def round(self, ios: Sequence["abi.SpreadIO"], stream: int):
        arr = (abi.SpreadIO * len(ios))(*ios)
        self.lib.check(self.lib.fn("spread_round")(self._h, arr, C.c_void_p(stream)))

    def dropped(self, k: int = 0) -> int:
        out = C.c_uint32(0)
        self.lib.check(self.lib.fn("spread_dropped")(self._h, C.c_uint32(k), C.byref(out)))
        return out.value

    def close(self):
        if self._h:
            self.lib.fn("spread_destroy")(self._h)
            self._h = C.c_void_p()
