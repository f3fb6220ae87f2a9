/*
 * gpx_kernels.cuh -- the four hot-path kernels (sm_100a) and the state-maintenance kernels.
 *
 *   k_propose  RequestBatcher.dequeueImpl (RequestBatcher.java:168-234) + PISM.handleProposal
 *              (:818-888) + PaxosCoordinatorState.propose (:233-263)
 *   k_accept   PISM.handleAccept (:1080-1166) incl. PaxosAcceptor.acceptAndUpdateBallot (:302-322),
 *              the toLog rule (:1146-1149) and the log append (AbstractPaxosLogger.logAndMessage :157)
 *   k_tally    PISM.handleBatchedAcceptReply (:1370) / PaxosCoordinator.handleAcceptReply (:210-250)
 *   k_commit   PISM.handleBatchedCommit (:1480-1528) / handleCommittedRequest (:1432-1478) /
 *              extractExecuteAndCheckpoint (:1619-1701)
 *
 * Work mapping: one thread per record; records of one group are adjacent in every stream
 * ("grouped by gid"), and the thread of the first record of a run processes the whole run
 * in order, which reproduces the per-instance `synchronized` of the reference
 * (PaxosAcceptor.java:302,325; PaxosCoordinator.java:210).  Different groups never share
 * state, so runs are independent.  All traffic is 128-bit: 48-byte ACCEPTs are 3 x LDG.128,
 * rows are one LDG.128, window entries two.  Fixed-position outputs (reply i*L+lane, log
 * image i, exec i*L+lane) need no atomics; variable outputs (ACCEPTs, DECISIONs) are
 * compacted with one atomic per block.
 */
#pragma once
#include "gpx_dev.cuh"

#define GPX_BLOCK 256

struct RoundCtl { /* device-resident per-round counters */
  uint32_t n_accepts;
  uint32_t n_decisions;
  uint32_t n_extra;
  uint32_t any_batched;
  unsigned long long blob1_used;
  uint32_t pad[2];
};

/* every thread of the block calls; returns the first index reserved for this thread */
__device__ __forceinline__ uint32_t block_reserve(uint32_t cnt, uint32_t* counter, uint32_t* s_scan /*[GPX_BLOCK/32+1]*/) {
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t incl = cnt;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= (uint32_t)d) incl += t;
  }
  if (lane == 31) s_scan[wid] = incl;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t tot = 0;
    for (uint32_t w = 0; w < GPX_BLOCK / 32; w++) {
      uint32_t v = s_scan[w];
      s_scan[w] = tot;
      tot += v;
    }
    uint32_t base = tot ? atomicAdd(counter, tot) : 0u;
    s_scan[GPX_BLOCK / 32] = base;
  }
  __syncthreads();
  uint32_t r = s_scan[GPX_BLOCK / 32] + s_scan[wid] + incl - cnt;
  __syncthreads();
  return r;
}

__device__ __forceinline__ void flush_counters(const DevState& S, unsigned int* s_ctr) {
  __syncthreads();
  if (threadIdx.x < C_NCTR) {
    unsigned int v = s_ctr[threadIdx.x];
    if (v) atomicAdd(&S.ctr[threadIdx.x], (unsigned long long)v);
  }
}

__device__ __forceinline__ bool usable(const DevState& S, uint32_t gid, uint32_t lane, uint32_t* aux_out) {
  uint32_t aux = S.acc_aux[row_idx(S, lane, gid)];
  *aux_out = aux;
  uint32_t st = GPX_AUX_STATE(aux);
  return st == GPX_ST_ACTIVE_1 || st == GPX_ST_ACTIVE_2;
}

/* ============================== k_propose ===================================== */
struct ProposeArgs {
  const gpx_request_rec* reqs;
  uint32_t n;
  unsigned long long payload_bytes_al; /* request payload arena size, 16-B aligned */
  gpx_accept_rec* accepts;
  int32_t* status;
  uint32_t* copy_tab;
  uint32_t* copy_dst;
  RoundCtl* ctl;
};

__device__ __forceinline__ uint32_t batch_end(const DevState& S, const gpx_request_rec* reqs, uint32_t n, uint32_t k,
                                              uint32_t gid) {
  long long bytes = (long long)reqs[k].payload_len + S.size_est;
  int cnt = 1;
  uint32_t b = k + 1;
  if (S.batching)
    while (b < n && reqs[b].gid == gid) {
      bytes += (long long)reqs[b].payload_len + S.size_est;
      if (bytes > S.max_batch_bytes) break;
      cnt += 1;
      if (cnt > S.max_batch_size) break;
      b++;
    }
  return b;
}

__device__ __noinline__ void propose_run(const DevState& S, const ProposeArgs& A, uint32_t i, uint32_t run_end,
                                         uint32_t nb, uint32_t base, unsigned int* s_ctr) {
  const gpx_request_rec* reqs = A.reqs;
  const uint32_t gid = reqs[i].gid;
  const uint32_t Wm = S.W - 1;
  uint32_t emitted = 0;
  int code = 0;
  uint32_t k = i;
  int clane = -1;
  const MsetInfo* ms = nullptr;
  uint32_t R = 0;
  do {
    uint32_t entry = (reqs[i].flags >> 8) & 0xfu;
    if (gid >= S.G || entry >= S.L) {
      code = GPX_RS_DROPPED;
      break;
    }
    uint32_t meta = S.grp_meta[gid];
    uint32_t aux;
    if (!(meta & GPX_META_LIVE) || !usable(S, gid, entry, &aux)) {
      code = GPX_RS_DROPPED;
      break;
    }
    ms = &S.msets[meta & 0xffffu];
    R = (meta >> 16) & 0xffu;
    int4 Ae = S.acc_row[row_idx(S, entry, gid)];
    int4 Ce = S.coord_row[row_idx(S, entry, gid)];
    if (((unsigned)Ce.w & GPX_CF_EXISTS) && bcmp(Ce.x, Ce.y, Ae.y, Ae.z) >= 0) { /* PaxosCoordinator.exists(c, ballot) */
      clane = (int)entry;
    } else {
      int fl = -1;
      for (uint32_t l = 0; l < S.L; l++)
        if (S.lane_node[l] == Ae.z) fl = (int)l;
      if (fl < 0) {
        code = GPX_RS_FORWARD;
        break;
      }
      if (fl == (int)entry) {
        code = GPX_RS_NOCOORD;
        break;
      }
      if (!usable(S, gid, (uint32_t)fl, &aux)) {
        code = GPX_RS_DROPPED;
        break;
      }
      int4 Af = S.acc_row[row_idx(S, fl, gid)];
      int4 Cf = S.coord_row[row_idx(S, fl, gid)];
      if (((unsigned)Cf.w & GPX_CF_EXISTS) && bcmp(Cf.x, Cf.y, Af.y, Af.z) >= 0)
        clane = fl;
      else {
        code = GPX_RS_NOCOORD;
        break;
      }
    }
  } while (false);

  if (code == 0) {
    int4 crow = S.coord_row[row_idx(S, clane, gid)];
    bool dirty = false;
    while (k < run_end) {
      uint32_t b = batch_end(S, reqs, A.n, k, gid);
      if (b > run_end) b = run_end;
      uint32_t nreq = b - k;
      /* PCS.propose :235-239 refuse after a STOP that is still outstanding */
      {
        int prev = (int)((unsigned)crow.z - 1u);
        int4 pe = S.prop_win[win_idx(S, clane, (uint32_t)prev & Wm, gid)];
        if (((unsigned)pe.y & GPX_PV_PRESENT) && pe.x == prev && ((unsigned)pe.y & GPX_PV_STOP)) {
          code = GPX_RS_REFUSED_STOP;
          break;
        }
      }
      uint32_t w = (uint32_t)crow.z & Wm;
      {
        int4 pe = S.prop_win[win_idx(S, clane, w, gid)];
        if ((unsigned)pe.y & GPX_PV_PRESENT) { /* window full: W proposals in flight */
          code = GPX_RS_BACKPRESSURE;
          break;
        }
      }
      bool stop = false;
      for (uint32_t q = k; q < b; q++) stop = stop || (reqs[q].flags & GPX_F_STOP);
      int slot = crow.z;
      crow.z = (int)((unsigned)crow.z + 1u);
      crow.w = (int)((unsigned)crow.w + (1u << 8));
      dirty = true;
      long long rid = reqs[k].req_id;
      S.prop_win[win_idx(S, clane, w, gid)] =
          make_int4(slot, (int)(GPX_PV_PRESENT | (stop ? GPX_PV_STOP : 0u)), (int)(unsigned)(rid & 0xffffffffll),
                    (int)(rid >> 32));
      if (!((unsigned)crow.w & GPX_CF_ACTIVE)) { /* pre-active: queued, no ACCEPT yet :254-261 */
        for (uint32_t q = k; q < b; q++) A.status[q] = GPX_RS_PREACTIVE;
        k = b;
        continue;
      }
      int median = median_minus(S, (uint32_t)clane, gid, R);
      uint32_t off, plen;
      if (nreq == 1) {
        off = reqs[k].payload_off; /* zero copy: the blob is the request's own payload */
        plen = reqs[k].payload_len;
      } else {
        unsigned long long total = 16ull * nreq;
        for (uint32_t q = k; q < b; q++) total += reqs[q].payload_len;
        unsigned long long o1 = atomicAdd(&A.ctl->blob1_used, (total + 15ull) & ~15ull);
        off = (uint32_t)(A.payload_bytes_al + o1);
        plen = (uint32_t)total;
        unsigned long long run = 0;
        for (uint32_t q = k; q < b; q++) {
          A.copy_tab[q] = off + 16u * (q - k);
          A.copy_dst[q] = (uint32_t)(off + 16ull * nreq + run);
          run += reqs[q].payload_len;
        }
        A.ctl->any_batched = 1u;
      }
      gpx_accept_rec a;
      a.h.gid = gid;
      a.h.slot = slot;
      a.h.bnum = crow.x;
      a.h.bcoord = crow.y;
      a.h.median_cp = median;
      a.h.flags = (uint16_t)(GPX_F_ACCEPT | (stop ? GPX_F_STOP : 0u));
      a.h.dst_mask = ms->lane_mask;
      a.h.req_id = rid;
      a.payload_off = off;
      a.payload_len = plen;
      a.nreq = nreq;
      a.sender = crow.y;
      int4* dst = reinterpret_cast<int4*>(&A.accepts[base + emitted]);
      const int4* src = reinterpret_cast<const int4*>(&a);
      dst[0] = src[0];
      dst[1] = src[1];
      dst[2] = src[2];
      emitted++;
      A.status[k] = slot;
      for (uint32_t q = k + 1; q < b; q++) A.status[q] = GPX_RS_BATCHED;
      atomicAdd(&s_ctr[C_PROPOSALS], 1u);
      atomicAdd(&s_ctr[C_REQS_BATCHED], nreq);
      k = b;
    }
    if (dirty) S.coord_row[row_idx(S, clane, gid)] = crow;
  }
  if (code != 0) {
    for (uint32_t q = k; q < run_end; q++) A.status[q] = code;
    atomicAdd(&s_ctr[C_REQS_REJECTED], run_end - k);
  }
  for (; emitted < nb; emitted++) { /* reserved but unused: VOID keeps the run adjacent */
    gpx_accept_rec a;
    memset(&a, 0, sizeof a);
    a.h.gid = gid;
    a.h.flags = GPX_F_VOID;
    int4* dst = reinterpret_cast<int4*>(&A.accepts[base + emitted]);
    const int4* src = reinterpret_cast<const int4*>(&a);
    dst[0] = src[0];
    dst[1] = src[1];
    dst[2] = src[2];
  }
}

__global__ void __launch_bounds__(GPX_BLOCK) k_propose(const __grid_constant__ DevState S,
                                                       const __grid_constant__ ProposeArgs A) {
  __shared__ unsigned int s_ctr[C_NCTR];
  __shared__ uint32_t s_scan[GPX_BLOCK / 32 + 1];
  if (threadIdx.x < C_NCTR) s_ctr[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t i = blockIdx.x * GPX_BLOCK + threadIdx.x;
  bool head = false;
  uint32_t nb = 0, run_end = i + 1;
  if (i < A.n) {
    uint32_t gid = A.reqs[i].gid;
    head = (i == 0) || (A.reqs[i - 1].gid != gid);
    if (head) {
      if (i + 1 < A.n && A.reqs[i + 1].gid == gid) {
        uint32_t k = i;
        while (k < A.n && A.reqs[k].gid == gid) {
          k = batch_end(S, A.reqs, A.n, k, gid);
          nb++;
        }
        run_end = k;
      } else
        nb = 1;
    }
  }
  uint32_t base = block_reserve(head ? nb : 0u, &A.ctl->n_accepts, s_scan);
  if (head) propose_run(S, A, i, run_end, nb, base, s_ctr);
  if (i == 0) atomicAdd(&s_ctr[C_KERNEL_LAUNCHES], 1u);
  flush_counters(S, s_ctr);
}

/* builds the blobs of batched slots: [nreq x gpx_batch_ent][payloads], one thread per request */
__global__ void __launch_bounds__(GPX_BLOCK) k_build_blobs(const __grid_constant__ ProposeArgs A,
                                                           const uint8_t* payload, uint8_t* blob1) {
  if (!A.ctl->any_batched) return;
  const uint32_t i = blockIdx.x * GPX_BLOCK + threadIdx.x;
  if (i >= A.n) return;
  int st = A.status[i];
  bool batched = (st == GPX_RS_BATCHED) ||
                 (st > 0 && i + 1 < A.n && A.status[i + 1] == GPX_RS_BATCHED && A.reqs[i + 1].gid == A.reqs[i].gid);
  if (!batched) return;
  gpx_request_rec r = A.reqs[i];
  gpx_batch_ent be;
  be.req_id = r.req_id;
  be.len = r.payload_len;
  be.flags = r.flags;
  *reinterpret_cast<int4*>(blob1 + (A.copy_tab[i] - A.payload_bytes_al)) = *reinterpret_cast<const int4*>(&be);
  uint8_t* d = blob1 + (A.copy_dst[i] - A.payload_bytes_al);
  const uint8_t* s = payload + r.payload_off;
  for (uint32_t b = 0; b < r.payload_len; b++) d[b] = s[b];
}

/* ============================== k_accept ====================================== */
struct AcceptArgs {
  const gpx_accept_rec* recs;
  const uint32_t* n_ptr; /* device count, or null */
  uint32_t n_max;        /* record slots reserved (grid covers these) */
  const uint8_t* blob0;  /* payload arena: offsets [0, blob0_bytes) */
  unsigned long long blob0_bytes;
  const uint8_t* blob1; /* constructed blobs: offsets [blob0_bytes, blob0_bytes+blob1_bytes) */
  unsigned long long blob1_bytes;
  const unsigned long long* blob1_used_ptr; /* device: bytes of blob1 actually used (overrides blob1_bytes) */
  gpx_accept_reply_rec* replies; /* [n_max][L] */
  gpx_exec_rec* extra;
  uint32_t extra_cap;
  uint32_t* n_extra;
};

__device__ __forceinline__ const uint8_t* blob_ptr(const AcceptArgs& A, unsigned long long off) {
  return off < A.blob0_bytes ? A.blob0 + off : A.blob1 + (off - A.blob0_bytes);
}

/* one ACCEPT at every addressed lane; returns the mask of lanes that log it */
__device__ __forceinline__ uint32_t accept_one(const DevState& S, const AcceptArgs& A, const int4 q0, const int4 q1,
                                               const int4 q2, uint32_t j, const unsigned long long* segb,
                                               unsigned long long pay_rel, unsigned int* s_ctr) {
  const uint32_t gid = (uint32_t)q0.x;
  const int slot = q0.y, bnum = q0.z, bcoord = q0.w;
  const int median_cp = q1.x;
  const uint32_t fl_dm = (uint32_t)q1.y; /* flags | dst_mask<<16 */
  const uint32_t rflags = fl_dm & 0xffffu, dst_mask = fl_dm >> 16;
  const uint32_t payload_off = (uint32_t)q2.x, plen = (uint32_t)q2.y, nreq = (uint32_t)q2.z;
  const int sender = q2.w;
  const uint32_t Wm = S.W - 1;
  uint32_t logmask = 0;
  uint32_t meta = 0;
  const MsetInfo* ms = nullptr;
  bool gid_ok = gid < S.G;
  if (gid_ok) {
    meta = S.grp_meta[gid];
    ms = &S.msets[meta & 0xffffu];
  }
  const bool live = gid_ok && (meta & GPX_META_LIVE);
  uint32_t dstIdx = 0xffu;
  if (live) {
    const uint32_t R = (meta >> 16) & 0xffu;
    for (uint32_t m = 0; m < R; m++)
      if (ms->nodes[m] == sender) dstIdx = m;
  }
  for (uint32_t l = 0; l < S.L; l++) {
    /* default: VOID reply, VOID image */
    int4 rep0 = make_int4((int)gid, slot, 0, 0);
    int4 rep1 = make_int4(0, (int)GPX_WHO(0xffu, 0xffu, GPX_F_VOID), 0, 0);
    uint32_t img_flags = GPX_F_VOID, img_dm = dst_mask;
    do {
      if (!((dst_mask >> l) & 1u) || (rflags & GPX_F_VOID)) break;
      uint32_t aux;
      if (!live || !usable(S, gid, l, &aux)) { /* PISM :456-460 */
        atomicAdd(&s_ctr[C_ACCEPTS_DROPPED], 1u);
        break;
      }
      const uint32_t myIdx = ms->idx_of_lane[l];
      if (myIdx == 0xffu) {
        atomicAdd(&s_ctr[C_ACCEPTS_DROPPED], 1u);
        break;
      }
      const size_t ri = row_idx(S, l, gid);
      int4 row = S.acc_row[ri];
      if (jsub(slot, row.x) >= (int)S.W) { /* beyond the in-flight window: drop + flag for the host */
        S.acc_aux[ri] = aux | (GPX_GF_OVERFLOW << 24);
        atomicAdd(&s_ctr[C_WINDOW_OVERFLOW], 1u);
        atomicAdd(&s_ctr[C_ACCEPTS_DROPPED], 1u);
        break;
      }
      atomicAdd(&s_ctr[C_ACCEPTS_HANDLED], 1u);
      const int4 row_in = row;
      /* prev = paxosState.getAccept(slot) :1123 */
      const size_t ai = 2 * win_idx(S, l, (uint32_t)slot & Wm, gid);
      const int4 e0 = S.acc_win[ai], e1 = S.acc_win[ai + 1];
      const bool ent_alive = ((unsigned)e1.w & GPX_ENT_VALID) && jsub(e0.x, row.w) > 0;
      const bool hasPrev = ent_alive && e0.x == slot;
      unsigned frame_ref = (unsigned)(((segb[l] + pay_rel + payload_off) & (S.ring_cap - 1)) >> 4);
      if (hasPrev && e0.y == bnum && e0.z == bcoord) frame_ref = (unsigned)e0.w; /* duplicate keeps its frame */
      /* acceptAndUpdateBallot :302-322 */
      bool store = false;
      if (bcmp(bnum, bcoord, row.y, row.z) >= 0) {
        row.y = bnum;
        row.z = bcoord;
        if (jsub(slot, row.w) > 0) {
          store = true;
          if (ent_alive && e0.x != slot) { /* ring conflict: never evict a live entry for a stale accept */
            bool staleNew = jsub(slot, row.x) < 0, occStale = jsub(e0.x, row.x) < 0;
            if (staleNew && !occStale) store = false;
          }
        }
      }
      gc_step(row, median_cp); /* :320 */
      /* AcceptReplyPacket :1139-1143 */
      int max_cp = row.x - 1;
      if (!S.gc_majority_executed) {
        int cpi = S.cpi_per_group ? S.grp_cpi[gid] : S.cpi_const;
        int s1 = row.x - 1;
        int lcp = s1 - s1 % cpi;
        if (lcp < 0) {
          lcp = jsub(lcp, cpi);
          if (lcp > 0) lcp = 2147483647 - 2147483647 % cpi;
        }
        max_cp = lcp;
      }
      /* toLog :1146-1149 */
      const bool toLog = bcmp(bnum, bcoord, row.y, row.z) >= 0 && jsub(slot, row.w) > 0 &&
                         (!hasPrev || bcmp(e0.y, e0.z, bnum, bcoord) < 0);
      const bool nack = bcmp(row.y, row.z, bnum, bcoord) > 0;
      rep0 = make_int4((int)gid, slot, row.y, row.z);
      rep1 = make_int4(max_cp, (int)GPX_WHO(myIdx, dstIdx, (toLog ? GPX_F_LOGGED : 0u) | (nack ? GPX_F_NACK : 0u)),
                       q1.z, q1.w);
      atomicAdd(&s_ctr[nack ? C_ACCEPTS_NACKED : C_ACCEPTS_ACKED], 1u);
      if (toLog) {
        atomicAdd(&s_ctr[C_ACCEPTS_LOGGED], 1u);
        logmask |= 1u << l;
        img_flags = rflags;
        img_dm = 1u << l;
      }
      if (store) {
        S.acc_win[ai] = make_int4(slot, bnum, bcoord, (int)frame_ref);
        S.acc_win[ai + 1] = make_int4(q1.z, q1.w, (int)plen,
                                      (int)(GPX_ENT_VALID | ((rflags & GPX_F_STOP) ? GPX_ENT_STOP : 0u) | (nreq << 16)));
      }
      /* reconstructDecision(slot) -> handleCommittedRequest :1158-1161 (rare: a commit overtook its accept) */
      const int dslot = jsub(slot, row.x);
      if (dslot >= 0 && dslot < (int)S.W && ((GPX_AUX_PRESENT(aux) >> ((uint32_t)slot & Wm)) & 1u)) {
        const uint32_t w = (uint32_t)slot & Wm;
        const size_t ci = 2 * win_idx(S, l, w, gid);
        const int4 c0 = S.com_win[ci], c1 = S.com_win[ci + 1];
        DPValue d;
        bool ok = false;
        if ((GPX_AUX_VALUED(aux) >> w) & 1u) {
          d.slot = slot;
          d.bnum = c0.x;
          d.bcoord = c0.y;
          d.median_cp = c0.z;
          d.frame_ref = (unsigned)c0.w;
          d.req_id = ((long long)c1.y << 32) | (unsigned)c1.x;
          d.plen = (unsigned)c1.z;
          d.fl = (unsigned)c1.w;
          d.valued = true;
          ok = true;
        } else {
          const int4 n0 = S.acc_win[ai], n1 = S.acc_win[ai + 1];
          const bool alive = ((unsigned)n1.w & GPX_ENT_VALID) && jsub(n0.x, row.w) > 0 && n0.x == slot;
          if (alive && n0.y == c0.x && n0.z == c0.y) {
            d.slot = slot;
            d.bnum = n0.y;
            d.bcoord = n0.z;
            d.median_cp = c0.z;
            d.frame_ref = (unsigned)n0.w;
            d.req_id = ((long long)n1.y << 32) | (unsigned)n1.x;
            d.plen = (unsigned)n1.z;
            d.fl = (unsigned)n1.w;
            d.valued = true;
            ok = true;
          }
        }
        if (ok) eec(S, l, gid, row, aux, d, nullptr, A.extra, A.extra_cap, A.n_extra, s_ctr, true);
        S.acc_aux[ri] = aux;
      }
      if (row.x != row_in.x || row.y != row_in.y || row.z != row_in.z || row.w != row_in.w) S.acc_row[ri] = row;
    } while (false);
    /* reply (fixed position) */
    int4* rp = reinterpret_cast<int4*>(&A.replies[(size_t)j * S.L + l]);
    st_stream4(rp, rep0);
    st_stream4(rp + 1, rep1);
    /* log image (fixed position in this launch's ACCEPT segment of lane l) */
    int4* ip = reinterpret_cast<int4*>(ring_ptr(S, l, segb[l] + 64 + (unsigned long long)j * 48));
    st_stream4(ip, q0);
    st_stream4(ip + 1, make_int4(q1.x, (int)(img_flags | (img_dm << 16)), q1.z, q1.w));
    st_stream4(ip + 2, q2);
  }
  return logmask;
}

__global__ void __launch_bounds__(GPX_BLOCK) k_accept(const __grid_constant__ DevState S,
                                                      const __grid_constant__ AcceptArgs A) {
  __shared__ unsigned int s_ctr[C_NCTR];
  if (threadIdx.x < C_NCTR) s_ctr[threadIdx.x] = 0;
  __syncthreads();
  uint32_t n = A.n_ptr ? *A.n_ptr : A.n_max;
  if (n > A.n_max) n = A.n_max;
  const uint32_t i = blockIdx.x * GPX_BLOCK + threadIdx.x;
  const unsigned long long pay_bytes = A.blob0_bytes + (A.blob1_used_ptr ? *A.blob1_used_ptr : A.blob1_bytes);
  const unsigned long long reserved = 64ull + (unsigned long long)A.n_max * 48ull + pay_bytes;
  const unsigned long long pay_rel = 64ull + (unsigned long long)A.n_max * 48ull;
  unsigned long long segb[GPX_MAX_LANES];
#pragma unroll
  for (uint32_t l = 0; l < GPX_MAX_LANES; l++) segb[l] = l < S.L ? seg_base(S, l, reserved) : 0ull;
  if (i == 0) {
    for (uint32_t l = 0; l < S.L; l++) {
      gpx_log_seg_hdr h;
      memset(&h, 0, sizeof h);
      h.magic = GPX_SEG_MAGIC;
      h.type = GPX_F_ACCEPT;
      h.lane = (uint16_t)l;
      h.n_slots = A.n_max;
      h.n_valid = n;
      h.payload_bytes = pay_bytes;
      h.seq = S.seg_seq[l];
      h.ring_off = segb[l];
      h.rec_bytes = 48;
      int4* hp = reinterpret_cast<int4*>(ring_ptr(S, l, segb[l]));
      const int4* sp = reinterpret_cast<const int4*>(&h);
      hp[0] = sp[0];
      hp[1] = sp[1];
      hp[2] = sp[2];
      hp[3] = sp[3];
    }
    atomicAdd(&s_ctr[C_KERNEL_LAUNCHES], 1u);
  }
  if (i < n) {
    const int4* rp = reinterpret_cast<const int4*>(&A.recs[i]);
    int4 q0 = ld_stream4(rp), q1 = ld_stream4(rp + 1), q2 = ld_stream4(rp + 2);
    const uint32_t gid = (uint32_t)q0.x;
    const bool head = (i == 0) || (A.recs[i - 1].h.gid != gid);
    if (head) {
      uint32_t j = i;
      while (true) {
        uint32_t logmask = accept_one(S, A, q0, q1, q2, j, segb, pay_rel, s_ctr);
        if (logmask) { /* append the blob to the payload area of every logging lane */
          const uint32_t off = (uint32_t)q2.x, plen = (uint32_t)q2.y;
          const uint8_t* src = blob_ptr(A, off);
          if (((off | (uint32_t)(uintptr_t)src) & 15u) == 0) {
            for (uint32_t b = 0; b < plen; b += 16) {
              int4 v = ld_stream4(src + b);
              for (uint32_t l = 0; l < S.L; l++)
                if ((logmask >> l) & 1u) st_stream4(ring_ptr(S, l, segb[l] + pay_rel + off + b), v);
            }
          } else {
            for (uint32_t b = 0; b < plen; b++) {
              uint8_t v = src[b];
              for (uint32_t l = 0; l < S.L; l++)
                if ((logmask >> l) & 1u) *ring_ptr(S, l, segb[l] + pay_rel + off + b) = v;
            }
          }
        }
        j++;
        if (j >= n) break;
        rp = reinterpret_cast<const int4*>(&A.recs[j]);
        int4 t0 = ld_stream4(rp);
        if ((uint32_t)t0.x != gid) break;
        q0 = t0;
        q1 = ld_stream4(rp + 1);
        q2 = ld_stream4(rp + 2);
      }
    }
  }
  flush_counters(S, s_ctr);
  /* last block publishes the new ring heads */
  __shared__ unsigned int s_last;
  __threadfence();
  if (threadIdx.x == 0) s_last = (atomicAdd(&S.tickets[1], 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last && threadIdx.x < S.L) {
    S.ring_head[threadIdx.x] = segb[threadIdx.x] + ((reserved + 15ull) & ~15ull);
    S.seg_seq[threadIdx.x] += 1ull;
    if (threadIdx.x == 0) S.tickets[1] = 0;
  }
}

/* ============================== k_tally ======================================= */
struct TallyArgs {
  const gpx_accept_reply_rec* replies;
  const uint32_t* n_ptr; /* device count of ACCEPTs (replies = n * mult), or null */
  uint32_t mult;
  uint32_t n_max; /* reply slots covered by the grid */
  gpx_decision_rec* decisions;
  uint32_t* n_decisions;
};

__global__ void __launch_bounds__(GPX_BLOCK) k_tally(const __grid_constant__ DevState S,
                                                     const __grid_constant__ TallyArgs A) {
  __shared__ unsigned int s_ctr[C_NCTR];
  __shared__ uint32_t s_scan[GPX_BLOCK / 32 + 1];
  if (threadIdx.x < C_NCTR) s_ctr[threadIdx.x] = 0;
  __syncthreads();
  uint32_t n = A.n_ptr ? (*A.n_ptr) * A.mult : A.n_max;
  if (n > A.n_max) n = A.n_max;
  const uint32_t i = blockIdx.x * GPX_BLOCK + threadIdx.x;
  const uint32_t Wm = S.W - 1;
  gpx_decision_rec dbuf[GPX_MAX_WINDOW];
  uint32_t nd = 0;
  if (i == 0) atomicAdd(&s_ctr[C_KERNEL_LAUNCHES], 1u);
  if (i < n) {
    const int4* rp = reinterpret_cast<const int4*>(&A.replies[i]);
    int4 q0 = ld_stream4(rp), q1 = ld_stream4(rp + 1);
    const uint32_t gid = (uint32_t)q0.x;
    const bool head = (i == 0) || (A.replies[i - 1].gid != gid);
    if (head) {
      /* coordinator rows touched by this run are cached in registers per lane */
      int cl = -1;
      int4 crow = make_int4(0, 0, 0, 0);
      bool dirty = false;
      uint32_t meta = 0, R = 0;
      const MsetInfo* ms = nullptr;
      const bool gid_ok = gid < S.G;
      if (gid_ok) {
        meta = S.grp_meta[gid];
        ms = &S.msets[meta & 0xffffu];
        R = (meta >> 16) & 0xffu;
      }
      const bool live = gid_ok && (meta & GPX_META_LIVE);
      uint32_t j = i;
      while (true) {
        const uint32_t who = (uint32_t)q1.y;
        const uint32_t wf = GPX_WHO_FLAGS(who);
        if (!(wf & GPX_F_VOID)) {
          const uint32_t dstIdx = GPX_WHO_DST(who), accIdx = GPX_WHO_ACC(who);
          int lane = -1;
          uint32_t aux;
          if (live && dstIdx < R && ms->lane_of_idx[dstIdx] != 0xffu && usable(S, gid, ms->lane_of_idx[dstIdx], &aux))
            lane = ms->lane_of_idx[dstIdx];
          if (lane < 0) {
            atomicAdd(&s_ctr[C_REPLIES_IGNORED], 1u);
          } else {
            atomicAdd(&s_ctr[C_REPLIES_HANDLED], 1u);
            if (lane != cl) {
              if (dirty) S.coord_row[row_idx(S, cl, gid)] = crow;
              cl = lane;
              crow = S.coord_row[row_idx(S, cl, gid)];
              dirty = false;
            }
            const int slot = q0.y, rb = q0.z, rc = q0.w, max_cp = q1.x;
            const uint32_t cf = (unsigned)crow.w & 0xffu;
            if ((cf & GPX_CF_EXISTS) && (cf & GPX_CF_ACTIVE)) { /* PaxosCoordinator.handleAcceptReply :212 */
              const int c = bcmp(rb, rc, crow.x, crow.y);
              const size_t pi = win_idx(S, cl, (uint32_t)slot & Wm, gid);
              if (c > 0) { /* handleAcceptReplyHigherBallot :661-675 */
                int4 pe = S.prop_win[pi];
                if (((unsigned)pe.y & GPX_PV_PRESENT) && pe.x == slot) {
                  pe.y = (int)((unsigned)pe.y & ~GPX_PV_PRESENT);
                  S.prop_win[pi] = pe;
                  crow.w = (int)((unsigned)crow.w - (1u << 8));
                  dirty = true;
                  atomicAdd(&s_ctr[C_PREEMPTED], 1u);
                }
              } else if (c == 0) { /* handleAcceptReplyMyBallot :597-640 */
                if (accIdx < R) { /* recordSlotNumber :809-825 (plain <) */
                  const size_t ni = ns_idx(S, cl, accIdx, gid);
                  if (S.node_slots[ni] < max_cp) S.node_slots[ni] = max_cp;
                }
                int4 pe = S.prop_win[pi];
                if (((unsigned)pe.y & GPX_PV_PRESENT) && pe.x == slot) {
                  uint32_t vf = (unsigned)pe.y;
                  if (accIdx < R) vf |= (1u << accIdx); /* WaitforUtility.updateHeardFrom :51-62 */
                  if (__popc(vf & 0xffffu) > (int)(R / 2)) { /* heardFromMajority :64-68 */
                    gpx_decision_rec d;
                    d.gid = gid;
                    d.slot = slot;
                    d.bnum = crow.x;
                    d.bcoord = crow.y;
                    d.median_cp = median_minus(S, cl, gid, R); /* makeDecision(getMajorityCommittedSlot()) */
                    d.flags = (uint16_t)(GPX_F_DECISION | ((vf & GPX_PV_STOP) ? GPX_F_STOP : 0u));
                    d.dst_mask = ms->lane_mask;
                    d.req_id = ((long long)pe.w << 32) | (unsigned)pe.z;
                    if (nd < GPX_MAX_WINDOW) dbuf[nd] = d;
                    nd++;
                    pe.y = (int)(vf & ~GPX_PV_PRESENT);
                    crow.w = (int)((unsigned)crow.w - (1u << 8));
                    dirty = true;
                    atomicAdd(&s_ctr[C_DECISIONS_MADE], 1u);
                  } else
                    pe.y = (int)vf;
                  S.prop_win[pi] = pe;
                }
              }
            }
            /* nullifyCoordinatorIfPreemptedFully :1353-1356 */
            if ((((unsigned)crow.w) & GPX_CF_EXISTS) && bcmp(rb, rc, crow.x, crow.y) > 0 && (((unsigned)crow.w) >> 8) == 0) {
              crow = make_int4(0, 0, 0, 0);
              dirty = true;
              atomicAdd(&s_ctr[C_COORD_RESIGNED], 1u);
            }
          }
        }
        j++;
        if (j >= n) break;
        rp = reinterpret_cast<const int4*>(&A.replies[j]);
        int4 t0 = ld_stream4(rp);
        if ((uint32_t)t0.x != gid) break;
        q0 = t0;
        q1 = ld_stream4(rp + 1);
      }
      if (dirty) S.coord_row[row_idx(S, cl, gid)] = crow;
    }
  }
  if (nd > GPX_MAX_WINDOW) nd = GPX_MAX_WINDOW; /* cannot happen: <= W proposals outstanding */
  uint32_t base = block_reserve(nd, A.n_decisions, s_scan);
  for (uint32_t k = 0; k < nd; k++) {
    int4* dp = reinterpret_cast<int4*>(&A.decisions[base + k]);
    const int4* sp = reinterpret_cast<const int4*>(&dbuf[k]);
    st_stream4(dp, sp[0]);
    st_stream4(dp + 1, sp[1]);
  }
  flush_counters(S, s_ctr);
}

/* ============================== k_commit ====================================== */
struct CommitArgs {
  const gpx_decision_rec* decisions;
  const uint32_t* n_ptr;
  uint32_t n_max;
  gpx_exec_rec* exec; /* [n_max][L] */
  gpx_exec_rec* extra;
  uint32_t extra_cap;
  uint32_t* n_extra;
};

__device__ __forceinline__ void commit_one(const DevState& S, const CommitArgs& A, const int4 q0, const int4 q1,
                                           uint32_t j, const unsigned long long* segb, unsigned int* s_ctr) {
  const uint32_t gid = (uint32_t)q0.x;
  const int slot = q0.y, bnum = q0.z, bcoord = q0.w, median_cp = q1.x;
  const uint32_t rflags = (uint32_t)q1.y & 0xffffu, dst_mask = (uint32_t)q1.y >> 16;
  const uint32_t Wm = S.W - 1;
  const bool gid_ok = gid < S.G;
  const bool live = gid_ok && (S.grp_meta[gid] & GPX_META_LIVE);
  for (uint32_t l = 0; l < S.L; l++) {
    gpx_exec_rec* ex = &A.exec[(size_t)j * S.L + l];
    gpx_exec_rec vx;
    vx.gid = gid;
    vx.slot = slot;
    vx.req_id = 0;
    vx.payload_off = 0;
    vx.flags = GPX_F_VOID | (l << 12);
    store_exec(ex, vx);
    int4 img0 = q0, img1 = make_int4(q1.x, (int)(GPX_F_VOID | (dst_mask << 16)), q1.z, q1.w);
    do {
      if (!((dst_mask >> l) & 1u) || (rflags & GPX_F_VOID)) break;
      uint32_t aux;
      if (!live || !usable(S, gid, l, &aux)) {
        atomicAdd(&s_ctr[C_DECISIONS_DROPPED], 1u);
        break;
      }
      const size_t ri = row_idx(S, l, gid);
      int4 row = S.acc_row[ri];
      if (jsub(slot, row.x) >= (int)S.W) {
        S.acc_aux[ri] = aux | ((GPX_GF_OVERFLOW | GPX_GF_NEEDS_SYNC) << 24);
        atomicAdd(&s_ctr[C_WINDOW_OVERFLOW], 1u);
        atomicAdd(&s_ctr[C_DECISIONS_DROPPED], 1u);
        break;
      }
      atomicAdd(&s_ctr[C_DECISIONS_HANDLED], 1u);
      const int4 row_in = row;
      const uint32_t aux_in = aux;
      const size_t ai = 2 * win_idx(S, l, (uint32_t)slot & Wm, gid);
      const int4 a0 = S.acc_win[ai], a1 = S.acc_win[ai + 1];
      const bool a_alive = ((unsigned)a1.w & GPX_ENT_VALID) && jsub(a0.x, row.w) > 0 && a0.x == slot;
      DPValue d;
      d.slot = slot;
      d.bnum = bnum;
      d.bcoord = bcoord;
      d.median_cp = median_cp;
      if (a_alive && a0.y == bnum && a0.z == bcoord) { /* :1488 decision := the accept we hold */
        d.req_id = ((long long)a1.y << 32) | (unsigned)a1.x;
        d.frame_ref = (unsigned)a0.w;
        d.plen = (unsigned)a1.z;
        d.fl = (unsigned)a1.w & ~GPX_ENT_VALID;
        d.valued = true;
      } else { /* placeholder :1514-1522 */
        d.req_id = 0;
        d.frame_ref = 0;
        d.plen = 0;
        d.fl = 0;
        d.valued = false;
        atomicAdd(&s_ctr[C_PLACEHOLDERS], 1u);
      }
      /* logDecision :1446-1466 */
      if (d.valued || S.log_meta) {
        const bool meta = S.log_meta && a_alive && bcmp(a0.y, a0.z, d.bnum, d.bcoord) >= 0;
        const uint32_t lf = GPX_F_DECISION | (meta ? GPX_F_META : 0u) | ((d.fl & GPX_ENT_STOP) ? GPX_F_STOP : 0u);
        img0 = make_int4((int)gid, slot, d.bnum, d.bcoord);
        img1 = make_int4(meta ? -1 : d.median_cp, (int)(lf | ((1u << l) << 16)), (int)(unsigned)(d.req_id & 0xffffffffll),
                         (int)(d.req_id >> 32));
      }
      const int slot_before = row.x;
      eec(S, l, gid, row, aux, d, ex, A.extra, A.extra_cap, A.n_extra, s_ctr, false);
      if (GPX_AUX_STATE(aux) != GPX_ST_STOPPED && !d.valued && jsub(slot, row.x) >= 0 && row.x == slot_before)
        aux |= (GPX_GF_NEEDS_SYNC << 24);
      if (aux != aux_in) S.acc_aux[ri] = aux;
      if (row.x != row_in.x || row.y != row_in.y || row.z != row_in.z || row.w != row_in.w) S.acc_row[ri] = row;
    } while (false);
    int4* ip = reinterpret_cast<int4*>(ring_ptr(S, l, segb[l] + 64 + (unsigned long long)j * 32));
    st_stream4(ip, img0);
    st_stream4(ip + 1, img1);
  }
}

__global__ void __launch_bounds__(GPX_BLOCK) k_commit(const __grid_constant__ DevState S,
                                                      const __grid_constant__ CommitArgs A) {
  __shared__ unsigned int s_ctr[C_NCTR];
  if (threadIdx.x < C_NCTR) s_ctr[threadIdx.x] = 0;
  __syncthreads();
  uint32_t n = A.n_ptr ? *A.n_ptr : A.n_max;
  if (n > A.n_max) n = A.n_max;
  const uint32_t i = blockIdx.x * GPX_BLOCK + threadIdx.x;
  const unsigned long long reserved = 64ull + (unsigned long long)A.n_max * 32ull;
  unsigned long long segb[GPX_MAX_LANES];
#pragma unroll
  for (uint32_t l = 0; l < GPX_MAX_LANES; l++) segb[l] = l < S.L ? seg_base(S, l, reserved) : 0ull;
  if (i == 0) {
    for (uint32_t l = 0; l < S.L; l++) {
      gpx_log_seg_hdr h;
      memset(&h, 0, sizeof h);
      h.magic = GPX_SEG_MAGIC;
      h.type = GPX_F_DECISION;
      h.lane = (uint16_t)l;
      h.n_slots = A.n_max;
      h.n_valid = n;
      h.payload_bytes = 0;
      h.seq = S.seg_seq[l];
      h.ring_off = segb[l];
      h.rec_bytes = 32;
      int4* hp = reinterpret_cast<int4*>(ring_ptr(S, l, segb[l]));
      const int4* sp = reinterpret_cast<const int4*>(&h);
      hp[0] = sp[0];
      hp[1] = sp[1];
      hp[2] = sp[2];
      hp[3] = sp[3];
    }
    atomicAdd(&s_ctr[C_KERNEL_LAUNCHES], 1u);
  }
  if (i < n) {
    const int4* rp = reinterpret_cast<const int4*>(&A.decisions[i]);
    int4 q0 = ld_stream4(rp), q1 = ld_stream4(rp + 1);
    const uint32_t gid = (uint32_t)q0.x;
    const bool head = (i == 0) || (A.decisions[i - 1].gid != gid);
    if (head) {
      uint32_t j = i;
      while (true) {
        commit_one(S, A, q0, q1, j, segb, s_ctr);
        j++;
        if (j >= n) break;
        rp = reinterpret_cast<const int4*>(&A.decisions[j]);
        int4 t0 = ld_stream4(rp);
        if ((uint32_t)t0.x != gid) break;
        q0 = t0;
        q1 = ld_stream4(rp + 1);
      }
    }
  }
  flush_counters(S, s_ctr);
  __shared__ unsigned int s_last;
  __threadfence();
  if (threadIdx.x == 0) s_last = (atomicAdd(&S.tickets[3], 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last && threadIdx.x < S.L) {
    S.ring_head[threadIdx.x] = segb[threadIdx.x] + reserved;
    S.seg_seq[threadIdx.x] += 1ull;
    if (threadIdx.x == 0) S.tickets[3] = 0;
  }
}

/* ============================== state maintenance ============================== */
struct InitRec { /* host-preprocessed gpx_group_desc */
  uint32_t gid;
  uint32_t mset;
  int32_t coord0; /* roundRobinCoordinator(name, members, 0) */
  int32_t cpi;
  int32_t init_mode;
  uint32_t R;
};

/* PaxosManager.createPaxosInstance batch form :664-691 -> createHRI paxosutil/HotRestoreInfo.java:145-157,
 * or PISM.initiateRecovery :591-675 + putInitialState :692-699 */
__global__ void k_init_groups(const __grid_constant__ DevState S, const InitRec* recs, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const InitRec r = recs[i];
  const MsetInfo* ms = &S.msets[r.mset];
  S.grp_meta[r.gid] = r.mset | (r.R << 16) | GPX_META_LIVE;
  S.grp_cpi[r.gid] = r.cpi;
  for (uint32_t l = 0; l < S.L; l++) {
    const size_t ri = row_idx(S, l, r.gid);
    for (uint32_t w = 0; w < S.W; w++) {
      const size_t wi = win_idx(S, l, w, r.gid);
      S.acc_win[2 * wi] = make_int4(0, 0, 0, 0);
      S.acc_win[2 * wi + 1] = make_int4(0, 0, 0, 0);
      S.com_win[2 * wi] = make_int4(0, 0, 0, 0);
      S.com_win[2 * wi + 1] = make_int4(0, 0, 0, 0);
      S.prop_win[wi] = make_int4(0, 0, 0, 0);
    }
    if (ms->idx_of_lane[l] == 0xffu) { /* lane is not a member: no instance here */
      S.acc_row[ri] = make_int4(0, -1, -1, -1);
      S.acc_aux[ri] = GPX_ST_FREE;
      S.coord_row[ri] = make_int4(0, 0, 0, 0);
      continue;
    }
    const bool batch = r.init_mode == GPX_INIT_BATCH;
    S.acc_row[ri] = make_int4(1, 0, r.coord0, batch ? -1 : 0);
    S.acc_aux[ri] = GPX_ST_ACTIVE_1;
    const bool am_coord = (r.coord0 == S.lane_node[l]);
    S.coord_row[ri] = am_coord ? make_int4(0, r.coord0, 1, (int)(GPX_CF_EXISTS | GPX_CF_ACTIVE)) : make_int4(0, 0, 0, 0);
    for (uint32_t m = 0; m < S.Rcap; m++) S.node_slots[ns_idx(S, l, m, r.gid)] = am_coord ? (batch ? 0 : -1) : 0;
  }
}

__global__ void k_destroy_groups(const __grid_constant__ DevState S, const uint32_t* gids, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t gid = gids[i];
  if (gid >= S.G) return;
  S.grp_meta[gid] = 0;
  for (uint32_t l = 0; l < S.L; l++) {
    const size_t ri = row_idx(S, l, gid);
    S.acc_row[ri] = make_int4(0, -1, -1, -1);
    S.acc_aux[ri] = GPX_ST_FREE;
    S.coord_row[ri] = make_int4(0, 0, 0, 0);
  }
}

__global__ void k_dump_rows(const __grid_constant__ DevState S, const uint32_t* gids, uint32_t n, uint32_t lane,
                            gpx_row* out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t gid = gids[i];
  gpx_row r;
  memset(&r, 0, sizeof r);
  r.gid = gid;
  r.lane = lane;
  if (gid < S.G) {
    const uint32_t meta = S.grp_meta[gid];
    const size_t ri = row_idx(S, lane, gid);
    const int4 row = S.acc_row[ri];
    const int4 c = S.coord_row[ri];
    const bool live = (meta & GPX_META_LIVE) != 0;
    r.acc_slot = row.x;
    r.acc_bnum = row.y;
    r.acc_bcoord = row.z;
    r.acc_gc_slot = row.w;
    r.state = live ? (int)GPX_AUX_STATE(S.acc_aux[ri]) : GPX_ST_FREE;
    const bool ex = ((unsigned)c.w & GPX_CF_EXISTS) != 0;
    r.coord_exists = ex;
    r.coord_active = ex && (((unsigned)c.w & GPX_CF_ACTIVE) != 0);
    r.coord_bnum = ex ? c.x : 0;
    r.coord_bcoord = ex ? c.y : 0;
    r.next_proposal_slot = ex ? c.z : 0;
    if (live) {
      const MsetInfo* ms = &S.msets[meta & 0xffffu];
      const uint32_t R = (meta >> 16) & 0xffu;
      r.n_members = (int)R;
      for (uint32_t m = 0; m < R; m++) {
        r.members[m] = ms->nodes[m];
        r.node_slots[m] = ex ? S.node_slots[ns_idx(S, lane, m, gid)] : 0;
      }
    }
  }
  out[i] = r;
}

struct LoadRec {
  gpx_row row;
  uint32_t mset;
  int32_t cpi;
};
__global__ void k_load_rows(const __grid_constant__ DevState S, const LoadRec* recs, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const gpx_row& r = recs[i].row;
  const uint32_t gid = r.gid, l = r.lane;
  S.grp_meta[gid] = recs[i].mset | ((uint32_t)r.n_members << 16) | GPX_META_LIVE;
  S.grp_cpi[gid] = recs[i].cpi;
  const size_t ri = row_idx(S, l, gid);
  S.acc_row[ri] = make_int4(r.acc_slot, r.acc_bnum, r.acc_bcoord, r.acc_gc_slot);
  S.acc_aux[ri] = (uint32_t)r.state & 0xffu;
  for (uint32_t w = 0; w < S.W; w++) {
    const size_t wi = win_idx(S, l, w, gid);
    S.acc_win[2 * wi + 1] = make_int4(0, 0, 0, 0);
    S.prop_win[wi] = make_int4(0, 0, 0, 0);
  }
  if (r.coord_exists) {
    S.coord_row[ri] =
        make_int4(r.coord_bnum, r.coord_bcoord, r.next_proposal_slot, (int)(GPX_CF_EXISTS | (r.coord_active ? GPX_CF_ACTIVE : 0u)));
    for (int m = 0; m < r.n_members && m < (int)S.Rcap; m++) S.node_slots[ns_idx(S, l, m, gid)] = r.node_slots[m];
  } else
    S.coord_row[ri] = make_int4(0, 0, 0, 0);
}

/* slow-path effects as state patches: distinct (gid,lane) targets per call */
__global__ void k_patch(const __grid_constant__ DevState S, const gpx_patch_rec* p, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const gpx_patch_rec r = p[i];
  if (r.gid >= S.G || r.lane >= S.L) return;
  const size_t ri = row_idx(S, r.lane, r.gid);
  int4 row = S.acc_row[ri];
  uint32_t aux = S.acc_aux[ri];
  const uint32_t Wm = S.W - 1;
  switch (r.op) {
    case GPX_PATCH_SET_BALLOT: /* handlePrepare :245-251 */
      if (bcmp(r.a, r.b, row.y, row.z) > 0) {
        row.y = r.a;
        row.z = r.b;
      }
      break;
    case GPX_PATCH_JUMP_SLOT: /* jumpSlot :564-578 */
      while (jsub(row.x, r.a) < 0) {
        const uint32_t w = (uint32_t)row.x & Wm;
        aux &= ~((1u << (8 + w)) | (1u << (16 + w)));
        if (S.journaling) {
          const size_t ai = 2 * win_idx(S, r.lane, w, r.gid);
          int4 a0 = S.acc_win[ai], a1 = S.acc_win[ai + 1];
          if (((unsigned)a1.w & GPX_ENT_VALID) && a0.x == row.x) {
            a1.w = (int)((unsigned)a1.w & ~GPX_ENT_VALID);
            S.acc_win[ai + 1] = a1;
          }
        }
        row.x = (int)((unsigned)row.x + 1u);
      }
      break;
    case GPX_PATCH_SET_STATE:
      aux = (aux & ~0xffu) | ((uint32_t)r.a & 0xffu);
      if (((uint32_t)r.a & 0xffu) == GPX_ST_STOPPED) aux &= ~0x00ffff00u;
      break;
    case GPX_PATCH_INSTALL_COORD: {
      S.coord_row[ri] = make_int4(r.a, r.b, r.c, (int)(GPX_CF_EXISTS | (r.d ? GPX_CF_ACTIVE : 0u)));
      for (uint32_t m = 0; m < S.Rcap; m++) S.node_slots[ns_idx(S, r.lane, m, r.gid)] = -1;
      for (uint32_t w = 0; w < S.W; w++) S.prop_win[win_idx(S, r.lane, w, r.gid)] = make_int4(0, 0, 0, 0);
      break;
    }
    case GPX_PATCH_RESIGN_COORD:
      S.coord_row[ri] = make_int4(0, 0, 0, 0);
      for (uint32_t w = 0; w < S.W; w++) S.prop_win[win_idx(S, r.lane, w, r.gid)] = make_int4(0, 0, 0, 0);
      break;
    case GPX_PATCH_SET_GC: row.w = r.a; break;
    default: break;
  }
  S.acc_row[ri] = row;
  S.acc_aux[ri] = aux;
}

__global__ void k_get_flags(const __grid_constant__ DevState S, uint32_t lane, const uint32_t* gids, uint32_t n,
                            uint8_t* out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = gids[i] < S.G ? (uint8_t)GPX_AUX_FLAGS(S.acc_aux[row_idx(S, lane, gids[i])]) : 0;
}
