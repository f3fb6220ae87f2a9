/*
 * gpx_route.cuh -- k_route: bucket inter-replica records by destination node for the exchange between
 * engines (replicas of a group living on different GPUs).
 *
 * Reference: PaxosPacketBatcher groups outgoing packets per destination (PaxosPacketBatcher.java:270-303) and
 * PaxosManager.send (:2098-2128) / MessagingTask.getNonLoopback (:196-224) split a multicast into unicasts;
 * here one launch turns a grouped-by-gid record stream (ACCEPTs out of the batcher, ACCEPT_REPLYs out of
 * the acceptors, DECISIONs out of the tally) into one contiguous run per destination node, ready for an
 * all-to-all.  The local node is just another destination (the reference's loopback).
 *
 * Work mapping as everywhere: the thread of the first record of a run of equal gids handles the run, so the
 * records of a group stay adjacent and in order inside every destination run (the receiving kernels rely on
 * that); the order BETWEEN groups is decided by block scheduling.  One atomicAdd per block and destination
 * (warp-shuffle scan inside the block).  ACCEPT blobs are re-packed per destination: payload_off of the routed
 * record is relative to that destination's blob.
 */
#pragma once
#include "gpx_kernels.cuh"

#define GPX_ROUTE_ND 8 /* destinations per launch (GPUs of one box) */

struct RouteArgs {
  const uint8_t* recs;  /* gpx_accept_rec[48] / gpx_accept_reply_rec[32] / gpx_decision_rec[32] */
  const uint32_t* n_ptr; /* device count, or null */
  uint32_t n_max;
  uint32_t kind;         /* GPX_F_ACCEPT, GPX_F_DECISION, or 0 = ACCEPT_REPLY */
  uint32_t n_dest;
  int32_t dest_node[GPX_ROUTE_ND];
  uint8_t* out_recs;     /* [n_dest][cap] records */
  uint32_t cap;
  uint32_t* out_counts;  /* [n_dest] device, zeroed by the caller: records per destination */
  /* ACCEPT only: source blob space = [blob0 | blob1], destination blobs [n_dest][blob_cap] */
  const uint8_t* blob0;
  unsigned long long blob0_bytes;
  const uint8_t* blob1;
  uint8_t* out_blob;
  unsigned long long blob_cap;
  uint32_t* out_blob_units; /* [n_dest] device, zeroed: blob bytes / 16 per destination */
  uint32_t* dropped;        /* device counter: records whose destination is not in dest_node / over capacity */
};

/* destination buckets of one record (bit d = dest_node[d]) */
__device__ __forceinline__ uint32_t route_mask(const DevState& S, const RouteArgs& A, const int4 q0, const int4 q1) {
  const uint32_t gid = (uint32_t)q0.x;
  if (gid >= S.G) return 0;
  const uint32_t meta = S.grp_meta[gid];
  if (!(meta & GPX_META_LIVE)) return 0;
  const MsetInfo* ms = &S.msets[meta & 0xffffu];
  const uint32_t R = (meta >> 16) & 0xffu;
  uint32_t members = 0; /* member indices the record goes to */
  if (A.kind == 0) {    /* ACCEPT_REPLY -> the coordinator that sent the ACCEPT */
    const uint32_t who = (uint32_t)q1.y;
    if (GPX_WHO_FLAGS(who) & GPX_F_VOID) return 0;
    const uint32_t dst = GPX_WHO_DST(who);
    if (dst < R) members = 1u << dst;
  } else { /* ACCEPT / DECISION: multicast to every member (dst_mask is only the sender's LOCAL lane mask) */
    const uint32_t fl = (uint32_t)q1.y;
    if (fl & GPX_F_VOID) return 0;
    members = (1u << R) - 1u;
  }
  uint32_t out = 0;
  for (uint32_t m = 0; m < R; m++)
    if ((members >> m) & 1u) {
      const int32_t node = ms->nodes[m];
#pragma unroll
      for (uint32_t d = 0; d < GPX_ROUTE_ND; d++)
        if (d < A.n_dest && A.dest_node[d] == node) out |= 1u << d;
    }
  return out;
}

__global__ void __launch_bounds__(GPX_BLOCK) k_route(const __grid_constant__ DevState S,
                                                     const __grid_constant__ RouteArgs A) {
  __shared__ uint32_t s_scan[GPX_BLOCK / 32 + 1];
  uint32_t n = A.n_ptr ? *A.n_ptr : A.n_max;
  if (n > A.n_max) n = A.n_max;
  const uint32_t rb = A.kind == GPX_F_ACCEPT ? 48u : 32u;
  const uint32_t i = blockIdx.x * GPX_BLOCK + threadIdx.x;
  bool head = false;
  uint32_t gid = 0, run = 0;
  uint32_t cnt[GPX_ROUTE_ND], units[GPX_ROUTE_ND];
#pragma unroll
  for (int d = 0; d < GPX_ROUTE_ND; d++) cnt[d] = units[d] = 0;
  if (i < n) {
    gid = *reinterpret_cast<const uint32_t*>(A.recs + (size_t)i * rb);
    head = (i == 0) || (*reinterpret_cast<const uint32_t*>(A.recs + (size_t)(i - 1) * rb) != gid);
    if (head) { /* pass 1: how much of the run goes where */
      for (uint32_t j = i; j < n; j++) {
        const int4* rp = reinterpret_cast<const int4*>(A.recs + (size_t)j * rb);
        const int4 q0 = rp[0], q1 = rp[1];
        if ((uint32_t)q0.x != gid) break;
        run++;
        const uint32_t m = route_mask(S, A, q0, q1);
        const bool is_void = A.kind == 0 ? (GPX_WHO_FLAGS((uint32_t)q1.y) & GPX_F_VOID) != 0 : ((uint32_t)q1.y & GPX_F_VOID) != 0;
        if (!m && !is_void && A.dropped) atomicAdd(A.dropped, 1u); /* no such group / destination not served */
        const uint32_t u = A.kind == GPX_F_ACCEPT ? (((uint32_t)rp[2].y + 15u) >> 4) : 0u;
#pragma unroll
        for (int d = 0; d < GPX_ROUTE_ND; d++)
          if ((m >> d) & 1u) {
            cnt[d]++;
            units[d] += u;
          }
      }
    }
  }
  /* one reservation per block and destination */
  uint32_t base[GPX_ROUTE_ND], ubase[GPX_ROUTE_ND];
#pragma unroll
  for (int d = 0; d < GPX_ROUTE_ND; d++) {
    base[d] = ubase[d] = 0;
    if ((uint32_t)d < A.n_dest) { /* uniform across the grid */
      base[d] = block_reserve(cnt[d], &A.out_counts[d], s_scan);
      if (A.kind == GPX_F_ACCEPT) ubase[d] = block_reserve(units[d], &A.out_blob_units[d], s_scan);
    }
  }
  if (!head) return;
  /* pass 2: copy the run */
  for (uint32_t j = i; j < i + run; j++) {
    const int4* rp = reinterpret_cast<const int4*>(A.recs + (size_t)j * rb);
    const int4 q0 = rp[0], q1 = rp[1];
    int4 q2 = make_int4(0, 0, 0, 0);
    if (rb == 48u) q2 = rp[2];
    const uint32_t m = route_mask(S, A, q0, q1);
    const uint32_t plen = (uint32_t)q2.y, u = (plen + 15u) >> 4;
#pragma unroll
    for (int d = 0; d < GPX_ROUTE_ND; d++)
      if ((m >> d) & 1u) {
        const uint32_t pos = base[d]++;
        const unsigned long long boff = (unsigned long long)ubase[d] << 4;
        ubase[d] += (A.kind == GPX_F_ACCEPT) ? u : 0u;
        if (pos >= A.cap || (A.kind == GPX_F_ACCEPT && boff + ((unsigned long long)u << 4) > A.blob_cap)) {
          if (A.dropped) atomicAdd(A.dropped, 1u); /* over capacity: the caller sized the buckets too small */
          continue;
        }
        int4* op = reinterpret_cast<int4*>(A.out_recs + ((size_t)d * A.cap + pos) * rb);
        op[0] = q0;
        op[1] = q1;
        if (rb == 48u) {
          op[2] = make_int4((int)(uint32_t)boff, q2.y, q2.z, q2.w);
          const unsigned long long off = (uint32_t)q2.x;
          const uint8_t* src = off < A.blob0_bytes ? A.blob0 + off : A.blob1 + (off - A.blob0_bytes);
          uint8_t* dst = A.out_blob + (unsigned long long)d * A.blob_cap + boff;
          uint32_t b = 0;
          if ((((uint32_t)(uintptr_t)src) & 15u) == 0)
            for (; b + 16 <= plen; b += 16) st_stream4(dst + b, ld_stream4(src + b));
          for (; b < plen; b++) dst[b] = src[b];
          for (; b < (u << 4); b++) dst[b] = 0; /* deterministic padding */
        }
      }
  }
}

/* Records received from other engines (ACCEPTs, DECISIONs): dst_mask is a LOCAL lane mask, so it is rewritten
 * to this engine's lanes that are members of the group; ACCEPT chunks are concatenated, so payload_off is made
 * relative to the concatenated blob. */
struct IngestArgs {
  uint8_t* recs;
  uint32_t rec_bytes; /* 48 (ACCEPT) or 32 (DECISION) */
  uint32_t n;
  uint32_t n_chunks;
  uint32_t rec_end[GPX_ROUTE_ND];              /* exclusive end index of chunk c */
  unsigned long long blob_base[GPX_ROUTE_ND];  /* where chunk c's blob starts in the concatenated blob */
};
__global__ void __launch_bounds__(GPX_BLOCK) k_ingest(const __grid_constant__ DevState S,
                                                      const __grid_constant__ IngestArgs A) {
  const uint32_t i = blockIdx.x * GPX_BLOCK + threadIdx.x;
  if (i >= A.n) return;
  uint8_t* r = A.recs + (size_t)i * A.rec_bytes;
  const uint32_t gid = *reinterpret_cast<const uint32_t*>(r);
  uint32_t lanes = 0;
  if (gid < S.G) {
    const uint32_t meta = S.grp_meta[gid];
    if (meta & GPX_META_LIVE) lanes = S.msets[meta & 0xffffu].lane_mask;
  }
  reinterpret_cast<gpx_pvalue_hdr*>(r)->dst_mask = (uint16_t)lanes;
  if (A.rec_bytes == 48u) {
    uint32_t c = 0;
#pragma unroll
    for (uint32_t k = 0; k + 1 < GPX_ROUTE_ND; k++)
      if (k + 1 < A.n_chunks && i >= A.rec_end[k]) c = k + 1;
    if (A.blob_base[c]) reinterpret_cast<gpx_accept_rec*>(r)->payload_off += (uint32_t)A.blob_base[c];
  }
}

/* ---- k_unpack: expand packed 16-byte requests (GPX_ROUND_PACKED_REQS) into gpx_request_rec; payload_off is the
 * exclusive prefix sum of payload_len.  Three small launches: per-block sums, scan of the block sums, expand. ---- */
#define GPX_UNPACK_PER_BLOCK 1024u /* requests per block: 256 threads x 4 */

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* s_w /*[GPX_BLOCK/32]*/, uint32_t* total) {
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= (uint32_t)d) incl += t;
  }
  if (lane == 31) s_w[wid] = incl;
  __syncthreads();
  uint32_t base = 0, tot = 0;
#pragma unroll
  for (uint32_t w = 0; w < GPX_BLOCK / 32; w++) {
    const uint32_t x = s_w[w];
    if (w < wid) base += x;
    tot += x;
  }
  __syncthreads();
  *total = tot;
  return base + incl - v;
}

__global__ void __launch_bounds__(GPX_BLOCK) k_unpack_sums(const gpx_request_packed* in, uint32_t n, uint32_t* bsum) {
  __shared__ uint32_t s_w[GPX_BLOCK / 32];
  const uint32_t b0 = blockIdx.x * GPX_UNPACK_PER_BLOCK + threadIdx.x * 4u;
  uint32_t v = 0;
#pragma unroll
  for (uint32_t k = 0; k < 4; k++)
    if (b0 + k < n) v += in[b0 + k].payload_len;
  uint32_t tot;
  block_exclusive_scan(v, s_w, &tot);
  if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(GPX_BLOCK) k_unpack_scan(uint32_t* bsum, uint32_t nb) {
  __shared__ uint32_t s_w[GPX_BLOCK / 32];
  uint32_t carry = 0;
  for (uint32_t base = 0; base < nb; base += GPX_BLOCK) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = i < nb ? bsum[i] : 0u;
    uint32_t tot;
    const uint32_t ex = block_exclusive_scan(v, s_w, &tot);
    if (i < nb) bsum[i] = carry + ex;
    carry += tot;
  }
}
__global__ void __launch_bounds__(GPX_BLOCK) k_unpack_expand(const __grid_constant__ DevState S, const gpx_request_packed* in,
                                                             uint32_t n, const uint32_t* bsum, gpx_request_rec* out) {
  __shared__ uint32_t s_w[GPX_BLOCK / 32];
  const uint32_t b0 = blockIdx.x * GPX_UNPACK_PER_BLOCK + threadIdx.x * 4u;
  gpx_request_packed r[4];
  uint32_t v = 0;
#pragma unroll
  for (uint32_t k = 0; k < 4; k++) {
    r[k].payload_len = 0;
    if (b0 + k < n) {
      r[k] = in[b0 + k];
      v += r[k].payload_len;
    }
  }
  uint32_t tot;
  uint32_t off = bsum[blockIdx.x] + block_exclusive_scan(v, s_w, &tot);
#pragma unroll
  for (uint32_t k = 0; k < 4; k++)
    if (b0 + k < n) {
      const uint32_t lane = (r[k].flags >> 8) & 0xfu;
      gpx_request_rec q;
      q.gid = r[k].gid;
      q.flags = r[k].flags;
      q.req_id = r[k].req_id;
      q.payload_off = off;
      q.payload_len = r[k].payload_len;
      q.entry_node = lane < S.L ? S.lane_node[lane] : S.lane_node[0];
      q.client = b0 + k;
      const int4* qp = reinterpret_cast<const int4*>(&q);
      st256(&out[b0 + k], qp[0], qp[1]);
      off += r[k].payload_len;
    }
}
