/*
 * gpx_wire.cpp -- big-endian wire / journal codec of the four byte-codec'd packet types
 * (see include/gpx_wire.h for the reference file:line of every layout).  Host only.
 */
#include "gpx_wire.h"

#include <algorithm>
#include <cstring>
#include <vector>

#include "gpx.h"

namespace {

struct W { /* bounded big-endian writer (java.nio.ByteBuffer default order) */
  uint8_t* p;
  size_t cap, n = 0;
  bool ok = true;
  W(uint8_t* b, size_t c) : p(b), cap(c) {}
  void raw(const void* s, size_t k) {
    if (n + k > cap) {
      ok = false;
      n += k;
      return;
    }
    if (k) memcpy(p + n, s, k);
    n += k;
  }
  void u8(uint8_t v) { raw(&v, 1); }
  void i16(uint16_t v) {
    uint8_t b[2] = {(uint8_t)(v >> 8), (uint8_t)v};
    raw(b, 2);
  }
  void i32(int32_t v) {
    uint32_t u = (uint32_t)v;
    uint8_t b[4] = {(uint8_t)(u >> 24), (uint8_t)(u >> 16), (uint8_t)(u >> 8), (uint8_t)u};
    raw(b, 4);
  }
  void i64(int64_t v) {
    uint64_t u = (uint64_t)v;
    uint8_t b[8];
    for (int i = 0; i < 8; i++) b[i] = (uint8_t)(u >> (56 - 8 * i));
    raw(b, 8);
  }
};

struct Rd {
  const uint8_t* p;
  size_t len, n = 0;
  bool ok = true;
  Rd(const uint8_t* b, size_t l) : p(b), len(l) {}
  bool need(size_t k) {
    if (n + k > len) {
      ok = false;
      return false;
    }
    return true;
  }
  uint8_t u8() {
    if (!need(1)) return 0;
    return p[n++];
  }
  uint16_t i16() {
    if (!need(2)) return 0;
    uint16_t v = (uint16_t)((p[n] << 8) | p[n + 1]);
    n += 2;
    return v;
  }
  int32_t i32() {
    if (!need(4)) return 0;
    uint32_t v = ((uint32_t)p[n] << 24) | ((uint32_t)p[n + 1] << 16) | ((uint32_t)p[n + 2] << 8) | p[n + 3];
    n += 4;
    return (int32_t)v;
  }
  int64_t i64() {
    if (!need(8)) return 0;
    uint64_t v = 0;
    for (int i = 0; i < 8; i++) v = (v << 8) | p[n + i];
    n += 8;
    return (int64_t)v;
  }
  const uint8_t* skip(size_t k) {
    if (!need(k)) return nullptr;
    const uint8_t* q = p + n;
    n += k;
    return q;
  }
};

size_t id_len(const char* id, uint32_t l) {
  if (!id) return 0;
  return l ? l : strlen(id);
}

/* PaxosPacket.toBytes(ByteBuffer) :459-476 */
void put_header(W& w, int32_t type, int32_t version, const char* id, size_t idl) {
  w.i32(GPX_PT_PAXOS_PACKET);
  w.i32(type);
  w.i32(version);
  w.u8((uint8_t)idl);
  w.raw(id, idl);
}

/* RequestPacket.toBytes(boolean) :819-949 */
void put_request(W& w, const gpx_wire_request* r, int32_t type) {
  size_t idl = id_len(r->paxos_id, r->paxos_id_len);
  if (idl > GPX_WIRE_MAX_ID) {
    w.ok = false;
    return;
  }
  put_header(w, type, r->version, r->paxos_id, idl);
  w.i64(r->request_id);
  w.u8(r->stop ? 1 : 0);
  static const uint8_t zero4[4] = {0, 0, 0, 0};
  w.raw(r->client_port ? r->client_ip : zero4, 4);
  w.i16(r->client_port);
  w.raw(r->listen_port ? r->listen_ip : zero4, 4);
  w.i16(r->listen_port);
  w.i32(r->entry_replica);
  w.i64(r->entry_time);
  w.u8(r->should_return_request_value ? 1 : 0);
  w.i32(r->forward_count);
  w.u8(r->broadcasted ? 1 : 0);
  w.i32((int32_t)(r->digest ? r->digest_len : 0));
  if (r->digest) w.raw(r->digest, r->digest_len);
  w.i32((int32_t)r->value_len);
  w.raw(r->value, r->value_len);
  w.i32((int32_t)r->response_len);
  w.raw(r->response, r->response_len);
  w.i32((int32_t)r->n_batched);
  for (uint32_t i = 0; i < r->n_batched; i++) {
    /* batched requests: int length + req.toBytes() (a REQUEST packet) :929-936 */
    size_t sz = gpx_wire_request_size(&r->batched[i]);
    w.i32((int32_t)sz);
    put_request(w, &r->batched[i], GPX_PT_REQUEST);
  }
}

bool get_header(Rd& r, int32_t* type, int32_t* version, const char** id, uint32_t* idl) {
  int32_t pp = r.i32();
  *type = r.i32();
  *version = r.i32();
  uint8_t l = r.u8();
  const uint8_t* p = r.skip(l);
  if (!r.ok || pp != GPX_PT_PAXOS_PACKET || l > GPX_WIRE_MAX_ID) return false;
  *id = (const char*)p;
  *idl = l;
  return true;
}

/* RequestPacket(ByteBuffer) :956-1024 */
bool get_request(Rd& r, gpx_wire_accept_view* v) {
  if (!get_header(r, &v->packet_type, &v->version, &v->paxos_id, &v->paxos_id_len)) return false;
  v->request_id = r.i64();
  v->stop = r.u8() == 1;
  r.skip(4 + 2 + 4 + 2);
  v->entry_replica = r.i32();
  v->entry_time = r.i64();
  r.u8();
  r.i32();
  r.u8();
  int32_t dl = r.i32();
  if (dl < 0) return false;
  if (dl > 0) r.skip((size_t)dl);
  int32_t vl = r.i32();
  if (vl < 0) return false;
  v->value = r.skip((size_t)vl);
  v->value_len = (uint32_t)vl;
  int32_t rl = r.i32();
  if (rl < 0) return false;
  r.skip((size_t)rl);
  int32_t nb = r.i32();
  if (nb < 0) return false;
  v->n_batched = (uint32_t)nb;
  for (int32_t i = 0; i < nb; i++) {
    int32_t l = r.i32();
    if (l < 0) return false;
    r.skip((size_t)l);
  }
  return r.ok;
}

}  // namespace

extern "C" {

size_t gpx_wire_request_size(const gpx_wire_request* r) {
  if (!r) return 0;
  size_t n = 13 + id_len(r->paxos_id, r->paxos_id_len);
  n += 8 + 1 + 4 + 2 + 4 + 2 + 4 + 8 + 1 + 4 + 1 + 4 + (r->digest ? r->digest_len : 0) + 4 + r->value_len + 4 +
       r->response_len + 4;
  for (uint32_t i = 0; i < r->n_batched; i++) n += 4 + gpx_wire_request_size(&r->batched[i]);
  return n;
}

size_t gpx_wire_encode_request(const gpx_wire_request* r, int32_t packet_type, uint8_t* out, size_t cap) {
  if (!r || !out) return 0;
  W w(out, cap);
  put_request(w, r, packet_type);
  return w.ok ? w.n : 0;
}

/* AcceptPacket.toBytes :95-138 */
size_t gpx_wire_encode_accept(const gpx_wire_request* r, int32_t slot, int32_t bnum, int32_t bcoord, uint8_t recovery,
                              int32_t median_cp, int32_t sender, uint8_t* out, size_t cap) {
  if (!r || !out) return 0;
  W w(out, cap);
  put_request(w, r, GPX_PT_ACCEPT);
  w.i32(slot);             /* ProposalPacket.slot */
  w.i32(bnum);             /* PValuePacket: ballot */
  w.i32(bcoord);
  w.u8(recovery ? 1 : 0);  /* recovery */
  w.i32(median_cp);        /* medianCheckpointedSlot */
  w.u8(0);                 /* noCoalesce is always written as 0 (:124) */
  w.i32(sender);           /* AcceptPacket.sender */
  return w.ok ? w.n : 0;
}

int gpx_wire_decode_request(const uint8_t* buf, size_t len, gpx_wire_accept_view* out) {
  if (!buf || !out) return GPX_EINVAL;
  memset(out, 0, sizeof *out);
  Rd r(buf, len);
  if (!get_request(r, out)) return GPX_EINVAL;
  out->request_bytes = r.n;
  return GPX_OK;
}

int gpx_wire_decode_accept(const uint8_t* buf, size_t len, gpx_wire_accept_view* out) {
  if (!buf || !out) return GPX_EINVAL;
  memset(out, 0, sizeof *out);
  Rd r(buf, len);
  if (!get_request(r, out)) return GPX_EINVAL;
  out->request_bytes = r.n;
  if (out->packet_type != GPX_PT_ACCEPT) return GPX_EINVAL;
  out->slot = r.i32();
  out->bnum = r.i32();
  out->bcoord = r.i32();
  out->recovery = r.u8() == 1;
  out->median_cp = r.i32();
  r.u8();
  out->sender = r.i32();
  return (r.ok && r.n == len) ? GPX_OK : GPX_EINVAL;
}

/* BatchedAcceptReply.toBytes :120-173 over AcceptReplyPacket.toBytes(ByteBuffer) :174-184 */
size_t gpx_wire_encode_batched_accept_reply(const char* paxos_id, uint32_t paxos_id_len, int32_t version,
                                            int32_t acceptor, int32_t bnum, int32_t bcoord, int32_t slot_number,
                                            int32_t max_cp, int64_t request_id, uint32_t n, const int32_t* slots,
                                            const int64_t* req_ids, uint8_t* out, size_t cap) {
  if (!out || (n && (!slots || !req_ids))) return 0;
  size_t idl = id_len(paxos_id, paxos_id_len);
  if (idl > GPX_WIRE_MAX_ID) return 0;
  W w(out, cap);
  put_header(w, GPX_PT_BATCHED_ACCEPT_REPLY, version, paxos_id, idl);
  w.i32(acceptor);
  w.i32(bnum);
  w.i32(bcoord);
  w.i32(slot_number);
  w.i32(max_cp);
  w.i64(request_id);
  w.u8(0); /* undigestRequest */
  /* TreeMap<Integer,Long>: ascending signed slot order, one entry per slot (last put wins) */
  std::vector<uint32_t> idx(n);
  for (uint32_t i = 0; i < n; i++) idx[i] = i;
  std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return slots[a] < slots[b]; });
  std::vector<uint32_t> uniq;
  for (uint32_t k = 0; k < n; k++) {
    if (!uniq.empty() && slots[uniq.back()] == slots[idx[k]])
      uniq.back() = idx[k];
    else
      uniq.push_back(idx[k]);
  }
  w.i32((int32_t)uniq.size());
  for (uint32_t i : uniq) {
    w.i32(slots[i]);
    w.i64(req_ids[i]);
  }
  return w.ok ? w.n : 0;
}

int gpx_wire_decode_batched_accept_reply(const uint8_t* buf, size_t len, int32_t* version, char* paxos_id,
                                         uint32_t* paxos_id_len, int32_t* acceptor, int32_t* bnum, int32_t* bcoord,
                                         int32_t* slot_number, int32_t* max_cp, uint32_t* n, int32_t* slots,
                                         int64_t* req_ids, uint32_t cap_slots) {
  if (!buf) return GPX_EINVAL;
  Rd r(buf, len);
  int32_t type, ver;
  const char* id;
  uint32_t idl;
  if (!get_header(r, &type, &ver, &id, &idl) || type != GPX_PT_BATCHED_ACCEPT_REPLY) return GPX_EINVAL;
  if (version) *version = ver;
  if (paxos_id) memcpy(paxos_id, id, idl);
  if (paxos_id_len) *paxos_id_len = idl;
  int32_t a = r.i32(), bn = r.i32(), bc = r.i32(), sn = r.i32(), mc = r.i32();
  r.i64();
  r.u8();
  int32_t cnt = r.i32();
  if (!r.ok || cnt < 0) return GPX_EINVAL;
  if (acceptor) *acceptor = a;
  if (bnum) *bnum = bn;
  if (bcoord) *bcoord = bc;
  if (slot_number) *slot_number = sn;
  if (max_cp) *max_cp = mc;
  if (n) *n = (uint32_t)cnt;
  for (int32_t i = 0; i < cnt; i++) {
    int32_t s = r.i32();
    int64_t q = r.i64();
    if ((uint32_t)i < cap_slots) {
      if (slots) slots[i] = s;
      if (req_ids) req_ids[i] = q;
    }
  }
  return (r.ok && r.n == len) ? GPX_OK : GPX_EINVAL;
}

/* BatchedCommit.toBytes :184-252 */
size_t gpx_wire_encode_batched_commit(const char* paxos_id, uint32_t paxos_id_len, int32_t version, int32_t bnum,
                                      int32_t bcoord, int32_t median_cp, uint32_t n_slots, const int32_t* slots,
                                      uint32_t n_group, const int32_t* group, uint8_t* out, size_t cap) {
  if (!out || (n_slots && !slots) || (n_group && !group)) return 0;
  size_t idl = id_len(paxos_id, paxos_id_len);
  if (idl > GPX_WIRE_MAX_ID) return 0;
  W w(out, cap);
  put_header(w, GPX_PT_BATCHED_COMMIT, version, paxos_id, idl);
  w.i32(bnum);
  w.i32(bcoord);
  w.i32(median_cp);
  std::vector<int32_t> s(slots, slots + n_slots); /* TreeSet<Integer> */
  std::sort(s.begin(), s.end());
  s.erase(std::unique(s.begin(), s.end()), s.end());
  w.i32((int32_t)s.size());
  for (int32_t v : s) w.i32(v);
  w.i32((int32_t)n_group);
  for (uint32_t i = 0; i < n_group; i++) w.i32(group[i]);
  return w.ok ? w.n : 0;
}

int gpx_wire_decode_batched_commit(const uint8_t* buf, size_t len, int32_t* version, char* paxos_id,
                                   uint32_t* paxos_id_len, int32_t* bnum, int32_t* bcoord, int32_t* median_cp,
                                   uint32_t* n_slots, int32_t* slots, uint32_t cap_slots, uint32_t* n_group,
                                   int32_t* group, uint32_t cap_group) {
  if (!buf) return GPX_EINVAL;
  Rd r(buf, len);
  int32_t type, ver;
  const char* id;
  uint32_t idl;
  if (!get_header(r, &type, &ver, &id, &idl) || type != GPX_PT_BATCHED_COMMIT) return GPX_EINVAL;
  if (version) *version = ver;
  if (paxos_id) memcpy(paxos_id, id, idl);
  if (paxos_id_len) *paxos_id_len = idl;
  int32_t bn = r.i32(), bc = r.i32(), mc = r.i32();
  int32_t ns = r.i32();
  if (!r.ok || ns < 0) return GPX_EINVAL;
  for (int32_t i = 0; i < ns; i++) {
    int32_t s = r.i32();
    if ((uint32_t)i < cap_slots && slots) slots[i] = s;
  }
  int32_t ng = r.i32();
  if (!r.ok || ng < 0) return GPX_EINVAL;
  for (int32_t i = 0; i < ng; i++) {
    int32_t g = r.i32();
    if ((uint32_t)i < cap_group && group) group[i] = g;
  }
  if (bnum) *bnum = bn;
  if (bcoord) *bcoord = bc;
  if (median_cp) *median_cp = mc;
  if (n_slots) *n_slots = (uint32_t)ns;
  if (n_group) *n_group = (uint32_t)ng;
  return (r.ok && r.n == len) ? GPX_OK : GPX_EINVAL;
}

size_t gpx_wire_journal_frame(const uint8_t* packet, size_t len, uint8_t* out, size_t cap) {
  if (!packet || !out || len > 0x7fffffffu) return 0;
  W w(out, cap);
  w.i32((int32_t)len);
  w.raw(packet, len);
  return w.ok ? w.n : 0;
}

/* PaxosPacketBatcher.fuseBatchedCommits :389-417 + BatchedCommit.addCommit :113-121 */
uint32_t gpx_wire_fuse_commits(uint32_t n, const gpx_pvalue_hdr* d, uint32_t* run_start, int32_t* run_median_cp) {
  uint32_t runs = 0;
  for (uint32_t i = 0; i < n; i++) {
    bool same = runs && d[i].gid == d[run_start[runs - 1]].gid && d[i].bnum == d[run_start[runs - 1]].bnum &&
                d[i].bcoord == d[run_start[runs - 1]].bcoord;
    if (!same) {
      run_start[runs] = i;
      run_median_cp[runs] = d[i].median_cp;
      runs++;
    } else if ((int32_t)((uint32_t)d[i].median_cp - (uint32_t)run_median_cp[runs - 1]) > 0) {
      run_median_cp[runs - 1] = d[i].median_cp; /* wrap-aware max */
    }
  }
  return runs;
}

} /* extern "C" */
