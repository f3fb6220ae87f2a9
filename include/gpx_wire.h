/*
 * gpx_wire.h -- host-side codec between engine records and the reference's big-endian
 * wire / journal bytes, so that a Java PaxosManager (or a journal reader) interoperates
 * unchanged.  Paths relative to
 * /root/reference/src/edu/umass/cs/gigapaxos/paxospackets/.
 *
 *   PaxosPacket header      PaxosPacket.java:459-476  {int 90 (PAXOS_PACKET), int type, int version,
 *                                                      byte idLen, id bytes (ISO-8859-1)}
 *   RequestPacket body      RequestPacket.java:779-798, toBytes :819-949, ctor :956-1024
 *   AcceptPacket            AcceptPacket.java:95-138   request bytes (header type ACCEPT) + slot,
 *                                                      ballot, recovery, medianCP, noCoalesce(0), sender
 *   AcceptReplyPacket       AcceptReplyPacket.java:121-184
 *   BatchedAcceptReply      BatchedAcceptReply.java:103-173
 *   BatchedCommit           BatchedCommit.java:156-252
 *   journal frame           SQLPaxosLogger.java:1000-1003  {int32 BE length}{packet bytes}
 *
 * Encoders return the number of bytes written, or 0 if `cap` is too small / arguments are
 * invalid.  Decoders return GPX_OK or GPX_EINVAL on malformed input.  Pure host code.
 */
#ifndef GPX_WIRE_H
#define GPX_WIRE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* PaxosPacket.PaxosPacketType codes (PaxosPacket.java:202-291) */
enum {
  GPX_PT_PAXOS_PACKET = 90,
  GPX_PT_REQUEST = 1,
  GPX_PT_ACCEPT = 3,
  GPX_PT_DECISION = 6,
  GPX_PT_ACCEPT_REPLY = 8,
  GPX_PT_BATCHED_ACCEPT_REPLY = 34,
  GPX_PT_BATCHED_COMMIT = 35,
  GPX_PT_BATCHED_ACCEPT = 36,
  GPX_PT_BATCHED_PAXOS_PACKET = 37
};

#define GPX_WIRE_MAX_ID 127 /* idLen is one signed byte; PC.MAX_PAXOS_ID_SIZE defaults to 40 */

/* one RequestPacket (fields in wire order, RequestPacket.java:819-949) */
typedef struct gpx_wire_request {
  const char* paxos_id; /* not NUL-terminated if paxos_id_len is given */
  uint32_t paxos_id_len;
  int32_t version;
  int64_t request_id;
  uint8_t stop;
  uint8_t client_ip[4];
  uint16_t client_port; /* 0 = no client address */
  uint8_t listen_ip[4];
  uint16_t listen_port;
  int32_t entry_replica;
  int64_t entry_time;
  uint8_t should_return_request_value;
  int32_t forward_count;
  uint8_t broadcasted;
  const uint8_t* digest;
  uint32_t digest_len;
  const uint8_t* value; /* requestValue bytes (ISO-8859-1) */
  uint32_t value_len;
  const uint8_t* response;
  uint32_t response_len;
  uint32_t n_batched;
  const struct gpx_wire_request* batched; /* RequestPacket.batched, each encoded as a REQUEST packet */
} gpx_wire_request;

/* encode with the given PaxosPacketType in the header (REQUEST for requests, ACCEPT when the
 * bytes are the prefix of an AcceptPacket, AcceptPacket.java:104) */
size_t gpx_wire_encode_request(const gpx_wire_request* r, int32_t packet_type, uint8_t* out, size_t cap);
size_t gpx_wire_request_size(const gpx_wire_request* r);

size_t gpx_wire_encode_accept(const gpx_wire_request* r, int32_t slot, int32_t bnum, int32_t bcoord, uint8_t recovery,
                              int32_t median_cp, int32_t sender, uint8_t* out, size_t cap);

/* decoded view of an ACCEPT / REQUEST (pointers point into the input buffer) */
typedef struct gpx_wire_accept_view {
  int32_t packet_type, version;
  const char* paxos_id;
  uint32_t paxos_id_len;
  int64_t request_id;
  uint8_t stop;
  int32_t entry_replica;
  int64_t entry_time;
  const uint8_t* value;
  uint32_t value_len;
  uint32_t n_batched;
  size_t request_bytes; /* bytes of the RequestPacket part */
  int32_t slot, bnum, bcoord, median_cp, sender; /* valid for ACCEPT */
  uint8_t recovery;
} gpx_wire_accept_view;
int gpx_wire_decode_accept(const uint8_t* buf, size_t len, gpx_wire_accept_view* out);
int gpx_wire_decode_request(const uint8_t* buf, size_t len, gpx_wire_accept_view* out);

/* BATCHED_ACCEPT_REPLY: slots/req_ids are written in ascending (signed) slot order (TreeMap) */
size_t gpx_wire_encode_batched_accept_reply(const char* paxos_id, uint32_t paxos_id_len, int32_t version,
                                            int32_t acceptor, int32_t bnum, int32_t bcoord, int32_t slot_number,
                                            int32_t max_checkpointed_slot, int64_t request_id, uint32_t n,
                                            const int32_t* slots, const int64_t* req_ids, uint8_t* out, size_t cap);
int gpx_wire_decode_batched_accept_reply(const uint8_t* buf, size_t len, int32_t* version, char* paxos_id,
                                         uint32_t* paxos_id_len, int32_t* acceptor, int32_t* bnum, int32_t* bcoord,
                                         int32_t* slot_number, int32_t* max_checkpointed_slot, uint32_t* n,
                                         int32_t* slots, int64_t* req_ids, uint32_t cap_slots);

/* BATCHED_COMMIT: slots ascending (TreeSet); group members in the order given */
size_t gpx_wire_encode_batched_commit(const char* paxos_id, uint32_t paxos_id_len, int32_t version, int32_t bnum,
                                      int32_t bcoord, int32_t median_cp, uint32_t n_slots, const int32_t* slots,
                                      uint32_t n_group, const int32_t* group, uint8_t* out, size_t cap);
int gpx_wire_decode_batched_commit(const uint8_t* buf, size_t len, int32_t* version, char* paxos_id,
                                   uint32_t* paxos_id_len, int32_t* bnum, int32_t* bcoord, int32_t* median_cp,
                                   uint32_t* n_slots, int32_t* slots, uint32_t cap_slots, uint32_t* n_group,
                                   int32_t* group, uint32_t cap_group);

/* journal frame {int32 BE len}{bytes} (SQLPaxosLogger.java:1000-1003) */
size_t gpx_wire_journal_frame(const uint8_t* packet, size_t len, uint8_t* out, size_t cap);

/* PaxosPacketBatcher coalescing (PaxosPacketBatcher.java:97-179, BatchedCommit.addCommit :113-121):
 * fuse decisions (same gid, same ballot) into runs; median_cp of a run is the wrap-aware max.
 * decs must be grouped by gid.  run_start[k] is the index of the first decision of run k;
 * returns the number of runs (<= n). */
struct gpx_pvalue_hdr;
uint32_t gpx_wire_fuse_commits(uint32_t n, const struct gpx_pvalue_hdr* decs, uint32_t* run_start,
                               int32_t* run_median_cp);

#ifdef __cplusplus
}
#endif
#endif /* GPX_WIRE_H */
