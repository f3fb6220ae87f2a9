/*
 * gpx_jni.c -- the JNI shim between the reference's Java host code and libgpx.so (SURVEY.md 8b / 8f rank 4,
 * INTEGRATION.md).  NOT compiled in this repository's image (no JDK, no jni.h -- profiles/r2_java_probe_gpu_box.txt);
 * it is the file a gigapaxos maintainer builds next to the jar:
 *
 *     gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude \
 *         gigapaxos_b200/jni/gpx_jni.c -Lgigapaxos_b200 -lgpx -o libgpxjni.so
 *
 * Every function is a 1:1 pass-through: records travel in direct ByteBuffers laid out exactly like the structs of
 * include/gpx.h (little-endian; ByteBuffer.order(ByteOrder.LITTLE_ENDIAN)), so nothing is copied or converted on
 * the way.  Java side: edu.umass.cs.gigapaxos.gpx.PaxosEngine, used from
 *   PaxosManager.createPaxosInstance  :632 / :664   -> createGroups
 *   PaxosManager.kill                 :2162          -> destroyGroups
 *   RequestBatcher.dequeueImpl        :168-234       -> roundSubmit   (the batch the RequestBatcher would hand to PISM)
 *   PaxosManager.executed / app.execute :311, PISM.execute :1802 <- roundWait (EXEC summaries / records)
 *   AbstractPaxosLogger.BatchedLogger :691-716       -> logDrainAsync / logDrainWait / logRelease
 *   PaxosMessenger / PaxosPacketBatcher (replicas on other GPUs) -> spreadRound
 */
#include <jni.h>
#include <stdint.h>

#include "gpx.h"

#define GPX_JNI(name) Java_edu_umass_cs_gigapaxos_gpx_PaxosEngine_##name

static void* buf(JNIEnv* env, jobject bb) { return bb ? (*env)->GetDirectBufferAddress(env, bb) : NULL; }

/* long create(String propertiesPath, ByteBuffer cfg): cfg = sizeof(gpx_config) bytes, may be pre-filled */
JNIEXPORT jlong JNICALL GPX_JNI(create)(JNIEnv* env, jclass cls, jstring props, jobject cfgbuf) {
  gpx_config* cfg = (gpx_config*)buf(env, cfgbuf);
  if (!cfg) return 0;
  if (cfg->abi_version == 0) gpx_config_defaults(cfg);
  if (props) {
    const char* p = (*env)->GetStringUTFChars(env, props, NULL);
    int rc = gpx_config_from_properties(p, cfg); /* the gigapaxos.properties keys the engine consumes */
    (*env)->ReleaseStringUTFChars(env, props, p);
    if (rc) return 0;
  }
  gpx_engine* e = NULL;
  return gpx_engine_create(cfg, &e) == GPX_OK ? (jlong)(intptr_t)e : 0;
}
JNIEXPORT void JNICALL GPX_JNI(destroy)(JNIEnv* env, jclass cls, jlong h) { gpx_engine_destroy((gpx_engine*)(intptr_t)h); }
JNIEXPORT jstring JNICALL GPX_JNI(lastError)(JNIEnv* env, jclass cls) { return (*env)->NewStringUTF(env, gpx_last_error()); }

/* int createGroups(long h, int n, ByteBuffer descs): descs = gpx_group_desc[n] */
JNIEXPORT jint JNICALL GPX_JNI(createGroups)(JNIEnv* env, jclass cls, jlong h, jint n, jobject descs) {
  return gpx_create_groups((gpx_engine*)(intptr_t)h, (uint32_t)n, (const gpx_group_desc*)buf(env, descs));
}
JNIEXPORT jint JNICALL GPX_JNI(destroyGroups)(JNIEnv* env, jclass cls, jlong h, jint n, jobject gids) {
  return gpx_destroy_groups((gpx_engine*)(intptr_t)h, (uint32_t)n, (const uint32_t*)buf(env, gids));
}
JNIEXPORT jint JNICALL GPX_JNI(dumpRows)(JNIEnv* env, jclass cls, jlong h, jint n, jobject gids, jint lane, jobject out) {
  return gpx_dump_rows((gpx_engine*)(intptr_t)h, (uint32_t)n, (const uint32_t*)buf(env, gids), (uint32_t)lane,
                       (gpx_row*)buf(env, out));
}
JNIEXPORT jint JNICALL GPX_JNI(loadRows)(JNIEnv* env, jclass cls, jlong h, jint n, jobject rows) {
  return gpx_load_rows((gpx_engine*)(intptr_t)h, (uint32_t)n, (const gpx_row*)buf(env, rows));
}
JNIEXPORT jint JNICALL GPX_JNI(patch)(JNIEnv* env, jclass cls, jlong h, jint n, jobject patches) {
  return gpx_patch((gpx_engine*)(intptr_t)h, (uint32_t)n, (const gpx_patch_rec*)buf(env, patches));
}

/* long roundSubmit(long h, int n, int flags, ByteBuffer reqs, ByteBuffer payload, long payloadBytes,
 *                  ByteBuffer status, ByteBuffer exec, ByteBuffer sum, ByteBuffer extra, int extraCap)
 * returns the ticket (>= 0) or a negative gpx error code.  All buffers are direct (pinned by the JVM for the
 * duration of the call chain; allocate them once and reuse them: cudaHostRegister them for overlapped copies). */
JNIEXPORT jlong JNICALL GPX_JNI(roundSubmit)(JNIEnv* env, jclass cls, jlong h, jint n, jint flags, jobject reqs,
                                             jobject payload, jlong payload_bytes, jobject status, jobject exec,
                                             jobject sum, jobject extra, jint extra_cap) {
  gpx_round_io io;
  io.n = (uint32_t)n;
  io.flags = (uint32_t)flags;
  io.reqs = (const gpx_request_rec*)buf(env, reqs);
  io.payload = (const uint8_t*)buf(env, payload);
  io.payload_bytes = (uint64_t)payload_bytes;
  io.status = (int32_t*)buf(env, status);
  io.exec = (gpx_exec_rec*)buf(env, exec);
  io.sum = (gpx_exec_sum*)buf(env, sum);
  io.extra = (gpx_exec_rec*)buf(env, extra);
  io.extra_cap = (uint32_t)extra_cap;
  uint64_t ticket = 0;
  int rc = gpx_round_submit((gpx_engine*)(intptr_t)h, &io, &ticket);
  return rc == GPX_OK ? (jlong)ticket : (jlong)rc;
}
/* long roundWait(long h, long ticket): (nExecSlots << 32) | nExtra, or a negative error code */
JNIEXPORT jlong JNICALL GPX_JNI(roundWait)(JNIEnv* env, jclass cls, jlong h, jlong ticket) {
  uint32_t ns = 0, nx = 0;
  int rc = gpx_round_wait((gpx_engine*)(intptr_t)h, (uint64_t)ticket, &ns, &nx);
  return rc == GPX_OK ? (((jlong)ns) << 32) | (jlong)nx : (jlong)rc;
}

/* journal: long logDrainAsync(long h, int lane, ByteBuffer dst, long[] fromAndBytes) */
JNIEXPORT jint JNICALL GPX_JNI(logDrainAsync)(JNIEnv* env, jclass cls, jlong h, jint lane, jobject dst, jlongArray out) {
  uint64_t from = 0, nb = 0;
  jlong cap = (*env)->GetDirectBufferCapacity(env, dst);
  int rc = gpx_log_drain_async((gpx_engine*)(intptr_t)h, (uint32_t)lane, buf(env, dst), (uint64_t)cap, &from, &nb, NULL);
  jlong v[2] = {(jlong)from, (jlong)nb};
  (*env)->SetLongArrayRegion(env, out, 0, 2, v);
  return rc;
}
JNIEXPORT jint JNICALL GPX_JNI(logDrainWait)(JNIEnv* env, jclass cls, jlong h) { return gpx_log_drain_wait((gpx_engine*)(intptr_t)h); }
JNIEXPORT jint JNICALL GPX_JNI(logRelease)(JNIEnv* env, jclass cls, jlong h, jint lane, jlong upto) {
  return gpx_log_release((gpx_engine*)(intptr_t)h, (uint32_t)lane, (uint64_t)upto);
}

/* The per-packet-type entry points: what PaxosInstanceStateMachine.handlePaxosMessage :423 (switch :498-558)
 * dispatches to, one call per batch of same-type packets (PaxosPacketBatcher / PaxosManager.handleIncomingPacket hand
 * them over demuxed).  An adapter that keeps the reference's messenger between replicas uses these, not roundSubmit.
 *   REQUEST / PROPOSAL -> propose            (PISM.handleProposal :818, PaxosCoordinatorState.propose :233)
 *   ACCEPT             -> handleAccepts      (PISM.handleAccept :1080, PaxosAcceptor.acceptAndUpdateBallot :302)
 *   ACCEPT_REPLY       -> handleAcceptReplies(PISM.handleAcceptReply :1248, PaxosCoordinatorState.handleAcceptReplyMyBallot :597)
 *   DECISION           -> handleDecisions    (PISM.handleCommittedRequest :1432, extractExecuteAndCheckpoint :1619)
 *   PREPARE            -> handlePrepares     (PISM.handlePrepare :900, PaxosAcceptor.handlePrepare :239)
 *   PREPARE_REPLY      -> handlePrepareReplies (PISM.handlePrepareReply :1017, PaxosCoordinatorState :264-587: tally,
 *                         carry-over, no-op fill, coordinator installed ACTIVE; the plan comes back for re-proposal)
 * counts come back through a direct IntBuffer / LongBuffer of one element. */
JNIEXPORT jint JNICALL GPX_JNI(propose)(JNIEnv* env, jclass cls, jlong h, jint n, jobject reqs, jobject payload,
                                        jlong payload_bytes, jobject out_accepts, jobject n_accepts, jobject out_blob,
                                        jobject blob_bytes, jobject status) {
  jlong cap = out_blob ? (*env)->GetDirectBufferCapacity(env, out_blob) : 0;
  return gpx_propose((gpx_engine*)(intptr_t)h, (uint32_t)n, (const gpx_request_rec*)buf(env, reqs),
                     (const uint8_t*)buf(env, payload), (uint64_t)payload_bytes, (gpx_accept_rec*)buf(env, out_accepts),
                     (uint32_t*)buf(env, n_accepts), (uint8_t*)buf(env, out_blob), (uint64_t)cap,
                     (uint64_t*)buf(env, blob_bytes), (int32_t*)buf(env, status));
}
JNIEXPORT jint JNICALL GPX_JNI(handleAccepts)(JNIEnv* env, jclass cls, jlong h, jint n, jobject accepts, jobject blob,
                                              jlong blob_bytes, jobject out_replies, jobject out_extra, jint extra_cap,
                                              jobject n_extra) {
  return gpx_handle_accepts((gpx_engine*)(intptr_t)h, (uint32_t)n, (const gpx_accept_rec*)buf(env, accepts),
                            (const uint8_t*)buf(env, blob), (uint64_t)blob_bytes, (gpx_accept_reply_rec*)buf(env, out_replies),
                            (gpx_exec_rec*)buf(env, out_extra), (uint32_t)extra_cap, (uint32_t*)buf(env, n_extra));
}
JNIEXPORT jint JNICALL GPX_JNI(handleAcceptReplies)(JNIEnv* env, jclass cls, jlong h, jint n, jobject replies,
                                                    jobject out_decisions, jobject n_decisions) {
  return gpx_handle_accept_replies((gpx_engine*)(intptr_t)h, (uint32_t)n, (const gpx_accept_reply_rec*)buf(env, replies),
                                   (gpx_decision_rec*)buf(env, out_decisions), (uint32_t*)buf(env, n_decisions));
}
JNIEXPORT jint JNICALL GPX_JNI(handleDecisions)(JNIEnv* env, jclass cls, jlong h, jint n, jobject decisions,
                                                jobject out_exec, jobject out_extra, jint extra_cap, jobject n_extra) {
  return gpx_handle_decisions((gpx_engine*)(intptr_t)h, (uint32_t)n, (const gpx_decision_rec*)buf(env, decisions),
                              (gpx_exec_rec*)buf(env, out_exec), (gpx_exec_rec*)buf(env, out_extra), (uint32_t)extra_cap,
                              (uint32_t*)buf(env, n_extra));
}
JNIEXPORT jint JNICALL GPX_JNI(handlePrepares)(JNIEnv* env, jclass cls, jlong h, jint n, jobject prepares, jobject out_replies) {
  return gpx_handle_prepares((gpx_engine*)(intptr_t)h, (uint32_t)n, (const gpx_pvalue_hdr*)buf(env, prepares),
                             (gpx_prepare_reply_rec*)buf(env, out_replies));
}
/* int selectGroups(long h, int lane, int mask, int value, ByteBuffer gidsOut [cap x 4 B], int cap, ByteBuffer nFound [4 B]):
 * the groups a sweep of PaxosManager has to look at (slow-path list, pause candidates) */
JNIEXPORT jint JNICALL GPX_JNI(selectGroups)(JNIEnv* env, jclass cls, jlong h, jint lane, jint mask, jint value, jobject gids_out,
                                             jint cap, jobject n_found) {
  return gpx_select_groups((gpx_engine*)(intptr_t)h, (uint32_t)lane, (uint32_t)mask, (uint32_t)value, (uint32_t*)buf(env, gids_out),
                           (uint32_t)cap, (uint32_t*)buf(env, n_found));
}
/* int missingDecisions(long h, int lane, int n, ByteBuffer gids, int sizeLimit, int tooMuchGap, ByteBuffer out [n x 48 B]):
 * the fields of the SYNC_DECISIONS_REQUESTs PISM.requestMissingDecisions :2292 would send for these groups */
JNIEXPORT jint JNICALL GPX_JNI(missingDecisions)(JNIEnv* env, jclass cls, jlong h, jint lane, jint n, jobject gids, jint size_limit,
                                                 jint too_much_gap, jobject out) {
  return gpx_missing_decisions((gpx_engine*)(intptr_t)h, (uint32_t)lane, (uint32_t)n, (const uint32_t*)buf(env, gids),
                               (int32_t)size_limit, (int32_t)too_much_gap, (gpx_missing_rec*)buf(env, out));
}
JNIEXPORT jint JNICALL GPX_JNI(clearGroupFlags)(JNIEnv* env, jclass cls, jlong h, jint lane, jint n, jobject gids, jint mask) {
  return gpx_clear_group_flags((gpx_engine*)(intptr_t)h, (uint32_t)lane, (uint32_t)n, (const uint32_t*)buf(env, gids), (uint32_t)mask);
}
/* int pauseGroups(long h, int n, ByteBuffer gids, ByteBuffer rowsOut [n x nLanes x gpx_row], ByteBuffer pausedOut [n]):
 * PaxosManager.pause(Map, dequeue) :2327-2366 for the Deactivator's batch; unpause is loadRows */
JNIEXPORT jint JNICALL GPX_JNI(pauseGroups)(JNIEnv* env, jclass cls, jlong h, jint n, jobject gids, jobject rows_out,
                                            jobject paused_out) {
  return gpx_pause_groups((gpx_engine*)(intptr_t)h, (uint32_t)n, (const uint32_t*)buf(env, gids), (gpx_row*)buf(env, rows_out),
                          (uint8_t*)buf(env, paused_out));
}
JNIEXPORT jint JNICALL GPX_JNI(handlePrepareReplies)(JNIEnv* env, jclass cls, jlong h, jint n, jobject elections,
                                                     jint n_reply_recs, jobject replies, jobject out_elections) {
  return gpx_handle_prepare_replies((gpx_engine*)(intptr_t)h, (uint32_t)n, (const gpx_election_rec*)buf(env, elections),
                                    (uint32_t)n_reply_recs, (const gpx_prepare_reply_rec*)buf(env, replies),
                                    (gpx_election_out*)buf(env, out_elections));
}
/* int logFind(long h, int lane, long from, int n, ByteBuffer wants [n x 16 B], ByteBuffer hitsOut [n x 16 x 96 B]):
 * AbstractPaxosLogger.getLoggedDecisions / getLoggedAccepts for a batch of (group, slot range) as a scan of the log ring */
JNIEXPORT jint JNICALL GPX_JNI(logFind)(JNIEnv* env, jclass cls, jlong h, jint lane, jlong from, jint n, jobject wants,
                                        jobject hits_out) {
  return gpx_log_find((gpx_engine*)(intptr_t)h, (uint32_t)lane, (uint64_t)from, (uint32_t)n, (const gpx_log_want*)buf(env, wants),
                      (gpx_log_hit*)buf(env, hits_out));
}
/* int logGather(long h, int lane, int n, ByteBuffer ranges [n x 16 B: pos, len, dstOff], ByteBuffer dst): the request
 * bodies of a batch of logFind hits in one device->host copy */
JNIEXPORT jint JNICALL GPX_JNI(logGather)(JNIEnv* env, jclass cls, jlong h, jint lane, jint n, jobject ranges, jobject dst) {
  jlong cap = dst ? (*env)->GetDirectBufferCapacity(env, dst) : 0;
  return gpx_log_gather((gpx_engine*)(intptr_t)h, (uint32_t)lane, (uint32_t)n, (const gpx_log_range*)buf(env, ranges), buf(env, dst),
                        (uint64_t)cap);
}
/* long logRead(long h, int lane, long from, ByteBuffer dst, long[] out {nCopied, head}): the synchronous journal read
 * (recovery / tests); the steady state uses logDrainAsync */
JNIEXPORT jint JNICALL GPX_JNI(logRead)(JNIEnv* env, jclass cls, jlong h, jint lane, jlong from, jobject dst, jlongArray out) {
  uint64_t nc = 0, head = 0;
  jlong cap = dst ? (*env)->GetDirectBufferCapacity(env, dst) : 0;
  int rc = gpx_log_read((gpx_engine*)(intptr_t)h, (uint32_t)lane, (uint64_t)from, buf(env, dst), (uint64_t)cap, &nc, &head);
  jlong v[2] = {(jlong)nc, (jlong)head};
  (*env)->SetLongArrayRegion(env, out, 0, 2, v);
  return rc;
}
/* PaxosInstanceStateMachine.getCPI / roundRobinCoordinator and String.hashCode, for an adapter that wants the engine's
 * arithmetic rather than its own (they are bit-identical: tests/test_oracle.py) */
JNIEXPORT jint JNICALL GPX_JNI(getCpi)(JNIEnv* env, jclass cls, jint cpi, jdouble noise, jint name_hash) {
  return gpx_get_cpi(cpi, noise, name_hash);
}

/* spread placement (one JVM per GPU): long spreadCreate(long h, ByteBuffer cfg, int rank, ByteBuffer ncclId128) */
JNIEXPORT jint JNICALL GPX_JNI(spreadUniqueId)(JNIEnv* env, jclass cls, jobject id128) { return gpx_spread_unique_id(buf(env, id128)); }
JNIEXPORT jlong JNICALL GPX_JNI(spreadCreate)(JNIEnv* env, jclass cls, jlong h, jobject cfg, jint rank, jobject id128) {
  gpx_spread* sp = NULL;
  int rc = gpx_spread_create_nccl((gpx_engine*)(intptr_t)h, (const gpx_spread_config*)buf(env, cfg), (uint32_t)rank,
                                  buf(env, id128), &sp);
  return rc == GPX_OK ? (jlong)(intptr_t)sp : (jlong)rc;
}
/* int spreadRound(long sp, ByteBuffer io, long stream): io = gpx_spread_io with DEVICE pointers (the adapter keeps its
 * request staging and result buffers in device memory and moves them with its own cudaMemcpyAsync calls) */
JNIEXPORT jint JNICALL GPX_JNI(spreadRound)(JNIEnv* env, jclass cls, jlong sp, jobject io, jlong stream) {
  return gpx_spread_round((gpx_spread*)(intptr_t)sp, (const gpx_spread_io*)buf(env, io), (void*)(intptr_t)stream);
}
JNIEXPORT jint JNICALL GPX_JNI(spreadPlanNode)(JNIEnv* env, jclass cls, jobject cfg, jint rank, jobject out_plan) {
  return gpx_spread_plan_node((const gpx_spread_config*)buf(env, cfg), (uint32_t)rank, (gpx_spread_plan*)buf(env, out_plan));
}
/* long spreadDropped(long sp, int localIndex): records a bucket could not hold (0 in a correctly sized plan) */
JNIEXPORT jlong JNICALL GPX_JNI(spreadDropped)(JNIEnv* env, jclass cls, jlong sp, jint local_index) {
  uint32_t d = 0;
  int rc = gpx_spread_dropped((gpx_spread*)(intptr_t)sp, (uint32_t)local_index, &d);
  return rc == GPX_OK ? (jlong)d : (jlong)rc;
}
JNIEXPORT void JNICALL GPX_JNI(spreadDestroy)(JNIEnv* env, jclass cls, jlong sp) { gpx_spread_destroy((gpx_spread*)(intptr_t)sp); }

JNIEXPORT jint JNICALL GPX_JNI(getCounters)(JNIEnv* env, jclass cls, jlong h, jobject out) {
  return gpx_get_counters((gpx_engine*)(intptr_t)h, (gpx_counters*)buf(env, out));
}
