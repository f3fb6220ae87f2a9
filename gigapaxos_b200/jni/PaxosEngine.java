/*
 * The Java half of the binding (INTEGRATION.md section 2): the class gpx_jni.c's functions are named after.  Not
 * compiled in this repository (no JDK in the image: profiles/r2_java_probe_gpu_box.txt); tests/test_jni_shim.py checks
 * that every native method below has a JNI function of the same name and parameter types in gpx_jni.c and the
 * other way round, so the two files cannot drift apart.
 *
 * Records travel in direct ByteBuffers (ByteOrder.LITTLE_ENDIAN) laid out like the structs of include/gpx.h; a
 * maintainer fills them the way PaxosPacket.toBytes fills its buffers (field offsets: gigapaxos_b200/abi.py dtypes
 * = include/gpx.h structs, checked by tests/test_abi.py).
 */
package edu.umass.cs.gigapaxos.gpx;

import java.nio.ByteBuffer;

public final class PaxosEngine implements AutoCloseable {
	static {
		System.loadLibrary("gpxjni"); // libgpxjni.so, linked against libgpx.so
	}

	private long handle; // gpx_engine*

	/**
	 * @param propertiesPath gigapaxos.properties (the keys the engine consumes: gpx_config_from_properties), or null
	 * @param cfg            sizeof(gpx_config) bytes; zeroed = defaults, or pre-filled by the caller
	 */
	public PaxosEngine(String propertiesPath, ByteBuffer cfg) {
		handle = create(propertiesPath, cfg);
		if (handle == 0)
			throw new IllegalStateException(lastError());
	}

	public long handle() {
		return handle;
	}

	@Override
	public void close() {
		if (handle != 0) {
			destroy(handle);
			handle = 0;
		}
	}

	// ---- life cycle ----
	static native long create(String propertiesPath, ByteBuffer cfg);
	static native void destroy(long h);
	static native String lastError();

	// ---- groups: PaxosManager.createPaxosInstance :632 / :664-691, kill :2162, pause / unpause (HotRestoreInfo) ----
	public static native int createGroups(long h, int n, ByteBuffer groupDescs);
	public static native int destroyGroups(long h, int n, ByteBuffer gids);
	public static native int dumpRows(long h, int n, ByteBuffer gids, int lane, ByteBuffer rowsOut);
	public static native int loadRows(long h, int n, ByteBuffer rows);
	public static native int patch(long h, int n, ByteBuffer patches);

	// ---- one round for co-located replicas, pipelined: RequestBatcher.dequeueImpl :168-234 hands the batch over ----
	/** @return the ticket (>= 0) or a negative gpx error code */
	public static native long roundSubmit(long h, int n, int flags, ByteBuffer reqs, ByteBuffer payload,
			long payloadBytes, ByteBuffer status, ByteBuffer exec, ByteBuffer sum, ByteBuffer extra, int extraCap);
	/** @return (nExecSlots << 32) | nExtra, or a negative gpx error code */
	public static native long roundWait(long h, long ticket);

	// ---- per packet type: PaxosInstanceStateMachine.handlePaxosMessage :423 ----
	public static native int propose(long h, int n, ByteBuffer reqs, ByteBuffer payload, long payloadBytes,
			ByteBuffer acceptsOut, ByteBuffer nAccepts, ByteBuffer blobOut, ByteBuffer blobBytes, ByteBuffer status);
	public static native int handleAccepts(long h, int n, ByteBuffer accepts, ByteBuffer blob, long blobBytes,
			ByteBuffer repliesOut, ByteBuffer extraExecOut, int extraCap, ByteBuffer nExtra);
	public static native int handleAcceptReplies(long h, int n, ByteBuffer replies, ByteBuffer decisionsOut,
			ByteBuffer nDecisions);
	public static native int handleDecisions(long h, int n, ByteBuffer decisions, ByteBuffer execOut,
			ByteBuffer extraExecOut, int extraCap, ByteBuffer nExtra);
	/** the live ACTIVE groups of a lane whose flag byte satisfies (flags & mask) == value, ascending; nFound: one int */
	public static native int selectGroups(long h, int lane, int mask, int value, ByteBuffer gidsOut, int cap, ByteBuffer nFound);
	/** per group: next slot, highest committed slot, the missing slots, isMissingTooMuch (48 B each): a SYNC_DECISIONS_REQUEST */
	public static native int missingDecisions(long h, int lane, int n, ByteBuffer gids, int sizeLimit, int tooMuchGap, ByteBuffer out);
	/** the host has dealt with these slow-path entries: clears OVERFLOW / NEEDS_SYNC (mask) at the lane */
	public static native int clearGroupFlags(long h, int lane, int n, ByteBuffer gids, int mask);
	/** the Deactivator's batch (PaxosManager.pause(Map, dequeue)): rowsOut n x nLanes x 188 B, pausedOut n bytes; unpause = loadRows */
	public static native int pauseGroups(long h, int n, ByteBuffer gids, ByteBuffer rowsOut, ByteBuffer pausedOut);
	public static native int handlePrepares(long h, int n, ByteBuffer prepares, ByteBuffer prepareRepliesOut);
	/** phase 1b for n elections (gpx_election_rec, 32 B each) over nReplyRecs gpx_prepare_reply_rec; electionsOut: n x 896 B */
	public static native int handlePrepareReplies(long h, int n, ByteBuffer elections, int nReplyRecs, ByteBuffer replies,
			ByteBuffer electionsOut);

	// ---- journal: AbstractPaxosLogger.BatchedLogger :691-716, SQLPaxosLogger.journal :965-1036 ----
	/** fromAndBytes = {ring position the copy starts at, bytes being copied} */
	public static native int logDrainAsync(long h, int lane, ByteBuffer dst, long[] fromAndBytes);
	public static native int logDrainWait(long h);
	public static native int logRelease(long h, int lane, long upto);
	/** copiedAndHead = {bytes copied, ring head} */
	/** getLoggedDecisions / getLoggedAccepts for n (gid, minSlot, nSlots <= 16) wants sorted by gid: hitsOut n x 16 x 96 B */
	public static native int logFind(long h, int lane, long from, int n, ByteBuffer wants, ByteBuffer hitsOut);
	/** the request bodies of n ranges {long pos, int len, int dstOff} of a lane's log ring into dst, one copy */
	public static native int logGather(long h, int lane, int n, ByteBuffer ranges, ByteBuffer dst);
	public static native int logRead(long h, int lane, long from, ByteBuffer dst, long[] copiedAndHead);

	// ---- replicas of a group on different GPUs: one engine (a single lane) per GPU process ----
	public static native int spreadUniqueId(ByteBuffer id128);
	public static native int spreadPlanNode(ByteBuffer cfg, int rank, ByteBuffer planOut);
	/** collective over the spread group; @return gpx_spread* or a negative gpx error code */
	public static native long spreadCreate(long h, ByteBuffer cfg, int rank, ByteBuffer id128);
	public static native int spreadRound(long spread, ByteBuffer io, long cudaStream);
	public static native long spreadDropped(long spread, int localIndex);
	public static native void spreadDestroy(long spread);

	// ---- odds and ends ----
	public static native int getCounters(long h, ByteBuffer countersOut);
	/** PaxosInstanceStateMachine.getCPI */
	public static native int getCpi(int cpi, double noise, int nameHash);
}
