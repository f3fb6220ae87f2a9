/*
 * gpx_pause.cuh -- k_pause_groups: the deactivation sweep as one launch (PaxosManager.Deactivator :2951 ->
 * syncAndDeactivate :2806-2900 -> pause(Map, dequeue) :2327-2366; the contract is the comment above gpx_pause_groups
 * in include/gpx.h).
 *
 * One thread per gid of the batch.  PISM.tryPause :2004-2035 at every local lane: the group is paused only if it is
 * live, not stopped, and every lane is caught up (PaxosAcceptor.caughtUp :452-459, PaxosCoordinator.caughtUp: nothing
 * committed-but-unexecuted, no proposal in flight, no live accepted pvalue when accepts are kept in memory).  Then the
 * HotRestoreInfo field set of every lane goes to out_rows[i * L + lane] (dump_row: what gpx_dump_rows writes) and the
 * gid is freed (free_group: what gpx_destroy_groups does -- forceStop + softCrash).  A group that does not pause is not
 * touched and its rows are not written.  gigapaxos keeps millions of idle groups by moving them out of memory
 * (PaxosConfig PAUSE_BATCH_SIZE :715, DEACTIVATION_PERIOD :291); here that is a sweep over the idle gids at HBM speed:
 * 4 + 188 * L bytes out and ~40 * L bytes in per paused group.
 *
 * Plain C++ over gpx_dev.cuh (no warp primitives, no shared memory, no inline PTX): tests/emu/ runs this source on the
 * host against the oracle.
 */
#pragma once
#include "gpx_dev.cuh"

struct PauseArgs {
  const uint32_t* gids;
  uint32_t n;
  gpx_row* rows;   /* [n * L] */
  uint8_t* paused; /* [n] */
};

#define GPX_PAUSE_BLOCK 128

__global__ void __launch_bounds__(GPX_PAUSE_BLOCK) k_pause_groups(const __grid_constant__ DevState S,
                                                                  const __grid_constant__ PauseArgs A) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) atomicAdd(&S.ctr[C_KERNEL_LAUNCHES], 1ull);
  if (i >= A.n) return;
  const uint32_t gid = A.gids[i];
  const GroupCtx g = group_ctx(S, gid);
  bool ok = g.live;
  for (uint32_t l = 0; ok && l < S.L; l++) {
    if (g.ms->idx_of_lane[l] == 0xffu) continue; /* this lane hosts no replica of the group */
    if (!st_usable(S.acc_aux[row_idx(S, l, gid)])) ok = false; /* stopped / recovering: not a pause candidate */
    else if (group_flags(S, l, gid) & GPX_GF_NOT_CAUGHT_UP_BIT) ok = false;
  }
  A.paused[i] = ok ? 1 : 0;
  if (!ok) return;
  for (uint32_t l = 0; l < S.L; l++) {
    gpx_row r;
    dump_row(S, l, gid, r);
    A.rows[(size_t)i * S.L + l] = r;
  }
  free_group(S, gid);
}

/* ---- k_select_groups: the groups a sweep has to look at (gpx_select_groups) -------------------------------------------
 * One thread per gid of the engine: live, this lane hosts a replica, its acceptor is ACTIVE and
 * (group_flags & mask) == value -> the gid is appended (one atomicAdd per match; a sweep's matches are normally few --
 * the groups that need a sync -- or the result is consumed in chunks of a pause batch).  The host sorts what comes
 * back. */
struct SelectArgs {
  uint32_t lane, mask, value, cap;
  uint32_t* gids;           /* [cap] */
  unsigned long long* n_found;
};

__global__ void __launch_bounds__(GPX_PAUSE_BLOCK) k_select_groups(const __grid_constant__ DevState S,
                                                                   const __grid_constant__ SelectArgs A) {
  const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid == 0) atomicAdd(&S.ctr[C_KERNEL_LAUNCHES], 1ull);
  if (gid >= S.G) return;
  const GroupCtx g = group_ctx(S, gid);
  if (!g.live || g.ms->idx_of_lane[A.lane] == 0xffu) return;
  if (!st_usable(S.acc_aux[row_idx(S, A.lane, gid)])) return;
  if ((group_flags(S, A.lane, gid) & A.mask) != A.value) return;
  const unsigned long long k = atomicAdd(A.n_found, 1ull);
  if (k < A.cap) A.gids[k] = gid;
}

/* gpx_clear_group_flags: the host has dealt with these slow-path entries */
__global__ void k_clear_flags(const __grid_constant__ DevState S, uint32_t lane, const uint32_t* gids, uint32_t n,
                              uint32_t mask) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || gids[i] >= S.G) return;
  const size_t ri = row_idx(S, lane, gids[i]);
  S.acc_aux[ri] &= ~((mask & (GPX_GF_OVERFLOW | GPX_GF_NEEDS_SYNC)) << 24);
}

/* ---- k_missing_decisions: the fields of a SYNC_DECISIONS_REQUEST for a batch of groups (gpx_missing_decisions) ---------
 * One thread per gid.  committedRequests is the aux word's present mask over the window (slot s lives at s mod W, s in
 * [slot, slot + W)), hasRequestValue its valued mask, acceptedProposals.containsKey(s) a live accepted-window entry of
 * exactly that slot. */
struct MissingArgs {
  uint32_t lane, n;
  const uint32_t* gids;
  int32_t size_limit, too_much_gap;
  gpx_missing_rec* out;
};

__global__ void __launch_bounds__(GPX_PAUSE_BLOCK) k_missing_decisions(const __grid_constant__ DevState S,
                                                                       const __grid_constant__ MissingArgs A) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) atomicAdd(&S.ctr[C_KERNEL_LAUNCHES], 1ull);
  if (i >= A.n) return;
  const uint32_t gid = A.gids[i];
  gpx_missing_rec r;
  memset(&r, 0, sizeof r);
  r.gid = gid;
  const GroupCtx g = group_ctx(S, gid);
  if (g.live && g.ms->idx_of_lane[A.lane] != 0xffu) {
    const size_t ri = row_idx(S, A.lane, gid);
    const uint32_t aux = S.acc_aux[ri];
    const int4 row = S.acc_row[ri];
    const int slot = row.x;
    const uint32_t Wm = S.W - 1, present = GPX_AUX_PRESENT(aux), valued = GPX_AUX_VALUED(aux);
    const bool stopped = GPX_AUX_STATE(aux) == GPX_ST_STOPPED;
    r.slot = slot;
    r.flags = (uint8_t)group_flags(S, A.lane, gid);
    int maxCommitted = (int)((unsigned)slot - 1u); /* getMaxCommittedSlot :425-438 */
    if (!stopped)
      for (uint32_t d = 0; d < S.W; d++) {
        const int s = (int)((unsigned)slot + d);
        if ((present >> ((uint32_t)s & Wm)) & 1u) maxCommitted = s;
      }
    r.max_decision_slot = maxCommitted;
    if (st_usable(aux)) { /* a stopped (or recovering) acceptor asks for nothing: :407-408 */
      uint32_t nm = 0;
      const int limit = (int)((unsigned)slot + (unsigned)A.size_limit);
      for (uint32_t d = 0; d < S.W; d++) { /* getMissingCommittedSlots :415-421 */
        const int s = (int)((unsigned)slot + d);
        if (!(jsub(s, maxCommitted) < 0 && jsub(s, limit) < 0)) break;
        const uint32_t w = (uint32_t)s & Wm;
        bool missing = !((present >> w) & 1u);
        if (!missing && !((valued >> w) & 1u)) { /* a value-less commit: is its accept here? */
          const size_t ai = 2 * win_idx(S, A.lane, w, gid);
          const int4 a0 = S.acc_win[ai], a1 = S.acc_win[ai + 1];
          missing = !(((unsigned)a1.w & GPX_ENT_VALID) && a0.x == s && jsub(s, row.w) > 0);
        }
        if (missing) r.missing[nm++] = s;
      }
      if (nm == 0) r.missing[nm++] = slot; /* :2297-2298 */
      r.n_missing = (uint16_t)nm;
      /* isMissingTooMuch :2367-2370 = shouldSync(maxCommitted, gap) :2341-2361, DEFAULT_SYNC */
      const int gap = jsub(maxCommitted, slot), th = A.too_much_gap;
      const bool nontrivialInitialGap = gap >= th / 100, smallGapThreshold = th <= 1;
      r.missing_too_much = (gap >= th) || ((slot == 0 || slot == 1) && (nontrivialInitialGap || smallGapThreshold));
    }
  }
  A.out[i] = r;
}
