"""GPU parity of the per-phase C-ABI calls (gpx_propose / gpx_handle_accepts /
gpx_handle_accept_replies / gpx_handle_decisions / gpx_patch) against the oracle under
adversarial schedules: every output record, every state row, every counter and every log byte
is compared (tests/fuzz.py)."""
import pytest

from fuzz import Fuzzer
from test_round_parity_gpu import compare_logs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed,R,W,fused", [(1, 3, 8, 0.0), (2, 3, 8, 0.5), (3, 5, 8, 0.0), (4, 3, 4, 0.5),
                                            (5, 5, 2, 0.5), (6, 3, 1, 0.0), (7, 3, 8, 1.0), (8, 5, 4, 1.0),
                                            (9, 4, 8, 0.5), (10, 2, 8, 0.5), (11, 1, 4, 0.5)])
def test_fuzz_parity(oracle_lib, cuda_lib, seed, R, W, fused):
    f = Fuzzer([oracle_lib, cuda_lib], G=64, R=R, W=W, seed=seed)
    decided = f.run(steps=60, check_every=5, fused_prob=fused)
    assert decided > 50
    c = f.engines[1].counters()
    assert c["accepts_nacked"] > 0 or R <= 2
    compare_logs(f.engines[0], f.engines[1], R)
    f.close()


def test_fuzz_parity_no_faults_many_groups(oracle_lib, cuda_lib):
    f = Fuzzer([oracle_lib, cuda_lib], G=3000, R=3, W=8, seed=9)
    decided = f.run(steps=12, rival=False, view_changes=False, stop_prob=0.002, check_every=4)
    assert decided > 3000
    f.close()


def test_fuzz_parity_gc_modes(oracle_lib, cuda_lib):
    """GC_MAJORITY_EXECUTED=false (maxCheckpointedSlot = lastCheckpointSlot), LOG_META_DECISIONS=false,
    ENABLE_JOURNALING=false (accepts stay in memory after execution) and a small CPI."""
    f = Fuzzer([oracle_lib, cuda_lib], G=64, R=3, W=8, seed=21, gc_majority_executed=0, log_meta_decisions=0,
               journaling_enabled=0, checkpoint_interval=3)
    f.run(steps=40, check_every=5, fused_prob=0.5)
    compare_logs(f.engines[0], f.engines[1], 3)
    f.close()


@pytest.mark.parametrize("seed,R,W,fn", [(41, 3, 8, "round"), (42, 3, 2, "round"), (43, 5, 4, "round"), (44, 3, 1, "round"),
                                         (45, 4, 8, "round"), (46, 3, 8, "round_phases"), (47, 2, 4, "round"),
                                         (48, 1, 2, "round")])
def test_fuzz_parity_with_fused_rounds(oracle_lib, cuda_lib, seed, R, W, fn):
    """gpx_round (k_round / k_round_slow) in the states adversarial schedules leave behind: outstanding proposals,
    NACKed ballots, resigned and re-installed coordinators, placeholders, duplicates, stopped groups"""
    f = Fuzzer([oracle_lib, cuda_lib], G=64, R=R, W=W, seed=seed)
    decided = f.run(steps=60, check_every=5, fused_prob=0.3, round_prob=0.5, round_fn=fn)
    assert decided > 50
    compare_logs(f.engines[0], f.engines[1], R)
    f.close()


def test_fuzz_parity_fused_rounds_gc_modes(oracle_lib, cuda_lib):
    f = Fuzzer([oracle_lib, cuda_lib], G=64, R=3, W=8, seed=51, gc_majority_executed=0, log_meta_decisions=0,
               journaling_enabled=0, checkpoint_interval=3)
    f.run(steps=40, check_every=5, fused_prob=0.5, round_prob=0.5)
    compare_logs(f.engines[0], f.engines[1], 3)
    f.close()


@pytest.mark.parametrize("seed,R,W,kw", [(71, 3, 8, {}), (72, 5, 4, {}), (73, 3, 2, {}), (74, 3, 8, {"journaling_enabled": 0}),
                                         (75, 1, 8, {}), (76, 4, 1, {})])
def test_fuzz_parity_with_prepares(oracle_lib, cuda_lib, seed, R, W, kw):
    """phase 1a on the device (k_prepare: PISM.handlePrepare / PaxosAcceptor.handlePrepare) inside adversarial
    schedules with fused rounds: replies, logged PREPARE images, state and ring heads must agree"""
    f = Fuzzer([oracle_lib, cuda_lib], G=48, R=R, W=W, seed=seed, **kw)
    decided = f.run(steps=60, check_every=5, fused_prob=0.3, round_prob=0.3, prepares=True)
    assert decided > 30
    compare_logs(f.engines[0], f.engines[1], R)
    f.close()


def test_handle_prepare_semantics_gpu(cuda_lib):
    from test_fuzz_cpu import test_handle_prepare_semantics
    test_handle_prepare_semantics(cuda_lib)
