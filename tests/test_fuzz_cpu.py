"""CPU: adversarial schedules against the oracle alone -- Paxos safety (agreement, gap-free
in-order execution, TESTPaxosApp.java:179-213) must hold under NACKs, preemption, duplicates,
reordering, lost multicasts, commits overtaking accepts, STOPs and coordinator changes; and the
bounded-window engine rules (W=8) must agree with the unbounded Java maps (W=0) whenever no
overflow was flagged."""
import numpy as np
import pytest

from fuzz import Fuzzer


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
@pytest.mark.parametrize("R,fused", [(3, 0.0), (5, 0.0), (3, 0.5), (5, 1.0)])
def test_safety_under_adversarial_schedules(oracle_lib, seed, R, fused):
    f = Fuzzer([oracle_lib], G=48, R=R, W=8, seed=seed)
    decided = f.run(steps=50, fused_prob=fused)
    assert decided > 100
    c = f.engines[0].counters()
    assert c["accepts_nacked"] > 0 and c["preempted"] >= 0
    assert fused == 1.0 or c["placeholders"] > 0
    f.close()


def test_unbounded_and_windowed_oracle_agree(oracle_lib):
    """Same schedule through W=0 (TreeMap semantics) and W=8; identical unless overflow fired."""
    class Two(Fuzzer):
        pass
    f = Fuzzer([oracle_lib, oracle_lib], G=40, R=3, W=8, seed=11)
    # re-create engine 0 unbounded
    from helpers import Engine, group_descs, make_config
    f.engines[0].close()
    f.engines[0] = Engine(oracle_lib, make_config(oracle_lib, max_groups=40, window=0, max_batch_recs=1 << 14,
                                                  max_batch_payload=1 << 20))
    f.engines[0].create_groups(group_descs(40))
    try:
        f.run(steps=40, rival=False, view_changes=False, stop_prob=0.0)
    except AssertionError as e:
        # only acceptable divergence: the bounded engine dropped records beyond the window
        c = f.engines[1].counters()
        assert c["window_overflow"] > 0, str(e)
    f.close()


@pytest.mark.parametrize("seed,R,fn", [(31, 3, "round"), (32, 5, "round"), (33, 3, "round_phases")])
def test_fused_rounds_inside_adversarial_schedules(oracle_lib, seed, R, fn):
    """whole rounds (gpxo_round) interleaved with lossy / duplicated / reordered phase steps, rival coordinators
    and view changes: safety holds, and the windowed and the unbounded oracle agree"""
    f = Fuzzer([oracle_lib], G=48, R=R, W=8, seed=seed)
    assert f.run(steps=60, fused_prob=0.3, round_prob=0.5, round_fn=fn) > 100
    f.close()
