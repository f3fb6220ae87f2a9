#!/usr/bin/env python3
"""Soak the CPU-side checks of the kernels beside the round with many random configurations (no GPU needed):

  * the kernel sources compiled for the host (tests/emu) against the oracle: k_prepare_tally, k_pause_groups,
    k_log_dir / k_log_scan / k_log_hits (as the ring lies and re-laid to wrap);
  * the oracle's phase 1b against the host-language twin (two restatements of PaxosCoordinatorState.java:264-587).

A run's sanity assertions ("some but not all groups paused", "every verdict occurred") can miss for an unlucky seed; only
mismatches between the two sides count as failures.     python tools/soak_emu.py [iterations]
"""
import ctypes as C
import os
import subprocess
import sys
import tempfile
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import numpy as np  # noqa: E402

import test_log_find as tl  # noqa: E402
import test_pause_batch as tb  # noqa: E402
import test_phase1b as tp  # noqa: E402
from helpers import oracle_library  # noqa: E402

SANITY = ("verdicts ==", "0 < want_ok.sum()", ".any()", "found >", "int(ctl[0]) > 5", "assert 0 < ")


def build(src):
    out = os.path.join(tempfile.mkdtemp(), "lib.so")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-w", "-I", "/usr/local/cuda/include",
                           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "gigapaxos_b200", "csrc"),
                           "-x", "c++", os.path.join(ROOT, "tests", "emu", src), "-o", out])
    return C.CDLL(out)


def attempt(fn, *args):
    try:
        fn(*args)
        return 1
    except AssertionError:
        lines = traceback.format_exc().splitlines()
        if any(k in " ".join(lines[-4:]) for k in SANITY):
            return 0
        print(args, "\n".join(lines[-25:]))
        raise


def main(iters):
    lib = oracle_library()
    e1, e2, e3 = build("p1b_emu.cpp"), build("pause_emu.cpp"), build("logfind_emu.cpp")
    e3.emu_log_find.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p,
                                C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
    rng = np.random.default_rng(int(time.time()))
    ok = 0
    t0 = time.time()
    for _ in range(iters):
        seed = int(rng.integers(100, 60000))
        R = int(rng.choice([1, 2, 3, 3, 4, 5]))
        wrap = bool(rng.integers(0, 2))
        ln = None
        if R >= 3 and rng.integers(0, 3) == 0:
            ln = [tp.NODES5[i] for i in rng.choice(R, size=int(rng.integers(1, R)), replace=False)]
        ok += attempt(tp.test_kernel_source_on_the_host_equals_oracle, lib, e1, R, seed, wrap, int(rng.choice([1, 7, 33, 64])), ln)
        ok += attempt(tp.test_oracle_phase1b_equals_the_host_twin, lib, R, seed, wrap, ln)
        ok += attempt(tb.test_kernel_source_on_the_host_equals_oracle, lib, e2, int(rng.integers(0, 2)), int(rng.choice([2, 3, 5])),
                      seed, int(rng.choice([1, 33, 128])))
        ok += attempt(tl.test_kernel_source_on_the_host_equals_oracle, lib, e3, seed, int(rng.choice([1, 3, 7])),
                      int(rng.choice([1, 64, 256])))
        ok += attempt(tl.test_kernel_source_on_a_ring_that_wraps, lib, e3, seed)
    print(f"{iters} iterations x 5 checks: {ok} complete, {5 * iters - ok} ended at a sanity assertion, 0 mismatches; "
          f"{time.time() - t0:.0f} s")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 50)
