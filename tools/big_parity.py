#!/usr/bin/env python
"""One-off large parity run: every window of several big synthetic batches against the CPU oracle."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hypo_b200 import native
from hypo_b200.hostlib import synth_batch
from tests.oracle_util import DEFAULT_SCORES, oracle_consensus

native.init(DEFAULT_SCORES, 0)
total_bad = 0
for seed, kw in ((101, dict(n_win=120000, length=120, n_arms=30, kind="internal", err=0.01)),
                 (102, dict(n_win=60000, length=100, n_arms=30, kind="mixed", err=0.02)),
                 (103, dict(n_win=60000, length=60, n_arms=20, kind="backbone", err=0.03)),
                 (104, dict(n_win=80000, length=10, n_arms=35, kind="mixed", err=0.02)),
                 (105, dict(n_win=30000, length=110, n_arms=30, kind="prefix", err=0.05)),
                 (106, dict(n_win=30000, length=110, n_arms=30, kind="suffix", err=0.05)),
                 (107, dict(n_win=8000, length=200, n_arms=12, kind="mixed", err=0.02, wtype=1))):
    b = synth_batch(seed, **kw)
    t0 = time.time(); got = native.consensus(b); t1 = time.time()
    want, _ = oracle_consensus(b); t2 = time.time()
    bad = sum(a != c for a, c in zip(got, want))
    total_bad += bad
    print(f"{kw}: {b.n_win} windows, {bad} mismatches, gpu {t1-t0:.2f}s oracle {t2-t1:.1f}s, tiers {native.last_timing()[2][:5]}, "
          f"abandoned {native.last_fail_hist()[1:8]}", flush=True)
print("TOTAL MISMATCHES", total_bad)
sys.exit(1 if total_bad else 0)
