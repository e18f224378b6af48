#!/usr/bin/env python
"""Developer aid (CPU, this container only: needs oracle/_ref): the C restatement against the compiled
unmodified reference on the mid-size and large window shapes the multi-tile and global-memory tiers run
(250-500 bp, 5-8 % read error, 100-200 reads, LONG windows with the reference's arm filter applied)."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hypo_b200.hostlib import synth_batch
from hypo_b200.batch import WINDOW_LONG
from tests.oracle_util import oracle_consensus, ref_consensus, drop_rejected_arms
tot = bad = 0
t0 = time.time()
seed = 7000
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    for kw in (dict(n_win=400, length=250, n_arms=30, kind="internal", err=0.01),
               dict(n_win=300, length=250, n_arms=30, kind="mixed", err=0.05),
               dict(n_win=150, length=500, n_arms=30, kind="internal", err=0.03),
               dict(n_win=120, length=500, n_arms=12, kind="prefix", err=0.05),
               dict(n_win=120, length=400, n_arms=20, kind="internal", err=0.04, wtype=WINDOW_LONG),
               dict(n_win=100, length=250, n_arms=30, kind="internal", err=0.08, wtype=WINDOW_LONG),
               dict(n_win=80, length=120, n_arms=100, kind="mixed", err=0.05),
               dict(n_win=60, length=50, n_arms=200, kind="suffix", err=0.03)):
        seed += 1
        b = synth_batch(seed, **kw)
        r, acc, _ = ref_consensus(b)
        if not acc.all():
            b = drop_rejected_arms(b, acc)
            r, acc, _ = ref_consensus(b)
        a, _ = oracle_consensus(b)
        n = sum(x != y for x, y in zip(a, r))
        tot += b.n_win; bad += n
        if n: print("MISMATCH", seed, kw, n, flush=True)
    print(rep, tot, bad, round(time.time() - t0, 1), flush=True)
print("done", tot, bad)
