#!/usr/bin/env python
"""Developer aid (run under gpurun): LONG windows whose DAG outgrows every shared-memory tier, i.e. the
bound-driven global-memory tier with LONG windows, against the CPU oracle; prints the tier histogram,
the abandonment reasons and the kernel time."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hypo_b200 import native  # noqa: E402
from hypo_b200.hostlib import synth_batch  # noqa: E402
from tests.oracle_util import oracle_consensus  # noqa: E402

native.init((5, -4, -8, 3, -5, -4), 0)
bad = 0
for seed, kw in ((71, dict(n_win=48, length=300, n_arms=30, kind="internal", err=0.06, wtype=1)),
                 (72, dict(n_win=24, length=500, n_arms=20, kind="mixed", err=0.05, wtype=1)),
                 (73, dict(n_win=3000, length=500, n_arms=30, kind="internal", err=0.05, wtype=1))):
    b = synth_batch(seed, **kw)
    t0 = time.time()
    got = native.consensus(b)
    dt = time.time() - t0
    k = min(b.n_win, 48 if b.n_win < 100 else 16)
    import numpy as np
    want, _ = oracle_consensus(b.select(np.arange(k)))
    n = sum(x != y for x, y in zip(got[:k], want))
    bad += n
    ms, _, tiers = native.last_timing()
    print(kw, "mismatches", n, "of", k, "tiers", tiers, "reasons", native.last_fail_hist()[1:12],
          "kernel ms", round(ms, 1), "Mbp/s", round(b.polished_bp / 1e3 / ms, 2), "wall", round(dt, 2), flush=True)
print("long_t2_check:", "OK" if bad == 0 else f"{bad} MISMATCHES")
