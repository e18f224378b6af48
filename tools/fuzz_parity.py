#!/usr/bin/env python
"""Randomised parity fuzz: random shapes, kinds, window types, error rates and score parameters,
every window compared with the CPU oracle.  python tools/fuzz_parity.py [seconds] [seed]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hypo_b200 import native
from hypo_b200.batch import concat_batches
from hypo_b200.hostlib import synth_batch
from tests.oracle_util import oracle_consensus

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 12345)
t_end = time.time() + budget
total = bad_total = rounds = 0
while time.time() < t_end:
    rounds += 1
    wtype = int(rng.random() < 0.2)
    length = int(rng.choice([rng.integers(1, 20), rng.integers(20, 126), rng.integers(126, 260), rng.integers(260, 520)],
                            p=[0.25, 0.45, 0.2, 0.1]))
    n_arms = int(rng.choice([rng.integers(2, 12), rng.integers(12, 45), rng.integers(45, 130)], p=[0.3, 0.55, 0.15]))
    kind = str(rng.choice(["internal", "backbone", "prefix", "suffix", "mixed"]))
    err = float(rng.choice([0.0, 0.01, 0.03, 0.08, 0.15]))
    u = rng.random()
    if u < 0.55:
        scores = (5, -4, -8, 3, -5, -4)
    elif u < 0.62:   # beyond the 16-bit DP range: the 32-bit fill of the last tier
        scores = (int(rng.integers(40, 128)), -int(rng.integers(40, 129)), -int(rng.integers(40, 129)),
                  int(rng.integers(40, 128)), -int(rng.integers(40, 129)), -int(rng.integers(40, 129)))
    else:
        scores = (int(rng.integers(1, 9)), -int(rng.integers(1, 9)), -int(rng.integers(0, 10)),
                  int(rng.integers(1, 6)), -int(rng.integers(1, 8)), -int(rng.integers(0, 8)))
    cells = n_arms * length * length * (1 + err * n_arms) * (2 if wtype else 1)
    n_win = int(max(8, min(4000, 6e8 / max(cells, 1))))
    b = synth_batch(int(rng.integers(1, 1 << 30)), n_win, length, n_arms, kind, err, wtype=wtype)
    if rng.random() < 0.6:
        # ragged: a few more shapes of a similar size class, shuffled together (the windows of a warp of the
        # group tiers then differ in everything; tier lists get re-ordered by size)
        parts = [b]
        for _ in range(int(rng.integers(1, 4))):
            l2 = int(max(1, min(519, length * float(rng.uniform(0.3, 1.6)))))
            a2 = int(max(2, min(129, n_arms * float(rng.uniform(0.3, 1.8)))))
            k2 = str(rng.choice(["internal", "backbone", "prefix", "suffix", "mixed"]))
            e2 = float(rng.choice([0.0, 0.01, 0.03, 0.08]))
            c2 = a2 * l2 * l2 * (1 + e2 * a2) * (2 if wtype else 1)
            parts.append(synth_batch(int(rng.integers(1, 1 << 30)), int(max(4, min(2000, 3e8 / max(c2, 1)))), l2, a2, k2, e2, wtype=wtype))
        b = concat_batches(parts, {})
        b = b.select(rng.permutation(b.n_win))
    native.init(scores, 0)
    native.set_option("first_tier", int(rng.choice([0, 0, 0, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 10])))   # every tier gets its share
    native.set_option("group_tiers", int(rng.choice([0, 2, 2, 2])))   # (the default, 1, keeps small batches out of the group tiers)
    native.set_option("group_sort", int(rng.random() < 0.7))
    native.set_option("teams", int(rng.random() < 0.8))
    try:
        got = native.consensus(b)
    except native.HypoGpuError as e:
        print(f"round {rounds}: GPU error {e} for len={length} arms={n_arms} kind={kind} err={err} wtype={wtype} scores={scores}", flush=True)
        bad_total += 1
        continue
    want, _ = oracle_consensus(b, scores)
    bad = sum(a != c for a, c in zip(got, want))
    total += b.n_win
    bad_total += bad
    if bad:
        i = next(k for k in range(b.n_win) if got[k] != want[k])
        print(f"round {rounds}: {bad}/{b.n_win} MISMATCH len={length} arms={n_arms} kind={kind} err={err} wtype={wtype} "
              f"scores={scores} first={i}\n  spec={b.spec(i)}\n  gpu ={got[i]}\n  want={want[i]}", flush=True)
print(f"fuzz: {rounds} rounds, {total} windows, {bad_total} problems")
sys.exit(1 if bad_total else 0)
