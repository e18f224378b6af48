"""Small mixed batch through the C ABI (run under compute-sanitizer by tools/sanitize.sh)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hypo_b200 import native
from hypo_b200.batch import build_batch, WINDOW_LONG
from hypo_b200.synth import edge_case_windows, random_window
from tests.oracle_util import oracle_consensus, DEFAULT_SCORES

# HYPO_SANITIZE_NO_TEAMS=1: keep away from the team tiers (T1m / T1 always run several warps per window and
# synchronise them through flags in shared memory - polling that racecheck reports as a hazard by design); used
# for the racecheck pass, memcheck runs everything.
NO_TEAMS = os.environ.get("HYPO_SANITIZE_NO_TEAMS") == "1"
rng = np.random.default_rng(3)
specs = edge_case_windows()
for kind in ("internal", "backbone", "prefix", "suffix", "mixed"):
    specs += [random_window(rng, length=int(rng.integers(5, 100 if NO_TEAMS else 125)), n_arms=int(rng.integers(3, 34)), kind=kind,
                            err=float(rng.choice([0.01, 0.05]))) for _ in range(6)]
specs += [random_window(rng, length=120, n_arms=30, kind="internal") for _ in range(8)]
specs += [random_window(rng, length=120, n_arms=60, kind="internal", err=0.1) for _ in range(2)]      # overflow tiers
wide_specs = [random_window(rng, length=200, n_arms=10, kind="mixed") for _ in range(3)]                   # two tiles
wide_specs += [random_window(rng, length=int(rng.integers(100, 300)), n_arms=8, kind="internal", wtype=WINDOW_LONG) for _ in range(3)]
if not NO_TEAMS:
    specs += wide_specs
# small windows of every kind: the group tiers (several windows per warp)
for i in range(60):
    specs.append(random_window(rng, length=int(rng.integers(1, 56)), n_arms=int(rng.integers(2, 40)),
                               kind=("internal", "backbone", "prefix", "suffix", "mixed")[i % 5], err=float(rng.choice([0.0, 0.02, 0.08]))))
b = build_batch(specs)
native.init(DEFAULT_SCORES, 0)
native.set_option("group_tiers", 2)   # always (by default only batches of >= 131072 windows use the group tiers)
if NO_TEAMS:
    native.set_option("teams", 0)
got = native.consensus(b)
want, _ = oracle_consensus(b)
bad = sum(a != c for a, c in zip(got, want))
print(f"sanitize_run: {b.n_win} windows, {bad} mismatches, tiers {native.last_timing()[2]}")
if NO_TEAMS:
    # windows of more than one tile: straight into the bound-driven tiers, one warp per window
    native.set_option("first_tier", 6)
    wb = build_batch(wide_specs)
    got = native.consensus(wb)
    want, _ = oracle_consensus(wb)
    bad += sum(a != c for a, c in zip(got, want))
    print(f"sanitize_run: {wb.n_win} multi-tile windows, tiers {native.last_timing()[2]}")
    native.set_option("first_tier", 0)
# the 32-bit fill of the last tier
sc = (127, -128, -128, 127, -128, -128)
native.init(sc, 0)
if NO_TEAMS:
    native.set_option("teams", 0)
    native.set_option("first_tier", 6)
small = build_batch(specs[:40])
got = native.consensus(small)
want, _ = oracle_consensus(small, sc)
bad += sum(a != c for a, c in zip(got, want))
print(f"sanitize_run: 32-bit tier, {small.n_win} windows, tiers {native.last_timing()[2]}")
# arm extraction, the fused call and support counting on a slice of the captured run
native.init(DEFAULT_SCORES, 0)
native.set_option("first_tier", 0)
from tests.arms_util import device_inputs, load_capture
regions, clen, dumped, recs = load_capture()
args = device_inputs(regions, clen, recs[:1500])
batch, win_region = native.extract_arms(*args, 9)
out = native.polish_alignments(*args, 9)
print(f"sanitize_run: extract_arms {batch.n_win} windows {batch.n_arms} arms, fused contig {len(out[0])} bp")
import gzip
from oracle import support_oracle as so
spos, kid, _, _ = so.read_kmer_dump(gzip.open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "cli_short_60kb.kmer_support.gz"), "rt"))
c, s_ = native.solid_kmer_support(np.array([0, len(spos)], np.uint64), np.array(spos, np.uint32), np.array(kid, np.uint64), args[3], args[4], args[5], 9)
print(f"sanitize_run: solid k-mer support, coverage sum {int(c.sum())} support sum {int(s_.sum())}")
sys.exit(1 if bad else 0)
