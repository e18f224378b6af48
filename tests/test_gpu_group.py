"""GPU tests of the group tiers (hypo_b200/csrc/poa_group.cu): several small SHORT windows per warp in
lock-step - 8 lanes per window (Tq, <= 31 symbols, 4 windows per warp) and 16 lanes per window (Th, <= 63
symbols, 2 per warp).  Everything goes through the C ABI and is compared byte for byte with the CPU oracle;
the tier histogram shows the windows really ran there.  Windows of one warp differ in every respect (reads,
lengths, kinds, error rates, trivial outcomes), so the lock-step predicates are exercised in every phase."""
import numpy as np
import pytest

from hypo_b200 import native
from hypo_b200.batch import WINDOW_LONG, WindowSpec, build_batch, concat_batches
from hypo_b200.hostlib import synth_batch
from hypo_b200.synth import edge_case_windows, random_window
from tests.oracle_util import DEFAULT_SCORES, oracle_consensus

pytestmark = pytest.mark.gpu

TQ, TH = native.TIER_QUAD, native.TIER_HALF
KINDS = ("internal", "backbone", "prefix", "suffix", "mixed")


@pytest.fixture(autouse=True)
def _init():
    native.init(DEFAULT_SCORES, 0)
    native.set_option("first_tier", 0)
    native.set_option("group_tiers", 2)   # (the default, 1, only uses them for batches of >= 131072 windows)
    yield
    native.set_option("first_tier", 0)
    native.set_option("group_tiers", 1)
    native.init(DEFAULT_SCORES, 0)


def _same(got, want, batch, label):
    bad = [i for i, (a, b) in enumerate(zip(got, want)) if a != b]
    assert not bad, f"{label}: {len(bad)}/{len(want)} windows differ; first={bad[0]} spec={batch.spec(bad[0])} got={got[bad[0]]!r} want={want[bad[0]]!r}"


def _ragged(seed, n, lo, hi, max_arms=48, errs=(0.0, 0.01, 0.03, 0.08)):
    """Windows that differ in everything, shuffled: neighbours in the tier's list share a warp."""
    rng = np.random.default_rng(seed)
    specs = []
    for _ in range(n):
        specs.append(random_window(rng, length=int(rng.integers(lo, hi + 1)), n_arms=int(rng.integers(2, max_arms + 1)),
                                   kind=KINDS[int(rng.integers(0, 5))], err=float(errs[int(rng.integers(0, len(errs)))]),
                                   draft_n=0.05 if rng.random() < 0.1 else 0.0))
    return build_batch(specs)


def test_quad_tier_ragged_windows():
    b = _ragged(201, 3000, 1, 24)
    want, _ = oracle_consensus(b)
    _same(native.consensus(b), want, b, "Tq")
    _, _, tiers = native.last_timing()
    assert tiers[TQ] > 2000, tiers


def test_half_tier_ragged_windows():
    b = _ragged(202, 2000, 26, 52, max_arms=40)
    want, _ = oracle_consensus(b)
    _same(native.consensus(b), want, b, "Th")
    _, _, tiers = native.last_timing()
    assert tiers[TH] > 1200, tiers


def test_half_tier_forced_takes_the_small_windows_too():
    native.set_option("first_tier", TH)
    b = _ragged(203, 1500, 1, 50, max_arms=30)
    want, _ = oracle_consensus(b)
    _same(native.consensus(b), want, b, "Th forced")
    _, _, tiers = native.last_timing()
    assert tiers[TQ] == 0 and tiers[TH] > 1000, tiers


def test_edge_cases_and_trivial_outcomes_inside_a_warp():
    """Every branch of Window::generate_consensus (reference src/Window.cpp:44-61,87-154) next to ordinary
    windows in the same warp: no arms, one arm, empties dominating, only zero-length arms, single base."""
    rng = np.random.default_rng(204)
    specs = []
    for rep in range(40):
        for e in edge_case_windows():
            specs.append(e)
            specs.append(random_window(rng, length=int(rng.integers(2, 28)), n_arms=int(rng.integers(2, 20)),
                                       kind=KINDS[rep % 5], err=0.03))
    # zero-length arms in every position of the reference's order
    specs += [WindowSpec("ACGTACGTAC", ["", "ACGTACGTAC", "", "ACGAACGTAC", ""], ["", "ACGTA", ""], ["", "CGTAC", ""], 0, 0),
              WindowSpec("ACGTACGTAC", ["", ""], ["ACGTA", "", "ACGTAC"], ["", "GTAC"], 0, 0),
              WindowSpec("ACGTACGTAC", [], ["", "ACGTA", "ACG"], ["", ""], 0, 0)] * 8
    b = build_batch(specs)
    want, _ = oracle_consensus(b)
    _same(native.consensus(b), want, b, "edge")
    _, _, tiers = native.last_timing()
    assert tiers[TQ] > 0, tiers


def test_windows_that_outgrow_a_group_tier_are_rerun_in_the_one_warp_tiers():
    """Noisy, deep windows overflow the group tiers' node / edge / aligned-list capacities at run time; their
    neighbours in the warp must not notice."""
    rng = np.random.default_rng(205)
    specs = []
    for i in range(1200):
        noisy = i % 3 == 0
        specs.append(random_window(rng, length=int(rng.integers(18, 29)) if noisy else int(rng.integers(3, 29)),
                                   n_arms=int(rng.integers(30, 60)) if noisy else int(rng.integers(2, 30)),
                                   kind=KINDS[i % 5], err=0.15 if noisy else 0.02))
    b = build_batch(specs)
    want, _ = oracle_consensus(b)
    _same(native.consensus(b), want, b, "overflow")
    _, _, tiers = native.last_timing()
    reasons = native.last_fail_hist()
    assert tiers[TQ] > 0 and tiers[0] > 0 and sum(reasons[3:7]) > 0, (tiers, reasons)


def test_group_tiers_leave_every_byte_as_the_one_warp_tiers_produce_it():
    """Too many windows for the oracle: the same batch with and without the group tiers, byte for byte; a strided
    sample against the oracle."""
    parts = [synth_batch(300 + i, 6000, ln, na, kind, err)
             for i, (ln, na, kind, err) in enumerate(((6, 11, "internal", 0.01), (9, 39, "mixed", 0.01), (14, 30, "internal", 0.03),
                                                      (25, 48, "mixed", 0.01), (40, 16, "internal", 0.01), (58, 30, "mixed", 0.02),
                                                      (9, 16, "backbone", 0.05), (20, 24, "prefix", 0.03), (30, 24, "suffix", 0.03)))]
    b = concat_batches(parts, {})
    b = b.select(np.random.default_rng(7).permutation(b.n_win))
    with_groups = native.consensus(b)
    _, _, tiers = native.last_timing()
    assert tiers[TQ] > 20000 and tiers[TH] > 5000, tiers
    native.set_option("group_tiers", 0)
    without = native.consensus(b)
    _, _, tiers0 = native.last_timing()
    assert tiers0[TQ] == 0 and tiers0[TH] == 0, tiers0
    _same(with_groups, without, b, "group vs one-warp tiers")
    sample = b.select(np.arange(0, b.n_win, 37))
    want, _ = oracle_consensus(sample)
    _same([with_groups[i] for i in range(0, b.n_win, 37)], want, sample, "sample vs oracle")


def test_alternative_and_extreme_scores_in_the_group_tiers():
    for sc in ((1, -1, -1, 1, -1, -1), (2, -6, -3, 3, -5, -4), (10, -9, -12, 3, -5, -4)):
        native.init(sc, 0)
        b = _ragged(206, 800, 1, 50, max_arms=24)
        want, _ = oracle_consensus(b, sc)
        _same(native.consensus(b), want, b, f"scores {sc}")
    # scores x size beyond int16: the group tiers hand every window on (reason 2), the last tier computes them
    sc = (127, -128, -128, 127, -128, -128)
    native.init(sc, 0)
    b = _ragged(207, 300, 1, 40, max_arms=16)
    want, _ = oracle_consensus(b, sc)
    _same(native.consensus(b), want, b, "int32")


def test_long_windows_never_start_in_a_group_tier():
    b = synth_batch(208, 64, 40, 12, "internal", 0.02, wtype=WINDOW_LONG)
    want, _ = oracle_consensus(b)
    _same(native.consensus(b), want, b, "LONG")
    _, _, tiers = native.last_timing()
    assert tiers[TQ] == 0 and tiers[TH] == 0, tiers
