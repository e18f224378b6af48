"""GPU tests of the capacity tiers' hand-over paths and of the limits of the C ABI: the bound-driven
tiers forced, the 32-bit DP of the last tier (scores x size beyond int16 are computed, never refused),
DFS-stack overflow, the remaining HYPO_E_CAPACITY corner, malformed descriptors, and the regression
cases of the round-1 review (LONG windows full of zero-length arms, the mis-aligned error counter).
Everything goes through the C ABI and is compared byte for byte with the CPU oracle."""
import numpy as np
import pytest

from hypo_b200 import native
from hypo_b200.batch import WINDOW_LONG, WindowSpec, build_batch
from hypo_b200.hostlib import synth_batch
from hypo_b200.synth import edge_case_windows, random_batch
from tests.oracle_util import DEFAULT_SCORES, oracle_consensus

pytestmark = pytest.mark.gpu

FAIL_RANGE, FAIL_STACK = 2, 7


@pytest.fixture(autouse=True)
def _init():
    native.init(DEFAULT_SCORES, 0)
    native.set_option("first_tier", 0)
    native.set_option("scap", 0)
    yield
    native.set_option("first_tier", 0)
    native.set_option("scap", 0)
    native.init(DEFAULT_SCORES, 0)


def _same(got, want, batch, label):
    bad = [i for i, (a, b) in enumerate(zip(got, want)) if a != b]
    assert not bad, f"{label}: {len(bad)}/{len(want)} windows differ; first={bad[0]} spec={batch.spec(bad[0])}"


def _mixed_batches():
    yield "edge", build_batch(edge_case_windows())
    yield "short-mixed", synth_batch(101, 400, 110, 30, "mixed", 0.03)
    yield "short-prefix", random_batch(102, 64, kind="prefix", length=60, n_arms=20, err=0.05)
    yield "tiny", synth_batch(103, 2000, 9, 25, "mixed", 0.03)
    yield "long", synth_batch(104, 48, 300, 16, "internal", 0.03, wtype=WINDOW_LONG)
    yield "wide-short", synth_batch(105, 32, 400, 12, "mixed", 0.02)


@pytest.mark.parametrize("tier", [6, 7])
def test_bound_driven_tiers_forced(tier):
    """Every kind of window, routed straight into T2 / T3 (DAG in global memory, capacities from the
    windows' bounds): same bytes as the oracle, and the tier histogram shows they really ran there."""
    native.set_option("first_tier", tier)
    for label, b in _mixed_batches():
        want, _ = oracle_consensus(b)
        _same(native.consensus(b), want, b, f"T{tier - 4}/{label}")
        _, _, tiers = native.last_timing()
        assert sum(tiers[:tier]) == 0 and tiers[tier] > 0, (label, tiers)


@pytest.mark.parametrize("tier", [1, 2, 3, 4, 5])
def test_shared_memory_tiers_forced(tier):
    """The larger shared-memory tiers with windows that would normally run in the compact tier."""
    native.set_option("first_tier", tier)
    for label, b in _mixed_batches():
        want, _ = oracle_consensus(b)
        _same(native.consensus(b), want, b, f"tier{tier}/{label}")
        _, _, tiers = native.last_timing()
        assert sum(tiers[:tier]) == 0


def test_scores_beyond_int16_are_computed_with_32_bit_cells():
    """(127, -128, -128): S * (rows + columns) leaves the 16-bit range for almost every read.  The
    reference computes in int32 (sisd_alignment_engine.cpp:263-342) and never refuses a window; neither
    does the last tier.  Every 16-bit tier hands the windows on with reason kFailRange."""
    sc = (127, -128, -128, 127, -128, -128)
    native.init(sc, 0)
    for label, b in (("edge", build_batch(edge_case_windows())),
                     ("short", synth_batch(111, 300, 100, 24, "mixed", 0.03)),
                     ("backbone", random_batch(112, 48, kind="backbone", length=80, n_arms=12, err=0.05)),
                     ("long", synth_batch(113, 24, 260, 12, "internal", 0.03, wtype=WINDOW_LONG))):
        want, _ = oracle_consensus(b, sc)
        _same(native.consensus(b), want, b, f"int32/{label}")
        _, _, tiers = native.last_timing()
        reasons = native.last_fail_hist()
        if label != "edge":
            assert reasons[FAIL_RANGE] > 0 and tiers[7] > 0, (label, tiers, reasons)
    # asymmetric extremes as well
    for sc in ((127, -1, -128, 1, -128, -1), (1, -128, -1, 127, -128, -128), (0, 0, 0, 0, 0, 0), (127, 127, 0, 5, -4, 0)):
        native.init(sc, 0)
        b = synth_batch(114, 96, 70, 14, "mixed", 0.04)
        want, _ = oracle_consensus(b, sc)
        _same(native.consensus(b), want, b, f"int32/{sc}")


def test_large_noisy_windows_mix_16_and_32_bit_reads():
    """200 reads x 500 bp at 12 % error per edit type (the largest shape of BASELINE.json configs[4]).
    With the default scores the DAG (~2 650 nodes) still fits 16 bits; with scores of magnitude 12 the
    guard trips once the DAG has grown, so early reads of a window are filled with 16-bit cells and
    late ones with 32-bit cells - in the same window."""
    b = synth_batch(61, 2, 500, 200, "internal", 0.12)
    want, _ = oracle_consensus(b)
    _same(native.consensus(b), want, b, "200x500@12%")
    sc = (9, -9, -12, 9, -9, -12)
    native.init(sc, 0)
    want, _ = oracle_consensus(b, sc)
    _same(native.consensus(b), want, b, "200x500@12%, |score| 12")
    _, _, tiers = native.last_timing()
    assert native.last_fail_hist()[FAIL_RANGE] > 0 and tiers[7] == b.n_win
    bl = synth_batch(62, 2, 500, 60, "internal", 0.08, wtype=WINDOW_LONG)
    want, _ = oracle_consensus(bl, sc)
    _same(native.consensus(bl), want, bl, "LONG 60x500@8%, |score| 12")


def test_dfs_stack_overflow_moves_the_window_on():
    """A DFS stack too small for the exact topological sort (forced: 8 entries in T2) abandons the
    window with reason kFailStack; the last tier, whose stack is sized from the bounds, finishes it."""
    native.set_option("first_tier", 6)
    native.set_option("scap", 8)
    b = synth_batch(121, 40, 200, 14, "internal", 0.04, wtype=WINDOW_LONG)   # LONG: always sorts exactly
    want, _ = oracle_consensus(b)
    _same(native.consensus(b), want, b, "kFailStack")
    _, _, tiers = native.last_timing()
    reasons = native.last_fail_hist()
    assert reasons[FAIL_STACK] > 0 and tiers[7] == reasons[FAIL_STACK] and tiers[6] == b.n_win


def test_capacity_error_is_loud_and_leaves_the_library_usable():
    """What the device really cannot hold (here: more than 32 000 reads in one window, the 16-bit edge
    weights) is reported as HYPO_E_CAPACITY for the batch - no silent wrong answer, no CPU fallback -
    and the next batch runs normally."""
    many = WindowSpec("ACGT", ["A"] * 33000, [], [], 0, 0)
    b = build_batch([many, WindowSpec("ACGTACGT", ["ACGTACGT", "ACGAACGT", "ACGTACGT"], [], [], 0, 0)])
    with pytest.raises(native.HypoGpuError) as e:
        native.consensus(b)
    assert e.value.code == 6
    ok = random_batch(122, 16)
    want, _ = oracle_consensus(ok)
    _same(native.consensus(ok), want, ok, "after capacity error")


def test_long_window_with_hundreds_of_zero_length_arms():
    """Round-1 review: the LONG path slot was sized by the non-empty arms but indexed by all arms.
    Zero-length arms are legal (reference src/Window.cpp:182 skips them) and take no slot now."""
    rng = np.random.default_rng(7)
    specs = []
    for k in range(24):
        truth = "".join("ACGT"[i] for i in rng.integers(0, 4, size=140))
        arms = []
        for r in range(10):
            arms += [""] * 30
            arms.append(truth[:70] + ("A" if r % 3 == 0 else "") + truth[70:])
        specs.append(WindowSpec(truth, arms[: 200 + k], arms[200 + k:], [""] * 40, 0, WINDOW_LONG))
    b = build_batch(specs)
    want, _ = oracle_consensus(b)
    _same(native.consensus(b), want, b, "LONG with zero-length arms")
    native.set_option("first_tier", 7)
    _same(native.consensus(b), want, b, "LONG with zero-length arms, last tier")


@pytest.mark.parametrize("n_win", [1, 3, 7, 64, 1001])
def test_malformed_descriptors_return_arg_error(n_win):
    """Round-1 review: the error counter sat at an address that was 4-byte aligned only for some batch
    sizes, so malformed input usually killed the context instead of returning HYPO_E_ARG."""
    good = random_batch(130 + n_win, n_win, length=30, n_arms=6)
    want, _ = oracle_consensus(good)
    out, off = np.empty(int(good.out_bound().sum()) + 16, np.uint8), np.zeros(n_win + 1, np.uint64)
    for field, value in (("first_arm", 1 << 40), ("draft_off", 1 << 50), ("draft_len", 0xFFFFFFFF), ("wtype", 9)):
        b = random_batch(130 + n_win, n_win, length=30, n_arms=6)
        b.win[field][n_win // 2] = value
        with pytest.raises(native.HypoGpuError) as e:
            native.consensus_batch_host(b, out, off)
        assert e.value.code == 3, field
    for field, value in (("off", (1 << 64) - 2), ("len", 0xFFFFFFFE), ("len", 0x80000000), ("reserved", 1)):
        b = random_batch(130 + n_win, n_win, length=30, n_arms=6)
        b.arms[field][b.n_arms // 2] = value
        with pytest.raises(native.HypoGpuError) as e:
            native.consensus_batch_host(b, out, off)
        assert e.value.code == 3, field
    _same(native.consensus(good), want, good, "after malformed input")


def test_reinit_on_the_same_device_keeps_working_and_changes_scores():
    b = synth_batch(140, 200, 60, 12, "mixed", 0.08)
    for sc in ((2, -3, -2, 1, -1, -1), DEFAULT_SCORES, (1, -1, -1, 1, -1, -1)):
        native.init(sc, 0)
        want, _ = oracle_consensus(b, sc)
        _same(native.consensus(b), want, b, f"re-init {sc}")
    native.shutdown()
    with pytest.raises(native.HypoGpuError):
        native.consensus(b)
    native.init(DEFAULT_SCORES, 0)


def test_window_bounds_hold_on_the_device_path():
    """hypo_gpu_window_bounds is the slot size the device entry point needs: run LONG windows whose
    round-1 consensus is much longer than the draft (arms unlike the draft) with slots of exactly that size
    and a guard pattern between them."""
    import torch
    rng = np.random.default_rng(3)
    specs = []
    for _ in range(40):
        a = "".join("ACGT"[i] for i in rng.integers(0, 4, size=int(rng.integers(60, 200))))
        specs.append(WindowSpec("ACGTAC", [a, a, a[:-1], a + "C"], [], [], 0, WINDOW_LONG))
    b = build_batch(specs)
    want, _ = oracle_consensus(b)
    bound = native.window_bounds(b).astype(np.int64)
    assert (bound == b.out_bound()).all()
    gap = 64
    pos = (np.concatenate([[0], np.cumsum(bound + gap)[:-1]]) + gap).astype(np.uint64)
    dev = torch.device("cuda", 0)

    def to_dev(x):
        return torch.from_numpy(x.view(np.uint8).reshape(-1)).to(dev)

    d_win, d_arms, d_packed, d_pos = to_dev(b.win), to_dev(b.arms), to_dev(b.packed), to_dev(pos)
    d_scr = torch.full((int((bound + gap).sum()) + gap,), 0x5A, dtype=torch.uint8, device=dev)
    d_len = torch.zeros(b.n_win, dtype=torch.int32, device=dev)
    native.consensus_batch_device(d_win.data_ptr(), b.n_win, d_arms.data_ptr(), b.n_arms, d_packed.data_ptr(),
                                  b.packed.size, d_scr.data_ptr(), d_pos.data_ptr(), d_len.data_ptr(), 0)
    torch.cuda.synchronize()
    scr, ln = d_scr.cpu().numpy(), d_len.cpu().numpy()
    for w in range(b.n_win):
        p = int(pos[w])
        assert scr[p:p + int(ln[w])].tobytes().decode() == want[w]
        assert (scr[p + int(bound[w]):p + int(bound[w]) + gap] == 0x5A).all(), "a window wrote beyond its bound"


def test_small_batches_need_few_host_round_trips():
    """The tier lists live on the device (a tier appends what it abandons to its successor's list), so a
    batch that stays in the shared-memory tiers costs a fixed, small number of launches however many
    tiers its windows visit."""
    b = synth_batch(150, 64, 120, 30, "internal", 0.05)   # some windows overflow Tc and T0
    want, _ = oracle_consensus(b)
    l0 = native.launch_count()
    _same(native.consensus(b), want, b, "small batch")
    assert native.launch_count() - l0 <= 12   # classify, 2 scans, widen, gather, <= 7 tier launches


@pytest.mark.parametrize("gather", [0, 2])
def test_one_process_many_devices_same_bytes(gather):
    """hypo_gpu_init_multi: one host process, G devices, contiguous cost-balanced window ranges; the
    result is byte-identical for every G (and for a batch that is not laid out in window order)."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    b = synth_batch(160, 6000, 100, 20, "mixed", 0.03)
    perm = np.random.default_rng(1).permutation(b.n_win)
    want, _ = oracle_consensus(b)
    for g in sorted({2, n}):
        native.init_multi(DEFAULT_SCORES, g)
        assert native.device_count() == g
        native.set_option("gather", gather)
        _same(native.consensus(b), want, b, f"{g} devices")
        got = native.consensus(b.select(perm))
        assert got == [want[i] for i in perm]
    native.set_option("gather", 0)


def test_probe_reroutes_a_list_its_tier_cannot_hold():
    """30 x 100 bp at 5 % error per edit type: the static estimate routes the windows to the compact tier,
    whose DAG capacity almost none of them fits.  The launcher runs the first 4096 alone, sees them leave, and
    hands the rest to the successor untried (which probes again).  Same bytes as without probing."""
    b = synth_batch(170, 24000, 100, 30, "internal", 0.05)
    native.set_option("probe", 0)
    plain = native.consensus(b)
    _, _, tiers0 = native.last_timing()
    native.set_option("probe", 1)
    probed = native.consensus(b)
    _, _, tiers1 = native.last_timing()
    assert probed == plain
    assert native.last_rerouted() > 0 and tiers1[0] == 4096 and tiers0[0] == b.n_win
    idx = np.arange(0, b.n_win, 97)
    want, _ = oracle_consensus(b.select(idx))
    assert [probed[i] for i in idx] == want


def test_estimate_driven_shared_memory_tier_forced():
    """T2s (tier 10): the DAG of a large window in shared memory, capacities from an estimate of its size, a team
    of four warps per window.  Forced: every kind of window starts there; same bytes as the oracle."""
    native.set_option("first_tier", native.TIER_BIG)
    for label, b in _mixed_batches():
        want, _ = oracle_consensus(b)
        _same(native.consensus(b), want, b, f"T2s/{label}")
        _, _, tiers = native.last_timing()
        assert tiers[native.TIER_BIG] > 0 and sum(tiers[:6]) == 0, (label, tiers)


def test_noisy_long_windows_run_in_the_estimate_driven_tier():
    """LONG windows of noisy reads outgrow T1 (1024 nodes) and are re-run in T2s; the noisiest outgrow its
    estimate as well and end in the bound-driven tiers.  Bytes never change."""
    for err, n in ((0.05, 40), (0.10, 24)):
        b = synth_batch(121, n, 420, 24, "internal", err, wtype=WINDOW_LONG)
        want, _ = oracle_consensus(b)
        _same(native.consensus(b), want, b, f"LONG err {err}")
        _, _, tiers = native.last_timing()
        assert tiers[native.TIER_BIG] > 0, (err, tiers)
    # large SHORT windows keep going to the bound-driven tier (faster for them at 16 warps / SM)
    b = synth_batch(122, 24, 480, 60, "mixed", 0.02)
    want, _ = oracle_consensus(b)
    _same(native.consensus(b), want, b, "SHORT 60 x 480")
    _, _, tiers = native.last_timing()
    assert tiers[native.TIER_BIG] == 0 and tiers[6] == b.n_win, tiers


def test_repeated_reads_are_not_aligned_twice():
    """A read identical to the one added last - when that one left the DAG's structure untouched - is not
    aligned again: its weights go along the stored node path (repeat_sequence).  Windows made of repeats only
    (error-free reads), windows where repeats and erroneous reads alternate, prefix / suffix arms, LONG windows
    (whose per-sequence node paths the repeats must still record): same bytes as the oracle, and the device's
    DP-cell counter shows that most fills were skipped."""
    rng = np.random.default_rng(131)
    specs = []
    for i in range(300):
        truth = "".join("ACGT"[int(x)] for x in rng.integers(0, 4, size=int(rng.integers(20, 110))))
        bad = truth[: len(truth) // 2] + "A" + truth[len(truth) // 2 + 1:]
        arms = [truth] * 6 + [bad] + [truth] * 5 + [bad, bad] + [truth] * 8
        if i % 3 == 0:
            specs.append(WindowSpec(truth, arms, [], [], 0, 0))
        elif i % 3 == 1:
            cut = len(truth) // 2
            specs.append(WindowSpec(truth, arms[:9], [truth[:cut + 3]] * 5 + [truth[:cut]] * 3, [truth[cut:]] * 6, 0, 0))
        else:
            specs.append(WindowSpec(truth, [], [truth[: len(truth) - 2]] * 7, [truth[2:]] * 7, 0, 0))
    b = build_batch(specs)
    want, _ = oracle_consensus(b)
    for first_tier, group_tiers in ((0, 1), (0, 2), (3, 1), (6, 1)):
        native.set_option("first_tier", first_tier)
        native.set_option("group_tiers", group_tiers)
        try:
            _same(native.consensus(b), want, b, f"repeats tier {first_tier} groups {group_tiers}")
        finally:
            native.set_option("group_tiers", 1)
    native.set_option("first_tier", 0)
    clean = build_batch([WindowSpec("ACGTTGCAAGGCTTAACCGGTTAA" * 4, ["ACGTTGCAAGGCTTAACCGGTTAA" * 4] * 30, [], [], 0, 0)] * 64)
    native.consensus(clean)
    cells_all = 64 * 30 * 99 * 99   # what 30 fills per window would count at the very least
    assert 0 < native.last_cells() < cells_all // 5, (native.last_cells(), cells_all)
    # LONG windows: both rounds, node paths recorded for repeated arms too
    lb = build_batch([WindowSpec(s.draft, list(s.internal) + list(s.internal[:4]), [], [], 0, WINDOW_LONG)
                      for s in (b.spec(i) for i in range(0, 300, 3))])
    want, _ = oracle_consensus(lb)
    _same(native.consensus(lb), want, lb, "LONG repeats")
