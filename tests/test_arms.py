"""Arm extraction: reads cut into per-window arms, windows filled and pruned.

CPU: the Python restatement (oracle/arms_oracle.py) against what the reference command-line program itself
produced - its dump lists every window's arms in insertion order after pruning.  GPU: hypo_gpu_extract_arms
against both, and the fused hypo_gpu_polish_alignments (alignments in, polished contig out) against the
reference CLI's polished FASTA."""
import gzip
import os

import numpy as np
import pytest

from oracle import arms_oracle as ao
from tests.arms_util import GOLDEN, device_inputs, load_capture

K = 9   # the capture ran with -s 60k: k = 9 (reference src/main.cpp:490-528)


def _windows_of_dump(regions, dumped):
    out = {}
    for i, d in enumerate(dumped):
        if d is not None and regions[i].type not in ao.SR_TYPES:
            out[i] = (d[0], d[1], d[2], d[3], list(d[5]))
    return out


def _windows_of_oracle(regions, wins):
    out = {}
    for i, w in wins.items():
        if w.dropped or len(w.internal) + len(w.pre) + len(w.suf) + w.n_empty == 0:
            continue
        out[i] = (len(w.internal), len(w.pre), len(w.suf), w.n_empty, w.internal + w.pre + w.suf)
    return out


def test_arms_oracle_reproduces_the_reference_cli_windows():
    regions, clen, dumped, recs = load_capture()
    alns = [ao.make_alignment(p, c, s) for _, p, c, s in recs]
    wins = ao.fill_and_prune(alns, K, regions, clen)
    assert _windows_of_oracle(regions, wins) == _windows_of_dump(regions, dumped)
    assert len(_windows_of_dump(regions, dumped)) == 1720


def test_find_bp_corner_cases():
    """Region starts that fall on a CIGAR boundary: an insertion right at the boundary goes to the window on
    the left unless that is a strong region (src/Alignment.cpp:383-391)."""
    regs = [ao.Region(0, "SR", "A" * 20), ao.Region(20, "SWS", "C" * 10), ao.Region(30, "SR", "G" * 20)]
    starts = [0, 20, 30, 50]
    types = [r.type for r in regs]
    a = ao.make_alignment(10, "10M3I10M2I15M", "A" * 40)      # insertions exactly at 20 and at 30
    assert ao.find_bp(a, starts, types, 0, 3) == [10, 25]      # first goes right (left is SR), second stays left
    a = ao.make_alignment(10, "5M10D20M", "A" * 25)            # deletion across the first boundary
    assert ao.find_bp(a, starts, types, 0, 3) == [5, 10]


@pytest.mark.gpu
def test_device_arm_extraction_equals_the_reference_cli_windows():
    from hypo_b200 import native
    regions, clen, dumped, recs = load_capture()
    native.init((5, -4, -8, 3, -5, -4), 0)
    batch, win_region = native.extract_arms(*device_inputs(regions, clen, recs), K)
    want = _windows_of_dump(regions, dumped)
    assert sorted(int(r) for r in win_region) == sorted(want)
    for w in range(batch.n_win):
        s = batch.spec(w)
        r = int(win_region[w])
        assert s.draft == regions[r].text and s.wtype == 0
        assert (len(s.internal), len(s.pre), len(s.suf), s.n_empty, list(s.internal) + list(s.pre) + list(s.suf)) == want[r], r
    # the batch is laid out like the host packer lays it out: window after window, draft then arms
    assert int(batch.win["draft_off"][0]) == 0 and (np.diff(batch.win["draft_off"].astype(np.int64)) > 0).all()
    # ... and polishes to the consensus strings the reference recorded
    got = native.consensus(batch)
    assert got == [dumped[int(r)][4] for r in win_region]


@pytest.mark.gpu
def test_fused_polish_from_alignments_equals_the_reference_cli_output():
    from hypo_b200 import native
    regions, clen, dumped, recs = load_capture()
    native.init((5, -4, -8, 3, -5, -4), 0)
    out = native.polish_alignments(*device_inputs(regions, clen, recs), K)
    want = gzip.open(os.path.join(GOLDEN, "cli_short_60kb.polished.fa.gz"), "rt").read().split("\n")[1]
    assert out == [want]


@pytest.mark.gpu
def test_device_arm_extraction_edge_cases():
    """Reads that are dropped (N in the aligned part), clipped reads, a read inside one region, no reads."""
    from hypo_b200 import native
    regions, clen, dumped, recs = load_capture()
    native.init((5, -4, -8, 3, -5, -4), 0)
    sub = recs[:400]
    # soft / hard clips around the same alignment, an N inside one read, a read with an N only in its clip
    mod = []
    for i, (c, p, cg, s) in enumerate(sub):
        if i % 5 == 1:
            mod.append((c, p, "3S" + cg + "2S", "TTT" + s + "GG"))
        elif i % 5 == 2:
            mod.append((c, p, "4H" + cg, s))
        elif i % 5 == 3:
            mod.append((c, p, cg, s[:40] + "N" + s[41:]))
        elif i % 5 == 4:
            mod.append((c, p, "1S" + cg, "N" + s))
        else:
            mod.append((c, p, cg, s))
    alns = [ao.make_alignment(p, c, s) for _, p, c, s in mod]
    want = _windows_of_oracle(regions, ao.fill_and_prune(alns, K, regions, clen))
    batch, win_region = native.extract_arms(*device_inputs(regions, clen, mod), K)
    got = {}
    for w in range(batch.n_win):
        s = batch.spec(w)
        got[int(win_region[w])] = (len(s.internal), len(s.pre), len(s.suf), s.n_empty, list(s.internal) + list(s.pre) + list(s.suf))
    assert got == want and len(want) > 5
    empty, _ = native.extract_arms(*device_inputs(regions, clen, []), K)
    assert empty.n_win == 0
