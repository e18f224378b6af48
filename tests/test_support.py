"""Support counting: coverage / support of every solid k-mer and of every minimiser.

CPU: the Python restatement (oracle/support_oracle.py) against the counters the reference command-line program
ended up with in the captured run (tools/capture hypo_dump2 dumps Contig::_kmerinfo and _minimserinfo).
GPU: hypo_gpu_solid_kmer_support / hypo_gpu_minimiser_support against the same counters."""
import gzip
import os

import numpy as np
import pytest

from oracle import arms_oracle as ao
from oracle import support_oracle as so
from tests.arms_util import GOLDEN, device_inputs, load_capture

K = 9


def _dumps():
    spos, kid, cov, sup = so.read_kmer_dump(gzip.open(os.path.join(GOLDEN, "cli_short_60kb.kmer_support.gz"), "rt"))
    mini = so.read_minimiser_dump(gzip.open(os.path.join(GOLDEN, "cli_short_60kb.minimiser_support.gz"), "rt"))
    return (spos, kid, cov, sup), mini


def _alignments(recs):
    alns = [ao.make_alignment(p, c, s) for _, p, c, s in recs]
    return [(a.rb, a.re, a.seq) for a in alns if a.valid]


def test_support_oracle_reproduces_the_reference_cli_counters():
    regions, clen, dumped, recs = load_capture()
    (spos, kid, cov, sup), (even, bounds, minfo, mcov, msup) = _dumps()
    A = _alignments(recs)
    c, s = so.solid_kmer_support(A, K, spos, kid)
    assert len(spos) == 9450 and c == cov and s == sup
    c, s = so.minimiser_support(A, bounds, even, minfo)
    assert sum(len(m) for m in minfo) == 1216 and c == mcov and s == msup


def _mini_tables(even, bounds, minfo):
    """Per bound index: CSR of absolute minimiser positions and values."""
    first = np.zeros(len(bounds) + 1, np.uint64)
    pos, val, order = [], [], []
    for i in range(len(bounds)):
        first[i] = len(pos)
        is_win = (i % 2 == 0) if even else (i % 2 == 1)
        m = i // 2 if even else (i - 1) // 2
        if is_win and 0 <= m < len(minfo):
            p = bounds[i]
            for j, (rel, v) in enumerate(minfo[m]):
                p += rel
                pos.append(p); val.append(v); order.append((m, j))
    first[len(bounds)] = len(pos)
    return first, np.array(pos, np.uint32), np.array(val, np.uint32), order


@pytest.mark.gpu
def test_device_support_counters_equal_the_reference_cli_counters():
    from hypo_b200 import native
    regions, clen, dumped, recs = load_capture()
    (spos, kid, cov, sup), (even, bounds, minfo, mcov, msup) = _dumps()
    native.init((5, -4, -8, 3, -5, -4), 0)
    _, _, _, alns, cigar, seqs = device_inputs(regions, clen, recs)
    c, s = native.solid_kmer_support(np.array([0, len(spos)], np.uint64), np.array(spos, np.uint32), np.array(kid, np.uint64),
                                     alns, cigar, seqs, K)
    assert c.tolist() == cov and s.tolist() == sup
    first, mpos, mval, order = _mini_tables(even, bounds, minfo)
    c, s = native.minimiser_support(np.array([0, len(bounds)], np.uint64), np.array([1 if even else 0], np.uint8),
                                    np.array(bounds, np.uint32), first, mpos, mval, alns, cigar, seqs)
    assert [int(c[i]) for i in range(len(order))] == [mcov[m][j] for m, j in order]
    assert [int(s[i]) for i in range(len(order))] == [msup[m][j] for m, j in order]


@pytest.mark.gpu
def test_device_support_counters_on_modified_reads_equal_the_oracle():
    """Clipped reads, reads with N (dropped), reads shifted so that spans end inside k-mers."""
    from hypo_b200 import native
    regions, clen, dumped, recs = load_capture()
    (spos, kid, _, _), (even, bounds, minfo, _, _) = _dumps()
    native.init((5, -4, -8, 3, -5, -4), 0)
    mod = []
    for i, (c, p, cg, s) in enumerate(recs[:3000]):
        if i % 4 == 1:
            mod.append((c, p, "2S" + cg + "3S", "GG" + s + "TTT"))
        elif i % 4 == 2:
            mod.append((c, p, cg, s[:70] + "N" + s[71:]))
        elif i % 4 == 3 and p > 5:
            mod.append((c, p - 3, cg, s))
        else:
            mod.append((c, p, cg, s))
    A = _alignments(mod)
    _, _, _, alns, cigar, seqs = device_inputs(regions, clen, mod)
    wc, ws = so.solid_kmer_support(A, K, spos, kid)
    c, s = native.solid_kmer_support(np.array([0, len(spos)], np.uint64), np.array(spos, np.uint32), np.array(kid, np.uint64),
                                     alns, cigar, seqs, K)
    assert c.tolist() == wc and s.tolist() == ws
    wc, ws = so.minimiser_support(A, bounds, even, minfo)
    first, mpos, mval, order = _mini_tables(even, bounds, minfo)
    c, s = native.minimiser_support(np.array([0, len(bounds)], np.uint64), np.array([1 if even else 0], np.uint8),
                                    np.array(bounds, np.uint32), first, mpos, mval, alns, cigar, seqs)
    assert [int(c[i]) for i in range(len(order))] == [wc[m][j] for m, j in order]
    assert [int(s[i]) for i in range(len(order))] == [ws[m][j] for m, j in order]
