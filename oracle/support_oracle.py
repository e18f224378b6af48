"""TEST INFRASTRUCTURE ONLY — CPU restatement (plain Python, small cases) of the support-counting step that precedes
windowing.  Only tests/ may import it.

Restated from the reference (paths relative to the reference root):
  Alignment::update_solidkmers_support   src/Alignment.cpp:65-131   coverage / support of every solid k-mer of the draft
  Alignment::update_minimisers_support   src/Alignment.cpp:133-220  coverage / support of every minimiser of the large
                                                                    weak regions
Inputs as the reference holds them: the solid positions of a contig with their k-mer ids (Contig::_solid_pos /
_kmerinfo, src/Contig.cpp:40-74), after prepare_for_division the boundaries between strong regions and the
regions in between (Contig::_reg_pos, _is_win_even) and the minimisers of every such region
(Contig::_minimserinfo, src/Contig.cpp:455-518); alignments as (reference begin, reference end, aligned read).

Parity pinned: tests/test_support.py feeds it the tables a run of the reference command-line program dumped
(tools/capture: hypo_dump2) together with the SAM records of that run and requires the very counters the
reference ended up with (tests/golden/cli_short_60kb.kmer_support.gz / .minimiser_support.gz).
"""
from __future__ import annotations

import bisect
from typing import List, Tuple

MINIMIZER_K = 10   # reference src/main.cpp:86
MINIMIZER_W = 10
_CODE = {"A": 0, "C": 1, "G": 2, "T": 3}
U32 = 0xFFFFFFFF
U64 = 0xFFFFFFFFFFFFFFFF


def solid_kmer_support(alns, k: int, spos: List[int], kid: List[int]) -> Tuple[List[int], List[int]]:
    """alns: iterable of (rb, re, aligned read).  Returns (coverage, support) per solid k-mer."""
    n = len(spos)
    cov, sup = [0] * n, [0] * n
    kmask = (1 << (2 * k)) - 1
    for rb, re, seq in alns:
        first = bisect.bisect_left(spos, rb)      # _Rsolid_pos(_rb): solid positions < rb
        last = bisect.bisect_left(spos, re)
        i = last
        while i > first:                          # discard those which do not wholly fall in the alignment
            if spos[i - 1] + k <= re:             # _Ssolid_pos(i) is 1-based
                last = i
                break
            i -= 1
        if last <= first:
            continue
        for i in range(first, last):
            cov[i] += 1
        num_cbases = re - rb
        pvs_kpos, pvs_rbind = -1, 0
        kmer, klen = 0, 0
        lo = first                                # candidates: solid positions within k of the expected place
        for r_ind, ch in enumerate(seq):
            kmer = ((kmer << 2) | _CODE[ch]) & kmask
            if klen < k:
                klen += 1
            if klen < k:
                continue
            r_bind = r_ind + 1 - k
            # every solid k-mer of the span with this id, in position order (ids repeat only where the draft repeats)
            for c in range(first, last):
                if kid[c] != kmer:
                    continue
                c_dist = spos[c] - rb
                left = c_dist - k if c_dist > k else 0
                right = min(num_cbases, c_dist + k)
                if left <= r_bind <= right:
                    update = True
                    if pvs_kpos > -1 and spos[c] <= k + pvs_kpos:
                        if ((r_bind - pvs_rbind) & U32) != ((spos[c] - pvs_kpos) & U64):
                            update = False
                    if update:
                        pvs_kpos, pvs_rbind = spos[c], r_bind
                        sup[c] += 1
    return cov, sup


def read_minimisers(seq: str) -> List[Tuple[int, int]]:
    """(minimiser, start position) of every window of MINIMIZER_W k-mers of the read, consecutive duplicates
    dropped (src/Alignment.cpp:146-186; the deque keeps the left-most smallest k-mer)."""
    k, w = MINIMIZER_K, MINIMIZER_W
    mask = (1 << (2 * k)) - 1
    out = []
    last_found = len(seq) + 1
    kmer, not_n, processed = 0, 0, 0
    dq: List[Tuple[int, int]] = []
    for i, ch in enumerate(seq):
        c = _CODE.get(ch, 4)
        if c < 4:
            not_n += 1
            kmer = ((kmer << 2) | c) & mask
            if not_n >= k:
                while dq and dq[-1][0] > kmer:
                    dq.pop()
                dq.append((kmer, i))
                while dq[0][1] + w <= i:
                    dq.pop(0)
                processed += 1
                if processed >= w:
                    start = dq[0][1] - k + 1
                    if start != last_found:
                        out.append((dq[0][0], start))
                    last_found = start
        else:
            not_n = 0
    return out


def minimiser_support(alns, bounds: List[int], even: bool, minfo: List[List[Tuple[int, int]]]):
    """minfo[m] = [(relative position, minimiser)] of the m-th region between strong regions.  Returns per region the
    lists (coverage, support)."""
    cov = [[0] * len(m) for m in minfo]
    sup = [[0] * len(m) for m in minfo]
    K = MINIMIZER_K
    for rb, re, seq in alns:
        first = bisect.bisect_right(bounds, rb) - 1       # _RMreg_pos(_rb + 1) - 1
        last = bisect.bisect_left(bounds, re)             # _RMreg_pos(_re)
        is_win = lambda x: (even and x % 2 == 0) or (not even and x % 2 == 1)
        fw = first if is_win(first) else first + 1
        lw = last if is_win(last) else last - 1
        if lw < fw:
            continue
        found = read_minimisers(seq)
        num_cbases = (re - rb) & 0xFFFF                   # UINT16 in the reference
        for i in range(fw, lw + 1, 2):
            m = i // 2 if even else (i - 1) // 2
            if m >= len(minfo) or i >= len(bounds):
                continue
            pos = bounds[i]
            for mi, (rel, mini) in enumerate(minfo[m]):
                pos += rel
                c_dist = (pos - rb) & U32
                left = c_dist - 2 * K if c_dist > 2 * K else 0
                right = min(num_cbases, (c_dist + 3 * K) & 0xFFFF)
                if rb <= pos < re:
                    cov[m][mi] += 1
                    for val, rpos in found:
                        if val == mini and left <= rpos <= right:
                            sup[m][mi] += 1
                if pos >= re:
                    break
    return cov, sup


def read_kmer_dump(lines):
    """pos, id, coverage, support of every solid k-mer (tools/capture hypo_dump2: aux/kmer_support_<contig>.txt)."""
    spos, kid, cov, sup = [], [], [], []
    for ln in lines:
        p, k, c, s = ln.split()
        spos.append(int(p)); kid.append(int(k)); cov.append(int(c)); sup.append(int(s))
    return spos, kid, cov, sup


def read_minimiser_dump(lines):
    """even flag, region boundaries, and per region [(rel_pos, minimiser)], coverage, support."""
    even, bounds, minfo, cov, sup = True, [], [], [], []
    for ln in lines:
        f = ln.rstrip("\n").split("\t")
        if f[0] == "even":
            even = f[1] == "1"
        elif f[0] == "bounds":
            bounds = [int(x) for x in f[1:]]
        elif f[0].startswith("#"):
            minfo.append([]); cov.append([]); sup.append([])
        elif len(f) == 4:
            minfo[-1].append((int(f[0]), int(f[1]))); cov[-1].append(int(f[2])); sup[-1].append(int(f[3]))
    return even, bounds, minfo, cov, sup
