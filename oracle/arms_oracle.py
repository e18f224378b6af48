"""TEST INFRASTRUCTURE ONLY — CPU restatement (plain Python, small cases) of the step that FEEDS the POA path:
cutting aligned reads into per-window arms and filling / pruning the windows.  Only tests/ may import it.

Restated from the reference (paths relative to the reference root):
  Alignment::initialise_pos / copy_data      src/Alignment.cpp:513-576   (clipping, reference span, aligned query)
  Alignment::find_short_arms                 src/Alignment.cpp:222-259   (which windows a read touches)
  Alignment::find_bp                         src/Alignment.cpp:321-404   (CIGAR walk: query position of every region start)
  Alignment::prepare_short_arm               src/Alignment.cpp:406-509   (anchor k-mer / minimiser validation)
  Alignment::add_arms                        src/Alignment.cpp:299-318   (insertion in alignment order)
  Contig::fill_short_windows (pruning part)  src/Contig.cpp:262-289      (window dropped / prefix+suffix arms cleared)
Region tables follow Contig::prepare_for_division / divide_into_regions (src/Contig.cpp:75-245): a strong region's
anchor k-mers are the k-mers at its first and last k bases (:128-129), a minimiser region's key is the minimiser
itself (:596,614).

Parity pinned: tests/test_arms.py rebuilds, from the SAM records and the region table of a run of the reference
command-line program, exactly the windows (arms in order, counters, which windows were dropped) that the
reference's own dump of the same run lists (tests/golden/cli_short_60kb.*).
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

# reference include/globalDefs.hpp:146-156, src/main.cpp:85-88
ARMS = dict(min_short_num=3, min_internal_num1=20, min_internal_num2=5, min_internal_num3=10, min_contrib=10,
            min_internal_contrib=0.4, short_arm_coef=10)
MINIMIZER_K = 10
SR_TYPES = ("SR", "MSR")
INTERNAL, PREFIX, SUFFIX, EMPTY = 0, 1, 2, 3
_CODE = {"A": 0, "C": 1, "G": 2, "T": 3}


def kmer_code(s: str) -> int:
    v = 0
    for c in s:
        v = (v << 2) | _CODE[c]
    return v


@dataclass
class Region:
    beg: int
    type: str
    text: str = ""          # draft sequence of the region
    key0: int = 0           # SR: first anchor k-mer; MSR: the minimiser
    key1: int = 0           # SR: last anchor k-mer


@dataclass
class WindowArms:
    internal: List[str] = field(default_factory=list)
    pre: List[str] = field(default_factory=list)
    suf: List[str] = field(default_factory=list)
    n_empty: int = 0
    maxlen_pre: int = 0
    maxlen_suf: int = 0
    dropped: bool = False


CIGAR_RE = re.compile(r"(\d+)([MIDNSHP=X])")
# bam_cigar_type: bit 0 consumes query, bit 1 consumes reference (htslib sam.h)
CIGAR_TYPE = {"M": 3, "I": 1, "D": 2, "N": 2, "S": 1, "H": 0, "P": 0, "=": 3, "X": 3}


@dataclass
class Aln:
    rb: int
    re: int
    qae: int                      # length of the aligned part of the query (clips removed)
    cigar: List[Tuple[str, int]]
    seq: str                      # the aligned part of the query
    valid: bool


def make_alignment(pos0: int, cigar: str, seq: str) -> Aln:
    """initialise_pos + copy_data: soft clips at both ends are cut off the query, a read with a base other
    than A/C/G/T in its aligned part is dropped (PackedSeq<2> cannot hold it)."""
    ops = [(op, int(n)) for n, op in CIGAR_RE.findall(cigar)]
    qab, qp, rp, clip_before, clip_end = 0, 0, pos0, True, 0
    for op, n in ops:
        if clip_before:
            if op == "S":
                qab += n
            elif op != "H":
                clip_before = False
        t = CIGAR_TYPE[op]
        if t & 3 == 3:
            rp += n; qp += n
        elif t & 2:
            rp += n
        elif t & 1:
            if not clip_before and op == "S":
                clip_end += n
            qp += n
    qae = qp - clip_end
    aligned = seq[qab:qae]
    return Aln(pos0, rp, qae - qab, ops, aligned, all(c in _CODE for c in aligned))


def find_bp(aln: Aln, reg_start: List[int], reg_type: List[str], beg_ind: int, end_ind: int) -> List[int]:
    """src/Alignment.cpp:321-404 — query position at which every region after the first begins."""
    res: List[int] = []
    cur_ref = aln.rb
    idx = beg_ind + 1
    next_ref = reg_start[idx]
    qpos = 0
    corner = False
    for op, n in aln.cigar:
        if op in "SH":
            continue
        t = CIGAR_TYPE[op]
        if t & 3 == 3:
            if corner:
                res.append(qpos); corner = False; idx += 1; next_ref = reg_start[idx]
            while cur_ref + n >= next_ref and not corner:
                d = next_ref - cur_ref
                cur_ref = next_ref; qpos += d; n -= d
                if n > 0:
                    res.append(qpos); idx += 1; next_ref = reg_start[idx]
                else:
                    corner = True
            if n > 0:
                cur_ref += n; qpos += n
        elif t & 2:
            if corner:
                res.append(qpos); corner = False; idx += 1; next_ref = reg_start[idx]
            while cur_ref + n >= next_ref and not corner:
                d = next_ref - cur_ref
                cur_ref = next_ref; n -= d
                if n > 0:
                    res.append(qpos); idx += 1; next_ref = reg_start[idx]
                else:
                    corner = True
            if n > 0:
                cur_ref += n
        elif t & 1:
            if corner:
                res.append(qpos if reg_type[idx - 1] in SR_TYPES else qpos + n)
                idx += 1; next_ref = reg_start[idx]; corner = False
            qpos += n
        if idx == end_ind:
            break
    return res


def _find_kmer(seq: str, target: int, k: int, left: int, right: int, first: bool) -> Optional[int]:
    """PackedSeq::find_kmer (src/PackedSeq.cpp:264-330): the k-mer lies wholly inside [left, right)."""
    found = None
    mask = (1 << (2 * k)) - 1
    v, run = 0, 0
    for i in range(left, right):
        v = ((v << 2) | _CODE[seq[i]]) & mask
        run = min(run + 1, k)
        if run == k and v == target:
            found = i + 1 - k
            if first:
                break
    return found


def _check_kmer(seq: str, target: int, k: int, at: int) -> bool:
    return at + k <= len(seq) and _find_kmer(seq, target, k, at, at + k, True) is not None


def prepare_short_arm(aln: Aln, k: int, windex: int, qb: int, qe: int, armtype: int, regions: List[Region],
                      reg_start: List[int]):
    """src/Alignment.cpp:406-509; returns (q_beg, q_end) or None."""
    mk = MINIMIZER_K
    if (reg_start[windex + 1] - reg_start[windex]) > ARMS["short_arm_coef"] * (qe - qb):
        return None
    wtype = regions[windex].type
    valid, q_beg, q_end, seq, qae = True, qb, qe, aln.seq, aln.qae
    if wtype in ("SWS", "SW", "SWM") and armtype != SUFFIX:
        if q_beg < k:
            valid = False
        else:
            anchor = regions[windex - 1].key1            # last k-mer of the preceding strong region
            if not _check_kmer(seq, anchor, k, q_beg - k):
                s0 = 0 if q_beg < 2 * k else q_beg - 2 * k
                s1 = q_end if q_end < q_beg + k else q_beg + k
                at = _find_kmer(seq, anchor, k, s0, s1, False)
                if at is not None:
                    q_beg = at + k
                else:
                    valid = False
    if wtype in ("SWS", "WS", "MWS") and armtype != PREFIX:
        if q_end + k > qae:
            valid = False
        else:
            anchor = regions[windex + 1].key0            # first k-mer of the following strong region
            if not _check_kmer(seq, anchor, k, q_end):
                s0 = q_beg if q_end < q_beg + k else q_end - k
                s1 = min(qae, q_end + 2 * k)
                at = _find_kmer(seq, anchor, k, s0, s1, True)
                if at is not None:
                    q_end = at
                else:
                    valid = False
    if wtype in ("MWM", "MW", "MWS") and armtype != SUFFIX:
        if q_beg < mk:
            valid = False
        else:
            mini = regions[windex - 1].key0
            if not _check_kmer(seq, mini, mk, q_beg - mk):
                s0 = 0 if q_beg < 3 * mk else q_beg - 3 * mk
                s1 = q_end if q_end < q_beg + 2 * mk else q_beg + 2 * mk
                at = _find_kmer(seq, mini, mk, s0, s1, False)
                if at is not None:
                    q_beg = at + mk
                else:
                    valid = False
    if wtype in ("MWM", "WM", "SWM") and armtype != PREFIX:
        if q_end + mk > qae:
            valid = False
        else:
            mini = regions[windex + 1].key0
            if not _check_kmer(seq, mini, mk, q_end):
                s0 = q_beg if q_end < q_beg + 2 * mk else q_end - 2 * mk
                s1 = min(qae, q_end + 3 * mk)
                at = _find_kmer(seq, mini, mk, s0, s1, True)
                if at is not None:
                    q_end = at
                else:
                    valid = False
    if valid and q_beg < q_end:
        return q_beg, q_end
    return None


def find_short_arms(aln: Aln, k: int, regions: List[Region], reg_start: List[int]):
    """src/Alignment.cpp:222-259; returns [(window region index, type, q_beg, q_end)] in the reference's order."""
    import bisect
    starts = reg_start
    is_start = lambda x: starts[bisect.bisect_left(starts, x)] == x if bisect.bisect_left(starts, x) < len(starts) else False
    b_ind = bisect.bisect_left(starts, aln.rb)          # number of region starts < rb
    if not is_start(aln.rb):
        b_ind -= 1
    e_ind = bisect.bisect_left(starts, aln.re)
    arms = []
    if e_ind - b_ind > 1:
        types = [r.type for r in regions]
        bp = find_bp(aln, starts, types, b_ind, e_ind)

        def window(i):
            return regions[i].type not in SR_TYPES

        armtype = INTERNAL if is_start(aln.rb) else SUFFIX
        if window(b_ind):
            r = prepare_short_arm(aln, k, b_ind, 0, bp[0], armtype, regions, starts)
            if r:
                arms.append((b_ind, armtype, r[0], r[1]))
        bi = 0
        for ind in range(b_ind + 1, e_ind - 1):
            if window(ind):
                if bp[bi + 1] == bp[bi]:
                    arms.append((ind, EMPTY, 0, 0))
                else:
                    r = prepare_short_arm(aln, k, ind, bp[bi], bp[bi + 1], INTERNAL, regions, starts)
                    if r:
                        arms.append((ind, INTERNAL, r[0], r[1]))
            bi += 1
        armtype = INTERNAL if is_start(aln.re) else PREFIX
        if window(e_ind - 1):
            r = prepare_short_arm(aln, k, e_ind - 1, bp[bi], aln.qae, armtype, regions, starts)
            if r:
                arms.append((e_ind - 1, armtype, r[0], r[1]))
    return arms


def fill_and_prune(alns: List[Aln], k: int, regions: List[Region], contig_len: int) -> Dict[int, WindowArms]:
    """add_arms in alignment order, then the pruning rules of Contig::fill_short_windows."""
    reg_start = [r.beg for r in regions] + [contig_len]
    wins: Dict[int, WindowArms] = {i: WindowArms() for i, r in enumerate(regions) if r.type not in SR_TYPES}
    for a in alns:
        if not a.valid:
            continue
        for windex, t, qb, qe in find_short_arms(a, k, regions, reg_start):
            w = wins[windex]
            s = a.seq[qb:qe]
            if t == INTERNAL:
                w.internal.append(s)
            elif t == PREFIX:
                w.pre.append(s); w.maxlen_pre = max(w.maxlen_pre, len(s))
            elif t == SUFFIX:
                w.suf.append(s); w.maxlen_suf = max(w.maxlen_suf, len(s))
            else:
                w.n_empty += 1
    for i, w in wins.items():
        internal = len(w.internal) + w.n_empty                  # Window::get_num_internal counts the empties
        if internal < ARMS["min_short_num"]:
            win_len = reg_start[i + 1] - reg_start[i]
            covered = w.maxlen_pre + w.maxlen_suf >= win_len
            enough = len(w.pre) >= ARMS["min_short_num"] and len(w.suf) >= ARMS["min_short_num"]
            if not (covered and enough):
                w.dropped = True
                continue
        contrib = internal + len(w.pre) + len(w.suf)
        import math
        cond0 = internal > ARMS["min_internal_num1"]
        cond1 = contrib >= ARMS["min_contrib"] and internal >= math.floor(ARMS["min_internal_contrib"] * contrib)
        cond2 = regions[i].type in ("SWS", "SW", "WS", "MWS", "SWM") and internal >= ARMS["min_internal_num2"]
        if cond0 or cond1 or cond2:
            w.pre, w.suf = [], []
    return wins


# ---- inputs from a run of the reference CLI (tools/capture): SAM records + the region table of its dump -----

def read_sam(path_or_lines) -> List[Tuple[str, int, str, str]]:
    """(contig, 0-based position, CIGAR, SEQ) of every record the reference would keep
    (src/Hypo.cpp:296-300: unmapped / secondary / QC-fail / duplicate records and MAPQ < threshold are skipped)."""
    out = []
    lines = open(path_or_lines) if isinstance(path_or_lines, str) else path_or_lines
    for ln in lines:
        if ln.startswith("@"):
            continue
        f = ln.rstrip("\n").split("\t")
        flag = int(f[1])
        if flag & (0x4 | 0x100 | 0x200 | 0x400) or int(f[4]) < 2:
            continue
        out.append((f[2], int(f[3]) - 1, f[5], f[9]))
    return out


def read_regions(inspect_lines, k: int) -> Tuple[List[Region], int, list]:
    """Region table + the windows the reference ended up with, from its per-contig dump
    (Contig::generate_inspect_file, src/Contig.cpp:368-453).  Returns (regions, contig length, dumped windows):
    dumped[i] is None for a region without arms, else (n_int, n_pre, n_suf, n_empty, consensus, [arms...])."""
    lines = [l.rstrip("\n") for l in inspect_lines]
    regions: List[Region] = []
    dumped = []
    i = 0
    while i < len(lines):
        ln = lines[i]
        if not ln.startswith("=========="):
            i += 1
            continue
        m = re.match(r"==========\((\d+)-(\d+)\)\t(\w+)\t(\d+)\t(\d+)\t(\d+)\t(\d+)", ln)
        beg, typ = int(m.group(1)), m.group(3)
        n_int, n_pre, n_suf, n_empty = (int(m.group(j)) for j in (4, 5, 6, 7))
        draft = lines[i + 1].split("\t", 1)[1] if "\t" in lines[i + 1] else ""
        cons = lines[i + 2].split("\t", 1)[1] if "\t" in lines[i + 2] else ""
        n_arms = n_int + n_pre + n_suf
        arms = lines[i + 3:i + 3 + n_arms]
        r = Region(beg, typ, draft)
        if typ == "SR":
            r.key0, r.key1 = kmer_code(draft[:k]), kmer_code(draft[-k:])
        elif typ == "MSR":
            r.key0 = kmer_code(draft[:MINIMIZER_K])
        regions.append(r)
        dumped.append(None if n_arms + n_empty == 0 else (n_int, n_pre, n_suf, n_empty, cons, arms))
        i += 3 + n_arms
    contig_len = regions[-1].beg + len(regions[-1].text)
    return regions, contig_len, dumped
