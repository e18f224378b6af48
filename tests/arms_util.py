"""Test helpers: the inputs of arm extraction (hypo_gpu_extract_arms) from a run of the reference CLI - its SAM
records and the region table of its dump - in the layouts of include/hypo_b200.h."""
import gzip
import os

import numpy as np

from hypo_b200.batch import pack4
from hypo_b200.native import ALN_DTYPE, CONTIG_DTYPE, REGION_DTYPE, REGION_TYPES
from oracle import arms_oracle as ao

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_BAM_NIB = {"A": 1, "C": 2, "G": 4, "T": 8}
_CIGAR_OP = {c: i for i, c in enumerate("MIDNSHP=X")}


def load_capture(name="cli_short_60kb", k=9):
    """(regions, contig length, dumped windows, SAM records) of a captured run."""
    regions, clen, dumped = ao.read_regions(gzip.open(os.path.join(GOLDEN, name + ".inspect.gz"), "rt"), k)
    recs = ao.read_sam(gzip.open(os.path.join(GOLDEN, name + ".sam.gz"), "rt"))
    return regions, clen, dumped, recs


def device_inputs(regions, clen, recs):
    """Region table, draft and alignments of ONE contig as the arrays hypo_gpu_extract_arms takes."""
    reg = np.zeros(len(regions), REGION_DTYPE)
    for i, r in enumerate(regions):
        reg[i] = (r.key0, r.key1, r.beg, REGION_TYPES[r.type])
    draft = "".join(r.text for r in regions)
    assert len(draft) == clen
    drafts = np.concatenate([pack4(draft), np.zeros(16, np.uint8)])
    contigs = np.zeros(1, CONTIG_DTYPE)
    contigs[0] = (0, 0, len(regions), clen)
    alns = np.zeros(len(recs), ALN_DTYPE)
    cig, seq_chunks, seq_pos = [], [], 0
    for i, (_, pos, cigar, seq) in enumerate(recs):
        ops = ao.CIGAR_RE.findall(cigar)
        alns[i] = (len(cig), seq_pos, 0, pos, len(ops), len(seq))
        cig.extend((int(n) << 4) | _CIGAR_OP[op] for n, op in ops)
        nibs = np.array([_BAM_NIB.get(c, 15) for c in seq] + ([0] if len(seq) & 1 else []), np.uint8).reshape(-1, 2)
        b = (nibs[:, 0] << 4 | nibs[:, 1]).astype(np.uint8)
        seq_chunks.append(b)
        seq_pos += b.size
    cigar_arr = np.array(cig, np.uint32)
    seqs = np.concatenate(seq_chunks + [np.zeros(16, np.uint8)])
    return contigs, reg, drafts, alns, cigar_arr, seqs
