#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY (developer tool).  Packs a captured run of the reference CLI (tools/capture/capture.sh)
into the arrays hypo_gpu_extract_arms / hypo_gpu_polish_alignments take (include/hypo_b200.h): region tables and
drafts from the reference's own dump, alignments from the SAM file the run read, plus the reference's polished
FASTA and its Monitor times for the phases the fused call replaces.

  python tools/capture/make_n3_input.py <dataset dir> <k> <out.npz>
"""
import json
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hypo_b200.batch import pack4  # noqa: E402
from hypo_b200.native import ALN_DTYPE, CONTIG_DTYPE, REGION_DTYPE, REGION_TYPES  # noqa: E402
from oracle import arms_oracle as ao  # noqa: E402

_NIB = np.full(256, 15, np.uint8)
for c, v in (("A", 1), ("C", 2), ("G", 4), ("T", 8)):
    _NIB[ord(c)] = v
_OP = {c: i for i, c in enumerate("MIDNSHP=X")}


def main():
    d, k, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    names = [l[1:].strip() for l in open(os.path.join(d, "draft.fa")) if l.startswith(">")]
    contigs = np.zeros(len(names), CONTIG_DTYPE)
    regs, drafts, polished = [], [], []
    pol = {}
    cur = None
    for l in open(os.path.join(d, "polished.fa")):
        if l.startswith(">"):
            cur = l[1:].strip()
        else:
            pol[cur] = pol.get(cur, "") + l.strip()
    dpos = 0
    for ci, nm in enumerate(names):
        regions, clen, _ = ao.read_regions(open(os.path.join(d, "aux", f"inspect_{nm}.txt")), k)
        r = np.zeros(len(regions), REGION_DTYPE)
        for i, g in enumerate(regions):
            r[i] = (g.key0, g.key1, g.beg, REGION_TYPES[g.type])
        contigs[ci] = (sum(len(x) for x in regs), dpos, len(regions), clen)
        regs.append(r)
        p = pack4("".join(g.text for g in regions))
        drafts.append(p)
        dpos += p.size
        polished.append(pol[nm])
    cid = {nm: i for i, nm in enumerate(names)}
    recs = ao.read_sam(os.path.join(d, "sr.sam"))
    alns = np.zeros(len(recs), ALN_DTYPE)
    cig, seqs, spos = [], [], 0
    for i, (c, pos, cigar, seq) in enumerate(recs):
        ops = ao.CIGAR_RE.findall(cigar)
        alns[i] = (len(cig), spos, cid[c], pos, len(ops), len(seq))
        cig.extend((int(n) << 4) | _OP[op] for n, op in ops)
        nb = _NIB[np.frombuffer(seq.encode(), np.uint8)]
        if nb.size & 1:
            nb = np.append(nb, 0)
        b = (nb[0::2] << 4 | nb[1::2]).astype(np.uint8)
        seqs.append(b)
        spos += b.size
    # support-counting tables and the reference's counters (hypo_dump2), if that run was made
    n2 = {}
    if all(os.path.exists(os.path.join(d, "aux", f"kmer_support_{nm}.txt")) for nm in names):
        from oracle import support_oracle as so
        kfirst, spos, kid, kcov, ksup = [0], [], [], [], []
        bfirst, even, bounds, rfirst, mpos, mval, mcov, msup = [0], [], [], [], [], [], [], []
        for nm in names:
            a, b, c, e = so.read_kmer_dump(open(os.path.join(d, "aux", f"kmer_support_{nm}.txt")))
            spos += a; kid += b; kcov += c; ksup += e
            kfirst.append(len(spos))
            ev, bd, minfo, mc, ms = so.read_minimiser_dump(open(os.path.join(d, "aux", f"minimiser_support_{nm}.txt")))
            even.append(1 if ev else 0)
            for i in range(len(bd)):
                rfirst.append(len(mpos))
                is_win = (i % 2 == 0) if ev else (i % 2 == 1)
                m = i // 2 if ev else (i - 1) // 2
                if is_win and 0 <= m < len(minfo):
                    pp = bd[i]
                    for j, (rel, v) in enumerate(minfo[m]):
                        pp += rel
                        mpos.append(pp); mval.append(v); mcov.append(mc[m][j]); msup.append(ms[m][j])
            bounds += bd
            bfirst.append(len(bounds))
        rfirst.append(len(mpos))
        n2 = dict(kfirst=np.array(kfirst, np.uint64), spos=np.array(spos, np.uint32), kid=np.array(kid, np.uint64),
                  kcov=np.array(kcov, np.uint32), ksup=np.array(ksup, np.uint32), bfirst=np.array(bfirst, np.uint64),
                  even=np.array(even, np.uint8), bounds=np.array(bounds, np.uint32), rfirst=np.array(rfirst, np.uint64),
                  mpos=np.array(mpos, np.uint32), mval=np.array(mval, np.uint32), mcov=np.array(mcov, np.uint32),
                  msup=np.array(msup, np.uint32))
    log = open(os.path.join(d, "hypo.log")).read()
    times = {m.group(1).strip(): float(m.group(2)) for m in re.finditer(r"\[Hypo:Hypo\]: ([^)]*?)\. \): TIME= ([0-9.e+-]+)", log)}
    meta = {"k": k, "reference_cli_threads": 8, "reference_phase_s": times,
            "replaced_phases": ["Short arms computing", "Short arms filling", "POA of windows", "Writing results"]}
    np.savez_compressed(out, contigs=contigs, regions=np.concatenate(regs), drafts=np.concatenate(drafts + [np.zeros(16, np.uint8)]),
                        alns=alns, cigar=np.array(cig, np.uint32), seqs=np.concatenate(seqs + [np.zeros(16, np.uint8)]),
                        polished=np.array(polished), meta=np.array(json.dumps(meta)), **n2)
    print(out, os.path.getsize(out) / 1e6, "MB;", len(recs), "alignments;", meta["reference_phase_s"])


if __name__ == "__main__":
    main()
