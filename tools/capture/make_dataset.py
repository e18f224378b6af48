#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY (developer tool).  Synthetic input for the reference CLI:

  <out>/draft.fa              draft assembly = truth genome with ONT-like errors (sub / ins / del,
                              homopolymer length errors), one or more contigs
  <out>/sr.sam                short reads (150 bp) sampled from the TRUTH, aligned to the DRAFT with the
                              CIGAR the two edit scripts imply (no aligner needed; htslib reads text SAM)
  <out>/lr.sam   (--long X)   long reads (ONT-like, indel-rich) at X-fold coverage, same way, with NM tags
  <out>/aux/solid_kmers.bvsd  the solid (unique, non-homopolymer-terminal) k-mers of the truth, both
                              strands, in sdsl's bit_vector file format (what suk would store after KMC;
                              reference external/suk/src/SolidKmers.cpp:47-58,160-190)
  <out>/aux/stage.txt         "a b c 1": the CLI resumes behind the KMC stage (reference src/main.cpp:327-337)
  <out>/truth.fa              the genome the reads come from (to count residual edits)

Everything is seeded.  Pure numpy / Python; a 1 Mb genome at 50x takes about a minute."""
import argparse
import math
import os
import sys

import numpy as np

B = np.frombuffer(b"ACGT", np.uint8)
COMP = np.array([3, 2, 1, 0], np.uint8)


def genome(rng, n, repeat_frac, hp_every):
    g = rng.integers(0, 4, size=n, dtype=np.uint8)
    # homopolymer runs (where ONT drafts are weakest)
    if hp_every:
        for p in range(hp_every, n - 20, hp_every):
            p += int(rng.integers(0, hp_every // 2))
            if p + 12 < n:
                g[p:p + int(rng.integers(4, 12))] = g[p]
    # 2-4 copy repeats, 1 % diverged
    done = 0
    while done < repeat_frac * n:
        ln = int(rng.integers(300, 3000))
        src = int(rng.integers(0, n - ln))
        for _ in range(int(rng.integers(1, 4))):
            dst = int(rng.integers(0, n - ln))
            seg = g[src:src + ln].copy()
            mut = rng.random(ln) < 0.01
            seg[mut] = (seg[mut] + rng.integers(1, 4, size=int(mut.sum()), dtype=np.uint8)) % 4
            g[dst:dst + ln] = seg
            done += ln
    return g


def edit_script(rng, truth, sub, ins, dele, hp_extra):
    """Per truth position: keep (0 = truth base absent from the product), the product's base, and how many
    bases the product inserts behind it (with those bases)."""
    n = truth.size
    r = rng.random(n)
    keep = r >= dele
    base = truth.copy()
    s = (r >= dele) & (r < dele + sub)
    base[s] = (base[s] + rng.integers(1, 4, size=int(s.sum()), dtype=np.uint8)) % 4
    n_ins = (rng.random(n) < ins).astype(np.int64)
    if hp_extra > 0:   # homopolymer length errors: a run loses or gains one base more often
        run = np.zeros(n, bool)
        run[1:] = truth[1:] == truth[:-1]
        run[2:] &= truth[2:] == truth[:-2]
        h = run & (rng.random(n) < hp_extra)
        drop = h & (rng.random(n) < 0.5)
        keep &= ~drop
        n_ins[h & ~drop] += 1
    ins_base = rng.integers(0, 4, size=n, dtype=np.uint8)
    return keep, base, n_ins, ins_base


def apply_script(keep, base, n_ins, ins_base):
    n = keep.size
    cnt = keep.astype(np.int64) + n_ins
    start = np.concatenate([[0], np.cumsum(cnt)])
    out = np.zeros(int(start[-1]), np.uint8)
    out[start[:-1][keep]] = base[keep]
    pos = start[:-1] + keep
    for k in range(int(n_ins.max()) if n else 0):
        m = n_ins > k
        out[pos[m] + k] = ins_base[m] if k == 0 else (ins_base[m] + k) % 4
    return out, start


def cigar_and_seq(a, b, dk, dn, rk, rb, rn, ri):
    """Alignment of a read (its own edit script rk / rb / rn / ri over truth[a:b)) against the draft (edit
    script dk / dn): column by column through the truth coordinates."""
    ops = []
    seq = bytearray()
    nm = 0
    last, run = "", 0

    def push(op, k=1):
        nonlocal last, run
        if k <= 0:
            return
        if op == last:
            run += k
        else:
            if run:
                ops.append(f"{run}{last}")
            last, run = op, k
    for i in range(a, b):
        r_has, d_has = rk[i], dk[i]
        if r_has:
            seq.append(rb[i])
        if r_has and d_has:
            push("M")
        elif r_has:
            push("I"); nm += 1
        elif d_has:
            push("D"); nm += 1
        if i + 1 < b:
            x, y = rn[i], dn[i]
            for k in range(x):
                seq.append((ri[i] + k) % 4)
            m = min(x, y)
            push("M", m)
            push("I", x - m)
            push("D", y - m)
            nm += abs(x - y)
    if run:
        ops.append(f"{run}{last}")
    return "".join(ops), seq, nm


def write_reads(path, rng, truth, names, bounds, dscripts, dstarts, cov, rlen_fn, sub, ins, dele, with_nm, header,
                gaps=None):
    """gaps: per contig, a boolean mask over truth positions no read may touch (coverage drop-outs: where the
    short reads leave a hole the reference falls back to LONG windows built from long reads)."""
    n_reads = 0
    with open(path, "w") as f:
        f.write(header)
        for c, (lo, hi) in enumerate(bounds):
            t = truth[lo:hi]
            n = t.size
            dk, _, dn, _ = dscripts[c]
            dkl, dnl = dk.tolist(), dn.tolist()
            dstart = dstarts[c]
            tot = 0
            starts = []
            while tot < cov * n:
                ln = rlen_fn()
                a = int(rng.integers(0, max(1, n - ln)))
                starts.append((a, min(n, a + ln)))
                tot += ln
            starts.sort()
            for a, b in starts:
                if gaps is not None and gaps[c][a:b].any():
                    continue
                # the alignment starts and ends on a column present in the read and in the draft
                rk, rb, rn, ri = edit_script(rng, t[a:b], sub, ins, dele, 0.0)
                rkl = rk.tolist()
                while a < b and not (dkl[a] and rkl[0]):
                    a += 1; rkl = rkl[1:]; rk = rk[1:]; rb = rb[1:]; rn = rn[1:]; ri = ri[1:]
                while b > a and not (dkl[b - 1] and rkl[-1]):
                    b -= 1; rkl = rkl[:-1]; rk = rk[:-1]; rb = rb[:-1]; rn = rn[:-1]; ri = ri[:-1]
                if b - a < 30:
                    continue
                off = a
                cig, seq, nm = cigar_and_seq(0, b - a, [dkl[off + i] for i in range(b - a)],
                                             [dnl[off + i] for i in range(b - a)], rkl, rb.tolist(), rn.tolist(), ri.tolist())
                s = B[np.frombuffer(bytes(seq), np.uint8)].tobytes().decode()
                flag = 16 if rng.random() < 0.5 else 0
                tags = f"\tNM:i:{nm}" if with_nm else ""
                f.write(f"r{n_reads}\t{flag}\t{names[c]}\t{int(dstart[a]) + 1}\t60\t{cig}\t*\t0\t0\t{s}\t{'I' * len(s)}{tags}\n")
                n_reads += 1
    return n_reads


def solid_kmers(truth_contigs, k, path):
    size = 1 << (2 * k)
    cnt = np.zeros(size, np.uint32)
    mask = size - 1
    codes = []
    for t in truth_contigs:
        t64 = t.astype(np.uint64)
        n = t.size - k + 1
        fw = np.zeros(n, np.uint64)
        rc = np.zeros(n, np.uint64)
        for j in range(k):
            fw = (fw << np.uint64(2)) | t64[j:j + n]
            rc |= (np.uint64(3) ^ t64[j:j + n]) << np.uint64(2 * j)
        codes.append((fw, rc, t))
        cnt += np.bincount(fw.astype(np.int64), minlength=size).astype(np.uint32)
    bv = np.zeros(size, bool)
    n_solid = 0
    for fw, rc, t in codes:
        n = fw.size
        total = cnt[fw.astype(np.int64)] + cnt[rc.astype(np.int64)]
        hp = (t[0:n] == t[1:n + 1]) | (t[k - 1:k - 1 + n] == t[k - 2:k - 2 + n])
        ok = (total == 1) & ~hp
        bv[fw[ok].astype(np.int64)] = True
        bv[rc[ok].astype(np.int64)] = True
        n_solid += int(ok.sum())
    words = np.packbits(bv.reshape(-1, 64)[:, ::-1], axis=1).view(">u8").astype("<u8").reshape(-1)
    with open(path, "wb") as f:
        f.write(np.uint64(size).tobytes())
        f.write(words.tobytes())
    return n_solid, mask


def kmer_len(size_str):
    """reference src/main.cpp:490-528"""
    unit = size_str[-1].upper()
    power = {"K": 10, "M": 20, "G": 30, "T": 40}.get(unit, 0)
    val = float(size_str[:-1] if power else size_str)
    k = power + math.ceil(math.log2(val))
    k = k // 2
    return k + 1 if k % 2 == 0 else k


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("out")
    ap.add_argument("--size", type=int, default=1_000_000, help="total genome size in bp")
    ap.add_argument("--contigs", type=int, default=1)
    ap.add_argument("--size-ref", default="", help="value passed to the CLI's -s (default: derived from --size)")
    ap.add_argument("--cov", type=float, default=50.0)
    ap.add_argument("--long", type=float, default=0.0, help="long-read coverage (0 = none)")
    ap.add_argument("--draft-err", type=float, default=0.025, help="total draft error rate (sub + ins + del)")
    ap.add_argument("--repeats", type=float, default=0.05)
    ap.add_argument("--sr-gaps", type=float, default=0.0,
                    help="fraction of the genome without short-read coverage (drop-outs of 200-1500 bp)")
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    os.makedirs(os.path.join(a.out, "aux"), exist_ok=True)
    truth = genome(rng, a.size, a.repeats, 700)
    cuts = np.linspace(0, a.size, a.contigs + 1).astype(int)
    bounds = [(int(cuts[i]), int(cuts[i + 1])) for i in range(a.contigs)]
    names = [f"ctg{i + 1}" for i in range(a.contigs)]
    e = a.draft_err
    dscripts, dstarts, drafts = [], [], []
    for lo, hi in bounds:
        sc = edit_script(rng, truth[lo:hi], 0.3 * e, 0.3 * e, 0.4 * e, 0.15)
        d, start = apply_script(*sc)
        dscripts.append(sc); dstarts.append(start); drafts.append(d)
    with open(os.path.join(a.out, "draft.fa"), "w") as f, open(os.path.join(a.out, "truth.fa"), "w") as ft:
        for (lo, hi), nm, d in zip(bounds, names, drafts):
            f.write(f">{nm}\n{B[d].tobytes().decode()}\n")
            ft.write(f">{nm}\n{B[truth[lo:hi]].tobytes().decode()}\n")
    header = "@HD\tVN:1.6\tSO:coordinate\n" + "".join(f"@SQ\tSN:{nm}\tLN:{d.size}\n" for nm, d in zip(names, drafts))
    gaps = None
    if a.sr_gaps > 0:
        gaps = []
        for lo, hi in bounds:
            m = np.zeros(hi - lo, bool)
            while m.mean() < a.sr_gaps:
                ln = int(rng.integers(200, 1500))
                p = int(rng.integers(0, max(1, hi - lo - ln)))
                m[p:p + ln] = True
            gaps.append(m)
    n_sr = write_reads(os.path.join(a.out, "sr.sam"), rng, truth, names, bounds, dscripts, dstarts, a.cov,
                       lambda: 150, 0.002, 0.0002, 0.0002, False, header, gaps)
    n_lr = 0
    if a.long > 0:
        n_lr = write_reads(os.path.join(a.out, "lr.sam"), rng, truth, names, bounds, dscripts, dstarts, a.long,
                           lambda: int(rng.integers(3000, 12000)), 0.02, 0.03, 0.03, True, header)
    size_ref = a.size_ref or (f"{max(1, round(a.size / 1e6))}m" if a.size >= 500_000 else f"{max(1, round(a.size / 1e3))}k")
    k = kmer_len(size_ref)
    n_solid, _ = solid_kmers([truth[lo:hi] for lo, hi in bounds], k, os.path.join(a.out, "aux", "solid_kmers.bvsd"))
    with open(os.path.join(a.out, "aux", "stage.txt"), "w") as f:
        f.write("a b c 1\n")
    with open(os.path.join(a.out, "cli_args.txt"), "w") as f:
        f.write(f"-s {size_ref} -c {int(a.cov)}\n")
    print(f"genome {a.size} bp in {a.contigs} contig(s), draft {sum(d.size for d in drafts)} bp, {n_sr} short reads, "
          f"{n_lr} long reads, k = {k}, {n_solid} solid k-mers; CLI: -s {size_ref} -c {int(a.cov)}", file=sys.stderr)


if __name__ == "__main__":
    main()
