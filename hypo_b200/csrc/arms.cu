// arms.cu — arm extraction on the device: from the alignments of a contig batch and the
// contigs' region tables to the packed window batch of the POA path, without the host ever building
// Window / PackedSeq objects or packing them.
//
// Reference code restated here (paths relative to the reference root):
//   Alignment::initialise_pos / copy_data   src/Alignment.cpp:513-576   aln_scan_kernel
//   Alignment::find_short_arms              src/Alignment.cpp:222-259   aln_scan_kernel (range), aln_arms_kernel
//   Alignment::find_bp                      src/Alignment.cpp:321-404   find_bp
//   Alignment::prepare_short_arm            src/Alignment.cpp:406-509   prepare_short_arm
//   Alignment::add_arms                     src/Alignment.cpp:299-318   the sort by (window, kind, alignment)
//   Contig::fill_short_windows (pruning)    src/Contig.cpp:262-289      window_kernel
//   PackedSeq sub-range copies              src/PackedSeq.cpp:155-194   fill_window_kernel / fill_arm_kernel
//
// Pipeline (all HBM-bound integer / byte work; one thread per alignment, window or arm):
//   1. aln_scan_kernel   clipping, reference span, validity, first / last region touched -> slots per alignment
//   2. exclusive scan of the slots
//   3. aln_arms_kernel   CIGAR walk (query position of every region start), anchor validation, one record per
//                        slot: key = window region << 34 | kind << 32 | alignment, query range
//   4. radix sort of the records by key: the arms of a window become contiguous, kinds in container order
//                        (internal, prefix, suffix, empty), alignment order inside a kind - exactly the order
//                        add_arms produces by walking the alignments
//   5. window_kernel     per window: counters, longest prefix / suffix, the reference's pruning rules
//   6. scans             window / arm / byte positions of what is kept
//   7. fill kernels      descriptors + draft sub-range (4-bit) + arm sub-ranges (BAM 4-bit -> 2-bit) into the slab
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/hypo_b200.h"

extern "C" int hypo_internal_fail(int code, const char* msg);          // api.cu: sets the thread's error text
extern "C" int hypo_internal_primary_device(void);                     // api.cu: ordinal of the first driven device, -1
extern "C" int hypo_internal_polish_device(const HypoWindowDesc* d_win, uint64_t n_win, const HypoArmDesc* d_arms,
                                           uint64_t n_arms, const uint8_t* d_packed, uint64_t packed_bytes,
                                           const HypoRegionDesc* d_regions,
                                           const uint32_t* d_reg_contig, uint64_t n_regions,
                                           const uint64_t* contig_first_region, uint64_t n_contigs,
                                           const uint8_t* d_drafts, const uint64_t* d_draft_off, char* out,
                                           uint64_t out_cap, uint64_t* out_off, void* stream);   // api.cu

namespace {

constexpr uint32_t kMinimizerK = 10;             // reference src/main.cpp:86
// reference src/main.cpp:88 (Arms_settings), include/globalDefs.hpp:146-156
constexpr uint32_t kMinShortNum = 3, kMinInternal1 = 20, kMinInternal2 = 5, kMinContrib = 10, kShortArmCoef = 10;
constexpr double kMinInternalContrib = 0.4;
enum ArmKind : uint32_t { kInternal = 0, kPrefix = 1, kSuffix = 2, kEmpty = 3 };
constexpr uint64_t kNoArm = ~0ull;

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        const size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
};

struct AlnInfo {
    uint32_t rb, re;        // reference span [rb, re)
    uint32_t qab, qlen;     // aligned part of the query: bases [qab, qab + qlen) of the read
    uint32_t b_ind, e_ind;  // regions of the contig the span touches: [b_ind, e_ind)
    uint32_t flags;         // bit 0 valid, bit 1 a region starts at rb, bit 2 a region starts at re
    uint32_t pad;
};

struct ArmRec {
    uint32_t qb, qe;        // query range of the arm inside the aligned part
};

struct WinTmp {
    uint64_t region;        // batch-wide region index
    uint32_t first;         // first sorted record of the window
    uint32_t n_int, n_pre, n_suf, n_empty;
    uint32_t flags;         // bit 0 dropped, bit 1 prefix / suffix arms cleared
};

struct Ctl {
    unsigned long long n_valid;   // records that carry an arm (or an empty arm)
    uint32_t bad;                 // alignments outside their contig
    uint32_t pad;
};

__device__ __forceinline__ uint32_t cigar_type(uint32_t op) { return (0x3C1A7u >> (op << 1)) & 3u; }   // htslib BAM_CIGAR_TYPE
__device__ __forceinline__ uint32_t nib(const uint8_t* seq, uint32_t i) { return (seq[i >> 1] >> ((i & 1) ? 0 : 4)) & 15u; }
// BAM nibble -> 2-bit code (A1 C2 G4 T8), 4 for anything else
__device__ __forceinline__ uint32_t code_of(uint32_t n) { return n == 1 ? 0u : n == 2 ? 1u : n == 4 ? 2u : n == 8 ? 3u : 4u; }

struct ContigView {
    const HypoRegionRec* reg;   // the contig's regions
    uint32_t n, len;
    __device__ uint32_t start(uint32_t i) const { return i < n ? reg[i].start : len; }   // (dummy region at the end)
    __device__ bool is_sr(uint32_t i) const { return i >= n || reg[i].type <= HYPO_REG_MSR; }
};

// number of region starts < x
__device__ uint32_t starts_below(const ContigView& c, uint32_t x) {
    uint32_t lo = 0, hi = c.n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (c.reg[mid].start < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// ---- 1. per alignment: clipping, span, validity, regions touched (initialise_pos, copy_data, find_short_arms) ----
__global__ void aln_scan_kernel(const HypoContigDesc* __restrict__ contigs, uint64_t n_contigs,
                                const HypoRegionRec* __restrict__ regions, const HypoAlnDesc* __restrict__ alns,
                                uint64_t n_alns, const uint32_t* __restrict__ cigar, const uint8_t* __restrict__ seqs,
                                AlnInfo* __restrict__ info, uint64_t* __restrict__ slots, Ctl* __restrict__ ctl) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n_alns) return;
    const HypoAlnDesc a = alns[i];
    AlnInfo o;
    memset(&o, 0, sizeof(o));
    slots[i] = 0;
    if (a.contig >= n_contigs) { atomicAdd(&ctl->bad, 1u); info[i] = o; return; }
    const HypoContigDesc cd = contigs[a.contig];
    const uint32_t* cg = cigar + a.cigar_off;
    uint32_t qab = 0, qp = 0, rp = a.pos, clip_end = 0;
    bool clip_before = true;
    for (uint32_t j = 0; j < a.n_cigar; ++j) {
        const uint32_t op = cg[j] & 15u, len = cg[j] >> 4;
        if (clip_before) {
            if (op == 4) qab += len;            // BAM_CSOFT_CLIP
            else if (op != 5) clip_before = false;   // BAM_CHARD_CLIP
        }
        const uint32_t t = cigar_type(op);
        if (t == 3) { rp += len; qp += len; }
        else if (t & 2) rp += len;
        else if (t & 1) { if (!clip_before && op == 4) clip_end += len; qp += len; }
    }
    const uint32_t qae = qp - clip_end;
    o.rb = a.pos; o.re = rp; o.qab = qab; o.qlen = qae - qab;
    if (o.rb >= cd.len || o.re > cd.len || qae > a.l_qseq || qae < qab) { atomicAdd(&ctl->bad, 1u); info[i] = o; return; }
    bool valid = true;
    const uint8_t* sq = seqs + a.seq_off;
    for (uint32_t q = qab; q < qae && valid; ++q) valid = code_of(nib(sq, q)) < 4;
    ContigView c{regions + cd.first_region, cd.n_regions, cd.len};
    uint32_t b = starts_below(c, o.rb);
    const bool start_at_rb = b < c.n && c.reg[b].start == o.rb;
    if (!start_at_rb) --b;   // (region 0 starts at 0, so b >= 1 here)
    const uint32_t e = starts_below(c, o.re);
    const bool start_at_re = o.re == c.len || (e < c.n && c.reg[e].start == o.re);
    o.b_ind = b; o.e_ind = e;
    o.flags = (valid ? 1u : 0u) | (start_at_rb ? 2u : 0u) | (start_at_re ? 4u : 0u);
    info[i] = o;
    slots[i] = (valid && e - b > 1) ? (uint64_t)(e - b) : 0;
}

// ---- find_bp (src/Alignment.cpp:321-404): query position at which every region after the first begins -----------
__device__ uint32_t find_bp(const ContigView& c, const AlnInfo& o, const uint32_t* __restrict__ cg, uint32_t n_cigar,
                            uint32_t* __restrict__ bp, uint32_t cap) {
    uint32_t n = 0;
    uint32_t cur_ref = o.rb, idx = o.b_ind + 1, next_ref = c.start(idx), qpos = 0;
    bool corner = false;
    for (uint32_t j = 0; j < n_cigar; ++j) {
        const uint32_t op = cg[j] & 15u;
        uint32_t len = cg[j] >> 4;
        if (op == 4 || op == 5) continue;
        const uint32_t t = cigar_type(op);
        if (t == 3) {
            if (corner) { if (n < cap) bp[n] = qpos; ++n; corner = false; ++idx; next_ref = c.start(idx); }
            while (cur_ref + len >= next_ref && !corner) {
                const uint32_t d = next_ref - cur_ref;
                cur_ref = next_ref; qpos += d; len -= d;
                if (len > 0) { if (n < cap) bp[n] = qpos; ++n; ++idx; next_ref = c.start(idx); }
                else corner = true;
            }
            if (len > 0) { cur_ref += len; qpos += len; }
        } else if (t & 2) {
            if (corner) { if (n < cap) bp[n] = qpos; ++n; corner = false; ++idx; next_ref = c.start(idx); }
            while (cur_ref + len >= next_ref && !corner) {
                const uint32_t d = next_ref - cur_ref;
                cur_ref = next_ref; len -= d;
                if (len > 0) { if (n < cap) bp[n] = qpos; ++n; ++idx; next_ref = c.start(idx); }
                else corner = true;
            }
            if (len > 0) cur_ref += len;
        } else if (t & 1) {
            if (corner) {
                if (n < cap) bp[n] = c.is_sr(idx - 1) ? qpos : qpos + len;
                ++n; ++idx; next_ref = c.start(idx); corner = false;
            }
            qpos += len;
        }
        if (idx == o.e_ind) break;
    }
    return n;
}

struct ReadView {
    const uint8_t* seq;   // bam_get_seq()
    uint32_t qab, len;    // aligned part
    __device__ uint32_t at(uint32_t i) const { return code_of(nib(seq, qab + i)); }
};

// PackedSeq::find_kmer (src/PackedSeq.cpp:264-330): the k-mer lies wholly inside [left, right)
__device__ bool find_kmer(const ReadView& r, uint64_t target, uint32_t k, uint32_t left, uint32_t right, bool first,
                          uint32_t* at) {
    const uint64_t mask = (k >= 32) ? ~0ull : ((1ull << (2 * k)) - 1);
    uint64_t v = 0;
    uint32_t run = 0;
    bool found = false;
    for (uint32_t i = left; i < right; ++i) {
        v = ((v << 2) | r.at(i)) & mask;
        if (run < k) ++run;
        if (run == k && v == target) {
            *at = i + 1 - k;
            found = true;
            if (first) break;
        }
    }
    return found;
}
__device__ bool check_kmer(const ReadView& r, uint64_t target, uint32_t k, uint32_t at) {
    uint32_t dummy;
    return at + k <= r.len && find_kmer(r, target, k, at, at + k, true, &dummy);
}

// prepare_short_arm (src/Alignment.cpp:406-509)
__device__ bool prepare_short_arm(const ContigView& c, const ReadView& r, uint32_t k, uint32_t windex, uint32_t qb,
                                  uint32_t qe, uint32_t armtype, uint32_t* out_b, uint32_t* out_e) {
    const uint32_t mk = kMinimizerK;
    if ((uint64_t)(c.start(windex + 1) - c.start(windex)) > (uint64_t)kShortArmCoef * (qe - qb)) return false;
    const uint32_t wt = c.reg[windex].type;
    bool valid = true;
    uint32_t q_beg = qb, q_end = qe, at = 0;
    const uint32_t qae = r.len;
    if ((wt == HYPO_REG_SWS || wt == HYPO_REG_SW || wt == HYPO_REG_SWM) && armtype != kSuffix) {
        if (q_beg < k) valid = false;
        else {
            const uint64_t anchor = c.reg[windex - 1].key1;   // last k-mer of the preceding strong region
            if (!check_kmer(r, anchor, k, q_beg - k)) {
                const uint32_t s0 = q_beg < 2 * k ? 0 : q_beg - 2 * k;
                const uint32_t s1 = q_end < q_beg + k ? q_end : q_beg + k;
                if (find_kmer(r, anchor, k, s0, s1, false, &at)) q_beg = at + k; else valid = false;
            }
        }
    }
    if ((wt == HYPO_REG_SWS || wt == HYPO_REG_WS || wt == HYPO_REG_MWS) && armtype != kPrefix) {
        if (q_end + k > qae) valid = false;
        else {
            const uint64_t anchor = c.reg[windex + 1].key0;   // first k-mer of the following strong region
            if (!check_kmer(r, anchor, k, q_end)) {
                const uint32_t s0 = q_end < q_beg + k ? q_beg : q_end - k;
                const uint32_t s1 = min(qae, q_end + 2 * k);
                if (find_kmer(r, anchor, k, s0, s1, true, &at)) q_end = at; else valid = false;
            }
        }
    }
    if ((wt == HYPO_REG_MWM || wt == HYPO_REG_MW || wt == HYPO_REG_MWS) && armtype != kSuffix) {
        if (q_beg < mk) valid = false;
        else {
            const uint64_t mini = c.reg[windex - 1].key0;
            if (!check_kmer(r, mini, mk, q_beg - mk)) {
                const uint32_t s0 = q_beg < 3 * mk ? 0 : q_beg - 3 * mk;
                const uint32_t s1 = q_end < q_beg + 2 * mk ? q_end : q_beg + 2 * mk;
                if (find_kmer(r, mini, mk, s0, s1, false, &at)) q_beg = at + mk; else valid = false;
            }
        }
    }
    if ((wt == HYPO_REG_MWM || wt == HYPO_REG_WM || wt == HYPO_REG_SWM) && armtype != kPrefix) {
        if (q_end + mk > qae) valid = false;
        else {
            const uint64_t mini = c.reg[windex + 1].key0;
            if (!check_kmer(r, mini, mk, q_end)) {
                const uint32_t s0 = q_end < q_beg + 2 * mk ? q_beg : q_end - 2 * mk;
                const uint32_t s1 = min(qae, q_end + 3 * mk);
                if (find_kmer(r, mini, mk, s0, s1, true, &at)) q_end = at; else valid = false;
            }
        }
    }
    if (valid && q_beg < q_end) { *out_b = q_beg; *out_e = q_end; return true; }
    return false;
}

// ---- 3. per alignment: one record per region touched (find_short_arms) ------------------------------------------
__global__ void aln_arms_kernel(const HypoContigDesc* __restrict__ contigs, const HypoRegionRec* __restrict__ regions,
                                const HypoAlnDesc* __restrict__ alns, uint64_t n_alns, const uint32_t* __restrict__ cigar,
                                const uint8_t* __restrict__ seqs, const AlnInfo* __restrict__ info,
                                const uint64_t* __restrict__ slot_off, uint32_t k, uint32_t* __restrict__ bp_buf,
                                uint64_t* __restrict__ keys, ArmRec* __restrict__ recs, uint32_t* __restrict__ vals,
                                Ctl* __restrict__ ctl) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n_alns) return;
    const AlnInfo o = info[i];
    const uint32_t n_slots = (uint32_t)(slot_off[i + 1] - slot_off[i]);
    if (n_slots == 0) return;
    const HypoAlnDesc a = alns[i];
    const HypoContigDesc cd = contigs[a.contig];
    const ContigView c{regions + cd.first_region, cd.n_regions, cd.len};
    const ReadView r{seqs + a.seq_off, o.qab, o.qlen};
    const uint64_t s0 = slot_off[i];
    uint32_t* bp = bp_buf + s0;
    const uint32_t n_bp_want = n_slots - 1;
    const uint32_t n_bp = find_bp(c, o, cigar + a.cigar_off, a.n_cigar, bp, n_bp_want);
    for (uint32_t j = n_bp; j < n_bp_want; ++j) bp[j] = o.qlen;   // (cannot happen for a consistent CIGAR)
    uint32_t n_valid = 0;
    for (uint32_t j = 0; j < n_slots; ++j) {
        const uint32_t ind = o.b_ind + j;
        uint64_t key = kNoArm;
        ArmRec rec{0, 0};
        if (!c.is_sr(ind)) {
            uint32_t kind, qb, qe;
            bool empty = false;
            if (j == 0) { kind = (o.flags & 2u) ? kInternal : kSuffix; qb = 0; qe = bp[0]; }
            else if (j == n_slots - 1) { kind = (o.flags & 4u) ? kInternal : kPrefix; qb = bp[j - 1]; qe = o.qlen; }
            else { kind = kInternal; qb = bp[j - 1]; qe = bp[j]; empty = qb == qe; }
            uint32_t ob = 0, oe = 0;
            if (empty) {
                key = ((cd.first_region + ind) << 34) | ((uint64_t)kEmpty << 32) | (uint64_t)(uint32_t)i;
            } else if (qe > qb && prepare_short_arm(c, r, k, ind, qb, qe, kind, &ob, &oe)) {
                key = ((cd.first_region + ind) << 34) | ((uint64_t)kind << 32) | (uint64_t)(uint32_t)i;
                rec.qb = ob; rec.qe = oe;
            }
        }
        keys[s0 + j] = key;
        recs[s0 + j] = rec;
        vals[s0 + j] = (uint32_t)(s0 + j);
        n_valid += key != kNoArm;
    }
    if (n_valid) atomicAdd(&ctl->n_valid, (unsigned long long)n_valid);
}

// ---- 5. windows: segment heads, counters, pruning (Contig::fill_short_windows, src/Contig.cpp:262-289) ----------
__global__ void head_kernel(const uint64_t* __restrict__ keys, uint64_t n, uint64_t* __restrict__ head) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    head[i] = (i == 0 || (keys[i] >> 34) != (keys[i - 1] >> 34)) ? 1 : 0;
}

__global__ void window_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                              const ArmRec* __restrict__ recs, uint64_t n, const uint64_t* __restrict__ head,
                              const uint64_t* __restrict__ wid, const HypoRegionRec* __restrict__ regions,
                              const uint32_t* __restrict__ reg_end, WinTmp* __restrict__ wt,
                              uint64_t* __restrict__ keep_win, uint64_t* __restrict__ draft_bytes) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n || !head[i]) return;
    const uint64_t region = keys[i] >> 34;
    uint32_t cnt[4] = {0, 0, 0, 0}, maxpre = 0, maxsuf = 0;
    for (uint64_t j = i; j < n && (keys[j] >> 34) == region; ++j) {
        const uint32_t kind = (uint32_t)(keys[j] >> 32) & 3u;
        ++cnt[kind];
        const ArmRec r = recs[vals[j]];
        if (kind == kPrefix) maxpre = max(maxpre, r.qe - r.qb);
        if (kind == kSuffix) maxsuf = max(maxsuf, r.qe - r.qb);
    }
    const HypoRegionRec rg = regions[region];
    const uint32_t win_len = reg_end[region] - rg.start;
    const uint32_t internal = cnt[kInternal] + cnt[kEmpty];   // Window::get_num_internal counts the empties
    bool dropped = false, cleared = false;
    if (internal < kMinShortNum) {
        const bool covered = maxpre + maxsuf >= win_len;
        const bool enough = cnt[kPrefix] >= kMinShortNum && cnt[kSuffix] >= kMinShortNum;
        dropped = !(covered && enough);
    }
    if (!dropped) {
        const uint32_t contrib = internal + cnt[kPrefix] + cnt[kSuffix];
        const bool cond0 = internal > kMinInternal1;
        const bool cond1 = contrib >= kMinContrib && (double)internal >= floor(kMinInternalContrib * (double)contrib);
        const bool cond2 = (rg.type == HYPO_REG_SWS || rg.type == HYPO_REG_SW || rg.type == HYPO_REG_WS ||
                            rg.type == HYPO_REG_MWS || rg.type == HYPO_REG_SWM) && internal >= kMinInternal2;
        cleared = cond0 || cond1 || cond2;
    }
    WinTmp w;
    w.region = region; w.first = (uint32_t)i;
    w.n_int = cnt[kInternal]; w.n_pre = cleared ? 0 : cnt[kPrefix]; w.n_suf = cleared ? 0 : cnt[kSuffix];
    w.n_empty = cnt[kEmpty];
    w.flags = (dropped ? 1u : 0u) | (cleared ? 2u : 0u);
    const uint64_t raw = wid[i];
    wt[raw] = w;
    keep_win[raw] = dropped ? 0 : 1;
    draft_bytes[raw] = dropped ? 0 : (win_len + 1) / 2;
}

// per sorted record: is the arm kept, and how many bytes does it take
__global__ void armflag_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                               const ArmRec* __restrict__ recs, uint64_t n, const uint64_t* __restrict__ wid,
                               const WinTmp* __restrict__ wt, uint64_t* __restrict__ keep_arm,
                               uint64_t* __restrict__ arm_bytes) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t kind = (uint32_t)(keys[i] >> 32) & 3u;
    const WinTmp w = wt[wid[i]];
    const bool keep = !(w.flags & 1u) && kind != kEmpty && !((w.flags & 2u) && kind != kInternal);
    const ArmRec r = recs[vals[i]];
    keep_arm[i] = keep ? 1 : 0;
    arm_bytes[i] = keep ? (r.qe - r.qb + 3) / 4 : 0;
}

// ---- 7. fill: descriptors and bytes ------------------------------------------------------------------------------
// One thread per kept window: descriptor + its draft, bases [start, end) of the contig's PackedSeq<4> re-packed from
// the window's first base (PackedSeq<4>(ps, left, right), src/PackedSeq.cpp:155-194).
__global__ void fill_window_kernel(const WinTmp* __restrict__ wt, uint64_t n_raw, const uint64_t* __restrict__ win_idx,
                                   const uint64_t* __restrict__ arm_idx, const uint64_t* __restrict__ arm_byte,
                                   const uint64_t* __restrict__ draft_byte, const HypoRegionRec* __restrict__ regions,
                                   const uint32_t* __restrict__ reg_end, const uint32_t* __restrict__ reg_contig,
                                   const HypoContigDesc* __restrict__ contigs, const uint8_t* __restrict__ drafts,
                                   HypoWindowDesc* __restrict__ win, uint64_t* __restrict__ win_region,
                                   uint8_t* __restrict__ packed) {
    const uint64_t raw = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (raw >= n_raw) return;
    const WinTmp w = wt[raw];
    if (w.flags & 1u) return;
    const HypoRegionRec rg = regions[w.region];
    const uint32_t len = reg_end[w.region] - rg.start;
    HypoWindowDesc d;
    d.draft_off = arm_byte[w.first] + draft_byte[raw];
    d.first_arm = arm_idx[w.first];
    d.draft_len = len;
    d.n_internal = w.n_int; d.n_pre = w.n_pre; d.n_suf = w.n_suf; d.n_empty = w.n_empty;
    d.wtype = HYPO_WINDOW_SHORT;
    const uint64_t o = win_idx[raw];
    win[o] = d;
    win_region[o] = w.region;
    const uint8_t* src = drafts + contigs[reg_contig[w.region]].draft_off;
    uint8_t* dst = packed + d.draft_off;
    for (uint32_t i = 0; i < len; i += 2) {
        const uint32_t p0 = rg.start + i, p1 = p0 + 1;
        const uint32_t hi = (src[p0 >> 1] >> ((p0 & 1) ? 0 : 4)) & 15u;
        const uint32_t lo = (i + 1 < len) ? (src[p1 >> 1] >> ((p1 & 1) ? 0 : 4)) & 15u : 0u;
        dst[i >> 1] = (uint8_t)((hi << 4) | lo);
    }
}

// One thread per kept arm: descriptor + bases [qb, qe) of the aligned read, BAM nibbles -> 2 bits per base
// (PackedSeq<2>(ps, left, right), src/PackedSeq.cpp:155-194).
__global__ void fill_arm_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                const ArmRec* __restrict__ recs, uint64_t n, const uint64_t* __restrict__ keep_arm,
                                const uint64_t* __restrict__ arm_idx, const uint64_t* __restrict__ arm_byte,
                                const uint64_t* __restrict__ wid, const uint64_t* __restrict__ draft_byte,
                                const HypoAlnDesc* __restrict__ alns,
                                const AlnInfo* __restrict__ info, const uint8_t* __restrict__ seqs,
                                HypoArmDesc* __restrict__ arms, uint8_t* __restrict__ packed) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n || !keep_arm[i]) return;
    const ArmRec r = recs[vals[i]];
    const uint32_t aln = (uint32_t)keys[i];
    const uint64_t raw = wid[i];
    HypoArmDesc d;
    d.off = arm_byte[i] + draft_byte[raw + 1];   // the drafts of all kept windows up to and including this one
    d.len = r.qe - r.qb;
    d.reserved = 0;
    arms[arm_idx[i]] = d;
    const uint8_t* sq = seqs + alns[aln].seq_off;
    const uint32_t q0 = info[aln].qab + r.qb;
    uint8_t* dst = packed + d.off;
    for (uint32_t b = 0; b < d.len; b += 4) {
        uint32_t v = 0;
        for (uint32_t t = 0; t < 4; ++t) {
            const uint32_t c = (b + t < d.len) ? code_of(nib(sq, q0 + b + t)) & 3u : 0u;
            v |= c << (6 - 2 * t);
        }
        dst[b >> 2] = (uint8_t)v;
    }
}

// end of every region (start of the next one on the same contig, or the contig's length) and its contig
__global__ void region_end_kernel(const HypoContigDesc* __restrict__ contigs, uint64_t n_contigs,
                                  const HypoRegionRec* __restrict__ regions, uint32_t* __restrict__ reg_end,
                                  uint32_t* __restrict__ reg_contig) {
    const uint64_t c = blockIdx.y;
    if (c >= n_contigs) return;
    const HypoContigDesc cd = contigs[c];
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < cd.n_regions; r += gridDim.x * blockDim.x) {
        reg_end[cd.first_region + r] = r + 1 < cd.n_regions ? regions[cd.first_region + r + 1].start : cd.len;
        reg_contig[cd.first_region + r] = (uint32_t)c;
    }
}

// region descriptors for the stitcher: a polished window, or a copy of the draft
__global__ void stitch_regions_kernel(const HypoRegionRec* __restrict__ regions, const uint32_t* __restrict__ reg_end,
                                      uint64_t n_regions, HypoRegionDesc* __restrict__ out) {
    const uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (r >= n_regions) return;
    HypoRegionDesc d;
    d.src = regions[r].start;
    d.len = reg_end[r] - regions[r].start;
    d.window = HYPO_REGION_DRAFT;
    out[r] = d;
}
__global__ void stitch_windows_kernel(const uint64_t* __restrict__ win_region, uint64_t n_win, HypoRegionDesc* __restrict__ out) {
    const uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (w < n_win) out[win_region[w]].window = (uint32_t)w;
}

struct ArmsCtx {
    int device = -1;
    cudaStream_t stream = nullptr;
    DevBuf contigs, regions, drafts, alns, cigar, seqs, info, slots, slot_off, bp, keys, keys2, vals, vals2, recs, head, wid,
        wt, keep_win, draft_bytes, keep_arm, arm_bytes, win_idx, draft_byte, arm_idx, arm_byte, tmp, ctl, reg_end, reg_contig,
        win, win_region, arms, packed, sregions, doff;
    void* pinned = nullptr;
    uint64_t n_win = 0, n_arms = 0, n_bytes = 0;
} A;
std::mutex a_mu;

#define CUDA_TRY(x)                                                                                        \
    do {                                                                                                   \
        cudaError_t e_ = (x);                                                                              \
        if (e_ != cudaSuccess) {                                                                           \
            char b_[384];                                                                                  \
            snprintf(b_, sizeof(b_), "%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return hypo_internal_fail(HYPO_E_CUDA, b_);                                                    \
        }                                                                                                  \
    } while (0)

int scan(DevBuf& tmp, const uint64_t* in, uint64_t* out, uint64_t n, cudaStream_t s) {
    size_t bytes = 0;
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n, s));
    CUDA_TRY(tmp.reserve(bytes + 256));
    bytes = tmp.cap;
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, in, out, n, s));
    return HYPO_OK;
}

// Everything up to the device-resident batch (A.win / A.arms / A.packed / A.bound / A.win_region).
int extract_device(const HypoContigDesc* contigs, uint64_t n_contigs, const HypoRegionRec* regions, uint64_t n_regions,
                   const uint8_t* drafts, uint64_t draft_bytes, const HypoAlnDesc* alns, uint64_t n_alns,
                   const uint32_t* cigar, uint64_t n_cigar, const uint8_t* seqs, uint64_t seq_bytes, uint32_t k) {
    const int dev = hypo_internal_primary_device();
    if (dev < 0) return hypo_internal_fail(HYPO_E_NOT_INIT, "hypo_gpu_init has not been called");
    if (k == 0 || k > 31) return hypo_internal_fail(HYPO_E_ARG, "k must be 1..31");
    if (n_regions >= (1ull << 30) || n_alns >= (1ull << 32))
        return hypo_internal_fail(HYPO_E_ARG, "too many regions (2^30) or alignments (2^32) in one call");
    CUDA_TRY(cudaSetDevice(dev));
    if (A.device != dev) {
        if (!A.stream) CUDA_TRY(cudaStreamCreateWithFlags(&A.stream, cudaStreamNonBlocking));
        if (!A.pinned) CUDA_TRY(cudaHostAlloc(&A.pinned, 256, cudaHostAllocDefault));
        A.device = dev;
    }
    cudaStream_t s = A.stream;
    A.n_win = A.n_arms = A.n_bytes = 0;
    const int tb = 128;
    auto up = [&](DevBuf& b, const void* src, size_t bytes) -> cudaError_t {
        cudaError_t e = b.reserve(bytes + 16);
        if (e != cudaSuccess || bytes == 0) return e;
        return cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, s);
    };
    CUDA_TRY(up(A.contigs, contigs, sizeof(HypoContigDesc) * n_contigs));
    CUDA_TRY(up(A.regions, regions, sizeof(HypoRegionRec) * n_regions));
    CUDA_TRY(up(A.drafts, drafts, draft_bytes));
    CUDA_TRY(up(A.alns, alns, sizeof(HypoAlnDesc) * n_alns));
    CUDA_TRY(up(A.cigar, cigar, sizeof(uint32_t) * n_cigar));
    CUDA_TRY(up(A.seqs, seqs, seq_bytes));
    CUDA_TRY(A.ctl.reserve(sizeof(Ctl)));
    CUDA_TRY(cudaMemsetAsync(A.ctl.p, 0, sizeof(Ctl), s));
    CUDA_TRY(A.reg_end.reserve(sizeof(uint32_t) * (n_regions + 1)));
    CUDA_TRY(A.reg_contig.reserve(sizeof(uint32_t) * (n_regions + 1)));
    if (n_regions && n_contigs) {
        dim3 grid(64, (unsigned)n_contigs);
        region_end_kernel<<<grid, 256, 0, s>>>((const HypoContigDesc*)A.contigs.p, n_contigs, (const HypoRegionRec*)A.regions.p,
                                             (uint32_t*)A.reg_end.p, (uint32_t*)A.reg_contig.p);
    }
    if (n_alns == 0 || n_regions == 0) return HYPO_OK;
    CUDA_TRY(A.info.reserve(sizeof(AlnInfo) * n_alns));
    CUDA_TRY(A.slots.reserve(sizeof(uint64_t) * (n_alns + 1)));
    CUDA_TRY(A.slot_off.reserve(sizeof(uint64_t) * (n_alns + 1)));
    aln_scan_kernel<<<(unsigned)((n_alns + tb - 1) / tb), tb, 0, s>>>(
        (const HypoContigDesc*)A.contigs.p, n_contigs, (const HypoRegionRec*)A.regions.p, (const HypoAlnDesc*)A.alns.p, n_alns,
        (const uint32_t*)A.cigar.p, (const uint8_t*)A.seqs.p, (AlnInfo*)A.info.p, (uint64_t*)A.slots.p, (Ctl*)A.ctl.p);
    CUDA_TRY(cudaMemsetAsync((uint64_t*)A.slots.p + n_alns, 0, sizeof(uint64_t), s));
    if (int rc = scan(A.tmp, (const uint64_t*)A.slots.p, (uint64_t*)A.slot_off.p, n_alns + 1, s)) return rc;
    uint64_t* h = (uint64_t*)A.pinned;
    CUDA_TRY(cudaMemcpyAsync(h, (uint64_t*)A.slot_off.p + n_alns, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(h + 1, A.ctl.p, sizeof(Ctl), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    const uint64_t n_slots = h[0];
    if (((Ctl*)(h + 1))->bad)
        return hypo_internal_fail(HYPO_E_ARG, "alignment(s) outside their contig: the reference in the alignment file is not the draft "
                                              "(reference src/Alignment.cpp:32-36)");
    if (n_slots == 0) return HYPO_OK;
    if (n_slots >= (1ull << 32)) return hypo_internal_fail(HYPO_E_ARG, "too many region crossings in one call (2^32)");
    CUDA_TRY(A.bp.reserve(sizeof(uint32_t) * n_slots));
    CUDA_TRY(A.keys.reserve(sizeof(uint64_t) * n_slots));
    CUDA_TRY(A.keys2.reserve(sizeof(uint64_t) * n_slots));
    CUDA_TRY(A.vals.reserve(sizeof(uint32_t) * n_slots));
    CUDA_TRY(A.vals2.reserve(sizeof(uint32_t) * n_slots));
    CUDA_TRY(A.recs.reserve(sizeof(ArmRec) * n_slots));
    aln_arms_kernel<<<(unsigned)((n_alns + tb - 1) / tb), tb, 0, s>>>(
        (const HypoContigDesc*)A.contigs.p, (const HypoRegionRec*)A.regions.p, (const HypoAlnDesc*)A.alns.p, n_alns,
        (const uint32_t*)A.cigar.p, (const uint8_t*)A.seqs.p, (const AlnInfo*)A.info.p, (const uint64_t*)A.slot_off.p, k,
        (uint32_t*)A.bp.p, (uint64_t*)A.keys.p, (ArmRec*)A.recs.p, (uint32_t*)A.vals.p, (Ctl*)A.ctl.p);
    {
        size_t bytes = 0;
        CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint64_t*)A.keys.p, (uint64_t*)A.keys2.p,
                                                 (const uint32_t*)A.vals.p, (uint32_t*)A.vals2.p, n_slots, 0, 64, s));
        CUDA_TRY(A.tmp.reserve(bytes + 256));
        bytes = A.tmp.cap;
        CUDA_TRY(cub::DeviceRadixSort::SortPairs(A.tmp.p, bytes, (const uint64_t*)A.keys.p, (uint64_t*)A.keys2.p,
                                                 (const uint32_t*)A.vals.p, (uint32_t*)A.vals2.p, n_slots, 0, 64, s));
    }
    CUDA_TRY(cudaMemcpyAsync(h + 1, A.ctl.p, sizeof(Ctl), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    const uint64_t n_valid = ((Ctl*)(h + 1))->n_valid;
    if (n_valid == 0) return HYPO_OK;
    const uint64_t* keys = (const uint64_t*)A.keys2.p;
    const uint32_t* vals = (const uint32_t*)A.vals2.p;
    CUDA_TRY(A.head.reserve(sizeof(uint64_t) * (n_valid + 1)));
    CUDA_TRY(A.wid.reserve(sizeof(uint64_t) * (n_valid + 1)));
    const unsigned gv = (unsigned)((n_valid + tb - 1) / tb);
    head_kernel<<<gv, tb, 0, s>>>(keys, n_valid, (uint64_t*)A.head.p);
    CUDA_TRY(cudaMemsetAsync((uint64_t*)A.head.p + n_valid, 0, sizeof(uint64_t), s));
    if (int rc = scan(A.tmp, (const uint64_t*)A.head.p, (uint64_t*)A.wid.p, n_valid + 1, s)) return rc;
    CUDA_TRY(cudaMemcpyAsync(h, (uint64_t*)A.wid.p + n_valid, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    const uint64_t n_raw = h[0];
    // (wid of a head record is its window; of a non-head record it is the window's id + 1: make it the id)
    CUDA_TRY(A.wt.reserve(sizeof(WinTmp) * n_raw));
    CUDA_TRY(A.keep_win.reserve(sizeof(uint64_t) * (n_raw + 1)));
    CUDA_TRY(A.draft_bytes.reserve(sizeof(uint64_t) * (n_raw + 1)));
    CUDA_TRY(A.win_idx.reserve(sizeof(uint64_t) * (n_raw + 1)));
    CUDA_TRY(A.draft_byte.reserve(sizeof(uint64_t) * (n_raw + 2)));
    CUDA_TRY(A.keep_arm.reserve(sizeof(uint64_t) * (n_valid + 1)));
    CUDA_TRY(A.arm_bytes.reserve(sizeof(uint64_t) * (n_valid + 1)));
    CUDA_TRY(A.arm_idx.reserve(sizeof(uint64_t) * (n_valid + 1)));
    CUDA_TRY(A.arm_byte.reserve(sizeof(uint64_t) * (n_valid + 1)));
    return (int)0x7fffffff;   // continue in extract_finish (split to keep the function readable)
}

// inclusive-scan fix-up: window id of every record = (exclusive scan of the head flags) + head - 1
__global__ void wid_fix_kernel(const uint64_t* __restrict__ head, uint64_t* __restrict__ wid, uint64_t n) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n) wid[i] = wid[i] + head[i] - 1;
}

int extract_finish(uint64_t n_regions, uint64_t n_valid, uint64_t n_raw) {
    cudaStream_t s = A.stream;
    const int tb = 128;
    uint64_t* h = (uint64_t*)A.pinned;
    const uint64_t* keys = (const uint64_t*)A.keys2.p;
    const uint32_t* vals = (const uint32_t*)A.vals2.p;
    const unsigned gv = (unsigned)((n_valid + tb - 1) / tb), gw = (unsigned)((n_raw + tb - 1) / tb);
    wid_fix_kernel<<<gv, tb, 0, s>>>((const uint64_t*)A.head.p, (uint64_t*)A.wid.p, n_valid);
    window_kernel<<<gv, tb, 0, s>>>(keys, vals, (const ArmRec*)A.recs.p, n_valid, (const uint64_t*)A.head.p,
                                   (const uint64_t*)A.wid.p, (const HypoRegionRec*)A.regions.p, (const uint32_t*)A.reg_end.p,
                                   (WinTmp*)A.wt.p, (uint64_t*)A.keep_win.p, (uint64_t*)A.draft_bytes.p);
    armflag_kernel<<<gv, tb, 0, s>>>(keys, vals, (const ArmRec*)A.recs.p, n_valid, (const uint64_t*)A.wid.p,
                                    (const WinTmp*)A.wt.p, (uint64_t*)A.keep_arm.p, (uint64_t*)A.arm_bytes.p);
    CUDA_TRY(cudaMemsetAsync((uint64_t*)A.keep_win.p + n_raw, 0, sizeof(uint64_t), s));
    CUDA_TRY(cudaMemsetAsync((uint64_t*)A.draft_bytes.p + n_raw, 0, sizeof(uint64_t), s));
    CUDA_TRY(cudaMemsetAsync((uint64_t*)A.keep_arm.p + n_valid, 0, sizeof(uint64_t), s));
    CUDA_TRY(cudaMemsetAsync((uint64_t*)A.arm_bytes.p + n_valid, 0, sizeof(uint64_t), s));
    if (int rc = scan(A.tmp, (const uint64_t*)A.keep_win.p, (uint64_t*)A.win_idx.p, n_raw + 1, s)) return rc;
    if (int rc = scan(A.tmp, (const uint64_t*)A.draft_bytes.p, (uint64_t*)A.draft_byte.p, n_raw + 1, s)) return rc;
    if (int rc = scan(A.tmp, (const uint64_t*)A.keep_arm.p, (uint64_t*)A.arm_idx.p, n_valid + 1, s)) return rc;
    if (int rc = scan(A.tmp, (const uint64_t*)A.arm_bytes.p, (uint64_t*)A.arm_byte.p, n_valid + 1, s)) return rc;
    CUDA_TRY(cudaMemcpyAsync(h, (uint64_t*)A.win_idx.p + n_raw, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(h + 1, (uint64_t*)A.draft_byte.p + n_raw, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(h + 2, (uint64_t*)A.arm_idx.p + n_valid, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(h + 3, (uint64_t*)A.arm_byte.p + n_valid, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    A.n_win = h[0]; A.n_arms = h[2]; A.n_bytes = h[1] + h[3];
    (void)n_regions;
    if (A.n_win == 0) return HYPO_OK;
    CUDA_TRY(A.win.reserve(sizeof(HypoWindowDesc) * A.n_win));
    CUDA_TRY(A.win_region.reserve(sizeof(uint64_t) * A.n_win));
    CUDA_TRY(A.arms.reserve(sizeof(HypoArmDesc) * std::max<uint64_t>(A.n_arms, 1)));
    CUDA_TRY(A.packed.reserve(A.n_bytes + 16));
    CUDA_TRY(cudaMemsetAsync((uint8_t*)A.packed.p + A.n_bytes, 0, 16, s));
    fill_window_kernel<<<gw, tb, 0, s>>>((const WinTmp*)A.wt.p, n_raw, (const uint64_t*)A.win_idx.p, (const uint64_t*)A.arm_idx.p,
                                        (const uint64_t*)A.arm_byte.p, (const uint64_t*)A.draft_byte.p,
                                        (const HypoRegionRec*)A.regions.p, (const uint32_t*)A.reg_end.p,
                                        (const uint32_t*)A.reg_contig.p, (const HypoContigDesc*)A.contigs.p,
                                        (const uint8_t*)A.drafts.p, (HypoWindowDesc*)A.win.p, (uint64_t*)A.win_region.p,
                                        (uint8_t*)A.packed.p);
    fill_arm_kernel<<<gv, tb, 0, s>>>(keys, vals, (const ArmRec*)A.recs.p, n_valid, (const uint64_t*)A.keep_arm.p,
                                     (const uint64_t*)A.arm_idx.p, (const uint64_t*)A.arm_byte.p, (const uint64_t*)A.wid.p,
                                     (const uint64_t*)A.draft_byte.p,
                                     (const HypoAlnDesc*)A.alns.p, (const AlnInfo*)A.info.p, (const uint8_t*)A.seqs.p,
                                     (HypoArmDesc*)A.arms.p, (uint8_t*)A.packed.p);
    CUDA_TRY(cudaGetLastError());
    return HYPO_OK;
}

int extract_all(const HypoContigDesc* contigs, uint64_t n_contigs, const HypoRegionRec* regions, uint64_t n_regions,
                const uint8_t* drafts, uint64_t draft_bytes, const HypoAlnDesc* alns, uint64_t n_alns,
                const uint32_t* cigar, uint64_t n_cigar, const uint8_t* seqs, uint64_t seq_bytes, uint32_t k) {
    if ((!contigs && n_contigs) || (!regions && n_regions) || (!alns && n_alns) || (!cigar && n_cigar) || (!seqs && seq_bytes) ||
        (!drafts && draft_bytes))
        return hypo_internal_fail(HYPO_E_ARG, "NULL input buffer");
    // host-side validation of the tables (cheap, and keeps the kernels free of bounds checks)
    for (uint64_t c = 0; c < n_contigs; ++c) {
        const HypoContigDesc& d = contigs[c];
        if (d.first_region + d.n_regions > n_regions || d.draft_off + ((uint64_t)d.len + 1) / 2 > draft_bytes)
            return hypo_internal_fail(HYPO_E_ARG, "contig descriptor out of range");
        for (uint32_t r = 0; r < d.n_regions; ++r) {
            const HypoRegionRec& g = regions[d.first_region + r];
            if (g.type > HYPO_REG_OTHER || g.start >= d.len || (r == 0 ? g.start != 0 : g.start <= regions[d.first_region + r - 1].start))
                return hypo_internal_fail(HYPO_E_ARG, "region table: starts must begin at 0 and increase, types must be HYPO_REG_*");
        }
    }
    for (uint64_t i = 0; i < n_alns; ++i)
        if (alns[i].cigar_off + alns[i].n_cigar > n_cigar || alns[i].seq_off + ((uint64_t)alns[i].l_qseq + 1) / 2 > seq_bytes)
            return hypo_internal_fail(HYPO_E_ARG, "alignment descriptor out of range");
    int rc = extract_device(contigs, n_contigs, regions, n_regions, drafts, draft_bytes, alns, n_alns, cigar, n_cigar, seqs,
                            seq_bytes, k);
    if (rc != (int)0x7fffffff) return rc;
    uint64_t* h = (uint64_t*)A.pinned;
    const uint64_t n_raw = h[0];
    const uint64_t n_valid = ((Ctl*)(h + 1))->n_valid;
    return extract_finish(n_regions, n_valid, n_raw);
}

}  // namespace

extern "C" {

int hypo_gpu_extract_arms(const HypoContigDesc* contigs, uint64_t n_contigs, const HypoRegionRec* regions, uint64_t n_regions,
                          const uint8_t* drafts, uint64_t draft_bytes, const HypoAlnDesc* alns, uint64_t n_alns,
                          const uint32_t* cigar, uint64_t n_cigar, const uint8_t* seqs, uint64_t seq_bytes, uint32_t k,
                          HypoWindowDesc* win, uint64_t win_cap, uint64_t* n_win, uint64_t* win_region, HypoArmDesc* arms,
                          uint64_t arm_cap, uint64_t* n_arms, uint8_t* packed, uint64_t packed_cap, uint64_t* packed_bytes) {
    std::lock_guard<std::mutex> lk(a_mu);
    hypo_internal_fail(HYPO_OK, "");
    if (!n_win || !n_arms || !packed_bytes) return hypo_internal_fail(HYPO_E_ARG, "NULL size output");
    *n_win = *n_arms = *packed_bytes = 0;
    if (int rc = extract_all(contigs, n_contigs, regions, n_regions, drafts, draft_bytes, alns, n_alns, cigar, n_cigar, seqs,
                             seq_bytes, k))
        return rc;
    *n_win = A.n_win; *n_arms = A.n_arms; *packed_bytes = A.n_bytes;
    if (A.n_win > win_cap || A.n_arms > arm_cap || A.n_bytes > packed_cap)
        return hypo_internal_fail(HYPO_E_OUT_CAP, "extracted batch is larger than the output buffers");
    cudaStream_t s = A.stream;
    if (A.n_win) {
        CUDA_TRY(cudaMemcpyAsync(win, A.win.p, sizeof(HypoWindowDesc) * A.n_win, cudaMemcpyDeviceToHost, s));
        if (win_region) CUDA_TRY(cudaMemcpyAsync(win_region, A.win_region.p, sizeof(uint64_t) * A.n_win, cudaMemcpyDeviceToHost, s));
        if (A.n_arms) CUDA_TRY(cudaMemcpyAsync(arms, A.arms.p, sizeof(HypoArmDesc) * A.n_arms, cudaMemcpyDeviceToHost, s));
        if (A.n_bytes) CUDA_TRY(cudaMemcpyAsync(packed, A.packed.p, A.n_bytes, cudaMemcpyDeviceToHost, s));
    }
    CUDA_TRY(cudaStreamSynchronize(s));
    return HYPO_OK;
}

int hypo_gpu_polish_alignments(const HypoContigDesc* contigs, uint64_t n_contigs, const HypoRegionRec* regions,
                               uint64_t n_regions, const uint8_t* drafts, uint64_t draft_bytes, const HypoAlnDesc* alns,
                               uint64_t n_alns, const uint32_t* cigar, uint64_t n_cigar, const uint8_t* seqs,
                               uint64_t seq_bytes, uint32_t k, char* out, uint64_t out_cap, uint64_t* out_off) {
    std::lock_guard<std::mutex> lk(a_mu);
    hypo_internal_fail(HYPO_OK, "");
    if (!out_off) return hypo_internal_fail(HYPO_E_ARG, "out_off == NULL");
    if (int rc = extract_all(contigs, n_contigs, regions, n_regions, drafts, draft_bytes, alns, n_alns, cigar, n_cigar, seqs,
                             seq_bytes, k))
        return rc;
    cudaStream_t s = A.stream;
    const int tb = 256;
    // region descriptors for the stitcher: every region copies its draft unless a window polishes it
    CUDA_TRY(A.sregions.reserve(sizeof(HypoRegionDesc) * std::max<uint64_t>(n_regions, 1)));
    CUDA_TRY(A.doff.reserve(sizeof(uint64_t) * std::max<uint64_t>(n_contigs, 1)));
    std::vector<uint64_t> doff(n_contigs), first(n_contigs + 1);
    for (uint64_t c = 0; c < n_contigs; ++c) { doff[c] = contigs[c].draft_off; first[c] = contigs[c].first_region; }
    first[n_contigs] = n_regions;
    for (uint64_t c = 0; c + 1 < n_contigs; ++c)
        if (contigs[c].first_region + contigs[c].n_regions != contigs[c + 1].first_region)
            return hypo_internal_fail(HYPO_E_ARG, "contigs must list their regions back to back, in order");
    if (n_contigs && (contigs[0].first_region != 0 || contigs[n_contigs - 1].first_region + contigs[n_contigs - 1].n_regions != n_regions))
        return hypo_internal_fail(HYPO_E_ARG, "contigs must cover the region table");
    if (n_contigs) CUDA_TRY(cudaMemcpyAsync(A.doff.p, doff.data(), sizeof(uint64_t) * n_contigs, cudaMemcpyHostToDevice, s));
    if (n_regions) {
        stitch_regions_kernel<<<(unsigned)((n_regions + tb - 1) / tb), tb, 0, s>>>(
            (const HypoRegionRec*)A.regions.p, (const uint32_t*)A.reg_end.p, n_regions, (HypoRegionDesc*)A.sregions.p);
        if (A.n_win)
            stitch_windows_kernel<<<(unsigned)((A.n_win + tb - 1) / tb), tb, 0, s>>>((const uint64_t*)A.win_region.p, A.n_win,
                                                                                 (HypoRegionDesc*)A.sregions.p);
    }
    CUDA_TRY(cudaStreamSynchronize(s));   // (doff / first live on this stack frame)
    return hypo_internal_polish_device((const HypoWindowDesc*)A.win.p, A.n_win, (const HypoArmDesc*)A.arms.p, A.n_arms,
                                       (const uint8_t*)A.packed.p, A.n_bytes,
                                       (const HypoRegionDesc*)A.sregions.p, (const uint32_t*)A.reg_contig.p, n_regions,
                                       first.data(), n_contigs, (const uint8_t*)A.drafts.p, (const uint64_t*)A.doff.p, out,
                                       out_cap, out_off, s);
}

}  // extern "C"
