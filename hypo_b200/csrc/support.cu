// support.cu — support counting on the device: how many reads cover / support every solid
// k-mer of the draft and every minimiser of its large weak regions.  One thread per alignment, device atomics
// instead of the reference's mutex per k-mer; HBM-bound integer work.
//
// Reference code restated here (paths relative to the reference root):
//   Alignment::initialise_pos / copy_data     src/Alignment.cpp:513-576   aln_span
//   Alignment::update_solidkmers_support      src/Alignment.cpp:65-131    solid_support_kernel
//   Alignment::update_minimisers_support      src/Alignment.cpp:133-220   minimiser_support_kernel
//   MinimizerDeque                            include/MinimizerDeque.hpp   MiniWindow
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/hypo_b200.h"

extern "C" int hypo_internal_fail(int code, const char* msg);
extern "C" int hypo_internal_primary_device(void);

namespace {

constexpr uint32_t kMiniK = 10, kMiniW = 10;   // reference src/main.cpp:86

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

__device__ __forceinline__ uint32_t cigar_type(uint32_t op) { return (0x3C1A7u >> (op << 1)) & 3u; }
__device__ __forceinline__ uint32_t nib(const uint8_t* seq, uint32_t i) { return (seq[i >> 1] >> ((i & 1) ? 0 : 4)) & 15u; }
__device__ __forceinline__ uint32_t code_of(uint32_t n) { return n == 1 ? 0u : n == 2 ? 1u : n == 4 ? 2u : n == 8 ? 3u : 4u; }

struct Span {
    uint32_t rb, re, qab, qlen;
    bool valid;
};

// initialise_pos + copy_data: reference span, aligned part of the query, validity
__device__ Span aln_span(const HypoAlnDesc& a, const uint32_t* __restrict__ cigar, const uint8_t* __restrict__ seqs) {
    const uint32_t* cg = cigar + a.cigar_off;
    uint32_t qab = 0, qp = 0, rp = a.pos, clip_end = 0;
    bool clip_before = true;
    for (uint32_t j = 0; j < a.n_cigar; ++j) {
        const uint32_t op = cg[j] & 15u, len = cg[j] >> 4;
        if (clip_before) {
            if (op == 4) qab += len;
            else if (op != 5) clip_before = false;
        }
        const uint32_t t = cigar_type(op);
        if (t == 3) { rp += len; qp += len; }
        else if (t & 2) rp += len;
        else if (t & 1) { if (!clip_before && op == 4) clip_end += len; qp += len; }
    }
    Span s;
    const uint32_t qae = qp - clip_end;
    s.rb = a.pos; s.re = rp; s.qab = qab; s.qlen = qae >= qab ? qae - qab : 0;
    s.valid = qae >= qab && qae <= a.l_qseq;
    const uint8_t* sq = seqs + a.seq_off;
    for (uint32_t q = qab; q < qae && s.valid; ++q) s.valid = code_of(nib(sq, q)) < 4;
    return s;
}

__device__ uint64_t lower_bound_u32(const uint32_t* __restrict__ v, uint64_t lo, uint64_t hi, uint32_t x) {
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (v[mid] < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// ---- solid k-mers (src/Alignment.cpp:65-131) --------------------------------------------------------------------
__global__ void solid_support_kernel(const uint64_t* __restrict__ contig_first, uint64_t n_contigs,
                                     const uint32_t* __restrict__ spos, const uint64_t* __restrict__ kid,
                                     const HypoAlnDesc* __restrict__ alns, uint64_t n_alns,
                                     const uint32_t* __restrict__ cigar, const uint8_t* __restrict__ seqs, uint32_t k,
                                     uint32_t* __restrict__ cov, uint32_t* __restrict__ sup) {
    const uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (t >= n_alns) return;
    const HypoAlnDesc a = alns[t];
    if (a.contig >= n_contigs) return;
    const Span s = aln_span(a, cigar, seqs);
    if (!s.valid) return;
    const uint64_t c0 = contig_first[a.contig], c1 = contig_first[a.contig + 1];
    const uint64_t first = lower_bound_u32(spos, c0, c1, s.rb);   // _Rsolid_pos(_rb)
    uint64_t last = lower_bound_u32(spos, c0, c1, s.re);
    // discard those which do not wholly fall in the alignment (if none does, `last` keeps its value: the
    // reference's loop then simply never breaks)
    for (uint64_t i = last; i > first; --i)
        if (spos[i - 1] + k <= s.re) { last = i; break; }
    if (last <= first) return;
    for (uint64_t i = first; i < last; ++i) atomicAdd(&cov[i], 1u);
    const uint64_t kmask = (1ull << (2 * k)) - 1;
    const uint32_t num_cbases = s.re - s.rb;
    const uint8_t* sq = seqs + a.seq_off;
    long long pvs_kpos = -1;
    uint32_t pvs_rbind = 0, klen = 0;
    uint64_t kmer = 0, lo = first;
    for (uint32_t r_ind = 0; r_ind < s.qlen; ++r_ind) {
        kmer = ((kmer << 2) | code_of(nib(sq, s.qab + r_ind))) & kmask;
        if (klen < k) ++klen;
        if (klen < k) continue;
        const uint32_t r_bind = r_ind + 1 - k;
        if (r_bind > num_cbases) break;   // beyond every srange_right
        // candidates: solid positions within k of where this k-mer is expected, in position order
        while (lo < last && (uint64_t)spos[lo] + k < (uint64_t)s.rb + r_bind) ++lo;
        for (uint64_t c = lo; c < last && (uint64_t)spos[c] <= (uint64_t)s.rb + r_bind + k; ++c) {
            if (kid[c] != kmer) continue;
            const uint32_t c_dist = spos[c] - s.rb;
            const uint32_t left = c_dist > k ? c_dist - k : 0;
            const uint32_t right = min(num_cbases, c_dist + k);
            if (r_bind < left || r_bind > right) continue;
            bool update = true;
            if (pvs_kpos > -1 && (long long)spos[c] <= (long long)k + pvs_kpos) {
                // an adjacent / overlapping neighbour was supported: the read must keep their distance
                if ((unsigned long long)(uint32_t)(r_bind - pvs_rbind) != (unsigned long long)((long long)spos[c] - pvs_kpos))
                    update = false;
            }
            if (update) {
                pvs_kpos = spos[c];
                pvs_rbind = r_bind;
                atomicAdd(&sup[c], 1u);
            }
        }
    }
}

// ---- minimisers (src/Alignment.cpp:133-220) ---------------------------------------------------------------------
// The sliding-window minimum of the reference's MinimizerDeque: k-mers leave from the back while they are larger
// than the new one (the left-most smallest stays), from the front when they fall out of the window.
struct MiniWindow {
    uint32_t val[kMiniW + 1], pos[kMiniW + 1];
    uint32_t n = 0;
    __device__ void push(uint32_t v, uint32_t p) {
        while (n && val[n - 1] > v) --n;
        val[n] = v; pos[n] = p; ++n;
        while (pos[0] + kMiniW <= p) {
            for (uint32_t j = 1; j < n; ++j) { val[j - 1] = val[j]; pos[j - 1] = pos[j]; }
            --n;
        }
    }
};

constexpr int kSpanMinis = 48;   // contig minimisers handled per pass over the read

__global__ void minimiser_support_kernel(const uint64_t* __restrict__ contig_first_bound, const uint8_t* __restrict__ contig_even,
                                         uint64_t n_contigs, const uint32_t* __restrict__ bounds,
                                         const uint64_t* __restrict__ region_first_mini, const uint32_t* __restrict__ mpos,
                                         const uint32_t* __restrict__ mval, const HypoAlnDesc* __restrict__ alns,
                                         uint64_t n_alns, const uint32_t* __restrict__ cigar, const uint8_t* __restrict__ seqs,
                                         uint32_t* __restrict__ cov, uint32_t* __restrict__ sup) {
    const uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (t >= n_alns) return;
    const HypoAlnDesc a = alns[t];
    if (a.contig >= n_contigs) return;
    const Span s = aln_span(a, cigar, seqs);
    if (!s.valid) return;
    const uint64_t b0 = contig_first_bound[a.contig], b1 = contig_first_bound[a.contig + 1];   // bounds of the contig
    const bool even = contig_even[a.contig] != 0;
    // first = _RMreg_pos(_rb + 1) - 1, last = _RMreg_pos(_re): region indices inside the contig
    const long long first = (long long)(lower_bound_u32(bounds, b0, b1, s.rb + 1) - b0) - 1;
    const long long last = (long long)(lower_bound_u32(bounds, b0, b1, s.re) - b0);
    auto is_win = [&](long long x) { return even ? (x % 2 == 0) : (x % 2 == 1); };
    const long long fw = is_win(first) ? first : first + 1;
    const long long lw = is_win(last) ? last : last - 1;
    if (lw < fw) return;
    const uint32_t num_cbases = (s.re - s.rb) & 0xffffu;   // UINT16 in the reference
    const uint8_t* sq = seqs + a.seq_off;
    // the contig minimisers the alignment covers, a few at a time; one pass over the read per group
    uint32_t gid[kSpanMinis], gval[kSpanMinis], gleft[kSpanMinis], gright[kSpanMinis];
    long long w = fw;
    uint64_t mi = 0;
    bool region_open = false, done = false;
    while (!done) {
        int n = 0;
        while (n < kSpanMinis && !done) {
            if (!region_open) {
                if (w > lw || (uint64_t)w + 1 >= b1 - b0) { done = true; break; }   // (no region starts at the last bound)
                mi = region_first_mini[b0 + w];
                region_open = true;
            }
            if (mi >= region_first_mini[b0 + w + 1]) { region_open = false; w += 2; continue; }
            const uint32_t p = mpos[mi];
            if (p >= s.rb && p < s.re) {
                atomicAdd(&cov[mi], 1u);
                const uint32_t c_dist = p - s.rb;
                gid[n] = (uint32_t)mi; gval[n] = mval[mi];
                gleft[n] = c_dist > 2 * kMiniK ? c_dist - 2 * kMiniK : 0;
                gright[n] = min(num_cbases, (c_dist + 3 * kMiniK) & 0xffffu);
                ++n;
            }
            if (p >= s.re) { region_open = false; w += 2; continue; }   // (the reference breaks out of this region)
            ++mi;
        }
        if (n == 0) continue;
        // the read's minimisers (consecutive duplicates dropped), each held against the group
        const uint32_t mask = (1u << (2 * kMiniK)) - 1;
        MiniWindow win;
        uint32_t kmer = 0, not_n = 0, processed = 0, last_found = s.qlen + 1;
        for (uint32_t i = 0; i < s.qlen; ++i) {
            const uint32_t c = code_of(nib(sq, s.qab + i));   // (always < 4: the read is valid)
            ++not_n;
            kmer = ((kmer << 2) | c) & mask;
            if (not_n < kMiniK) continue;
            win.push(kmer, i);
            if (++processed < kMiniW) continue;
            const uint32_t start = win.pos[0] - kMiniK + 1;
            if (start != last_found) {
                const uint32_t v = win.val[0];
                for (int j = 0; j < n; ++j)
                    if (gval[j] == v && start >= gleft[j] && start <= gright[j]) atomicAdd(&sup[gid[j]], 1u);
            }
            last_found = start;
        }
    }
}

struct SupCtx {
    int device = -1;
    cudaStream_t stream = nullptr;
    DevBuf first, even, a, b, c, d, alns, cigar, seqs, cov, sup;
} S;
std::mutex s_mu;

#define CUDA_TRY(x)                                                                                        \
    do {                                                                                                   \
        cudaError_t e_ = (x);                                                                              \
        if (e_ != cudaSuccess) {                                                                           \
            char b_[384];                                                                                  \
            snprintf(b_, sizeof(b_), "%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return hypo_internal_fail(HYPO_E_CUDA, b_);                                                    \
        }                                                                                                  \
    } while (0)

int prepare(cudaStream_t* s) {
    const int dev = hypo_internal_primary_device();
    if (dev < 0) return hypo_internal_fail(HYPO_E_NOT_INIT, "hypo_gpu_init has not been called");
    CUDA_TRY(cudaSetDevice(dev));
    if (!S.stream) CUDA_TRY(cudaStreamCreateWithFlags(&S.stream, cudaStreamNonBlocking));
    S.device = dev;
    *s = S.stream;
    return HYPO_OK;
}

cudaError_t up(DevBuf& b, const void* src, size_t bytes, cudaStream_t s) {
    cudaError_t e = b.reserve(bytes + 16);
    if (e != cudaSuccess || bytes == 0) return e;
    return cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, s);
}

int check_alns(const HypoAlnDesc* alns, uint64_t n_alns, uint64_t n_cigar, uint64_t seq_bytes) {
    for (uint64_t i = 0; i < n_alns; ++i)
        if (alns[i].cigar_off + alns[i].n_cigar > n_cigar || alns[i].seq_off + ((uint64_t)alns[i].l_qseq + 1) / 2 > seq_bytes)
            return hypo_internal_fail(HYPO_E_ARG, "alignment descriptor out of range");
    return HYPO_OK;
}

}  // namespace

extern "C" {

int hypo_gpu_solid_kmer_support(const uint64_t* contig_first_kmer, uint64_t n_contigs, const uint32_t* solid_pos,
                                const uint64_t* kmer_id, uint64_t n_kmers, const HypoAlnDesc* alns, uint64_t n_alns,
                                const uint32_t* cigar, uint64_t n_cigar, const uint8_t* seqs, uint64_t seq_bytes, uint32_t k,
                                uint32_t* coverage, uint32_t* support) {
    std::lock_guard<std::mutex> lk(s_mu);
    hypo_internal_fail(HYPO_OK, "");
    if (!contig_first_kmer || (!solid_pos && n_kmers) || (!kmer_id && n_kmers) || (!alns && n_alns) || !coverage || !support)
        return hypo_internal_fail(HYPO_E_ARG, "NULL buffer");
    if (k == 0 || k > 31) return hypo_internal_fail(HYPO_E_ARG, "k must be 1..31");
    if (contig_first_kmer[0] != 0 || contig_first_kmer[n_contigs] != n_kmers)
        return hypo_internal_fail(HYPO_E_ARG, "contig_first_kmer must run from 0 to n_kmers");
    for (uint64_t c = 0; c < n_contigs; ++c)
        if (contig_first_kmer[c] > contig_first_kmer[c + 1]) return hypo_internal_fail(HYPO_E_ARG, "contig_first_kmer must not decrease");
    if (int rc = check_alns(alns, n_alns, n_cigar, seq_bytes)) return rc;
    cudaStream_t s;
    if (int rc = prepare(&s)) return rc;
    CUDA_TRY(up(S.first, contig_first_kmer, sizeof(uint64_t) * (n_contigs + 1), s));
    CUDA_TRY(up(S.a, solid_pos, sizeof(uint32_t) * n_kmers, s));
    CUDA_TRY(up(S.b, kmer_id, sizeof(uint64_t) * n_kmers, s));
    CUDA_TRY(up(S.alns, alns, sizeof(HypoAlnDesc) * n_alns, s));
    CUDA_TRY(up(S.cigar, cigar, sizeof(uint32_t) * n_cigar, s));
    CUDA_TRY(up(S.seqs, seqs, seq_bytes, s));
    CUDA_TRY(S.cov.reserve(sizeof(uint32_t) * (n_kmers + 1)));
    CUDA_TRY(S.sup.reserve(sizeof(uint32_t) * (n_kmers + 1)));
    CUDA_TRY(cudaMemsetAsync(S.cov.p, 0, sizeof(uint32_t) * (n_kmers + 1), s));
    CUDA_TRY(cudaMemsetAsync(S.sup.p, 0, sizeof(uint32_t) * (n_kmers + 1), s));
    if (n_alns && n_kmers) {
        const int tb = 128;
        solid_support_kernel<<<(unsigned)((n_alns + tb - 1) / tb), tb, 0, s>>>(
            (const uint64_t*)S.first.p, n_contigs, (const uint32_t*)S.a.p, (const uint64_t*)S.b.p, (const HypoAlnDesc*)S.alns.p,
            n_alns, (const uint32_t*)S.cigar.p, (const uint8_t*)S.seqs.p, k, (uint32_t*)S.cov.p, (uint32_t*)S.sup.p);
        CUDA_TRY(cudaGetLastError());
    }
    if (n_kmers) {
        CUDA_TRY(cudaMemcpyAsync(coverage, S.cov.p, sizeof(uint32_t) * n_kmers, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaMemcpyAsync(support, S.sup.p, sizeof(uint32_t) * n_kmers, cudaMemcpyDeviceToHost, s));
    }
    CUDA_TRY(cudaStreamSynchronize(s));
    return HYPO_OK;
}

int hypo_gpu_minimiser_support(const uint64_t* contig_first_bound, const uint8_t* contig_even, uint64_t n_contigs,
                               const uint32_t* bounds, uint64_t n_bounds, const uint64_t* region_first_mini,
                               const uint32_t* mini_pos, const uint32_t* mini_val, uint64_t n_minis, const HypoAlnDesc* alns,
                               uint64_t n_alns, const uint32_t* cigar, uint64_t n_cigar, const uint8_t* seqs, uint64_t seq_bytes,
                               uint32_t* coverage, uint32_t* support) {
    std::lock_guard<std::mutex> lk(s_mu);
    hypo_internal_fail(HYPO_OK, "");
    if (!contig_first_bound || !contig_even || (!bounds && n_bounds) || !region_first_mini || (!alns && n_alns) || !coverage || !support)
        return hypo_internal_fail(HYPO_E_ARG, "NULL buffer");
    if (contig_first_bound[0] != 0 || contig_first_bound[n_contigs] != n_bounds)
        return hypo_internal_fail(HYPO_E_ARG, "contig_first_bound must run from 0 to n_bounds");
    for (uint64_t r = 0; r < n_bounds; ++r)
        if (region_first_mini[r] > region_first_mini[r + 1] || region_first_mini[r + 1] > n_minis)
            return hypo_internal_fail(HYPO_E_ARG, "region_first_mini must not decrease and end at n_minis");
    if (int rc = check_alns(alns, n_alns, n_cigar, seq_bytes)) return rc;
    cudaStream_t s;
    if (int rc = prepare(&s)) return rc;
    CUDA_TRY(up(S.first, contig_first_bound, sizeof(uint64_t) * (n_contigs + 1), s));
    CUDA_TRY(up(S.even, contig_even, n_contigs, s));
    CUDA_TRY(up(S.a, bounds, sizeof(uint32_t) * n_bounds, s));
    CUDA_TRY(up(S.b, region_first_mini, sizeof(uint64_t) * (n_bounds + 1), s));
    CUDA_TRY(up(S.c, mini_pos, sizeof(uint32_t) * n_minis, s));
    CUDA_TRY(up(S.d, mini_val, sizeof(uint32_t) * n_minis, s));
    CUDA_TRY(up(S.alns, alns, sizeof(HypoAlnDesc) * n_alns, s));
    CUDA_TRY(up(S.cigar, cigar, sizeof(uint32_t) * n_cigar, s));
    CUDA_TRY(up(S.seqs, seqs, seq_bytes, s));
    CUDA_TRY(S.cov.reserve(sizeof(uint32_t) * (n_minis + 1)));
    CUDA_TRY(S.sup.reserve(sizeof(uint32_t) * (n_minis + 1)));
    CUDA_TRY(cudaMemsetAsync(S.cov.p, 0, sizeof(uint32_t) * (n_minis + 1), s));
    CUDA_TRY(cudaMemsetAsync(S.sup.p, 0, sizeof(uint32_t) * (n_minis + 1), s));
    if (n_alns && n_minis) {
        const int tb = 128;
        minimiser_support_kernel<<<(unsigned)((n_alns + tb - 1) / tb), tb, 0, s>>>(
            (const uint64_t*)S.first.p, (const uint8_t*)S.even.p, n_contigs, (const uint32_t*)S.a.p, (const uint64_t*)S.b.p,
            (const uint32_t*)S.c.p, (const uint32_t*)S.d.p, (const HypoAlnDesc*)S.alns.p, n_alns, (const uint32_t*)S.cigar.p,
            (const uint8_t*)S.seqs.p, (uint32_t*)S.cov.p, (uint32_t*)S.sup.p);
        CUDA_TRY(cudaGetLastError());
    }
    if (n_minis) {
        CUDA_TRY(cudaMemcpyAsync(coverage, S.cov.p, sizeof(uint32_t) * n_minis, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaMemcpyAsync(support, S.sup.p, sizeof(uint32_t) * n_minis, cudaMemcpyDeviceToHost, s));
    }
    CUDA_TRY(cudaStreamSynchronize(s));
    return HYPO_OK;
}

}  // extern "C"
