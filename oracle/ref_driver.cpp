// TEST INFRASTRUCTURE ONLY — never linked into or called by the product path.
//
// Thin driver around the UNMODIFIED reference implementation of the hot path.
// It is compiled by oracle/Makefile straight from the sources under
// /root/reference (hypo::Window + hypo::PackedSeq + the adapted spoa fork) with
// the reference's default flags (-O3, no -march  =>  spoa's SISD engine, the
// parity oracle) into oracle/_ref/libhypo_ref.so.
// No reference source is copied into this repository.
//
// It speaks the same batch layout as include/hypo_b200.h so that tests can
// hand the very same buffers to the reference, to the C restatement
// (oracle/poa_oracle.c) and to the CUDA library.
//
// Reference call sequence reproduced here (reference src/Hypo.cpp:236-248):
//   Window::prepare_for_poa(score_params, threads);
//   #pragma omp parallel for schedule(static,1)
//   for w: window[w]->generate_consensus(omp_get_thread_num());
#include <omp.h>

#include <chrono>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <memory>
#include <string>
#include <vector>

// The engines live in private static vectors that prepare_for_poa only ever
// appends to (reference src/Window.cpp:31-42).  To be able to switch score
// parameters between calls inside one test process the driver needs to clear
// them; the access specifier is lifted for this translation unit only.
// Everything Window.hpp includes is pulled in first (include guards), so the
// define only affects class hypo::Window itself.
#include "spoa/spoa.hpp"
#include "globalDefs.hpp"
#include "PackedSeq.hpp"
#include "Filter.hpp"
#define private public
#include "Window.hpp"
#undef private

#include "../include/hypo_b200.h"

namespace {

const char kNib[16] = {'A', 'C', 'G', 'T', 'N', 'N', 'N', 'N',
                       'N', 'N', 'N', 'N', 'N', 'N', 'N', 'N'};

std::string unpack2(const uint8_t* p, uint32_t len) {
    std::string s(len, 'A');
    for (uint32_t i = 0; i < len; ++i) s[i] = "ACGT"[(p[i >> 2] >> (6 - 2 * (i & 3))) & 3];
    return s;
}

std::string unpack4(const uint8_t* p, uint32_t len) {
    std::string s(len, 'A');
    for (uint32_t i = 0; i < len; ++i) s[i] = kNib[(p[i >> 1] >> ((i & 1) ? 0 : 4)) & 15];
    return s;
}

int8_t g_scores[6] = {0, 0, 0, 0, 0, 0};
int g_threads = 0;

void ensure_engines(const int8_t scores[6], int threads) {
    if (g_threads >= threads && std::memcmp(scores, g_scores, 6) == 0) return;
    hypo::Window::_alignment_engines.clear();
    hypo::Window::_alignment_engines_long.clear();
    hypo::ScoreParams sp;
    sp.sr_match_score = scores[0];
    sp.sr_misMatch_score = scores[1];
    sp.sr_gap_penalty = scores[2];
    sp.lr_match_score = scores[3];
    sp.lr_misMatch_score = scores[4];
    sp.lr_gap_penalty = scores[5];
    hypo::Window::prepare_for_poa(sp, (hypo::UINT32)threads);
    std::memcpy(g_scores, scores, 6);
    g_threads = threads;
}

}  // namespace

extern "C" {

// n_threads <= 0: omp_get_max_threads().  schedule: 0 = schedule(static,1) as
// shipped (reference src/Hypo.cpp:240), 1 = schedule(dynamic,1).
// arm_accepted (nullable, n_arms bytes): 1 if Window::add_* kept the arm (LONG
// windows run every arm through hypo::Filter::is_good at insert time, reference
// include/Window.hpp:66-101, which is outside the hot path).
// seconds (nullable): wall time of the consensus loop only.
int hypo_ref_consensus_batch(const int8_t scores[6], const HypoWindowDesc* win, uint64_t n_win,
                             const HypoArmDesc* arms, uint64_t n_arms, const uint8_t* packed,
                             uint64_t packed_bytes, char* out, uint64_t out_cap,
                             uint64_t* out_off, uint8_t* arm_accepted, int n_threads,
                             int schedule, double* seconds) {
    (void)n_arms;
    (void)packed_bytes;
    if (n_threads <= 0) n_threads = omp_get_max_threads();
    ensure_engines(scores, n_threads);

    std::vector<std::unique_ptr<hypo::Window>> windows(n_win);
#pragma omp parallel for schedule(dynamic, 64) num_threads(n_threads)
    for (uint64_t w = 0; w < n_win; ++w) {
        const HypoWindowDesc& d = win[w];
        hypo::PackedSeq<4> draft(unpack4(packed + d.draft_off, d.draft_len));
        auto wt = d.wtype == HYPO_WINDOW_LONG ? hypo::WindowType::LONG : hypo::WindowType::SHORT;
        windows[w].reset(new hypo::Window(draft, 0, d.draft_len, wt));
        hypo::Window& W = *windows[w];
        uint64_t a = d.first_arm;
        for (uint32_t i = 0; i < d.n_internal; ++i, ++a) {
            auto before = W._internal_arms.size();
            W.add_internal(hypo::PackedSeq<2>(unpack2(packed + arms[a].off, arms[a].len)));
            if (arm_accepted) arm_accepted[a] = W._internal_arms.size() != before;
        }
        for (uint32_t i = 0; i < d.n_pre; ++i, ++a) {
            auto before = W._pre_arms.size();
            W.add_prefix(hypo::PackedSeq<2>(unpack2(packed + arms[a].off, arms[a].len)));
            if (arm_accepted) arm_accepted[a] = W._pre_arms.size() != before;
        }
        for (uint32_t i = 0; i < d.n_suf; ++i, ++a) {
            auto before = W._suf_arms.size();
            W.add_suffix(hypo::PackedSeq<2>(unpack2(packed + arms[a].off, arms[a].len)));
            if (arm_accepted) arm_accepted[a] = W._suf_arms.size() != before;
        }
        for (uint32_t i = 0; i < d.n_empty; ++i) W.add_empty();
    }

    auto t0 = std::chrono::steady_clock::now();
    if (schedule == 0) {
#pragma omp parallel for schedule(static, 1) num_threads(n_threads)
        for (uint64_t w = 0; w < n_win; ++w) windows[w]->generate_consensus(omp_get_thread_num());
    } else {
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
        for (uint64_t w = 0; w < n_win; ++w) windows[w]->generate_consensus(omp_get_thread_num());
    }
    auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();

    uint64_t pos = 0;
    for (uint64_t w = 0; w < n_win; ++w) {
        out_off[w] = pos;
        std::string c = windows[w]->get_consensus();
        if (pos + c.size() > out_cap) return HYPO_E_OUT_CAP;
        std::memcpy(out + pos, c.data(), c.size());
        pos += c.size();
    }
    out_off[n_win] = pos;
    return HYPO_OK;
}

// Raw spoa entry used to pin the oracle on the reference's own known-answer test
// SpoaAlignmentTest.GlobalConsensus (reference external/spoa/test/spoa_test.cpp:220-239):
// kNW, linear gaps, sequences added in order with weight 1 per base,
// consensus = Graph::generate_consensus().
// seqs: concatenated ASCII sequences; seq_off: n_seq+1 offsets.
// align_type: 0 kNW, 1 kLOV, 2 kROV.
int hypo_ref_spoa_consensus(int8_t m, int8_t n, int8_t g, const char* seqs,
                            const uint64_t* seq_off, uint32_t n_seq, const uint8_t* align_type,
                            char* out, uint64_t out_cap, uint64_t* out_len) {
    auto engine = spoa::createAlignmentEngine(spoa::AlignmentType::kNW, m, n, g);
    auto graph = spoa::createGraph();
    for (uint32_t i = 0; i < n_seq; ++i) {
        std::string s(seqs + seq_off[i], seqs + seq_off[i + 1]);
        auto t = spoa::AlignmentType::kNW;
        if (align_type && align_type[i] == 1) t = spoa::AlignmentType::kLOV;
        if (align_type && align_type[i] == 2) t = spoa::AlignmentType::kROV;
        engine->changeAlignType(t);
        auto alignment = engine->align(s, graph);
        graph->add_alignment(alignment, s);
    }
    std::string c = graph->generate_consensus();
    if (c.size() > out_cap) return HYPO_E_OUT_CAP;
    std::memcpy(out, c.data(), c.size());
    *out_len = c.size();
    return HYPO_OK;
}

// Golden window streams: runs the reference on a batch and dumps every window with the REFERENCE'S
// OWN Window::operator<< (reference src/Window.cpp:63-84), framed the way
// Contig::generate_inspect_file frames it (reference src/Contig.cpp:373-451: ">name", "#regions",
// "==========(beg-end)<TAB>TYPE<TAB>" before each window; strong regions with zero counters and the
// sequence twice).  Used by tests/golden/make_golden.py to produce tests/golden/inspect_ref.txt.gz.
int hypo_ref_inspect_dump(const char* path, const char* contig, const int8_t scores[6],
                          const HypoWindowDesc* win, uint64_t n_win, const HypoArmDesc* arms,
                          const uint8_t* packed) {
    ensure_engines(scores, 1);
    std::ofstream ofs(path);
    if (!ofs.is_open()) return HYPO_E_ARG;
    ofs << ">" << contig << std::endl;
    ofs << "#" << 2 * n_win << std::endl;
    uint64_t curr = 0;
    const std::string sr = "ACGTTGCA";
    for (uint64_t w = 0; w < n_win; ++w) {
        const HypoWindowDesc& d = win[w];
        ofs << "==========(" << curr << "-" << curr + sr.size() - 1 << ")\t" << "SR" << "\t" << 0 << "\t" << 0
            << "\t" << 0 << "\t" << 0 << std::endl;
        ofs << "++\t" << sr << std::endl;
        ofs << "++\t" << sr << std::endl;
        curr += sr.size();
        hypo::PackedSeq<4> draft(unpack4(packed + d.draft_off, d.draft_len));
        const bool lng = d.wtype == HYPO_WINDOW_LONG;
        hypo::Window W(draft, 0, d.draft_len, lng ? hypo::WindowType::LONG : hypo::WindowType::SHORT);
        uint64_t a = d.first_arm;
        for (uint32_t i = 0; i < d.n_internal; ++i, ++a) W.add_internal(hypo::PackedSeq<2>(unpack2(packed + arms[a].off, arms[a].len)));
        for (uint32_t i = 0; i < d.n_pre; ++i, ++a) W.add_prefix(hypo::PackedSeq<2>(unpack2(packed + arms[a].off, arms[a].len)));
        for (uint32_t i = 0; i < d.n_suf; ++i, ++a) W.add_suffix(hypo::PackedSeq<2>(unpack2(packed + arms[a].off, arms[a].len)));
        for (uint32_t i = 0; i < d.n_empty; ++i) W.add_empty();
        W.generate_consensus(0);
        ofs << "==========(" << curr << "-" << curr + W.get_window_len() - 1 << ")\t" << (lng ? "LNG" : "OTH") << "\t";
        ofs << W;
        curr += W.get_window_len();
    }
    return ofs.good() ? HYPO_OK : HYPO_E_ARG;
}

// Behaviour of the reference's PackedSeq on the data-format side of the path (reference
// src/PackedSeq.cpp:122-229): a read as htslib stores it (4-bit codes, two per byte) packed from `offset`,
// and sub-ranges of packed sequences, as hypo::Alignment / hypo::Contig build arms and window drafts.
// Outputs (ASCII, caller-sized): out2 / out4 = the read as PackedSeq<2> / PackedSeq<4> (seq_len chars),
// sub2 = PackedSeq<2>(read2, left, right), sub24 = PackedSeq<2>(read4, left, right), sub4 =
// PackedSeq<4>(read4, left, right), rng4 = read4.unpack(left, right) (right-left chars each).
// flags bit 0: PackedSeq<2> of the read is valid (out2 / sub2 filled); bit 1: [left, right) of the read holds
// only A/C/G/T (sub24 filled - the reference exits on anything else).
int hypo_ref_packedseq_probe(const uint8_t* hts, uint32_t seq_len, uint32_t offset, uint32_t left, uint32_t right,
                             char* out2, char* out4, char* sub2, char* sub24, char* sub4, char* rng4) {
    using namespace hypo;
    int flags = 0;
    PackedSeq<4> r4(seq_len, offset, hts);
    const std::string s4 = r4.unpack();
    memcpy(out4, s4.data(), s4.size());
    PackedSeq<2> r2(seq_len, offset, hts);
    if (r2.is_valid()) {
        flags |= 1;
        const std::string s2 = r2.unpack();
        memcpy(out2, s2.data(), s2.size());
        PackedSeq<2> p(r2, left, right);
        const std::string t = p.unpack();
        memcpy(sub2, t.data(), t.size());
    }
    bool clean = true;
    for (uint32_t i = left; i < right; ++i) clean = clean && r4.enc_base_at(i) < 4;
    if (clean) {
        flags |= 2;
        PackedSeq<2> p(r4, left, right);
        const std::string t = p.unpack();
        memcpy(sub24, t.data(), t.size());
    }
    PackedSeq<4> q(r4, left, right);
    const std::string u = q.unpack();
    memcpy(sub4, u.data(), u.size());
    const std::string v = r4.unpack(left, right);
    memcpy(rng4, v.data(), v.size());
    return flags;
}

// The reference's k-mer search on packed sequences (reference src/PackedSeq.cpp:264-414), used by arm
// extraction to anchor read segments.  nb = 2 / 4: PackedSeq<2> / PackedSeq<4> built from `seq`;
// mode 0 find_kmer, 1 check_kmer (at `left`), 2 find_canonical_kmer, 3 check_canonical_kmer.
// Returns found; *result = position for the find modes (all ones when not found / check modes).
int hypo_ref_kmer_probe(const char* seq, uint32_t len, int nb, int mode, uint64_t target, uint32_t k, uint32_t left,
                        uint32_t right, int is_first, uint64_t* result) {
    using namespace hypo;
    size_t at = (size_t)-1;
    bool found;
    const std::string s(seq, len);
    if (nb == 2) {
        PackedSeq<2> p(s);
        found = mode == 0 ? p.find_kmer(target, k, left, right, is_first != 0, at)
              : mode == 1 ? p.check_kmer(target, k, left)
              : mode == 2 ? p.find_canonical_kmer(target, k, left, right, is_first != 0, at)
                          : p.check_canonical_kmer(target, k, left);
    } else {
        PackedSeq<4> p(s);
        found = mode == 0 ? p.find_kmer(target, k, left, right, is_first != 0, at)
              : mode == 1 ? p.check_kmer(target, k, left)
              : mode == 2 ? p.find_canonical_kmer(target, k, left, right, is_first != 0, at)
                          : p.check_canonical_kmer(target, k, left);
    }
    *result = found ? (uint64_t)at : ~0ull;
    return found ? 1 : 0;
}

// The bookkeeping side of the reference's Window (reference include/Window.hpp:61-120): what the
// counters report after the arms of a batch went through add_* (LONG windows filter them), and after
// clear_pre_suf.  counts[8 * w + ...] = num_pre, num_suf, num_internal (incl. empty), num_total,
// maxlen_pre, maxlen_suf, window_len, num_total after clear_pre_suf.
void hypo_ref_window_counts(const HypoWindowDesc* win, uint64_t n_win, const HypoArmDesc* arms, const uint8_t* packed,
                            uint32_t* counts) {
    for (uint64_t w = 0; w < n_win; ++w) {
        const HypoWindowDesc& d = win[w];
        hypo::PackedSeq<4> draft(unpack4(packed + d.draft_off, d.draft_len));
        hypo::Window W(draft, 0, d.draft_len, d.wtype == HYPO_WINDOW_LONG ? hypo::WindowType::LONG : hypo::WindowType::SHORT);
        uint64_t a = d.first_arm;
        for (uint32_t i = 0; i < d.n_internal; ++i, ++a) W.add_internal(hypo::PackedSeq<2>(unpack2(packed + arms[a].off, arms[a].len)));
        for (uint32_t i = 0; i < d.n_pre; ++i, ++a) W.add_prefix(hypo::PackedSeq<2>(unpack2(packed + arms[a].off, arms[a].len)));
        for (uint32_t i = 0; i < d.n_suf; ++i, ++a) W.add_suffix(hypo::PackedSeq<2>(unpack2(packed + arms[a].off, arms[a].len)));
        for (uint32_t i = 0; i < d.n_empty; ++i) W.add_empty();
        uint32_t* c = counts + 8 * w;
        c[0] = W.get_num_pre(); c[1] = W.get_num_suf(); c[2] = W.get_num_internal(); c[3] = W.get_num_total();
        c[4] = W.get_maxlen_pre(); c[5] = W.get_maxlen_suf(); c[6] = (uint32_t)W.get_window_len();
        W.clear_pre_suf();
        c[7] = W.get_num_total();
    }
}

int hypo_ref_max_threads(void) { return omp_get_max_threads(); }

// 1 if this build of the reference selected spoa's SIMD engine (built with
// -march=native), 0 for the SISD engine (default flags; the parity oracle).
int hypo_ref_is_simd(void) {
#if defined(__AVX2__) || defined(__SSE4_1__)
    return 1;
#else
    return 0;
#endif
}

}  // extern "C"
