// host_capi.cpp — C entry points of libhypo_host.so used by bench.py and the tests:
//   * hypo_synth_*      seeded synthetic window generator (BASELINE.md §3 / SURVEY.md §8d shapes)
//   * hypo_host_run     drives the hypo::Window mirror through its PUBLIC API (add_* /
//                       prepare_for_poa / generate_consensus_batch / get_consensus) from a flat
//                       batch, the way oracle/ref_driver.cpp drives the reference's Window.
#include <omp.h>

#include <cstdint>
#include <chrono>
#include <cstring>
#include <fstream>
#include <memory>
#include <string>
#include <vector>

#include "../../include/hypo_b200.h"
#include "Window.hpp"
#include "WindowBatch.hpp"
#include "WindowStream.hpp"

namespace {

struct Rng {   // splitmix64
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    uint64_t next() {
        uint64_t z = (s += 0x9e3779b97f4a7c15ull);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        return z ^ (z >> 31);
    }
    double uni() { return (next() >> 11) * (1.0 / 9007199254740992.0); }
    uint32_t below(uint32_t n) { return (uint32_t)(next() % n); }
};

// truth -> noisy copy: per base delete / substitute / keep, then maybe insert a random base
void mutate(Rng& r, const uint8_t* src, uint32_t n, double sub, double ins, double del, std::vector<uint8_t>& dst) {
    dst.clear();
    for (uint32_t i = 0; i < n; ++i) {
        const double u = r.uni();
        if (u >= del) dst.push_back(u < del + sub ? (uint8_t)((src[i] + 1 + r.below(3)) & 3) : src[i]);
        if (r.uni() < ins) dst.push_back((uint8_t)r.below(4));
    }
}

struct GenCfg {
    uint64_t seed;
    uint32_t len, n_arms;
    int kind;   // 0 internal, 1 backbone, 2 prefix-heavy, 3 suffix-heavy, 4 mixed
    double err, draft_err;
    uint32_t wtype;
};

int arm_kind(const GenCfg& c, uint32_t r) {   // 0 internal, 1 prefix, 2 suffix
    switch (c.kind) {
        case 0: return 0;
        case 1: return r < c.n_arms / 2 ? 1 : 2;
        case 2: return r < 3 ? 0 : 1;
        case 3: return r < 3 ? 0 : 2;
        default: { static const int k[5] = {0, 0, 0, 1, 2}; return k[r % 5]; }
    }
}

// Generates window w; calls sink(kind, codes, n) for the draft (kind -1) and every arm in
// generation order.  Deterministic in (seed, w) so the two passes agree and the result is
// independent of the thread count.
template <class Sink>
void gen_window(const GenCfg& c, uint64_t w, Sink&& sink) {
    Rng r(c.seed * 0x9e3779b97f4a7c15ull + w * 0xd1342543de82ef95ull + 1);
    std::vector<uint8_t> truth(c.len), buf;
    for (auto& b : truth) b = (uint8_t)r.below(4);
    mutate(r, truth.data(), c.len, c.draft_err, c.draft_err, c.draft_err, buf);
    if (buf.empty()) buf.push_back(0);
    sink(-1, buf.data(), (uint32_t)buf.size());
    for (uint32_t a = 0; a < c.n_arms; ++a) {
        const int k = arm_kind(c, a);
        if (k == 0) {
            mutate(r, truth.data(), c.len, c.err, c.err, c.err, buf);
        } else {
            const uint32_t lo = c.len / 2 > 0 ? c.len / 2 : 1;
            const uint32_t cut = c.len > 1 ? lo + r.below(c.len - lo) : 1;
            if (k == 1) mutate(r, truth.data(), cut, c.err, c.err, c.err, buf);
            else mutate(r, truth.data() + (c.len - cut), cut, c.err, c.err, c.err, buf);
        }
        sink(k, buf.data(), (uint32_t)buf.size());
    }
}

std::string unpack2(const uint8_t* p, uint32_t len) {
    std::string s(len, 'A');
    for (uint32_t i = 0; i < len; ++i) s[i] = "ACGT"[(p[i >> 2] >> (6 - 2 * (i & 3))) & 3];
    return s;
}
std::string unpack4(const uint8_t* p, uint32_t len) {
    std::string s(len, 'A');
    for (uint32_t i = 0; i < len; ++i) { int v = (p[i >> 1] >> ((i & 1) ? 0 : 4)) & 15; s[i] = "ACGTN"[v > 4 ? 4 : v]; }
    return s;
}

}  // namespace

extern "C" {

// Worst-case packed bytes for the configuration (every base followed by an insertion).
uint64_t hypo_synth_packed_bound(uint64_t n_win, uint32_t len, uint32_t n_arms) {
    return n_win * ((uint64_t)(2 * len + 2) / 2 + 1 + (uint64_t)n_arms * ((2 * len + 2) / 4 + 1)) + 64;
}

// Fills win[n_win], arms[n_win*n_arms] (container order: internal, prefix, suffix) and the
// packed slab.  Returns 0, or 1 if packed_cap is too small.
int hypo_synth_generate(uint64_t seed, uint64_t n_win, uint32_t len, uint32_t n_arms, int kind, double err,
                        double draft_err, uint32_t wtype, HypoWindowDesc* win, HypoArmDesc* arms, uint8_t* packed,
                        uint64_t packed_cap, uint64_t* packed_used, int n_threads) {
    GenCfg c{seed, len, n_arms, kind, err, draft_err, wtype};
    if (n_threads <= 0) n_threads = omp_get_max_threads();
    std::vector<uint64_t> bytes(n_win + 1, 0);
    // pass 1: sizes
#pragma omp parallel for schedule(static, 256) num_threads(n_threads)
    for (uint64_t w = 0; w < n_win; ++w) {
        uint64_t b = 0;
        gen_window(c, w, [&](int k, const uint8_t*, uint32_t n) { b += k < 0 ? (n + 1) / 2 : (n + 3) / 4; });
        bytes[w + 1] = b;
    }
    for (uint64_t w = 0; w < n_win; ++w) bytes[w + 1] += bytes[w];
    *packed_used = bytes[n_win];
    if (bytes[n_win] > packed_cap) return 1;
    // pass 2: fill
#pragma omp parallel for schedule(static, 256) num_threads(n_threads)
    for (uint64_t w = 0; w < n_win; ++w) {
        uint64_t pos = bytes[w];
        HypoWindowDesc& d = win[w];
        memset(&d, 0, sizeof(d));
        d.first_arm = w * n_arms;
        d.wtype = wtype;
        uint32_t cnt[3] = {0, 0, 0};
        for (uint32_t a = 0; a < n_arms; ++a) cnt[arm_kind(c, a)]++;
        d.n_internal = cnt[0]; d.n_pre = cnt[1]; d.n_suf = cnt[2];
        uint32_t slot[3] = {0, cnt[0], cnt[0] + cnt[1]};
        gen_window(c, w, [&](int k, const uint8_t* s, uint32_t n) {
            if (k < 0) {
                d.draft_off = pos; d.draft_len = n;
                const uint64_t nb = (n + 1) / 2;
                memset(packed + pos, 0, nb);
                for (uint32_t i = 0; i < n; ++i) packed[pos + (i >> 1)] |= (uint8_t)(s[i] << ((i & 1) ? 0 : 4));
                pos += nb;
            } else {
                HypoArmDesc& ad = arms[w * n_arms + slot[k]++];
                ad.off = pos; ad.len = n; ad.reserved = 0;
                const uint64_t nb = (n + 3) / 4;
                memset(packed + pos, 0, nb);
                for (uint32_t i = 0; i < n; ++i) packed[pos + (i >> 2)] |= (uint8_t)(s[i] << (6 - 2 * (i & 3)));
                pos += nb;
            }
        });
    }
    return 0;
}

// Drives the Window mirror through its public API.  Mirrors hypo_ref_consensus_batch.
}  // extern "C"

namespace {
// A flat batch rebuilt as hypo::Window objects through the public API (what the reference's pipeline does
// while it walks the alignments).
std::vector<std::unique_ptr<hypo::Window>> build_windows(const HypoWindowDesc* win, uint64_t n_win,
                                                         const HypoArmDesc* arms, const uint8_t* packed) {
    using namespace hypo;
    std::vector<std::unique_ptr<Window>> ws(n_win);
#pragma omp parallel for schedule(static, 64)
    for (uint64_t w = 0; w < n_win; ++w) {
        const HypoWindowDesc& d = win[w];
        PackedSeq<4> draft(unpack4(packed + d.draft_off, d.draft_len));
        ws[w].reset(new Window(draft, 0, d.draft_len, d.wtype == HYPO_WINDOW_LONG ? WindowType::LONG : WindowType::SHORT));
        uint64_t a = d.first_arm;
        for (uint32_t i = 0; i < d.n_internal; ++i, ++a) ws[w]->add_internal(PackedSeq<2>(unpack2(packed + arms[a].off, arms[a].len)));
        for (uint32_t i = 0; i < d.n_pre; ++i, ++a) ws[w]->add_prefix(PackedSeq<2>(unpack2(packed + arms[a].off, arms[a].len)));
        for (uint32_t i = 0; i < d.n_suf; ++i, ++a) ws[w]->add_suffix(PackedSeq<2>(unpack2(packed + arms[a].off, arms[a].len)));
        for (uint32_t i = 0; i < d.n_empty; ++i) ws[w]->add_empty();
    }
    return ws;
}
}  // namespace

extern "C" {

// CPU-only check and timing of the batch packer: flat batch -> hypo::Window objects -> WindowBatch::pack
// with `threads` threads -> flat buffers again (win/arms/packed must be as large as the inputs).  The
// packer writes the slab in container order (draft, internal, prefix, suffix arms); for a batch that is
// laid out that way already the outputs equal the inputs byte for byte.  *seconds = time of pack() alone.
int hypo_host_pack(const HypoWindowDesc* win, uint64_t n_win, const HypoArmDesc* arms, uint64_t n_arms,
                   const uint8_t* packed, uint64_t packed_bytes, int threads, HypoWindowDesc* out_win,
                   HypoArmDesc* out_arms, uint8_t* out_packed, double* seconds) {
    using namespace hypo;
    auto ws = build_windows(win, n_win, arms, packed);
    WindowBatch b;
    b.reserve(n_win);
    for (auto& w : ws) b.add(w.get());
    const double t0 = omp_get_wtime();
    b.pack(threads);
    if (seconds) *seconds = omp_get_wtime() - t0;
    if (b.n_arms() != n_arms || b.packed_bytes() != packed_bytes) return HYPO_E_ARG;
    memcpy(out_win, b.win_desc(), n_win * sizeof(HypoWindowDesc));
    memcpy(out_arms, b.arm_desc(), n_arms * sizeof(HypoArmDesc));
    memcpy(out_packed, b.packed(), packed_bytes);
    return HYPO_OK;
}

// CPU only: the PackedSeq mirror on the data-format side of the path - same contract as
// hypo_ref_packedseq_probe (oracle/ref_driver.cpp), which answers with the reference's own class.
int hypo_host_packedseq_probe(const uint8_t* hts, uint32_t seq_len, uint32_t offset, uint32_t left, uint32_t right,
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

// CPU only: the Window mirror's counters with the reference's LONG-window arm filter switched on - same
// contract as hypo_ref_window_counts.
void hypo_host_window_counts(const HypoWindowDesc* win, uint64_t n_win, const HypoArmDesc* arms, const uint8_t* packed,
                             uint32_t* counts) {
    hypo::Window::use_reference_long_filter(true);
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
    hypo::Window::use_reference_long_filter(false);
}

// CPU only: k-mer search of the PackedSeq mirror - same contract as hypo_ref_kmer_probe.
int hypo_host_kmer_probe(const char* seq, uint32_t len, int nb, int mode, uint64_t target, uint32_t k, uint32_t left,
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

// CPU only: which arms the Window mirror keeps when it filters the arms of LONG windows like the
// reference's Window does (Window::use_reference_long_filter; reference include/Window.hpp:66-101,
// include/Filter.hpp).  accepted[a] = 1/0 per arm descriptor; SHORT windows keep every arm.  Returns
// HYPO_E_ARG if the windows built through add_* disagree with the filter applied arm by arm.
int hypo_host_long_filter(const HypoWindowDesc* win, uint64_t n_win, const HypoArmDesc* arms, const uint8_t* packed,
                          uint8_t* accepted) {
    using namespace hypo;
    Window::use_reference_long_filter(true);
    auto ws = build_windows(win, n_win, arms, packed);
    Window::use_reference_long_filter(false);
    int bad = 0;
#pragma omp parallel for schedule(static, 64) reduction(+ : bad)
    for (uint64_t w = 0; w < n_win; ++w) {
        const HypoWindowDesc& d = win[w];
        const uint64_t n = (uint64_t)d.n_internal + d.n_pre + d.n_suf;
        MinimizerFilter f;
        if (d.wtype == HYPO_WINDOW_LONG) f.init(unpack4(packed + d.draft_off, d.draft_len));
        uint32_t kept[3] = {0, 0, 0};
        for (uint64_t i = 0; i < n; ++i) {
            const HypoArmDesc& a = arms[d.first_arm + i];
            const bool ok = d.wtype != HYPO_WINDOW_LONG || f.accepts(unpack2(packed + a.off, a.len));
            accepted[d.first_arm + i] = ok;
            kept[i < d.n_internal ? 0 : i < (uint64_t)d.n_internal + d.n_pre ? 1 : 2] += ok;
        }
        bad += ws[w]->get_num_internal() != kept[0] + d.n_empty || ws[w]->get_num_pre() != kept[1] ||
               ws[w]->get_num_suf() != kept[2];
    }
    return bad ? HYPO_E_ARG : HYPO_OK;
}

int hypo_host_run(const int8_t scores[6], int device, const HypoWindowDesc* win, uint64_t n_win,
                  const HypoArmDesc* arms, const uint8_t* packed, char* out, uint64_t out_cap, uint64_t* out_off) {
    using namespace hypo;
    ScoreParams sp{scores[0], scores[1], scores[2], scores[3], scores[4], scores[5]};
    Window::prepare_for_poa(sp, 1, device);
    auto ws = build_windows(win, n_win, arms, packed);
    std::vector<Window*> ptrs(n_win);
    for (uint64_t w = 0; w < n_win; ++w) ptrs[w] = ws[w].get();
    Window::generate_consensus_batch(ptrs);
    uint64_t pos = 0;
    for (uint64_t w = 0; w < n_win; ++w) {
        out_off[w] = pos;
        const std::string c = ws[w]->get_consensus();
        if (pos + c.size() > out_cap) return HYPO_E_OUT_CAP;
        memcpy(out + pos, c.data(), c.size());
        pos += c.size();
    }
    out_off[n_win] = pos;
    return HYPO_OK;
}

// ---- the plugin call a maintainer makes, for bench.py: hypo::Window objects in, consensus strings in the
// same objects out (pack + copies + kernels + scatter), reference src/Hypo.cpp:236-248 ----------------
struct HostWindows {
    std::vector<std::unique_ptr<hypo::Window>> ws;
    hypo::WindowBatch batch;   // kept: its page-locked buffer sets are reused from step to step
};

void* hypo_host_windows_create(const HypoWindowDesc* win, uint64_t n_win, const HypoArmDesc* arms,
                               const uint8_t* packed) {
    HostWindows* h = new HostWindows;
    h->ws = build_windows(win, n_win, arms, packed);
    return h;
}

void hypo_host_windows_free(void* h) { delete static_cast<HostWindows*>(h); }

// One pass: every window through WindowBatch::run (hypo_gpu_init / hypo_gpu_init_multi must have been
// called).  timing[0..4] = pack, device call, scatter, total seconds, chunks.
int hypo_host_windows_run(void* hp, uint64_t chunk_windows, double timing[5]) {
    HostWindows& h = *static_cast<HostWindows*>(hp);
    h.batch.clear();
    h.batch.reserve(h.ws.size());
    for (auto& w : h.ws) h.batch.add(w.get());
    h.batch.run((size_t)chunk_windows);
    const hypo::WindowBatch::Timing& t = h.batch.last_timing();
    if (timing) { timing[0] = t.pack; timing[1] = t.device; timing[2] = t.scatter; timing[3] = t.total; timing[4] = (double)t.chunks; }
    return HYPO_OK;
}

// Concatenated Window::get_consensus() of all windows; returns the bytes needed (nothing is written
// beyond out_cap).
uint64_t hypo_host_windows_consensus(void* hp, char* out, uint64_t out_cap, uint64_t* out_off) {
    HostWindows& h = *static_cast<HostWindows*>(hp);
    uint64_t pos = 0;
    for (size_t w = 0; w < h.ws.size(); ++w) {
        const std::string c = h.ws[w]->get_consensus();
        if (out_off) out_off[w] = pos;
        if (out && pos + c.size() <= out_cap) memcpy(out + pos, c.data(), c.size());
        pos += c.size();
    }
    if (out_off) out_off[h.ws.size()] = pos;
    return pos;
}

// ---- window streams in the reference's inspect-file format (WindowStream.hpp) ---------------------

// Writes a flat batch as one contig's inspect file: every window preceded by a short strong region,
// as the pipeline interleaves them; `cons`/`cons_off` are the consensus strings to record.
int hypo_host_inspect_write(const char* path, const char* contig, const HypoWindowDesc* win, uint64_t n_win,
                            const HypoArmDesc* arms, const uint8_t* packed, const char* cons,
                            const uint64_t* cons_off) {
    using namespace hypo;
    WindowStream ws;
    ws.contig = contig ? contig : "ctg";
    uint64_t pos = 0;
    for (uint64_t w = 0; w < n_win; ++w) {
        const HypoWindowDesc& d = win[w];
        const std::string sr = "ACGTTGCA";
        ws.add_plain("SR", pos, sr);
        pos += sr.size();
        const bool lng = d.wtype == HYPO_WINDOW_LONG;
        PackedSeq<4> draft(unpack4(packed + d.draft_off, d.draft_len));
        std::unique_ptr<Window> wp(new Window(draft, 0, d.draft_len, lng ? WindowType::LONG : WindowType::SHORT));
        uint64_t a = d.first_arm;
        for (uint32_t i = 0; i < d.n_internal; ++i, ++a) wp->add_internal(PackedSeq<2>(unpack2(packed + arms[a].off, arms[a].len)));
        for (uint32_t i = 0; i < d.n_pre; ++i, ++a) wp->add_prefix(PackedSeq<2>(unpack2(packed + arms[a].off, arms[a].len)));
        for (uint32_t i = 0; i < d.n_suf; ++i, ++a) wp->add_suffix(PackedSeq<2>(unpack2(packed + arms[a].off, arms[a].len)));
        for (uint32_t i = 0; i < d.n_empty; ++i) wp->add_empty();
        ws.add_window(std::move(wp), lng ? "LNG" : "OTH", pos, std::string(cons + cons_off[w], cons + cons_off[w + 1]));
        pos += d.draft_len;
    }
    std::ofstream os(path);
    if (!os.is_open()) return HYPO_E_ARG;
    ws.write(os, true);
    return os.good() ? HYPO_OK : HYPO_E_ARG;
}

// Parses an inspect file.  Returns an opaque handle (nullptr on error, message in err[0..err_cap)).
void* hypo_host_inspect_open(const char* path, char* err, uint64_t err_cap) {
    using namespace hypo;
    std::ifstream in(path);
    std::string msg;
    std::unique_ptr<WindowStream> ws(new WindowStream);
    if (!in.is_open()) msg = std::string("cannot open ") + path;
    else if (ws->read(in, &msg)) return ws.release();
    if (err && err_cap) { strncpy(err, msg.c_str(), err_cap - 1); err[err_cap - 1] = 0; }
    return nullptr;
}

void hypo_host_inspect_close(void* h) { delete static_cast<hypo::WindowStream*>(h); }

// counts[0..5] = regions, windows, arms, packed bytes (as the batch packer lays them out),
// recorded consensus bytes, polished bp
void hypo_host_inspect_sizes(void* h, uint64_t counts[6]) {
    using namespace hypo;
    WindowStream& ws = *static_cast<WindowStream*>(h);
    WindowBatch b;
    for (auto& w : ws.windows) b.add(w.get());
    uint64_t cb = 0;
    for (auto& c : ws.recorded) cb += c.size();
    counts[0] = ws.regions.size(); counts[1] = ws.windows.size(); counts[2] = b.n_arms();
    counts[3] = b.packed_bytes(); counts[4] = cb; counts[5] = ws.polished_bp();
}

// The stream's windows as the flat batch of the C ABI plus the recorded consensus strings.
void hypo_host_inspect_fill(void* h, HypoWindowDesc* win, HypoArmDesc* arms, uint8_t* packed, char* cons,
                            uint64_t* cons_off) {
    using namespace hypo;
    WindowStream& ws = *static_cast<WindowStream*>(h);
    WindowBatch b;
    for (auto& w : ws.windows) b.add(w.get());
    memcpy(win, b.win_desc(), b.size() * sizeof(HypoWindowDesc));
    memcpy(arms, b.arm_desc(), b.n_arms() * sizeof(HypoArmDesc));
    memcpy(packed, b.packed(), b.packed_bytes());
    uint64_t pos = 0;
    for (size_t i = 0; i < ws.recorded.size(); ++i) {
        cons_off[i] = pos;
        memcpy(cons + pos, ws.recorded[i].data(), ws.recorded[i].size());
        pos += ws.recorded[i].size();
    }
    cons_off[ws.recorded.size()] = pos;
}

// Replays the stream on the device: one generate_consensus_batch over all windows.  Returns the number
// of windows whose consensus differs from the recorded one (-1: not initialised); *seconds = wall
// time of the batch call; if out_path is given the stream is written back with the NEW consensus.
int64_t hypo_host_inspect_replay(void* h, const int8_t scores[6], int device, double* seconds, const char* out_path) {
    using namespace hypo;
    WindowStream& ws = *static_cast<WindowStream*>(h);
    ScoreParams sp{scores[0], scores[1], scores[2], scores[3], scores[4], scores[5]};
    Window::prepare_for_poa(sp, 1, device);
    const auto t0 = std::chrono::steady_clock::now();
    const size_t bad = ws.replay();
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (out_path && *out_path) {
        std::ofstream os(out_path);
        if (os.is_open()) ws.write(os, false);
    }
    return (int64_t)bad;
}

// The stream's contig stitched (Contig::operator<<, reference src/Contig.cpp:345-366) from the windows'
// current consensus strings: mode 0 on the host, 1 by hypo_gpu_stitch, 2 by hypo_gpu_stitch from the device-
// resident result of the replay that just ran, 3 on the host from the RECORDED consensus strings.  Returns the length (nothing is written beyond out_cap).
uint64_t hypo_host_inspect_stitch(void* h, int mode, char* out, uint64_t out_cap) {
    hypo::WindowStream& ws = *static_cast<hypo::WindowStream*>(h);
    std::string s;
    if (mode == 3) {   // from the consensus strings the stream RECORDED (no device involved)
        for (const auto& r : ws.regions) s += r.window < 0 ? r.text : ws.recorded[r.window];
    } else {
        s = mode == 0 ? ws.stitched() : ws.stitched_on_device(mode == 2);
    }
    if (out && s.size() <= out_cap) memcpy(out, s.data(), s.size());
    return s.size();
}

}  // extern "C"
