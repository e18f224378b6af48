// Window.cpp — see Window.hpp / WindowBatch.hpp.
#include "Window.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <omp.h>

#include <algorithm>
#include <thread>

#include "WindowBatch.hpp"

namespace hypo {

Window::ArmFilter Window::_long_filter;

static void die(const char* what) {
    fprintf(stderr, "[Hypo::GPU] Error: %s: %s\n", what, hypo_gpu_last_error());
    exit(1);
}

void Window::prepare_for_poa(const ScoreParams& sp, const UINT32 /*num_threads*/, int device) {
    const int8_t s[6] = {sp.sr_match_score, sp.sr_misMatch_score, sp.sr_gap_penalty,
                         sp.lr_match_score, sp.lr_misMatch_score, sp.lr_gap_penalty};
    if (hypo_gpu_init(s, device) != HYPO_OK) die("prepare_for_poa");
}

void Window::generate_consensus(const UINT32 /*engine_idx*/) {
    std::vector<Window*> one{this};
    generate_consensus_batch(one);
}

void Window::generate_consensus_batch(const std::vector<Window*>& windows) {
    WindowBatch b;
    for (Window* w : windows) b.add(w);
    b.run();
}

// reference src/Window.cpp:63-84 (the dump format of Contig::generate_inspect_file)
std::ostream& operator<<(std::ostream& os, const Window& wnd) {
    os << wnd._num_internal << "\t" << wnd._num_pre << "\t" << wnd._num_suf << "\t" << wnd._num_empty << std::endl;
    os << "++\t" << wnd._draft.unpack() << std::endl;
    os << "++\t" << wnd._consensus << std::endl;
    wnd.for_each_arm([&](const BYTE* p, UINT32 n) { os << ArmStore::unpack(p, n) << std::endl; });
    return os;
}

WindowBatch::~WindowBatch() {
    for (Slot& sl : _slot) sl.release();
}

void WindowBatch::Slot::release() {
    hypo_gpu_host_free(win); hypo_gpu_host_free(arms); hypo_gpu_host_free(packed);
    hypo_gpu_host_free(out); hypo_gpu_host_free(off);
    *this = Slot();
}

void WindowBatch::clear() {
    _windows.clear(); _win.clear(); _arms.clear(); _packed.reset(); _packed_bytes = 0; _n_packed = 0; _bp = 0;
}

void WindowBatch::reserve(size_t n_windows, size_t /*n_arms*/, size_t /*bytes*/) {
    _windows.reserve(n_windows);
}

void WindowBatch::add(Window* w) {
    _windows.push_back(w);
    _bp += w->_draft.get_seq_size();
}

namespace {

// Sizes of windows [first, first + n): prefix sums of arms and bytes (n + 1 entries each).
// The windows are separate heap objects, each with three arm stores of their own: walking them is a chain of
// cache misses (the packer ran at ~1 GB/s per thread).  Both passes therefore prefetch ahead: the object of
// the window a few places on, and the arm bytes / lengths of the one in between (whose object has arrived).
inline void prefetch_window(Window* const* ws, size_t i, size_t n, bool data) {
    constexpr size_t kFar = 8, kNear = 4;
    if (i + kFar < n) {
        const char* o = reinterpret_cast<const char*>(ws[i + kFar]);
        __builtin_prefetch(o); __builtin_prefetch(o + 64); __builtin_prefetch(o + 128);
    }
    if (data && i + kNear < n) {
        const Window* w = ws[i + kNear];
        __builtin_prefetch(w->draft().data());
        for (int k = 0; k < 3; ++k) {
            const ArmStore& st = w->arms(k);
            if (st.empty()) continue;
            const char* b = reinterpret_cast<const char*>(st.bytes.data());
            const size_t nb = st.bytes.size();
            for (size_t off = 0; off < nb && off < 1024; off += 64) __builtin_prefetch(b + off);
            __builtin_prefetch(st.len.data());
            __builtin_prefetch(reinterpret_cast<const char*>(st.len.data()) + 64);
        }
    }
}

void measure(Window* const* ws, size_t n, int threads, std::vector<uint64_t>& arm0, std::vector<uint64_t>& byte0) {
    arm0.assign(n + 1, 0);
    byte0.assign(n + 1, 0);
#pragma omp parallel for schedule(static, 512) num_threads(threads)
    for (size_t i = 0; i < n; ++i) {
        prefetch_window(ws, i, n, false);
        const Window* w = ws[i];
        uint64_t bytes = w->draft().data_size(), arms = 0;
        for (int k = 0; k < 3; ++k) { bytes += w->arms(k).bytes.size(); arms += w->arms(k).size(); }
        arm0[i + 1] = arms;
        byte0[i + 1] = bytes;
    }
    for (size_t i = 0; i < n; ++i) { arm0[i + 1] += arm0[i]; byte0[i + 1] += byte0[i]; }
}

// Every window fills its own slice (arms in container order: internal, prefix, suffix).
// Returns the sum of the windows' output bounds (the rule of hypo_gpu_window_bounds).
uint64_t fill(Window* const* ws, size_t n, int threads, const std::vector<uint64_t>& arm0, const std::vector<uint64_t>& byte0,
              HypoWindowDesc* win, HypoArmDesc* arms, uint8_t* slab) {
    uint64_t bound = 0;
#pragma omp parallel for schedule(static, 512) num_threads(threads) reduction(+ : bound)
    for (size_t i = 0; i < n; ++i) {
        prefetch_window(ws, i, n, true);
        const Window* w = ws[i];
        uint64_t pos = byte0[i];
        HypoWindowDesc d;
        memset(&d, 0, sizeof(d));
        d.draft_off = pos;
        d.draft_len = (uint32_t)w->draft().get_seq_size();
        memcpy(slab + pos, w->draft().data(), w->draft().data_size());
        pos += w->draft().data_size();
        d.first_arm = arm0[i];
        w->counts(d.n_internal, d.n_pre, d.n_suf, d.n_empty);
        d.wtype = w->get_type() == WindowType::LONG ? HYPO_WINDOW_LONG : HYPO_WINDOW_SHORT;
        HypoArmDesc* ad = arms + arm0[i];
        uint64_t bases = 0;
        for (int k = 0; k < 3; ++k) {   // one copy per kind; the descriptors follow from the lengths
            const ArmStore& st = w->arms(k);
            if (st.empty()) continue;
            memcpy(slab + pos, st.bytes.data(), st.bytes.size());
            for (UINT32 len : st.len) {
                ad->off = pos;
                ad->len = len;
                ad->reserved = 0;
                pos += ArmStore::arm_bytes(len);
                bases += len;
                ++ad;
            }
        }
        win[i] = d;
        const uint64_t n_arms = arm0[i + 1] - arm0[i];
        bound += d.wtype == HYPO_WINDOW_LONG ? 2 * bases + d.draft_len + 2 : bases + 2 * n_arms + d.draft_len + 2;
    }
    return bound;
}

template <class T>
void grow(T*& p, size_t& cap, size_t need) {
    if (need <= cap) return;
    hypo_gpu_host_free(p);
    cap = need + need / 4 + 64;
    p = static_cast<T*>(hypo_gpu_host_alloc(cap * sizeof(T)));
    if (!p) {
        fprintf(stderr, "[Hypo::GPU] Error: cannot allocate %zu bytes of page-locked host memory\n", cap * sizeof(T));
        exit(1);
    }
}

}  // namespace

void WindowBatch::pack(int threads) {
    const size_t n = _windows.size();
    if (_n_packed == n) return;
    if (threads <= 0) threads = omp_get_max_threads();
    std::vector<uint64_t> arm0, byte0;
    measure(_windows.data(), n, threads, arm0, byte0);
    _win.resize(n);
    _arms.resize(arm0[n]);
    _packed_bytes = byte0[n];
    _packed.reset(new uint8_t[_packed_bytes + 16]);
    fill(_windows.data(), n, threads, arm0, byte0, _win.data(), _arms.data(), _packed.get());
    _n_packed = n;
}

void WindowBatch::pack_chunk(Slot& s, size_t first, size_t n, int threads) {
    std::vector<uint64_t> arm0, byte0;
    measure(_windows.data() + first, n, threads, arm0, byte0);
    grow(s.win, s.win_cap, n);
    grow(s.arms, s.arms_cap, (size_t)arm0[n] + 1);
    grow(s.packed, s.packed_cap, (size_t)byte0[n] + 16);
    grow(s.off, s.off_cap, n + 1);
    const uint64_t bound = fill(_windows.data() + first, n, threads, arm0, byte0, s.win, s.arms, s.packed);
    s.first = first; s.n_win = n; s.n_arms = arm0[n]; s.n_bytes = byte0[n];
    grow(s.out, s.out_cap, (size_t)bound + 16);
}

void WindowBatch::scatter_chunk(const Slot& s, int threads) {
#pragma omp parallel for schedule(static, 512) num_threads(threads)
    for (size_t i = 0; i < s.n_win; ++i)
        _windows[s.first + i]->set_consensus(std::string(s.out + s.off[i], s.out + s.off[i + 1]));
}

void WindowBatch::run(size_t chunk_windows) {
    _timing = Timing();
    const size_t n = _windows.size();
    if (n == 0) return;
    const double t_begin = omp_get_wtime();
    const int threads = omp_get_max_threads();
    const int threads_ov = std::max(1, threads - 1);   // while the device thread spins in its synchronisations
    // Chunk boundaries.  A batch of fewer than 128 K windows per driven device goes in one piece (cutting it
    // starves the persistent kernels).  Otherwise a small first chunk (32 K windows per device: its pack is
    // the only host work the device cannot hide), then equal chunks of at most 256 K windows per device, at
    // least three: with two calls in flight the per-call costs (copy of a head, compaction, result copy,
    // kernel tails) hide behind the other call's kernels, so chunks may be small enough to keep the two
    // lanes' buffers modest (measured: tools/chunk_sweep.py, profiles/).
    // The counts below are for windows of the headline shape (30 reads x 120 bp); a chunk is meant to be a
    // certain amount of device WORK - per-chunk costs (tier launches and their tails, classification, copies)
    // are fixed - so for cheaper windows they scale up: `scale` = cost of a headline window / average cost
    // here, from a strided sample of reads x (length^2 + a per-read constant).
    double scale = 1.0;
    {
        double sum = 0;
        size_t cnt = 0;
        for (size_t i = 0; i < n; i += 61, ++cnt) {
            const Window* w = _windows[i];
            double arms = 0;
            for (int k = 0; k < 3; ++k) arms += (double)w->arms(k).size();
            const double len = (double)w->draft().get_seq_size() + 2.0;
            sum += arms * (len * len + 300.0) + 300.0;
        }
        const double avg = sum / (double)std::max<size_t>(cnt, 1), headline = 30.0 * (122.0 * 122.0 + 300.0);
        scale = std::min(4.0, std::max(1.0, headline / std::max(avg, 1.0)));
    }
    const size_t ndev = (size_t)std::max(1, hypo_gpu_device_count()) * (size_t)1;
    const size_t unit = (size_t)((double)ndev * scale + 0.5);   // "devices x scale": what the counts are multiplied by
    std::vector<size_t> cut{0};
    if (chunk_windows) {
        while (cut.back() < n) cut.push_back(std::min(n, cut.back() + chunk_windows));
    } else if (n < 131072 * unit && !(threads <= 8 && n >= 49152 * unit)) {
        cut.push_back(n);
    } else if (n < 131072 * unit) {
        // few host threads (several ranks share the host): packing is a visible part of the call, so even a
        // medium batch is cut - a small first chunk, then three equal ones - to overlap it with the device
        const size_t first = std::max<size_t>(8192, n / 8);
        cut.push_back(first);
        for (size_t i = 1; i <= 3; ++i) cut.push_back(first + (n - first) * i / 3);
    } else {
        const size_t first = 32768 * std::min(unit, 4 * ndev);
        cut.push_back(first);
        const size_t rest = n - first;
        const size_t k = rest < 196608 * unit ? 1 : std::max<size_t>(3, (rest + 262144 * unit - 1) / (262144 * unit));
        for (size_t i = 1; i <= k; ++i) cut.push_back(first + rest * i / k);
    }
    const size_t n_chunks = cut.size() - 1;
    auto chunk_lo = [&](size_t k) { return cut[std::min(k, n_chunks)]; };

    // Two chunks are in flight at any time (the library runs two calls side by side, one per lane): while
    // chunk k is on the device, chunk k+1 is already queued behind it - its copies and classification run
    // during chunk k's kernels, its CTAs fill the SMs chunk k's kernel frees in its tail - and this thread
    // packs chunk k+2 into the third slot, then scatters chunk k-1... in short: the device never waits.
    struct Flight {
        std::thread th;
        int rc = HYPO_OK;
        std::string err;
        double sec = 0;
    };
    std::vector<Flight> fl(n_chunks);
    auto start = [&](size_t k) {
        Slot& cur = _slot[k % 3];
        Flight& f = fl[k];
        f.th = std::thread([&cur, &f]() {
            const double d0 = omp_get_wtime();
            f.rc = hypo_gpu_consensus_batch(cur.win, cur.n_win, cur.arms, cur.n_arms, cur.packed, cur.n_bytes, cur.out,
                                            cur.out_cap, cur.off);
            if (f.rc != HYPO_OK) f.err = hypo_gpu_last_error();
            f.sec = omp_get_wtime() - d0;
        });
    };
    double t0 = omp_get_wtime();
    pack_chunk(_slot[0], 0, chunk_lo(1), threads);
    _timing.pack += omp_get_wtime() - t0;
    start(0);
    if (n_chunks > 1) {
        t0 = omp_get_wtime();
        pack_chunk(_slot[1], chunk_lo(1), chunk_lo(2) - chunk_lo(1), threads_ov);
        _timing.pack += omp_get_wtime() - t0;
        start(1);
    }
    for (size_t k = 0; k < n_chunks; ++k) {
        fl[k].th.join();
        _timing.device += fl[k].sec;
        if (fl[k].rc != HYPO_OK) {
            for (size_t j = k + 1; j < n_chunks; ++j) if (fl[j].th.joinable()) fl[j].th.join();
            fprintf(stderr, "[Hypo::GPU] Error: POA of windows: %s\n", fl[k].err.c_str());
            exit(1);
        }
        if (k + 2 < n_chunks) {   // slot (k+2) % 3 == (k-1) % 3 was scattered one iteration ago
            t0 = omp_get_wtime();
            pack_chunk(_slot[(k + 2) % 3], chunk_lo(k + 2), chunk_lo(k + 3) - chunk_lo(k + 2), threads_ov);
            _timing.pack += omp_get_wtime() - t0;
            start(k + 2);
        }
        t0 = omp_get_wtime();
        scatter_chunk(_slot[k % 3], k + 1 < n_chunks ? threads_ov : threads);
        _timing.scatter += omp_get_wtime() - t0;
    }
    _timing.chunks = n_chunks;
    _timing.total = omp_get_wtime() - t_begin;
}

}  // namespace hypo
