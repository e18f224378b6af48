// Window.cpp — see Window.hpp / WindowBatch.hpp.
#include "Window.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <omp.h>

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
    for (const auto& a : wnd._internal_arms) os << a.unpack() << std::endl;
    for (const auto& a : wnd._pre_arms) os << a.unpack() << std::endl;
    for (const auto& a : wnd._suf_arms) os << a.unpack() << std::endl;
    return os;
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

void WindowBatch::pack(int threads) {
    const size_t n = _windows.size();
    if (_n_packed == n) return;
    if (threads <= 0) threads = omp_get_max_threads();
    // pass 1: arms and bytes per window
    std::vector<uint64_t> arm0(n + 1), byte0(n + 1);
    arm0[0] = 0; byte0[0] = 0;
#pragma omp parallel for schedule(static, 512) num_threads(threads)
    for (size_t i = 0; i < n; ++i) {
        const Window* w = _windows[i];
        uint64_t bytes = w->_draft.data_size();
        for (const auto* v : {&w->_internal_arms, &w->_pre_arms, &w->_suf_arms})
            for (const auto& a : *v) bytes += a.data_size();
        arm0[i + 1] = w->_internal_arms.size() + w->_pre_arms.size() + w->_suf_arms.size();
        byte0[i + 1] = bytes;
    }
    for (size_t i = 0; i < n; ++i) { arm0[i + 1] += arm0[i]; byte0[i + 1] += byte0[i]; }
    _win.resize(n);
    _arms.resize(arm0[n]);
    _packed_bytes = byte0[n];
    _packed.reset(new uint8_t[_packed_bytes + 16]);
    uint8_t* const slab = _packed.get();
    // pass 2: every window fills its own slice (arms in container order: internal, prefix, suffix)
#pragma omp parallel for schedule(static, 512) num_threads(threads)
    for (size_t i = 0; i < n; ++i) {
        const Window* w = _windows[i];
        uint64_t pos = byte0[i];
        HypoWindowDesc d;
        memset(&d, 0, sizeof(d));
        d.draft_off = pos;
        d.draft_len = (uint32_t)w->_draft.get_seq_size();
        memcpy(slab + pos, w->_draft.data(), w->_draft.data_size());
        pos += w->_draft.data_size();
        d.first_arm = arm0[i];
        d.n_internal = (uint32_t)w->_internal_arms.size();
        d.n_pre = (uint32_t)w->_pre_arms.size();
        d.n_suf = (uint32_t)w->_suf_arms.size();
        d.n_empty = w->_num_empty;
        d.wtype = w->_wtype == WindowType::LONG ? HYPO_WINDOW_LONG : HYPO_WINDOW_SHORT;
        HypoArmDesc* ad = _arms.data() + arm0[i];
        for (const auto* v : {&w->_internal_arms, &w->_pre_arms, &w->_suf_arms})
            for (const auto& a : *v) {
                ad->off = pos;
                ad->len = (uint32_t)a.get_seq_size();
                ad->reserved = 0;
                memcpy(slab + pos, a.data(), a.data_size());
                pos += a.data_size();
                ++ad;
            }
        _win[i] = d;
    }
    _n_packed = n;
}

void WindowBatch::run() {
    if (_windows.empty()) return;
    pack();
    const uint64_t cap = hypo_gpu_out_bound(_win.data(), _win.size(), _arms.data(), _arms.size());
    _out.resize(cap + 16);
    _off.resize(_win.size() + 1);
    if (hypo_gpu_consensus_batch(_win.data(), _win.size(), _arms.data(), _arms.size(), _packed.get(), _packed_bytes,
                                 _out.data(), _out.size(), _off.data()) != HYPO_OK)
        die("POA of windows");
    const size_t n = _windows.size();
#pragma omp parallel for schedule(static, 512)
    for (size_t i = 0; i < n; ++i)
        _windows[i]->set_consensus(std::string(_out.data() + _off[i], _out.data() + _off[i + 1]));
}

}  // namespace hypo
