// Window.cpp — see Window.hpp / WindowBatch.hpp.
#include "Window.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>

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
    _windows.clear(); _win.clear(); _arms.clear(); _packed.clear(); _bp = 0;
}

void WindowBatch::reserve(size_t n_windows, size_t n_arms, size_t bytes) {
    _windows.reserve(n_windows); _win.reserve(n_windows); _arms.reserve(n_arms); _packed.reserve(bytes);
}

uint64_t WindowBatch::put(const uint8_t* p, size_t n) {
    const uint64_t off = _packed.size();
    _packed.insert(_packed.end(), p, p + n);
    return off;
}

void WindowBatch::add(Window* w) {
    HypoWindowDesc d;
    memset(&d, 0, sizeof(d));
    d.draft_off = put(w->_draft.data(), w->_draft.data_size());
    d.draft_len = (uint32_t)w->_draft.get_seq_size();
    d.first_arm = _arms.size();
    d.n_internal = (uint32_t)w->_internal_arms.size();
    d.n_pre = (uint32_t)w->_pre_arms.size();
    d.n_suf = (uint32_t)w->_suf_arms.size();
    d.n_empty = w->_num_empty;
    d.wtype = w->_wtype == WindowType::LONG ? HYPO_WINDOW_LONG : HYPO_WINDOW_SHORT;
    for (const auto* v : {&w->_internal_arms, &w->_pre_arms, &w->_suf_arms})
        for (const auto& a : *v) {
            HypoArmDesc ad;
            ad.off = put(a.data(), a.data_size());
            ad.len = (uint32_t)a.get_seq_size();
            ad.reserved = 0;
            _arms.push_back(ad);
        }
    _win.push_back(d);
    _windows.push_back(w);
    _bp += d.draft_len;
}

void WindowBatch::run() {
    if (_windows.empty()) return;
    const uint64_t cap = hypo_gpu_out_bound(_win.data(), _win.size(), _arms.data(), _arms.size());
    _out.resize(cap + 16);
    _off.resize(_win.size() + 1);
    if (hypo_gpu_consensus_batch(_win.data(), _win.size(), _arms.data(), _arms.size(), _packed.data(), _packed.size(),
                                 _out.data(), _out.size(), _off.data()) != HYPO_OK)
        die("POA of windows");
    for (size_t i = 0; i < _windows.size(); ++i)
        _windows[i]->set_consensus(std::string(_out.data() + _off[i], _out.data() + _off[i + 1]));
}

}  // namespace hypo
