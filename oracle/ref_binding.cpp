// TEST INFRASTRUCTURE ONLY — never linked into or called by the product path.
//
// The reference-side binding of INTEGRATION.md §3, COMPILED against the reference's own headers
// (/root/reference/include/Window.hpp, PackedSeq.hpp) and linked with the reference's own objects and with
// hypo_b200/libhypo_b200.so: reference hypo::Window objects are filled through the reference's public
// add_* interface, flattened by the packer below, polished through hypo_gpu_consensus_batch, and the very
// same objects are then polished by the reference's own Window::generate_consensus for comparison.
//
// The two accessors + friend declaration a maintainer adds to the reference headers are
// stood in for by lifting the access specifiers for this translation unit only; nothing else about the
// reference classes is touched, and no reference source is copied into this repository.
#include <omp.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "spoa/spoa.hpp"
#include "globalDefs.hpp"
#define private public
#include "PackedSeq.hpp"
#include "Filter.hpp"
#include "Window.hpp"
#undef private

#include "../include/hypo_b200.h"

namespace {

// ---- the packer, over the REFERENCE classes ------------------------------------------
class WindowBatch {
public:
    void add(hypo::Window* w) { _windows.push_back(w); }
    int run() {
        const size_t n = _windows.size();
        const int threads = omp_get_max_threads();
        std::vector<uint64_t> arm0(n + 1, 0), byte0(n + 1, 0);
#pragma omp parallel for schedule(static, 512) num_threads(threads)
        for (size_t i = 0; i < n; ++i) {                           // pass 1: arms and bytes per window
            const hypo::Window* w = _windows[i];
            uint64_t bytes = w->_draft._data.size();
            for (auto* v : {&w->_internal_arms, &w->_pre_arms, &w->_suf_arms}) for (auto& a : *v) bytes += a._data.size();
            arm0[i + 1] = w->_internal_arms.size() + w->_pre_arms.size() + w->_suf_arms.size();
            byte0[i + 1] = bytes;
        }
        for (size_t i = 0; i < n; ++i) { arm0[i + 1] += arm0[i]; byte0[i + 1] += byte0[i]; }
        _win.resize(n); _arms.resize(arm0[n]); _packed.assign(byte0[n] + 16, 0);
#pragma omp parallel for schedule(static, 512) num_threads(threads)
        for (size_t i = 0; i < n; ++i) {                           // pass 2: every window fills its own slice
            const hypo::Window* w = _windows[i];
            uint64_t pos = byte0[i];
            HypoWindowDesc d;
            memset(&d, 0, sizeof(d));
            d.draft_off = pos;  d.draft_len = (uint32_t)w->_draft.get_seq_size();  d.first_arm = arm0[i];
            memcpy(_packed.data() + pos, w->_draft._data.data(), w->_draft._data.size());  pos += w->_draft._data.size();
            d.n_internal = (uint32_t)w->_internal_arms.size();  d.n_pre = (uint32_t)w->_pre_arms.size();
            d.n_suf = (uint32_t)w->_suf_arms.size();            d.n_empty = w->_num_empty;
            d.wtype = w->_wtype == hypo::WindowType::LONG ? HYPO_WINDOW_LONG : HYPO_WINDOW_SHORT;
            HypoArmDesc* ad = _arms.data() + arm0[i];
            for (auto* v : {&w->_internal_arms, &w->_pre_arms, &w->_suf_arms})
                for (auto& a : *v) {
                    ad->off = pos; ad->len = (uint32_t)a.get_seq_size(); ad->reserved = 0; ++ad;
                    memcpy(_packed.data() + pos, a._data.data(), a._data.size());  pos += a._data.size();
                }
            _win[i] = d;
        }
        _out.resize(hypo_gpu_out_bound(_win.data(), _win.size(), _arms.data(), _arms.size()) + 16);
        _off.resize(n + 1);
        const int rc = hypo_gpu_consensus_batch(_win.data(), n, _arms.data(), _arms.size(), _packed.data(), byte0[n],
                                                _out.data(), _out.size(), _off.data());
        if (rc) { fprintf(stderr, "[Hypo::GPU] Error: POA of windows: %s\n", hypo_gpu_last_error()); return rc; }
#pragma omp parallel for schedule(static, 512)
        for (size_t i = 0; i < n; ++i)
            _windows[i]->_consensus = std::string(_out.data() + _off[i], _out.data() + _off[i + 1]);   // set_consensus_from_gpu
        return 0;
    }

private:
    std::vector<hypo::Window*> _windows;
    std::vector<HypoWindowDesc> _win;
    std::vector<HypoArmDesc> _arms;
    std::vector<uint8_t> _packed;
    std::vector<char> _out;
    std::vector<uint64_t> _off;
};

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

int collect(std::vector<std::unique_ptr<hypo::Window>>& ws, char* out, uint64_t cap, uint64_t* off) {
    uint64_t pos = 0;
    for (size_t w = 0; w < ws.size(); ++w) {
        off[w] = pos;
        const std::string c = ws[w]->get_consensus();
        if (pos + c.size() > cap) return HYPO_E_OUT_CAP;
        memcpy(out + pos, c.data(), c.size());
        pos += c.size();
    }
    off[ws.size()] = pos;
    return HYPO_OK;
}

}  // namespace

extern "C" {

// The flat batch is rebuilt as REFERENCE hypo::Window objects (public add_* API; LONG windows filter their arms
// like the reference does), polished on the GPU through the binding (out_gpu / off_gpu), then polished by the
// reference's own Window::generate_consensus (out_ref / off_ref).
int hypo_refbind_run(const int8_t scores[6], int device, const HypoWindowDesc* win, uint64_t n_win,
                     const HypoArmDesc* arms, const uint8_t* packed, char* out_gpu, uint64_t* off_gpu,
                     char* out_ref, uint64_t* off_ref, uint64_t cap) {
    if (hypo_gpu_init(scores, device)) { fprintf(stderr, "[Hypo::GPU] Error: %s\n", hypo_gpu_last_error()); return HYPO_E_CUDA; }
    std::vector<std::unique_ptr<hypo::Window>> ws(n_win);
    for (uint64_t w = 0; w < n_win; ++w) {
        const HypoWindowDesc& d = win[w];
        hypo::PackedSeq<4> draft(unpack4(packed + d.draft_off, d.draft_len));
        ws[w].reset(new hypo::Window(draft, 0, d.draft_len, d.wtype == HYPO_WINDOW_LONG ? hypo::WindowType::LONG : hypo::WindowType::SHORT));
        uint64_t a = d.first_arm;
        for (uint32_t i = 0; i < d.n_internal; ++i, ++a) ws[w]->add_internal(hypo::PackedSeq<2>(unpack2(packed + arms[a].off, arms[a].len)));
        for (uint32_t i = 0; i < d.n_pre; ++i, ++a) ws[w]->add_prefix(hypo::PackedSeq<2>(unpack2(packed + arms[a].off, arms[a].len)));
        for (uint32_t i = 0; i < d.n_suf; ++i, ++a) ws[w]->add_suffix(hypo::PackedSeq<2>(unpack2(packed + arms[a].off, arms[a].len)));
        for (uint32_t i = 0; i < d.n_empty; ++i) ws[w]->add_empty();
    }
    // the drop-in: what replaces the loop of reference src/Hypo.cpp:238-247
    WindowBatch batch;
    for (auto& w : ws) batch.add(w.get());
    if (int rc = batch.run()) return rc;
    if (int rc = collect(ws, out_gpu, cap, off_gpu)) return rc;
    // the reference's own loop on the same objects
    hypo::Window::_alignment_engines.clear();
    hypo::Window::_alignment_engines_long.clear();
    hypo::ScoreParams sp;
    sp.sr_match_score = scores[0]; sp.sr_misMatch_score = scores[1]; sp.sr_gap_penalty = scores[2];
    sp.lr_match_score = scores[3]; sp.lr_misMatch_score = scores[4]; sp.lr_gap_penalty = scores[5];
    const int threads = omp_get_max_threads();
    hypo::Window::prepare_for_poa(sp, (hypo::UINT32)threads);
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
    for (uint64_t w = 0; w < n_win; ++w) ws[w]->generate_consensus(omp_get_thread_num());
    return collect(ws, out_ref, cap, off_ref);
}

}  // extern "C"
