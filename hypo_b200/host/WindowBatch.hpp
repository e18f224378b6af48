// WindowBatch.hpp — flattens hypo::Window objects into the three buffers of the C ABI
// (include/hypo_b200.h) and scatters the consensus strings back.
//
// This is the packer SURVEY.md §8b describes: it walks the windows a contig batch produced
// (reference src/Hypo.cpp:238-247 visits them through Contig::is_valid_window /
// Contig::generate_consensus), copies the PackedSeq bytes verbatim — arms in container order
// _internal_arms, _pre_arms, _suf_arms (reference include/Window.hpp:131-133) — and makes ONE
// FFI call for the whole batch.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/hypo_b200.h"
#include "Window.hpp"

namespace hypo {

class WindowBatch {
public:
    void clear();
    void reserve(size_t n_windows, size_t n_arms, size_t bytes);
    // Appends a window; the Window must outlive run().
    void add(Window* w);
    size_t size() const { return _windows.size(); }
    // Sum of Window::get_window_len() — the numerator of the Mbp-polished/s metric.
    uint64_t polished_bp() const { return _bp; }
    // One hypo_gpu_consensus_batch call + scatter into Window::_consensus.
    // On failure prints "[Hypo::GPU] Error: ..." and exits(1), the reference's convention.
    void run();

    const std::vector<HypoWindowDesc>& win_desc() const { return _win; }
    const std::vector<HypoArmDesc>& arm_desc() const { return _arms; }
    const std::vector<uint8_t>& packed() const { return _packed; }

private:
    uint64_t put(const uint8_t* p, size_t n);
    std::vector<Window*> _windows;
    std::vector<HypoWindowDesc> _win;
    std::vector<HypoArmDesc> _arms;
    std::vector<uint8_t> _packed;
    std::vector<char> _out;
    std::vector<uint64_t> _off;
    uint64_t _bp = 0;
};

}  // namespace hypo
