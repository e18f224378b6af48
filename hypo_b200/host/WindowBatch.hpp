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
#include <memory>
#include <string>
#include <vector>

#include "../../include/hypo_b200.h"
#include "Window.hpp"

namespace hypo {

// Packing is two passes over the windows, both OpenMP-parallel: sizes (arms and bytes per window), a
// serial prefix sum, then every window copies its own bytes and writes its own descriptors at the
// offsets the prefix sum assigned - the layout is the one a serial walk would produce, whatever the
// thread count (a million 30-arm windows are 31 M small copies: ~1 s on one thread, which would be
// twice the GPU's time for them).  The scatter of the consensus strings is parallel as well.
class WindowBatch {
public:
    void clear();
    void reserve(size_t n_windows, size_t n_arms = 0, size_t bytes = 0);
    // Appends a window (O(1), nothing is copied yet); the Window must outlive run().
    void add(Window* w);
    size_t size() const { return _windows.size(); }
    // Sum of Window::get_window_len() - the numerator of the Mbp-polished/s metric.
    uint64_t polished_bp() const { return _bp; }
    // Flattens the windows added so far (idempotent; run() and the accessors call it).
    // threads <= 0: all OpenMP threads.
    void pack(int threads = 0);
    // pack() + one hypo_gpu_consensus_batch call + scatter into Window::_consensus.
    // On failure prints "[Hypo::GPU] Error: ..." and exits(1), the reference's convention.
    void run();

    const HypoWindowDesc* win_desc() { pack(); return _win.data(); }
    const HypoArmDesc* arm_desc() { pack(); return _arms.data(); }
    const uint8_t* packed() { pack(); return _packed.get(); }
    size_t n_arms() { pack(); return _arms.size(); }
    size_t packed_bytes() { pack(); return _packed_bytes; }

private:
    std::vector<Window*> _windows;
    std::vector<HypoWindowDesc> _win;
    std::vector<HypoArmDesc> _arms;
    std::unique_ptr<uint8_t[]> _packed;   // (not a vector: no zero-fill of a slab that is overwritten anyway)
    size_t _packed_bytes = 0;
    size_t _n_packed = 0;                 // windows covered by the buffers above
    std::vector<char> _out;
    std::vector<uint64_t> _off;
    uint64_t _bp = 0;
};

}  // namespace hypo
