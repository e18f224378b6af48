// WindowBatch.hpp — flattens hypo::Window objects into the three buffers of the C ABI
// (include/hypo_b200.h) and scatters the consensus strings back.
//
// This is the packer SURVEY.md §8b describes: it walks the windows a contig batch produced
// (reference src/Hypo.cpp:238-247 visits them through Contig::is_valid_window /
// Contig::generate_consensus), copies the PackedSeq bytes verbatim — arms in container order
// _internal_arms, _pre_arms, _suf_arms (reference include/Window.hpp:131-133) — and sends them
// through the FFI.
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
//
// run() streams: the windows are cut into chunks and TWO chunks are in flight at a time (the library runs two
// batch calls side by side, one per lane; a worker thread sits in hypo_gpu_consensus_batch for each) while
// the calling thread packs the next chunk into the third of three page-locked buffer sets and scatters the
// consensus strings of the chunk that just finished into their Window objects.  Packing, copies, compaction,
// result copies and scatter hide behind the kernels as long as the host keeps up.
class WindowBatch {
public:
    struct Timing {
        double pack = 0, device = 0, scatter = 0, total = 0;   // seconds; pack / scatter overlap `device`
        uint64_t chunks = 0;
    };

    WindowBatch() = default;
    ~WindowBatch();
    WindowBatch(const WindowBatch&) = delete;
    WindowBatch& operator=(const WindowBatch&) = delete;

    void clear();
    void reserve(size_t n_windows, size_t n_arms = 0, size_t bytes = 0);
    // Appends a window (O(1), nothing is copied yet); the Window must outlive run().
    void add(Window* w);
    size_t size() const { return _windows.size(); }
    // Sum of Window::get_window_len() - the numerator of the Mbp-polished/s metric.
    uint64_t polished_bp() const { return _bp; }
    // Flattens ALL windows added so far into one buffer set (idempotent; the accessors call it).
    // threads <= 0: all OpenMP threads.
    void pack(int threads = 0);
    // Chunked pack -> hypo_gpu_consensus_batch -> scatter into Window::_consensus, double-buffered.
    // On failure prints "[Hypo::GPU] Error: ..." and exits(1), the reference's convention.
    // chunk_windows = 0: chosen from the batch size and the number of driven devices (see run()).
    void run(size_t chunk_windows = 0);
    const Timing& last_timing() const { return _timing; }

    const HypoWindowDesc* win_desc() { pack(); return _win.data(); }
    const HypoArmDesc* arm_desc() { pack(); return _arms.data(); }
    const uint8_t* packed() { pack(); return _packed.get(); }
    size_t n_arms() { pack(); return _arms.size(); }
    size_t packed_bytes() { pack(); return _packed_bytes; }

private:
    // One page-locked buffer set of the streaming run (grow-only, reused across run() calls).
    struct Slot {
        HypoWindowDesc* win = nullptr; size_t win_cap = 0;
        HypoArmDesc* arms = nullptr; size_t arms_cap = 0;
        uint8_t* packed = nullptr; size_t packed_cap = 0;
        char* out = nullptr; size_t out_cap = 0;
        uint64_t* off = nullptr; size_t off_cap = 0;
        size_t n_win = 0, n_arms = 0, n_bytes = 0, first = 0;
        void release();
    };
    void pack_chunk(Slot& s, size_t first, size_t n, int threads);
    void scatter_chunk(const Slot& s, int threads);

    std::vector<Window*> _windows;
    std::vector<HypoWindowDesc> _win;
    std::vector<HypoArmDesc> _arms;
    std::unique_ptr<uint8_t[]> _packed;   // (not a vector: no zero-fill of a slab that is overwritten anyway)
    size_t _packed_bytes = 0;
    size_t _n_packed = 0;                 // windows covered by the buffers above
    Slot _slot[3];   // one being packed / scattered, two in flight
    Timing _timing;
    uint64_t _bp = 0;
};

}  // namespace hypo
