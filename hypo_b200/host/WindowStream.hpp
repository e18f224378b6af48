// WindowStream.hpp — capture / replay of window streams in the reference's own dump format
// (SURVEY.md §8f N1).
//
// The reference can dump every region of a contig — window counters, draft, consensus and all arms —
// with Contig::generate_inspect_file (reference src/Contig.cpp:368-453, per-window part
// Window::operator<<, src/Window.cpp:63-84; enabled by un-commenting src/Hypo.cpp:262,265,271).  That
// text file is at the same time the input of the POA hot path (drafts + arms as the pipeline really
// produced them) and its expected output (the reference's consensus).  WindowStream reads such a file
// back into hypo::Window objects through the public add_* API, replays them through
// Window::generate_consensus_batch (one C-ABI call) and counts the windows whose consensus differs from
// the recorded one; it also writes the same format, so captured or synthetic streams can be stored as
// reproducible benchmark inputs.
#pragma once
#include <cstdint>
#include <iosfwd>
#include <memory>
#include <string>
#include <vector>

#include "Window.hpp"

namespace hypo {

class WindowStream {
public:
    struct Region {
        uint64_t beg = 0, end = 0;   // "(beg-end)" of the header line, inclusive coordinates on the contig
        std::string type;            // SR, MSR, SWS, WS, SW, MWM, WM, MW, SWM, MWS, OTH, LNG
        int window = -1;             // index into windows / recorded, -1 for a region without arms
        std::string text;            // window == -1: the region's sequence (copied to the output verbatim)
    };

    // Parses one inspect file.  Returns false and sets *err on malformed input.
    bool read(std::istream& in, std::string* err);
    // Writes the regions in the reference's format; a window's consensus line is the recorded
    // consensus (recorded_consensus = true) or its current get_consensus().
    void write(std::ostream& os, bool recorded_consensus) const;

    // Appends a region that carries a window (type "LNG" for WindowType::LONG is the caller's business).
    void add_window(std::unique_ptr<Window> w, const std::string& type, uint64_t beg, const std::string& recorded);
    // Appends a region without arms (strong region, or a window nobody mapped to).
    void add_plain(const std::string& type, uint64_t beg, const std::string& text);

    // Window::generate_consensus_batch over every window (prepare_for_poa must have been called);
    // returns how many consensus strings differ from the recorded ones.
    size_t replay();
    // Polished contig as Contig::operator<< would stitch it (reference src/Contig.cpp:345-366): plain
    // regions verbatim, windows replaced by their consensus.
    std::string stitched() const;

    // The same contig stitched on the device by hypo_gpu_stitch (the contig writer over the consensus slab,
    // SURVEY.md §8f N4): the regions become HypoRegionDesc records over the contig's PackedSeq<4> draft.
    // resident = true: the consensus bytes are not sent again - the result of the batch call that replay()
    // just made is still on the device (single-chunk batches on one device only).
    std::string stitched_on_device(bool resident) const;

    uint64_t polished_bp() const;   // sum of Window::get_window_len()

    std::string contig;
    uint64_t declared_regions = 0;   // the "#n" line (can exceed regions.size(), see read())
    std::vector<Region> regions;
    std::vector<std::unique_ptr<Window>> windows;
    std::vector<std::string> recorded;
};

}  // namespace hypo
