// MinimizerFilter.hpp — the arm filter of LONG windows.
//
// The reference's Window runs every arm of a LONG window through hypo::Filter::is_good before it stores it
// (reference include/Window.hpp:66-101; the filter itself: include/Filter.hpp:32-101, ring buffer
// include/MinimizerDeque.hpp): an arm is kept iff enough of its (w = 10, k = 10) canonical minimizers also
// occur among the minimizers of the window's draft — one per 50 bases of the arm.  That decision is taken
// on the host while the windows are filled, upstream of the POA path; it is restated here so that the
// hypo::Window mirror can behave like the reference's Window at its add_* interface
// (Window::use_reference_long_filter).  tests/test_host_cpu.py pins it against the acceptance flags of the
// compiled reference.
//
// Quirks kept on purpose (they decide borderline arms): a non-ACGT character only restarts the count of
// valid bases — the window of recent k-mers and the count of processed k-mers carry on; ties between equal
// k-mers keep the older one; when forward and reverse-complement k-mer are equal the value is the same
// either way.
#pragma once
#include <cstdint>
#include <string>
#include <unordered_set>

namespace hypo {

class MinimizerFilter {
public:
    static constexpr unsigned kK = 10;            // k-mer length
    static constexpr unsigned kW = 10;            // k-mers per minimizer window
    static constexpr unsigned kBasesPerHit = 50;  // one shared minimizer per this many bases of the arm

    void init(const std::string& draft) {
        _draft.clear();
        scan(draft, [&](uint64_t v, uint32_t) { _draft.insert(v); });
    }

    bool accepts(const std::string& arm) const {
        uint32_t hits = 0;
        int64_t last_pos = -1;
        scan(arm, [&](uint64_t v, uint32_t pos) {
            if ((int64_t)pos == last_pos) return;   // same minimizer occurrence as the previous window
            last_pos = pos;
            hits += (uint32_t)_draft.count(v);
        });
        return (uint64_t)hits * kBasesPerHit >= arm.size();
    }

private:
    static unsigned code(char ch) {
        switch (ch) {
            case 'A': case 'a': case 0: return 0;
            case 'C': case 'c': case 1: return 1;
            case 'G': case 'g': case 2: return 2;
            case 'T': case 't': case 'U': case 'u': case 3: return 3;
            default: return 4;
        }
    }

    // Calls emit(value, end position of the k-mer) with the minimizer of every full window of kW k-mers.
    template <class F>
    static void scan(const std::string& s, F&& emit) {
        constexpr uint64_t mask = (1ull << (2 * kK)) - 1;
        constexpr unsigned top = 2 * (kK - 1);
        struct Entry { uint64_t v; uint32_t pos; };
        Entry ring[kW + 1];   // values never decrease from head to tail
        unsigned head = 0, held = 0;
        uint64_t fwd = 0, rev = 0;
        unsigned valid_run = 0, kmers = 0;
        for (size_t i = 0; i < s.size(); ++i) {
            const unsigned c = code(s[i]);
            if (c > 3) { valid_run = 0; continue; }
            ++valid_run;
            fwd = ((fwd << 2) | c) & mask;
            rev = (rev >> 2) | ((uint64_t)(3u ^ c) << top);
            if (valid_run < kK) continue;
            const uint64_t canon = fwd < rev ? fwd : rev;
            while (held != 0 && ring[(head + held - 1) % (kW + 1)].v > canon) --held;
            ring[(head + held) % (kW + 1)] = Entry{canon, (uint32_t)i};
            ++held;
            while ((size_t)ring[head].pos + kW <= i) { head = (head + 1) % (kW + 1); --held; }
            if (++kmers >= kW) emit(ring[head].v, ring[head].pos);
        }
    }

    std::unordered_set<uint64_t> _draft;
};

}  // namespace hypo
