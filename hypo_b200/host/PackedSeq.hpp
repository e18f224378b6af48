// PackedSeq.hpp — host-side mirror of hypo::PackedSeq<NB> for the POA path.
//
// Same public surface and bit layout as the reference class (reference
// include/PackedSeq.hpp:84-158, src/PackedSeq.cpp:28-262) for everything the POA path and
// its callers use: construction from a string / from an htslib 4-bit sequence / from a
// sub-range of another PackedSeq, unpack, get_seq_size, enc_base_at, base_at, is_valid, and the k-mer
// search helpers arm extraction uses to anchor a read segment on its window (find_kmer, check_kmer and
// their canonical variants, reference src/PackedSeq.cpp:264-414).  tests/test_host_cpu.py pins all of it
// against the reference's own class (compiled into oracle/_ref).
//
// One addition: data()/remainder() give read-only access to the packed bytes so the batch
// packer can hand them to the device verbatim (the reference keeps _data private,
// include/PackedSeq.hpp:152; INTEGRATION.md shows the two-line accessor to add there).
#pragma once
#include <cassert>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

namespace hypo {

using BYTE = uint8_t;
using UINT8 = uint8_t;
using UINT32 = uint32_t;
using UINT64 = uint64_t;
using UINT = unsigned int;
using INT8 = int8_t;

// A:0 C:1 G:2 T:3 everything else 4 (reference include/globalDefs.hpp:158-178)
inline BYTE nt4(unsigned char c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': case 'U': case 'u': return 3;
        default: return 4;
    }
}

template <int NB>
class PackedSeq {
    static_assert(NB == 2 || NB == 4, "[Hypo::PackedSequence] Packed sequence base can only be 2 or 4.");

public:
    static constexpr BYTE byte_mask = (NB == 2) ? 0x03 : 0x0f;

    PackedSeq() : _len(0), _valid(true) {}

    explicit PackedSeq(const std::string& s) : _len(0), _valid(true) {
        if (s.size() > 0xffffffffu) {
            fprintf(stderr, "[Hypo::PackedSeq] Error: Length exceed limit: The length of a sequence is %lu which exceeds the limit of %u !\n",
                    (unsigned long)s.size(), 0xffffffffu);
            exit(1);
        }
        _data.assign(bytes_for(s.size()), 0);
        for (size_t i = 0; i < s.size(); ++i) {
            BYTE b = nt4((unsigned char)s[i]);
            if (NB == 2 && b > 3) {
                fprintf(stderr, "[Hypo::PackedSeq] Error: Wrong base (Can not pack in 2 bits): Base %c in a sequence is not A, C, G, or T !\n", s[i]);
                _valid = false;
                break;
            }
            put(i, b);
        }
        _len = s.size();
    }

    // htslib 4-bit encoded read (bam_get_seq layout): base i in the high nibble of byte i>>1
    // when i is even; codes 1,2,4,8 = A,C,G,T (reference src/PackedSeq.cpp:122-152).
    PackedSeq(const UINT32 seq_len, const UINT32 offset, const UINT8* hts_seq) : _len(seq_len), _valid(true) {
        static const UINT8 hts2nt[16] = {4, 0, 1, 4, 2, 4, 4, 4, 3, 4, 4, 4, 4, 4, 4, 4};
        _data.assign(bytes_for(seq_len), 0);
        for (UINT32 i = 0; i < seq_len; ++i) {
            const UINT32 k = i + offset;
            BYTE b = hts2nt[(hts_seq[k >> 1] >> ((~k & 1) << 2)) & 0xf];
            if (NB == 2 && b > 3) { _valid = false; break; }
            put(i, b);
        }
    }

    // sub-range [left, right) of a sequence packed with MB bits per base
    template <int MB>
    PackedSeq(const PackedSeq<MB>& ps, const size_t left_ind, const size_t right_ind) : _len(right_ind - left_ind), _valid(true) {
        assert(right_ind <= ps.get_seq_size() && left_ind <= right_ind);
        _data.assign(bytes_for(_len), 0);
        for (size_t i = 0; i < _len; ++i) {
            BYTE b = ps.enc_base_at(left_ind + i);
            if (NB == 2 && b > 3) {
                fprintf(stderr, "[Hypo::PackedSeq] Error: Wrong base (Can not pack in 2 bits): Base code at %lu in a sequence is not A, C, G, or T !\n",
                        (unsigned long)(left_ind + i));
                exit(1);
            }
            put(i, b);
        }
    }

    PackedSeq(const PackedSeq&) = default;
    PackedSeq& operator=(const PackedSeq&) = delete;
    PackedSeq(PackedSeq&&) = default;
    PackedSeq& operator=(PackedSeq&&) = default;

    bool is_valid() const { return _valid; }
    size_t get_seq_size() const { return _len; }
    BYTE enc_base_at(size_t ind) const {
        assert(ind < _len);
        return BYTE(_data[ind / per_byte] >> shift_of(ind)) & byte_mask;
    }
    char base_at(size_t ind) const { return "ACGTN"[enc_base_at(ind) > 4 ? 4 : enc_base_at(ind)]; }

    std::string unpack() const { return unpack(0, _len); }
    std::string unpack(const size_t left_ind, const size_t right_ind) const {
        assert(right_ind <= _len && left_ind <= right_ind);
        std::string s(right_ind - left_ind, 'N');
        for (size_t i = left_ind; i < right_ind; ++i) s[i - left_ind] = base_at(i);
        return s;
    }

    // ---- k-mer search (k-mers as 2-bit codes, first base in the highest bits; a non-ACGT base restarts) ----
    // Is `target` one of the k-mers inside [left, right)?  `result` = start of the first (is_first) or of
    // the last occurrence; untouched when there is none.
    bool find_kmer(const UINT64 target, const UINT k, const size_t left_ind, const size_t right_ind, const bool is_first,
                   size_t& result) const {
        assert(right_ind <= _len && left_ind <= right_ind);
        bool found = false;
        const UINT64 mask = (1ULL << (2 * k)) - 1;
        UINT64 fwd = 0;
        UINT run = 0;
        for (size_t i = left_ind; i < right_ind; ++i) {
            const BYTE b = enc_base_at(i);
            if (b > 3) { run = 0; fwd = 0; continue; }
            fwd = ((fwd << 2) | b) & mask;
            if (run < k) ++run;
            if (run == k && fwd == target) {
                result = i + 1 - k;
                found = true;
                if (is_first) break;
            }
        }
        return found;
    }
    // Does the k-mer that starts at `ind` equal `target`?
    bool check_kmer(const UINT64 target, const UINT k, const size_t ind) const {
        size_t at;
        return find_kmer(target, k, ind, ind + k, true, at);
    }
    // Same for canonical k-mers (the smaller of a k-mer and its reverse complement).  The reference
    // computes these in 32-bit registers (k <= 16) and so does this.
    bool find_canonical_kmer(const UINT64 target, const UINT k, const size_t left_ind, const size_t right_ind,
                             const bool is_first, size_t& result) const {
        assert(right_ind <= _len && left_ind <= right_ind);
        bool found = false;
        const UINT32 top = 2 * (k - 1);
        const UINT32 mask = (UINT32)((1ULL << (2 * k)) - 1);
        UINT32 fwd = 0, rev = 0;
        UINT run = 0;
        for (size_t i = left_ind; i < right_ind; ++i) {
            const BYTE b = enc_base_at(i);
            if (b > 3) { run = 0; continue; }   // (the registers keep their bits; k valid bases flush them)
            ++run;
            fwd = (UINT32)((fwd << 2ull | b) & mask);
            rev = (UINT32)((rev >> 2ull) | (3ULL ^ b) << top);
            if (run >= k && (fwd < rev ? fwd : rev) == target) {
                result = i + 1 - k;
                found = true;
                if (is_first) break;
            }
        }
        return found;
    }
    bool check_canonical_kmer(const UINT64 target, const UINT k, const size_t ind) const {
        size_t at;
        return find_canonical_kmer(target, k, ind, ind + k, true, at);
    }

    // read-only view of the packed bytes (device consumes them verbatim)
    const BYTE* data() const { return _data.data(); }
    size_t data_size() const { return _data.size(); }

private:
    static constexpr size_t per_byte = (NB == 2) ? 4 : 2;
    static size_t bytes_for(size_t n) { return (n + per_byte - 1) / per_byte; }
    static unsigned shift_of(size_t i) { return (NB == 2) ? 6 - 2 * (unsigned)(i & 3) : ((i & 1) ? 0 : 4); }
    void put(size_t i, BYTE b) { _data[i / per_byte] |= BYTE(b << shift_of(i)); }

    std::vector<BYTE> _data;
    size_t _len;
    bool _valid;
};

}  // namespace hypo
