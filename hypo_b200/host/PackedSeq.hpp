// PackedSeq.hpp — host-side mirror of hypo::PackedSeq<NB> for the POA path.
//
// Same public surface and bit layout as the reference class (reference
// include/PackedSeq.hpp:84-158, src/PackedSeq.cpp:28-262) for everything the POA path and
// its callers use: construction from a string / from an htslib 4-bit sequence / from a
// sub-range of another PackedSeq, unpack, get_seq_size, enc_base_at, base_at, is_valid.
// The k-mer search helpers (find_kmer, check_kmer, ...) belong to the windowing code, which
// is out of scope, and are not mirrored.
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
