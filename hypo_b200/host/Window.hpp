// Window.hpp — host-side drop-in for hypo::Window on the POA path.
//
// Same public interface as the reference class (reference include/Window.hpp:41-146): a
// window is filled with add_internal/add_prefix/add_suffix/add_empty exactly as
// Contig::fill_short_windows / fill_long_windows do, `prepare_for_poa` fixes the score
// parameters, `generate_consensus` produces `_consensus`.  What changes is WHERE the
// consensus is computed: instead of one spoa graph per OpenMP thread, windows are flattened
// into one batch (WindowBatch.hpp) and sent through the C ABI (include/hypo_b200.h) to the
// B200.  `generate_consensus(engine_idx)` is kept for source compatibility and runs a batch
// of one; the intended call is Window::generate_consensus_batch (see INTEGRATION.md for the
// 12-line replacement of reference src/Hypo.cpp:236-248).
#pragma once
#include <functional>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#include "MinimizerFilter.hpp"
#include "PackedSeq.hpp"

namespace hypo {

// reference include/globalDefs.hpp:58-66
struct ScoreParams {
    INT8 sr_match_score;
    INT8 sr_misMatch_score;
    INT8 sr_gap_penalty;
    INT8 lr_match_score;
    INT8 lr_misMatch_score;
    INT8 lr_gap_penalty;
};

enum class WindowType : UINT8 { SHORT, LONG };

class WindowBatch;
class WindowStream;

// Arms of one kind (internal / prefix / suffix) of one window: PackedSeq<2> bytes back to back + lengths.
struct ArmStore {
    std::vector<BYTE> bytes;
    std::vector<UINT32> len;   // bases per arm
    size_t size() const { return len.size(); }
    bool empty() const { return len.empty(); }
    void add(const PackedSeq<2>& ps) {
        bytes.insert(bytes.end(), ps.data(), ps.data() + ps.data_size());
        len.push_back((UINT32)ps.get_seq_size());
    }
    void clear() { bytes.clear(); len.clear(); bytes.shrink_to_fit(); len.shrink_to_fit(); }
    static size_t arm_bytes(UINT32 n) { return (n + 3) / 4; }
    // f(pointer to the arm's packed bytes, length in bases) for every arm in insertion order
    template <class F>
    void for_each(F&& f) const {
        const BYTE* p = bytes.data();
        for (UINT32 n : len) { f(p, n); p += arm_bytes(n); }
    }
    static std::string unpack(const BYTE* p, UINT32 n) {   // PackedSeq<2>::unpack (reference src/PackedSeq.cpp:231-262)
        std::string s(n, 'A');
        for (UINT32 i = 0; i < n; ++i) s[i] = "ACGT"[(p[i >> 2] >> (6 - 2 * (i & 3))) & 3];
        return s;
    }
};

class Window {
public:
    Window() : _wtype(WindowType::SHORT), _num_internal(0), _num_pre(0), _num_suf(0), _num_empty(0),
               _longest_pre_len(0), _longest_suf_len(0) {}
    Window(const PackedSeq<4>& ps, const size_t left_ind, const size_t right_ind, WindowType wt)
        : _wtype(wt), _num_internal(0), _num_pre(0), _num_suf(0), _num_empty(0), _longest_pre_len(0),
          _longest_suf_len(0), _draft(ps, left_ind, right_ind) {}

    Window(const Window&) = delete;
    Window& operator=(const Window&) = delete;
    Window(Window&&) = delete;
    Window& operator=(Window&&) = delete;
    ~Window() = default;

    // reference src/Window.cpp:31-42.  num_threads is accepted for source compatibility; the
    // device replaces the per-thread engines.  `device` selects the CUDA ordinal.
    static void prepare_for_poa(const ScoreParams& sp, const UINT32 num_threads, int device = 0);
    // reference src/Window.cpp:44-61 (a batch of one; prefer generate_consensus_batch)
    void generate_consensus(const UINT32 engine_idx);
    // Replaces the OpenMP loop of reference src/Hypo.cpp:238-247 for a set of windows.
    static void generate_consensus_batch(const std::vector<Window*>& windows);

    std::string get_consensus() const { return _consensus; }
    size_t get_window_len() const { return _draft.get_seq_size(); }

    // LONG windows run every arm through hypo::Filter::is_good at insert time in the
    // reference (include/Window.hpp:66-101).  The minimiser filter is upstream of the hot
    // path; by default this mirror accepts every arm.  use_reference_long_filter(true) makes
    // add_* behave like the reference's (MinimizerFilter.hpp, pinned against the compiled
    // reference); set_long_arm_filter plugs in any other predicate.
    using ArmFilter = std::function<bool(const Window&, const PackedSeq<2>&)>;
    static void set_long_arm_filter(ArmFilter f) { _long_filter = std::move(f); }
    static void use_reference_long_filter(bool on) {
        if (!on) { _long_filter = nullptr; return; }
        _long_filter = [](const Window& w, const PackedSeq<2>& ps) {
            if (!w._ref_filter) {   // (the reference builds it in the constructor of a LONG window, :51)
                w._ref_filter.reset(new MinimizerFilter());
                w._ref_filter->init(w._draft.unpack());
            }
            return w._ref_filter->accepts(ps.unpack());
        };
    }

    void add_prefix(const PackedSeq<2>& ps) {
        if (!accept(ps)) return;
        UINT arm_len = (UINT)ps.get_seq_size();
        ++_num_pre;
        if (arm_len > _longest_pre_len) _longest_pre_len = arm_len;
        _pre_arms.add(ps);
    }
    void add_suffix(const PackedSeq<2>& ps) {
        if (!accept(ps)) return;
        UINT arm_len = (UINT)ps.get_seq_size();
        ++_num_suf;
        if (arm_len > _longest_suf_len) _longest_suf_len = arm_len;
        _suf_arms.add(ps);
    }
    void add_internal(const PackedSeq<2>& ps) {
        if (!accept(ps)) return;
        ++_num_internal;
        _internal_arms.add(ps);
    }
    void add_empty() { ++_num_empty; }

    UINT32 get_num_pre() const { return _num_pre; }
    UINT32 get_num_suf() const { return _num_suf; }
    UINT32 get_num_internal() const { return _num_internal + _num_empty; }
    UINT32 get_num_total() const { return _num_internal + _num_empty + _num_pre + _num_suf; }
    UINT32 get_maxlen_pre() const { return _longest_pre_len; }
    UINT32 get_maxlen_suf() const { return _longest_suf_len; }
    void clear_pre_suf() {
        _num_pre = 0;
        _num_suf = 0;
        _pre_arms.clear();
        _suf_arms.clear();
    }
    WindowType get_type() const { return _wtype; }
    const PackedSeq<4>& draft() const { return _draft; }
    // Read-only views for the batch packer (the reference keeps these containers private,
    // include/Window.hpp:130-134; INTEGRATION.md lists the accessors its Window needs).
    // f(packed bytes, length in bases) for every arm, container order: internal, prefix, suffix
    template <class F>
    void for_each_arm(F&& f) const {
        _internal_arms.for_each(f);
        _pre_arms.for_each(f);
        _suf_arms.for_each(f);
    }
    const ArmStore& arms(int kind) const { return kind == 0 ? _internal_arms : kind == 1 ? _pre_arms : _suf_arms; }
    void counts(uint32_t& n_internal, uint32_t& n_pre, uint32_t& n_suf, uint32_t& n_empty) const {
        n_internal = (uint32_t)_internal_arms.size(); n_pre = (uint32_t)_pre_arms.size();
        n_suf = (uint32_t)_suf_arms.size(); n_empty = _num_empty;
    }

    friend std::ostream& operator<<(std::ostream&, const Window&);
    friend class WindowBatch;
    friend class WindowStream;

private:
    bool accept(const PackedSeq<2>& ps) const {
        return _wtype != WindowType::LONG || !_long_filter || _long_filter(*this, ps);
    }
    void set_consensus(std::string con) { _consensus = std::move(con); }

    WindowType _wtype;
    UINT32 _num_internal, _num_pre, _num_suf, _num_empty;
    UINT32 _longest_pre_len, _longest_suf_len;
    PackedSeq<4> _draft;
    // The reference keeps a std::vector<PackedSeq<2>> per kind (include/Window.hpp:131-133): one heap block
    // per arm.  Here the packed bytes of a kind's arms lie back to back in one block (exactly the bytes the
    // PackedSeq<2> objects held, in insertion order), so the batch packer copies a window with three
    // memcpys instead of chasing 30 pointers - that is what lets one host feed several GPUs.
    ArmStore _internal_arms;
    ArmStore _pre_arms;
    ArmStore _suf_arms;
    std::string _consensus;
    mutable std::unique_ptr<MinimizerFilter> _ref_filter;   // LONG windows, use_reference_long_filter only
    static ArmFilter _long_filter;
};

}  // namespace hypo
