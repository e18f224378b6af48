// WindowStream.cpp — see WindowStream.hpp.
#include "WindowStream.hpp"

#include "../../include/hypo_b200.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <istream>
#include <ostream>
#include <sstream>

namespace hypo {

namespace {

bool fail(std::string* err, const std::string& what, size_t line_no) {
    if (err) *err = "inspect file, line " + std::to_string(line_no) + ": " + what;
    return false;
}

// "++\t<seq>" -> seq (the sequence may be empty)
bool marked_line(const std::string& line, std::string* seq) {
    if (line.size() < 2 || line[0] != '+' || line[1] != '+') return false;
    *seq = line.size() >= 3 ? line.substr(3) : std::string();
    return line.size() < 3 || line[2] == '\t';
}

void chomp(std::string& s) {
    while (!s.empty() && (s.back() == '\r' || s.back() == '\n')) s.pop_back();
}

}  // namespace

bool WindowStream::read(std::istream& in, std::string* err) {
    contig.clear(); regions.clear(); windows.clear(); recorded.clear(); declared_regions = 0;
    std::string line;
    size_t ln = 0;
    uint64_t declared = 0;
    bool have_count = false;
    while (std::getline(in, line)) {
        ++ln;
        chomp(line);
        if (line.empty()) continue;
        if (line[0] == '>') { contig = line.substr(1); continue; }
        if (line[0] == '#') { declared = strtoull(line.c_str() + 1, nullptr, 10); have_count = true; continue; }
        // ==========(beg-end)\tTYPE\tn_int\tn_pre\tn_suf\tn_empty   (reference src/Contig.cpp:424-451)
        if (line.compare(0, 11, "==========(") != 0) return fail(err, "expected a region header", ln);
        Region r;
        char* p = nullptr;
        r.beg = strtoull(line.c_str() + 11, &p, 10);
        if (!p || *p != '-') return fail(err, "bad coordinates", ln);
        r.end = strtoull(p + 1, &p, 10);
        if (!p || *p != ')') return fail(err, "bad coordinates", ln);
        std::istringstream rest(std::string(p + 1));
        uint32_t n_int = 0, n_pre = 0, n_suf = 0, n_empty = 0;
        if (!(rest >> r.type >> n_int >> n_pre >> n_suf >> n_empty)) return fail(err, "bad region counters", ln);
        std::string draft, cons;
        if (!std::getline(in, line)) return fail(err, "missing draft line", ln);
        ++ln; chomp(line);
        if (!marked_line(line, &draft)) return fail(err, "expected '++<TAB>draft'", ln);
        if (!std::getline(in, line)) return fail(err, "missing consensus line", ln);
        ++ln; chomp(line);
        if (!marked_line(line, &cons)) return fail(err, "expected '++<TAB>consensus'", ln);
        if (n_int + n_pre + n_suf + n_empty == 0) {
            // strong region, or a window without any arm: the sequence passes through unchanged
            r.text = draft;
            regions.push_back(std::move(r));
            continue;
        }
        PackedSeq<4> pd(draft);
        std::unique_ptr<Window> w(new Window(pd, 0, draft.size(), r.type == "LNG" ? WindowType::LONG : WindowType::SHORT));
        for (uint32_t k = 0; k < n_int + n_pre + n_suf; ++k) {
            if (!std::getline(in, line)) return fail(err, "missing arm line", ln);
            ++ln; chomp(line);
            PackedSeq<2> arm(line);
            if (!arm.is_valid()) return fail(err, "arm with a base other than A/C/G/T", ln);
            if (k < n_int) w->add_internal(arm);
            else if (k < n_int + n_pre) w->add_prefix(arm);
            else w->add_suffix(arm);
        }
        for (uint32_t k = 0; k < n_empty; ++k) w->add_empty();
        r.window = (int)windows.size();
        windows.push_back(std::move(w));
        recorded.push_back(cons);
        regions.push_back(std::move(r));
    }
    // The header announces _reg_type.size() - 1 regions, but with long reads the regions a LONG
    // pseudo-window swallowed (reference src/Contig.cpp:292-343: null window pointers) are not printed
    // (:424-451 has no branch for them): fewer region records than announced is the reference's own format.
    if (have_count && declared < regions.size())
        return fail(err, "header announces " + std::to_string(declared) + " regions, file holds " +
                             std::to_string(regions.size()), ln);
    declared_regions = have_count ? declared : regions.size();
    return true;
}

void WindowStream::write(std::ostream& os, bool recorded_consensus) const {
    os << ">" << contig << std::endl;
    os << "#" << std::max<uint64_t>(declared_regions, regions.size()) << std::endl;
    for (const Region& r : regions) {
        os << "==========(" << r.beg << "-" << r.end << ")\t" << r.type << "\t";
        if (r.window < 0) {
            os << 0 << "\t" << 0 << "\t" << 0 << "\t" << 0 << std::endl;
            os << "++\t" << r.text << std::endl;
            os << "++\t" << r.text << std::endl;
        } else {
            Window& w = *windows[r.window];
            if (recorded_consensus) {
                const std::string keep = w.get_consensus();
                w.set_consensus(recorded[r.window]);
                os << w;
                w.set_consensus(keep);
            } else {
                os << w;
            }
        }
    }
}

void WindowStream::add_window(std::unique_ptr<Window> w, const std::string& type, uint64_t beg,
                              const std::string& rec) {
    Region r;
    r.beg = beg;
    r.end = beg + (w->get_window_len() ? w->get_window_len() - 1 : 0);
    r.type = type;
    r.window = (int)windows.size();
    windows.push_back(std::move(w));
    recorded.push_back(rec);
    regions.push_back(std::move(r));
}

void WindowStream::add_plain(const std::string& type, uint64_t beg, const std::string& text) {
    Region r;
    r.beg = beg;
    r.end = beg + (text.empty() ? 0 : text.size() - 1);
    r.type = type;
    r.text = text;
    regions.push_back(std::move(r));
}

size_t WindowStream::replay() {
    std::vector<Window*> ptrs;
    ptrs.reserve(windows.size());
    for (auto& w : windows) ptrs.push_back(w.get());
    Window::generate_consensus_batch(ptrs);
    size_t bad = 0;
    for (size_t i = 0; i < windows.size(); ++i) bad += windows[i]->get_consensus() != recorded[i];
    return bad;
}

std::string WindowStream::stitched() const {
    std::string s;
    for (const Region& r : regions) s += r.window < 0 ? r.text : windows[r.window]->get_consensus();
    return s;
}

std::string WindowStream::stitched_on_device(bool resident) const {
    // the contig's draft as one PackedSeq<4> (what Contig::_pseq holds): every region's draft bases in order
    std::string draft;
    std::vector<HypoRegionDesc> reg;
    reg.reserve(regions.size());
    for (const Region& r : regions) {
        HypoRegionDesc d;
        d.src = draft.size();
        if (r.window < 0) {
            d.len = (uint32_t)r.text.size();
            d.window = HYPO_REGION_DRAFT;
            draft += r.text;
        } else {
            d.len = 0;
            d.window = (uint32_t)r.window;
            draft += windows[r.window]->draft().unpack();
        }
        reg.push_back(d);
    }
    const PackedSeq<4> pseq(draft);
    std::string cons;
    std::vector<uint64_t> off(windows.size() + 1, 0);
    for (size_t i = 0; i < windows.size(); ++i) {
        cons += windows[i]->get_consensus();
        off[i + 1] = cons.size();
    }
    const uint64_t first[2] = {0, reg.size()};
    const uint64_t doff[1] = {0};
    std::string out(draft.size() + cons.size() + 16, '\0');
    uint64_t out_off[2] = {0, 0};
    const int rc = hypo_gpu_stitch(reg.data(), reg.size(), first, 1, pseq.data(), doff, pseq.data_size(),
                                   resident ? nullptr : cons.data(), resident ? nullptr : off.data(), windows.size(),
                                   &out[0], out.size(), out_off);
    if (rc != HYPO_OK) {
        fprintf(stderr, "[Hypo::GPU] Error: stitching: %s\n", hypo_gpu_last_error());
        exit(1);
    }
    out.resize(out_off[1]);
    return out;
}

uint64_t WindowStream::polished_bp() const {
    uint64_t bp = 0;
    for (const auto& w : windows) bp += w->get_window_len();
    return bp;
}

}  // namespace hypo
