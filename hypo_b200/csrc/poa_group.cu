// poa_group.cu — the POA-consensus path for SMALL windows: several windows per warp, in lock-step.
//
// The windows the reference pipeline produces are mostly tiny (median draft length 9 bp,
// 10-40 arms).  One warp per window (poa_kernel.cu) leaves most of a 128-column tile idle there and,
// more importantly, pays every per-read phase - decode, end cell, traceback, graph fusion, order
// maintenance, row records, and the serial walks on one lane - once per window.  Here a warp is cut into
// groups of G = 8 or 16 lanes, each group owns one window (lane l of a group holds DP columns [4l, 4l+4):
// 32 or 64 columns), and the 4 or 2 windows of a warp run the SAME instruction stream: every phase is
// written once for the whole warp with a per-group "active" predicate, loops that contain a warp
// collective run to the maximum trip count over the groups, and the collectives are full-mask shuffles /
// ballots of width G (a sub-mask would cost a MATCH + divergence check per collective).  A serial phase
// (the reference's DFS, the heaviest-bundle pass) runs on the leader lane of every group at once.
//
// Semantics are those of poa_kernel.cu (same arena layout, same row records, same preference orders), so
// results are bit-identical; the group tiers only run SHORT windows (reference src/Window.cpp:87-154)
// whose sequences fit the group's columns, and hand a window that outgrows their capacities on to the
// one-warp-per-window tiers like every other tier does.
#include "poa_kernel.cuh"
#include "poa_common.cuh"

namespace hypo_b200 {

namespace {

template <int G>
struct Grp {
    static_assert(G == 8 || G == 16, "group width");
    static constexpr int kGroups = 32 / G;
    static constexpr unsigned kBits = (1u << G) - 1u;
    static __device__ __forceinline__ int lane() { return threadIdx.x & (G - 1); }
    static __device__ __forceinline__ int shift() { return threadIdx.x & 31 & ~(G - 1); }
    // ballot over the lanes of the group (bit k = group lane k); executed by the whole warp
    static __device__ __forceinline__ unsigned ballot(bool p) { return (__ballot_sync(kFull, p) >> shift()) & kBits; }
    static __device__ __forceinline__ bool any(bool p) { return ballot(p) != 0u; }
    template <typename T>
    static __device__ __forceinline__ T shfl(T v, int src) { return __shfl_sync(kFull, v, src, G); }
    template <typename T>
    static __device__ __forceinline__ T shfl_up(T v, int d) { return __shfl_up_sync(kFull, v, d, G); }
    template <typename T>
    static __device__ __forceinline__ T shfl_down(T v, int d) { return __shfl_down_sync(kFull, v, d, G); }
    template <typename T>
    static __device__ __forceinline__ T shfl_xor(T v, int d) { return __shfl_xor_sync(kFull, v, d, G); }
};
__device__ __forceinline__ bool warp_any(bool p) { return __any_sync(kFull, p); }
__device__ __forceinline__ int warp_max(int v) { return __reduce_max_sync(kFull, v); }

// The arena of the group this lane belongs to (shared memory; derived from the extern array in every phase so
// that the accesses stay LDS/STS).
template <int kTier>
__device__ __forceinline__ Graph group_graph() {
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr ArenaLayout L = arena_layout(fixed_caps(kTier));
    constexpr int G = group_lanes(kTier);
    return bind_graph_at(smem + (threadIdx.x / G) * L.total, L);
}
template <int kTier>
__device__ __forceinline__ WarpState* group_state() {
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr ArenaLayout L = arena_layout(fixed_caps(kTier));
    constexpr int G = group_lanes(kTier);
    return (WarpState*)(smem + (threadIdx.x / G) * L.total + L.state);
}

// Records why a window leaves the tier (called under a group-uniform condition).
template <int G>
__device__ __forceinline__ void note_fail(uint32_t* fail_hist, int why) {
    if (Grp<G>::lane() == 0 && fail_hist) atomicAdd(fail_hist + why, 1u);
}

struct GSeq {
    const uint8_t* bytes;
    int len;        // bases without markers
    int nb;         // 2 or 4 bits per base
    bool head, tail;
    int type;
};

// Exclusive prefix max of the lane totals across the group (see warp_excl_max in poa_kernel.cu).
template <int G>
__device__ __forceinline__ uint32_t group_excl_max(uint32_t tot2, uint32_t lane0, uint32_t neg2) {
    uint32_t e = Grp<G>::shfl_up(tot2, 1);
    e = lane0 ? neg2 : e;
    {
        const uint32_t a1 = Grp<G>::shfl_up(e, 1), a2 = Grp<G>::shfl_up(e, 2), a3 = Grp<G>::shfl_up(e, 3);
        e = __vmaxs2(__vimax3_s16x2(e, a1, a2), a3);
    }
    if constexpr (G == 8) {
        return __vmaxs2(e, Grp<G>::shfl_up(e, 4));
    } else {
        const uint32_t b1 = Grp<G>::shfl_up(e, 4), b2 = Grp<G>::shfl_up(e, 8), b3 = Grp<G>::shfl_up(e, 12);
        return __vmaxs2(__vimax3_s16x2(e, b1, b2), b3);
    }
}

// ------------------------------------------------------------------------------------------
// DP fill (reference sisd_alignment_engine.cpp:263-342), one tile of 4 G columns per group; the rows of
// the groups advance together (a group with fewer rows idles through the others' last rows).  The last
// three rows stay in registers as in poa_kernel.cu; a row with a predecessor further back reads its
// predecessor rows from the matrix (its own four columns; the left neighbour's last column by shuffle).
// ------------------------------------------------------------------------------------------
struct GRow {
    uint32_t x[kNR];
    uint32_t left;
};

template <int kTier>
__device__ __noinline__ EndCell fill_g(bool act, int16_t* __restrict__ H, int len, int type, Scores sc) {
    constexpr int G = group_lanes(kTier);
    constexpr unsigned kStride = 4u * G;
    const Graph g = group_graph<kTier>();
    const int gl = Grp<G>::lane();
    const int n = act ? g.n_nodes : 0;
    const int nmax = warp_max(n);
    // (loop constants are made opaque so that ptxas keeps them in registers instead of re-deriving them per row)
    const uint32_t g2 = opaque(bcast16(sc.g)), mm2 = opaque(bcast16(sc.m - sc.g)), nn2 = opaque(bcast16(sc.n - sc.g));
    const uint32_t neg2 = opaque(kNegInf2);
    const uint32_t lane0 = opaque(gl == 0 ? 1u : 0u);
    const uint32_t row0_left = gl == 0 ? kNegInf2 : 0u;
    const uint32_t xinit0 = (type == kROV && gl == 0) ? (kNegInf2 & 0xffff0000u) : kNegInf2;
    uint32_t let4 = 0x07070707u;
    if (act) let4 = *reinterpret_cast<const uint32_t*>(g.colseq + gl * 4);
    const int16_t* Hl = H + gl * 4;
    int16_t* Hrow = H + gl * 4;
    if (act) stg64(Hrow, 0u, 0u);   // row 0: H^[0][j] = 0
    GRow A, B, C;
    A.x[0] = A.x[1] = 0u; A.left = row0_left;
    B = A; C = A;
    // one DP row: d1 / d2 / d3 hold the rows 1 / 2 / 3 ranks back, the new row replaces d3
    auto dp_row = [&](int rk, const GRow& d1, const GRow& d2, GRow& d3) {
        const bool ra = rk < n;
        const uint32_t info = ra ? g.rowinfo[rk] : (1u << kRowNearShift);
        uint32_t pf[kNR];
        profile_regs(let4, (info >> 24) & 7u, mm2, nn2, pf);
        uint32_t x[kNR] = {xinit0, neg2};
        const uint32_t near = info >> kRowNearShift;
        if (near & 1u) relax(x, d1.x, d1.left, pf, g2);
        if (near & 2u) relax(x, d2.x, d2.left, pf, g2);
        if (near & 4u) relax(x, d3.x, d3.left, pf, g2);
        const bool far = ra && near == 0u;
        if (warp_any(far)) {
            const int np = (int)((info >> 16) & 0xffu);
            const int cnt = far ? (np ? np : 1) : 0;   // no predecessor: the virtual row 0 (reference :300-301)
            const int cmax = warp_max(cnt);
            const unsigned off = info & 0xffffu;
#pragma unroll 1
            for (int k = 0; k < cmax; ++k) {
                const bool valid = k < cnt;
                unsigned prow = 0;
                if (valid && np) prow = g.prows[off + k];
                uint2 q = make_uint2(0u, 0u);
                if (valid) q = ldg64(Hl + prow * kStride);
                uint32_t left = Grp<G>::shfl_up(q.y, 1);
                if (lane0) left = neg2;
                const uint32_t p[kNR] = {q.x, q.y};
                if (valid) relax(x, p, left, pf, g2);
            }
        }
        const uint32_t cb = group_excl_max<G>(scan_inlane(x, neg2), lane0, neg2);
        d3.x[0] = __vmaxs2(x[0], cb);
        d3.x[1] = __vmaxs2(x[1], cb);
        d3.left = cb;
        Hrow += kStride;
        if (ra) stg64(Hrow, d3.x[0], d3.x[1]);
    };
    // two rows per trip: the register sets are rotated once per pair (an odd row count ends with an idle row)
#pragma unroll 1
    for (int rk = 0; rk < nmax; rk += 2) {
        dp_row(rk, A, B, C);       // new row -> C
        dp_row(rk + 1, C, A, B);   // new row -> B
        const GRow r = A;
        A = B; B = C; C = r;
    }
    __syncwarp();   // end cell and traceback read the matrix across lanes

    // End cell (reference :276-288,328-340): best last-column score over the candidate rows (NW / ROV: nodes
    // without out-edges, LOV: every node); strictly greater => lowest rank wins.
    int best = INT_MIN, brow = 0x7fffffff, cnt = 0;
#pragma unroll 1
    for (int r = gl; r < n; r += G) {
        const bool cand = (type == kLOV) || ((g.rowinfo[r] >> 27) & 1);
        if (cand) {
            const int v = (int)H[(unsigned)(r + 1) * kStride + (unsigned)len];
            if (v > best) { best = v; brow = r + 1; cnt = 1; }
            else if (v == best) ++cnt;
        }
    }
#pragma unroll
    for (int d = G / 2; d >= 1; d >>= 1) {
        const int ob = Grp<G>::shfl_xor(best, d);
        const int orow = Grp<G>::shfl_xor(brow, d);
        const int ocnt = Grp<G>::shfl_xor(cnt, d);
        if (ob > best) { best = ob; brow = orow; cnt = ocnt; }
        else if (ob == best) { brow = min(brow, orow); cnt += ocnt; }
    }
    EndCell ec;
    const bool found = brow != 0x7fffffff;
    ec.row = found ? brow : 0;
    ec.col = found ? len : 0;
    ec.score = found ? best : 0;
    ec.tie = found && cnt > 1;
    return ec;
}

// ------------------------------------------------------------------------------------------
// Traceback (reference sisd_alignment_engine.cpp:344-437), the scheme of poa_kernel.cu with G lanes: a run
// of "diagonal through the first in-edge" moves is verified G steps at a time through the jump pointers;
// any other move is a general step whose candidates are spread over the lanes in the reference's
// preference order (lanes [0, KD): diagonal via in-edge k; [KD, 2 KD): vertical via in-edge k; lane 2 KD:
// horizontal; lowest matching lane wins).  Every trip of the loop takes one of the two for every group.
// ------------------------------------------------------------------------------------------
template <int kTier>
__device__ __noinline__ AlnSpan traceback_g(bool act, const int16_t* __restrict__ H, EndCell ec, int type,
                                            Scores sc, int max_steps) {
    constexpr int G = group_lanes(kTier);
    constexpr unsigned ucols = 4u * G;
    constexpr int KD = (G - 2) / 2;
    const Graph g = group_graph<kTier>();
    const int gl = Grp<G>::lane();
    int i = ec.row, j = ec.col, hij = ec.score;
    AlnSpan span;
    span.first = -1; span.last = -1;
    const int mm = sc.m - sc.g, nn = sc.n - sc.g;
    int steps = 0;
    bool spec = true, dead = false;
#pragma unroll 1
    for (;;) {
        const bool cont = act && !dead && (type == kROV ? (i != 0 && j != 0) : (i != 0 || j != 0)) && steps < max_steps;
        if (!warp_any(cont)) break;
        ++steps;
        // ---- speculative diagonal run through first predecessors: lane k takes step k (row fp^k(i))
        const bool do_spec = cont && spec && i != 0 && j != 0;
        int my_r = 0, my_rn = 0;
        const int jj = j - gl;
        bool ok = false;
        int hp = 0;
        if (do_spec) {
            my_r = i;
            const int a = gl >> 2, b = gl & 3;
#pragma unroll
            for (int s = 0; s < G / 4 - 1; ++s) {
                const int t = g.fp4[my_r];
                if (a > s) my_r = t;
            }
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                const int t = g.fp[my_r];
                if (b > s) my_r = t;
            }
            my_rn = g.fp[my_r];
            ok = my_r != 0 && jj >= 1;
            if (ok) {
                const int hc = (int)H[(unsigned)my_r * ucols + (unsigned)jj];
                hp = (int)H[(unsigned)my_rn * ucols + (unsigned)(jj - 1)];
                const int s = (((g.rowinfo[my_r - 1] >> 24) & 7) == g.seq[jj - 1]) ? mm : nn;
                ok = hc == hp + s;
            }
        }
        const unsigned okm = Grp<G>::ballot(ok);
        const int run = __ffs(~okm) - 1;   // leading all-true lanes (okm has G bits: run <= G)
        const int rsrc = run > 0 ? run - 1 : 0;
        const int run_i = Grp<G>::shfl(my_rn, rsrc);
        const int run_h = Grp<G>::shfl(hp, rsrc);
        const bool took = do_spec && run > 0;
        if (took) {
            if (gl < run) g.cur[jj - 1] = g.r2n[my_r - 1];
            if (span.last < 0) span.last = j - 1;
            span.first = j - run;
            i = run_i; hij = run_h;
            j -= run;
            steps += run - 1;
            // a run cut below G means the next cell fails this very test (or row 0 / column 0 was reached)
            spec = run == G;
        }
        // ---- one general step
        const bool gen = cont && !took;
        uint32_t info = 0;
        int ps = 0, deg = 0;
        if (gen && i != 0) {
            info = g.rowinfo[i - 1];
            ps = (int)(info & 0xffffu);
            deg = (int)((info >> 16) & 0xffu);
        }
        const bool par = gen && deg <= KD;
        int pi = i, h = 0;
        bool match = false;
        if (par) {
            const int slots = deg == 0 ? 1 : deg;   // no in-edge: virtual row 0 (reference :300-301)
            if (i != 0 && gl < 2 * KD) {
                const int k = gl < KD ? gl : gl - KD;
                const bool diag = gl < KD;
                if (k < slots && (!diag || j != 0)) {
                    pi = deg == 0 ? 0 : (int)g.prows[ps + k];
                    if (diag) {
                        const int s = (((info >> 24) & 7) == g.seq[j - 1]) ? mm : nn;
                        h = (int)H[(unsigned)pi * ucols + (unsigned)(j - 1)];
                        match = hij == h + s;
                    } else {
                        h = (int)H[(unsigned)pi * ucols + (unsigned)j];
                        match = hij == h + sc.g;
                    }
                }
            } else if (gl == 2 * KD && j != 0) {
                h = (int)H[(unsigned)i * ucols + (unsigned)(j - 1)];
                match = hij == h;
            }
        }
        const unsigned mt = Grp<G>::ballot(match);
        const int wl = mt ? __ffs(mt) - 1 : 0;
        const int sel_i = Grp<G>::shfl(pi, wl);
        const int sel_h = Grp<G>::shfl(h, wl);
        if (gen) {
            int ni = i, nj = j, nh = hij;
            bool found = false;
            if (par) {
                if (mt != 0u) {
                    ni = sel_i; nh = sel_h;
                    nj = (wl < KD || wl == 2 * KD) ? j - 1 : j;
                    found = true;
                }
            } else {
                // in-degree beyond the lanes: the reference's serial walk (every lane of the group, redundantly)
                const int pe = ps + deg;
                if (j != 0) {
                    const int s = (((info >> 24) & 7) == g.seq[j - 1]) ? mm : nn;
#pragma unroll 1
                    for (int k = ps; k < pe; ++k) {
                        const int q = g.prows[k];
                        const int hh = (int)H[(unsigned)q * ucols + (unsigned)(j - 1)];
                        if (hij == hh + s) { ni = q; nj = j - 1; nh = hh; found = true; break; }
                    }
                }
                if (!found) {
#pragma unroll 1
                    for (int k = ps; k < pe; ++k) {
                        const int q = g.prows[k];
                        const int hh = (int)H[(unsigned)q * ucols + (unsigned)j];
                        if (hij == hh + sc.g) { ni = q; nj = j; nh = hh; found = true; break; }
                    }
                }
                if (!found && j != 0) {
                    const int hh = (int)H[(unsigned)i * ucols + (unsigned)(j - 1)];
                    if (hij == hh) { ni = i; nj = j - 1; nh = hh; found = true; }
                }
            }
            if (!found) {
                dead = true;   // impossible for a consistent H; never spin
            } else {
                if (nj != j) {
                    if (gl == 0) g.cur[j - 1] = (ni != i) ? g.r2n[i - 1] : kNone;
                    if (span.last < 0) span.last = j - 1;
                    span.first = j - 1;
                }
                i = ni; j = nj; hij = nh;
                spec = true;
            }
        }
    }
    __syncwarp();
    return span;
}

// ------------------------------------------------------------------------------------------
// Graph fusion (reference graph.cpp:154-291), group-parallel over sequence positions.  Returns false for
// a group whose window exceeded a capacity (it is handed on to the next tier).
// ------------------------------------------------------------------------------------------
template <int kTier>
__device__ __noinline__ bool add_to_graph_g(bool act, uint32_t* fail_hist, int len, AlnSpan span) {
    constexpr int G = group_lanes(kTier);
    constexpr Caps caps = fixed_caps(kTier);
    Graph g = group_graph<kTier>();
    const int gl = Grp<G>::lane();
    const unsigned lt_mask = (1u << gl) - 1u;
    bool ok = act;
    int first = span.first, last = span.last;
    if (first < 0) { first = len; last = len - 1; }   // empty alignment: the whole read is a chain (:174-182)
    const int head_n = first, tail_n = len - 1 - last;
    const int base = ok ? g.n_nodes : 0;
    if (ok && base + head_n + tail_n > caps.ncap) { note_fail<G>(fail_hist, kFailNodes); ok = false; }

    // head chain [0, first) and tail chain (last, len): fresh nodes, allocated FIRST (:194-200)
    if (ok) {
#pragma unroll 1
        for (int p = gl; p < len; p += G) {
            int id = -1;
            if (p < first) id = base + p;
            else if (p > last) id = base + head_n + (p - last - 1);
            if (id >= 0) {
                init_node(g, id, g.seq[p]);
                g.cur[p] = (uint16_t)id;
            }
        }
    }
    int n_nodes = base + head_n + tail_n;
    int n_al = ok ? g.n_al : 0;
    __syncwarp();

    // aligned part [first, last]: reuse / clique lookup / new node (:206-245)
    {
        const int iters = ok ? (last - first + G) / G : 0;
        const int itmax = warp_max(iters);
#pragma unroll 1
        for (int it = 0; it < itmax; ++it) {
            const int p = first + it * G + gl;
            const bool actp = ok && p <= last;
            int x = kNone, code = 0, res = -1;
            bool need_new = false;
            if (actp) {
                x = g.cur[p];
                code = g.seq[p];
                if (x == kNone) {
                    need_new = true;
                } else if ((g.ninfo[x] & 7) == code) {
                    res = x;
                } else {
                    need_new = true;
                    const int blk = g.al_blk[x];
                    if (blk != kNone) {
                        const int cnt = g.al_cnt[x];
#pragma unroll 1
                        for (int k = 0; k < cnt; ++k) {
                            const int a = g.al_pool[blk * g.als + k];
                            if ((g.ninfo[a] & 7) == code) { res = a; need_new = false; break; }
                        }
                    }
                }
            }
            const unsigned newmask = Grp<G>::ballot(need_new);
            const int n_new = __popc(newmask);
            if (need_new) res = n_nodes + __popc(newmask & lt_mask);
            // aligned-list blocks: the new node needs one; so does x if it had none
            const bool link = need_new && x != kNone;
            const bool x_needs_blk = link && g.al_blk[x] == kNone;
            const unsigned m1 = Grp<G>::ballot(link);
            const unsigned m2 = Grp<G>::ballot(x_needs_blk);
            const bool clique_full = Grp<G>::any(link && g.al_cnt[x] + 1 > g.als);
            if (ok) {
                if (n_nodes + n_new > caps.ncap) { note_fail<G>(fail_hist, kFailNodes); ok = false; }
                else if (n_al + __popc(m1) + __popc(m2) > caps.acap) { note_fail<G>(fail_hist, kFailAligned); ok = false; }
                else if (clique_full) { note_fail<G>(fail_hist, kFailClique); ok = false; }
            }
            if (ok) {
                if (need_new) init_node(g, res, code);
                if (link) {
                    const int yb = n_al + __popc(m1 & lt_mask);
                    int xb = g.al_blk[x];
                    if (x_needs_blk) {
                        xb = n_al + __popc(m1) + __popc(m2 & lt_mask);
                        g.al_blk[x] = (uint16_t)xb;
                    }
                    g.al_blk[res] = (uint16_t)yb;
                    const int cnt = g.al_cnt[x];
                    // y.list = x.list + [x]; every a in x.list gets y appended; x.list += y (:228-240)
#pragma unroll 1
                    for (int k = 0; k < cnt; ++k) {
                        const int a = g.al_pool[xb * g.als + k];
                        g.al_pool[yb * g.als + k] = (uint16_t)a;
                        const int ab = g.al_blk[a];
                        const int ac = g.al_cnt[a];
                        g.al_pool[ab * g.als + ac] = (uint16_t)res;
                        g.al_cnt[a] = (uint8_t)(ac + 1);
                    }
                    g.al_pool[yb * g.als + cnt] = (uint16_t)x;
                    g.al_cnt[res] = (uint8_t)(cnt + 1);
                    g.al_pool[xb * g.als + cnt] = (uint16_t)res;
                    g.al_cnt[x] = (uint8_t)(cnt + 1);
                }
                n_nodes += n_new;
                n_al += __popc(m1) + __popc(m2);
                if (actp) g.cur[p] = (uint16_t)res;
            }
            __syncwarp();
        }
    }

    // edges (cur[p-1] -> cur[p]), weight 1+1 per traversal (:99-115,251-265,283-288).  Every node of a
    // sequence is distinct, so lanes touch disjoint in-lists / source nodes.
    int n_edges = ok ? g.n_edges : 0;
    {
        const int iters = ok ? (len + G - 1) / G : 0;
        const int itmax = warp_max(iters);
#pragma unroll 1
        for (int it = 0; it < itmax; ++it) {
            const int p = it * G + gl;
            bool need_edge = false, sat = false;
            int src = 0, dst = 0, tail = kNone;
            if (ok && p < len && p >= 1) {
                dst = g.cur[p];
                src = g.cur[p - 1];
                need_edge = true;
#pragma unroll 1
                for (int e = g.in_head[dst]; e != kNone; e = g.e_next[e]) {
                    if (g.e_src[e] == src) {
                        g.e_w[e] = (uint16_t)(g.e_w[e] + 2);
                        need_edge = false;
                        break;
                    }
                    tail = e;
                }
                sat = need_edge && g.in_deg[dst] >= 254;
            }
            const unsigned em = Grp<G>::ballot(need_edge);
            const bool any_sat = Grp<G>::any(sat);
            if (ok && (n_edges + __popc(em) > caps.ecap || any_sat)) { note_fail<G>(fail_hist, kFailEdges); ok = false; }
            if (ok) {
                if (need_edge) {
                    const int e = n_edges + __popc(em & lt_mask);
                    g.e_src[e] = (uint16_t)src;
                    g.e_w[e] = 2;
                    g.e_next[e] = kNone;
                    if (tail == kNone) g.in_head[dst] = (uint16_t)e; else g.e_next[tail] = (uint16_t)e;
                    g.in_deg[dst] = (uint8_t)(g.in_deg[dst] + 1);
                    g.ninfo[src] = (uint8_t)(g.ninfo[src] | 8);
                }
                n_edges += __popc(em);
            }
        }
    }
    if (ok && gl == 0) {
        g.ws->n_nodes = n_nodes;
        g.ws->n_al = n_al;
        g.ws->n_edges = n_edges;
        g.ws->n_seq = g.n_seq + 1;
    }
    __syncwarp();
    return ok;
}

// ------------------------------------------------------------------------------------------
// Exact topological order (reference graph.cpp:293-353): the reference's DFS on the leader lane of every
// group that needs it.  Returns false for a group whose DFS stack overflowed.
// ------------------------------------------------------------------------------------------
template <int kTier>
__device__ __noinline__ bool topo_sort_g(bool act, uint32_t* fail_hist) {
    constexpr int G = group_lanes(kTier);
    constexpr Caps caps = fixed_caps(kTier);
    const Graph g = group_graph<kTier>();
    const int gl = Grp<G>::lane();
    const int n = act ? g.n_nodes : 0;
    __syncwarp();   // the sort's scratch aliases the row records other lanes may still be reading
#pragma unroll 1
    for (int i = gl; i < n; i += G) g.mark[i] = 0;
    __syncwarp();
    int ok = 1;
    if (act && gl == 0) {
        int nr = 0;
#pragma unroll 1
        for (int id = 0; id < n && ok; ++id)
            if ((g.mark[id] & 3) != 2) ok = dfs_from(g, caps, id, nr) ? 1 : 0;
    }
    ok = Grp<G>::shfl(ok, 0);
    if (act && !ok) note_fail<G>(fail_hist, kFailStack);
    __syncwarp();
    if (ok) {
#pragma unroll 1
        for (int r = gl; r < n; r += G) g.n2r[g.r2n[r]] = (uint16_t)r;
    }
    __syncwarp();
    return act && ok != 0;
}

// Incremental order maintenance (see order_update in poa_kernel.cu for the scheme).
template <int kTier>
__device__ __noinline__ void order_update_g(bool act, int len, int nb) {
    constexpr int G = group_lanes(kTier);
    const Graph g = group_graph<kTier>();
    const int gl = Grp<G>::lane();
    const int n = act ? g.n_nodes : 0;
    uint16_t* anch = g.anch;
    uint16_t* newa = g.newa;
    const int itmax = warp_max(act ? (len + G - 1) / G : 0);

    // pass A (reverse): anchor of every position that holds a new node
    int carry = nb;
#pragma unroll 1
    for (int it = itmax - 1; it >= 0; --it) {
        const int p = it * G + gl;
        int colmin = 0x7fffffff;
        int mate_anchor = -1;
        bool is_new = false;
        if (act && p < len) {
            const int v = g.cur[p];
            is_new = v >= nb;
            const int cnt = g.al_cnt[v];
            if (!is_new || cnt > 0) {
                int mn = is_new ? 0x7fffffff : (int)g.n2r[v];
                int mx = is_new ? -1 : (int)g.n2r[v];
                const int blk = g.al_blk[v];
#pragma unroll 1
                for (int k = 0; k < cnt; ++k) {
                    const int m = g.al_pool[blk * g.als + k];
                    if (m < nb) { const int r = g.n2r[m]; mn = min(mn, r); mx = max(mx, r); }
                }
                colmin = mn;
                if (is_new) mate_anchor = mx + 1;
            }
        }
        int suf = colmin;
#pragma unroll
        for (int d = 1; d < G; d <<= 1) suf = min(suf, Grp<G>::shfl_down(suf, d));
        int excl = Grp<G>::shfl_down(suf, 1);
        if (gl == G - 1) excl = 0x7fffffff;
        excl = min(excl, carry);
        if (is_new) anch[p] = (uint16_t)(mate_anchor >= 0 ? mate_anchor : excl);
        carry = min(carry, Grp<G>::shfl(suf, 0));
    }
    __syncwarp();
    // pass B (forward): compact the new nodes in path order, give them their ranks
    int K = 0;
#pragma unroll 1
    for (int it = 0; it < itmax; ++it) {
        const int p = it * G + gl;
        const bool in = act && p < len;
        const int v = in ? (int)g.cur[p] : 0;
        const bool is_new = in && v >= nb;
        const unsigned m = Grp<G>::ballot(is_new);
        if (is_new) {
            const int i = K + __popc(m & ((1u << gl) - 1u));
            const int a = anch[p];
            newa[i] = (uint16_t)a;
            g.n2r[v] = (uint16_t)(a + i);
        }
        K += __popc(m);
    }
    __syncwarp();
    // old nodes move right by the number of new nodes anchored at or before them
    if (act && K > 0) {
#pragma unroll 1
        for (int v = gl; v < nb; v += G) {
            const int r = g.n2r[v];
            int lo = 0, hi = K;
#pragma unroll 1
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if ((int)newa[mid] <= r) lo = mid + 1; else hi = mid;
            }
            g.n2r[v] = (uint16_t)(r + lo);
        }
    }
    __syncwarp();
    if (act && K > 0) {
#pragma unroll 1
        for (int v = gl; v < n; v += G) g.r2n[g.n2r[v]] = (uint16_t)v;
    }
    __syncwarp();
}

// Row records for the DP and the traceback (see build_rows in poa_kernel.cu).
template <int kTier>
__device__ __noinline__ void build_rows_g(bool act) {
    constexpr int G = group_lanes(kTier);
    const Graph g = group_graph<kTier>();
    const int gl = Grp<G>::lane();
    const int n = act ? g.n_nodes : 0;
    const int itmax = warp_max((n + G - 1) / G);
    int base = 0;
#pragma unroll 1
    for (int it = 0; it < itmax; ++it) {
        const int r = it * G + gl;
        int v = 0, deg = 0, code = 0;
        if (r < n) {
            v = g.r2n[r];
            deg = g.in_deg[v];
            const int info = g.ninfo[v];
            code = (info & 7) | ((info & 8) ? 0 : 8);   // bit 3 = sink (no out-edges)
        }
        int off = deg;
#pragma unroll
        for (int d = 1; d < G; d <<= 1) {
            const int y = Grp<G>::shfl_up(off, d);
            if (gl >= d) off += y;
        }
        const int total = Grp<G>::shfl(off, G - 1);
        off = base + off - deg;
        if (r < n) {
            int k = off, first = 0;
            uint32_t near = (deg == 0 && r == 0) ? 1u : 0u;
            bool far = deg == 0 && r != 0;
#pragma unroll 1
            for (int e = g.in_head[v]; e != kNone; e = g.e_next[e]) {
                const int prow = g.n2r[g.e_src[e]] + 1;
                if (k == off) first = prow;
                g.prows[k++] = (uint16_t)prow;
                const int dist = r + 1 - prow;
                if (dist >= 1 && dist <= 3) near |= 1u << (dist - 1); else far = true;
            }
            if (far) near = 0u;
            g.fp[r + 1] = (uint16_t)first;
            g.rowinfo[r] = (uint32_t)off | ((uint32_t)deg << 16) | ((uint32_t)code << 24) | (near << kRowNearShift);
        }
        base += total;
    }
    if (act && gl == 0) { g.fp[0] = 0; g.fp4[0] = 0; }
    __syncwarp();
#pragma unroll 1
    for (int r = 1 + gl; r <= n; r += G) g.fp4[r] = g.fp[g.fp[g.fp[g.fp[r]]]];
    __syncwarp();
}

// Heaviest bundle + branch completion (reference graph.cpp:610-705; see heaviest_bundle in poa_kernel.cu).
// Returns the consensus length (nodes in g.cons), -1 if the exact order is needed first.
template <int kTier>
__device__ __noinline__ int heaviest_bundle_g(bool act, bool exact_order) {
    constexpr int G = group_lanes(kTier);
    const Graph g = group_graph<kTier>();
    const int gl = Grp<G>::lane();
    const int n = act ? g.n_nodes : 0;
    int len = 0;
    __syncwarp();
    constexpr int kTiePred = 0xFFFE;
#pragma unroll 1
    for (int v = gl; v < n; v += G) {
        int wmax = -1, pv = kNone, cnt = 0;
#pragma unroll 1
        for (int e = g.in_head[v]; e != kNone; e = g.e_next[e]) {
            const int w = g.e_w[e];
            if (w > wmax) { wmax = w; pv = g.e_src[e]; cnt = 1; }
            else if (w == wmax) ++cnt;
        }
        g.score[v] = -1;
        g.pred[v] = (uint16_t)(cnt > 1 ? kTiePred : pv);
        g.cons[v] = (uint16_t)wmax;
    }
    __syncwarp();
    if (act && gl == 0) {
        int best = 0, ties = 0, sb = -1;
#pragma unroll 1
        for (int r = 0; r < n; ++r) {
            const int v = g.r2n[r];
            int pv = g.pred[v];
            int sv = -1;
            if (pv == kTiePred) {
                pv = kNone;
#pragma unroll 1
                for (int e = g.in_head[v]; e != kNone; e = g.e_next[e]) {
                    const int s = g.e_src[e];
                    const int w = g.e_w[e];
                    if (sv < w || (sv == w && g.score[pv] <= g.score[s])) { sv = w; pv = s; }
                }
                g.pred[v] = (uint16_t)pv;
                sv += g.score[pv];
            } else if (pv != kNone) {
                sv = (int)g.cons[v] + g.score[pv];
            }
            g.score[v] = sv;
            if (v == best) sb = sv;
            if (sb < sv) { best = v; sb = sv; ties = 1; }
            else if (sb == sv) ++ties;
        }
        if (!exact_order && (ties > 1 || (g.ninfo[best] & 8))) best = -1;
        int guard = 0;
#pragma unroll 1
        while (best >= 0 && (g.ninfo[best] & 8) && guard++ <= n) best = branch_completion(g, g.n2r[best]);
        int k = -1;
        if (best >= 0) {
            k = 0;
#pragma unroll 1
            while (g.pred[best] != kNone && k < n) { g.cons[k++] = (uint16_t)best; best = g.pred[best]; }
            g.cons[k++] = (uint16_t)best;
#pragma unroll 1
            for (int a = 0, b = k - 1; a < b; ++a, --b) {
                uint16_t t = g.cons[a]; g.cons[a] = g.cons[b]; g.cons[b] = t;
            }
        }
        len = k;
    }
    len = Grp<G>::shfl(len, 0);
    __syncwarp();
    return len;
}

// ------------------------------------------------------------------------------------------
// One sequence for every active group: decode, align, fuse, re-order.  Returns false for a group whose
// window has to leave the tier.
// ------------------------------------------------------------------------------------------
template <int kTier>
__device__ __noinline__ bool seq_step(bool act, uint32_t* fail_hist, int16_t* H, GSeq s, Scores sc) {
    constexpr int G = group_lanes(kTier);
    constexpr Caps caps = fixed_caps(kTier);
    constexpr int kCols = 4 * G;
    const Graph g = group_graph<kTier>();
    const int gl = Grp<G>::lane();
    const int len = s.len + (s.head ? 1 : 0) + (s.tail ? 1 : 0);
    bool ok = act;
    if (ok && len > caps.lcap) { note_fail<G>(fail_hist, kFailLen); ok = false; }
    if (ok) {
        uint8_t* dst = g.seq + (s.head ? 1 : 0);
        if (s.nb == 2) {
#pragma unroll 1
            for (int p = gl; p < s.len; p += G) dst[p] = (s.bytes[p >> 2] >> (6 - 2 * (p & 3))) & 3;
        } else {
#pragma unroll 1
            for (int p = gl; p < s.len; p += G) {
                const int v = (s.bytes[p >> 1] >> ((p & 1) ? 0 : 4)) & 15;
                dst[p] = v > 4 ? 4 : v;
            }
        }
        if (gl == 0) {
            if (s.head) g.seq[0] = kCodeJ;
            if (s.tail) g.seq[len - 1] = kCodeO;
            g.colseq[0] = 7;   // column 0 and the padding columns match no letter
        }
#pragma unroll 1
        for (int j = len + 1 + gl; j < kCols; j += G) g.colseq[j] = 7;
    }
    __syncwarp();

    AlnSpan span;
    span.first = -1; span.last = -1;
    WarpState* const ws = g.ws;
    const int nodes_before = ok ? g.n_nodes : 0, edges_before = ok ? g.n_edges : 0;
    bool dp = ok && nodes_before > 0;   // reference sisd_alignment_engine.cpp:249-251
    if (dp) {
        // 16-bit range guard: |H^| <= S*(rows+cols) and <= 2*S*cols
        const int S = max(max(abs(sc.m), abs(sc.n)), abs(sc.g));
        if (S * (nodes_before + 1 + kCols) > kMaxH16 || 2 * S * kCols > kMaxH16) {
            note_fail<G>(fail_hist, kFailRange);
            ok = false; dp = false;
        }
    }
    if (warp_any(dp)) {
        if (dp && gl == 0) ws->cells += (unsigned long long)(nodes_before + 1) * (unsigned long long)(len + 1);
        EndCell ec = fill_g<kTier>(dp, H, len, s.type, sc);
        bool redo = dp && ec.tie && !ws->exact;
        if (warp_any(redo)) {
            // the reference breaks this tie by rank in ITS order: derive it and redo the fill
            const bool sorted = topo_sort_g<kTier>(redo, fail_hist);
            if (redo && !sorted) { ok = false; dp = false; redo = false; }
            if (redo && gl == 0) ws->exact = 1;
            __syncwarp();
            build_rows_g<kTier>(redo);
            const EndCell ec2 = fill_g<kTier>(redo, H, len, s.type, sc);
            if (redo) ec = ec2;
        }
        span = traceback_g<kTier>(dp, H, ec, s.type, sc, nodes_before + len + 4);
        if (dp) {
            // the matrix of this read is dead: drop its lines from L2 instead of writing them back
            const unsigned lines = (unsigned)(nodes_before + 1) * (unsigned)kCols / 64u;
            char* hb = reinterpret_cast<char*>(H);
#pragma unroll 1
            for (unsigned l = gl; l < lines; l += G)
                asm volatile("discard.global.L2 [%0], 128;" ::"l"(hb + (size_t)l * 128) : "memory");
        }
    }
    ok = add_to_graph_g<kTier>(ok, fail_hist, len, span) && ok;
    const int nodes_after = ok ? ws->n_nodes : 0;
    // a read that only re-walks existing nodes and edges leaves the DAG's structure, hence every order, unchanged
    const bool changed = ok && (nodes_after != nodes_before || ws->n_edges != edges_before);
    if (warp_any(changed)) {
        const bool grew = changed && nodes_after != nodes_before;
        if (warp_any(grew)) order_update_g<kTier>(grew, len, nodes_before);
        if (changed && gl == 0) ws->exact = 0;
        __syncwarp();
        build_rows_g<kTier>(changed);
    }
    // (see repeat_sequence in poa_kernel.cu: a sequence that left the structure untouched can be repeated)
    if (ok && gl == 0) ws->clean = changed ? 0 : 1;
    __syncwarp();
    return ok;
}

// ------------------------------------------------------------------------------------------
// Kernel: every group pulls windows from the tier's queue; one trip of the main loop = one sequence
// (reference src/Window.cpp:95-132 order: draft backbone if there are no internal arms, internal arms,
// prefix arms last to first (kLOV), suffix arms (kROV)) for every group, the consensus
// (:134-149) for the groups whose window is complete.
// ------------------------------------------------------------------------------------------
template <int kTier>
__global__ void __launch_bounds__(256, 3) poa_group_kernel(const Params P) {
    constexpr int G = group_lanes(kTier);
    constexpr int NG = 32 / G;
    const int gl = Grp<G>::lane();
    const int group = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * NG + ((threadIdx.x & 31) / G);
    int16_t* const H = P.H + (size_t)group * P.h_slot;
    const Scores sc = {P.sr_m, P.sr_n, P.sr_g};
    const uint32_t n_work = __ldg(P.n_work);
    WarpState* const ws = group_state<kTier>();

    int phase = 0;   // 0: needs a window, 1: running one, 2: the queue is empty
    uint32_t widx = 0;
    uint64_t first_arm = 0, draft_off = 0;
    uint32_t draft_len = 0;
    int ni = 0, np = 0, ns = 0;
    int k = 0;       // next sequence in the reference's order: -1 = the draft backbone, then the arms
    char* out = nullptr;
    // the arm added last: its packed byte `gl`, length, alignment type (decides the markers too)
    uint32_t memo_byte = 0xffffffffu;
    int memo_len = -1, memo_type = -1;

    // window `widx` leaves this group: res >= 0 consensus length, -1 copy the draft, -2 hand it to the next tier
    auto finish = [&](int res) {
        if (res == -1) {   // draft copy (reference src/Window.cpp:58-60,150-152)
            const uint8_t* src = P.packed + draft_off;
#pragma unroll 1
            for (int p = gl; p < (int)draft_len; p += G) {
                const int v = (src[p >> 1] >> ((p & 1) ? 0 : 4)) & 15;
                out[p] = code_to_char(v > 4 ? 4 : v);
            }
            res = (int)draft_len;
        }
        if (gl == 0) {
            if (res == -2) {
                const uint32_t at = atomicAdd(P.next_count, 1u);
                if (P.abandoned) atomicAdd(P.abandoned, 1u);
                P.next_list[at] = widx;
            } else {
                P.out_len[widx] = (uint32_t)res;
                if (P.cells) atomicAdd(P.cells, ws->cells);
            }
        }
    };
    // storage index of the k-th arm in the reference's order (prefix arms run last to first)
    auto arm_at = [&](int q) { return q < ni ? q : q < ni + np ? ni + (ni + np - 1 - q) : q; };

#pragma unroll 1
    for (;;) {
        // ---- groups without a window fetch one
        const bool want = phase == 0;
        uint32_t wi = 0;
        if (want && gl == 0) wi = atomicAdd(P.queue, 1u);
        wi = Grp<G>::shfl(wi, 0);
        int res = -3;   // -3: run it
        bool added_l = false;
        if (want) {
            if (wi >= n_work) {
                phase = 2;
            } else {
                widx = P.work[wi];
                const WinDesc w = P.win[widx];
                out = P.out + P.out_pos[widx];
                first_arm = w.first_arm; draft_off = w.draft_off; draft_len = w.draft_len;
                ni = (int)w.n_internal; np = (int)w.n_pre; ns = (int)w.n_suf;
                const uint32_t n = w.n_internal + w.n_pre + w.n_suf;
                if (gl == 0) ws->cells = 0;
                if (w.n_empty > n) res = 0;            // reference src/Window.cpp:47-49
                else if (n < 2) res = -1;
                else if (w.wtype != 0) res = -2;        // LONG windows run in the one-warp tiers
                else {
                    const ArmDesc* a = P.arms + first_arm;
#pragma unroll 1
                    for (uint32_t q = gl; q < n; q += G) {
                        const ArmDesc d = a[q];
                        added_l |= d.len > 0;
                        if (d.len) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.packed + d.off));
                    }
                }
            }
        }
        const bool added = Grp<G>::any(added_l);
        if (want && phase != 2) {
            if (res == -3 && !added) res = -1;   // no arm was added: the draft (:150-152)
            if (res == -3) {
                if (gl == 0) { ws->n_nodes = 0; ws->n_edges = 0; ws->n_al = 0; ws->n_seq = 0; ws->exact = 1; ws->clean = 0; }
                phase = 1;
                k = ni == 0 ? -1 : 0;
                memo_len = -1;
            } else {
                finish(res);
            }
        }
        __syncwarp();
        if (!warp_any(phase != 2)) break;

        // ---- the next sequence of every running group.  A read that repeats the one added last (same bases,
        // length and kind, and that one left the DAG's structure untouched) aligns identically: its weights are
        // added along the node path `cur` still holds and the group moves on to its next read, without waiting
        // for a lock-step trip (repeat_sequence in poa_kernel.cu has the argument).
        const ArmDesc* a = P.arms + first_arm;
        const int n_arms = ni + np + ns;
        GSeq s;
        bool has_seq;
        uint32_t my_byte = 0xffffffffu;
#pragma unroll 1
        for (;;) {
            s.bytes = nullptr; s.len = 0; s.nb = 2; s.head = false; s.tail = false; s.type = kNW;
            has_seq = false;
            if (phase == 1) {
#pragma unroll 1
                while (k >= 0 && k < n_arms && a[arm_at(k)].len == 0) ++k;
                if (k < 0) {   // draft as backbone only without internal arms (:95-101)
                    s.bytes = P.packed + draft_off; s.len = (int)draft_len; s.nb = 4;
                    s.head = true; s.tail = true; s.type = kNW;
                    has_seq = true;
                } else if (k < n_arms) {
                    const int q = arm_at(k);
                    const ArmDesc d = a[q];
                    s.bytes = P.packed + d.off; s.len = (int)d.len; s.nb = 2;
                    s.head = q < ni + np; s.tail = q < ni || q >= ni + np;
                    s.type = q < ni ? kNW : q < ni + np ? kLOV : kROV;
                    has_seq = true;
                }
            }
            // (lane b of the group keeps packed byte b of the read added last in a register: a group's reads have
            // at most 4 G symbols = G bytes, so the comparison costs one byte load per lane - of the bytes the
            // decode is about to read anyway)
            my_byte = 0xffffffffu;
            if (has_seq && s.nb == 2 && gl < (s.len + 3) / 4) my_byte = s.bytes[gl];
            const bool cand = has_seq && s.nb == 2 && memo_len == s.len && memo_type == s.type && s.len <= 4 * G && ws->clean != 0;
            const bool any_diff = Grp<G>::any(my_byte != memo_byte);   // (a collective: evaluated by every lane, whatever `cand`)
            const bool rep = cand && !any_diff;
            if (!warp_any(rep)) break;
            if (rep) {
                const Graph v = group_graph<kTier>();
                const int len = s.len + (s.head ? 1 : 0) + (s.tail ? 1 : 0);
#pragma unroll 1
                for (int p = 1 + gl; p < len; p += G) {
                    const int dst = v.cur[p], src = v.cur[p - 1];
#pragma unroll 1
                    for (int e = v.in_head[dst]; e != kNone; e = v.e_next[e])
                        if (v.e_src[e] == src) { v.e_w[e] = (uint16_t)(v.e_w[e] + 2); break; }
                }
                ++k;
            }
            __syncwarp();   // (every lane has taken its snapshot of the counts)
            if (rep && gl == 0) ws->n_seq += 1;
            __syncwarp();
        }
        // the read that goes through the DP now is the one a later read may repeat
        memo_len = -1;
        if (has_seq && s.nb == 2) {
            memo_len = s.len; memo_type = s.type;
            memo_byte = my_byte;
        }
        if (warp_any(has_seq)) {
            const bool ok = seq_step<kTier>(has_seq, P.fail_hist, H, s, sc);
            if (has_seq) {
                if (!ok) { finish(-2); phase = 0; }
                else {
                    ++k;
#pragma unroll 1
                    while (k < n_arms && a[arm_at(k)].len == 0) ++k;
                }
            }
        }
        // ---- groups whose window is complete: heaviest bundle, marker strip (include/Window.hpp:144)
        const bool fin = phase == 1 && k >= n_arms;
        if (warp_any(fin)) {
            bool ok = fin;
            if (ok && (long long)ws->n_nodes * 2ll * (long long)ws->n_seq > 0x7fffffffll) {
                note_fail<G>(P.fail_hist, kFailRange);
                ok = false;
            }
            int nc = heaviest_bundle_g<kTier>(ok, ok && ws->exact != 0);
            bool again = ok && nc < 0;
            if (warp_any(again)) {
                // the heaviest bundle depends on WHICH valid order the ranks are in: spoa's exact order first
                const bool sorted = topo_sort_g<kTier>(again, P.fail_hist);
                if (again && !sorted) { ok = false; again = false; }
                if (again && gl == 0) ws->exact = 1;
                __syncwarp();
                const int nc2 = heaviest_bundle_g<kTier>(again, true);
                if (again) nc = nc2;
            }
            if (fin) {
                if (ok) {
                    const int n = nc >= 2 ? nc - 2 : 0;
                    const Graph v = group_graph<kTier>();
#pragma unroll 1
                    for (int p = gl; p < n; p += G) out[p] = code_to_char(v.ninfo[v.cons[p + 1]] & 7);
                    finish(n);
                } else {
                    finish(-2);
                }
                phase = 0;
            }
        }
        __syncwarp();
    }
}

}  // namespace

cudaError_t launch_poa_group(const Params& P, int tier, int blocks, int warps_per_block, size_t smem_bytes,
                             cudaStream_t stream) {
    void (*k)(const Params) = nullptr;
    if (warps_per_block * 32 > 256) return cudaErrorInvalidConfiguration;
    switch (tier) {
        case kTierQuad: k = poa_group_kernel<kTierQuad>; break;
        case kTierHalf: k = poa_group_kernel<kTierHalf>; break;
        default: return cudaErrorInvalidConfiguration;
    }
    cudaError_t err = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (err != cudaSuccess) return err;
    k<<<blocks, warps_per_block * 32, smem_bytes, stream>>>(P);
    return cudaGetLastError();
}

}  // namespace hypo_b200
