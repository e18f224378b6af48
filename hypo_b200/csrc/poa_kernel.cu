// poa_kernel.cu — sm_100a kernels for the POA-consensus hot path.  See poa_kernel.cuh for
// the overview and DESIGN.md for the data layout and the roofline accounting.
#include "poa_kernel.cuh"
#include "poa_common.cuh"

namespace hypo_b200 {

namespace {



// Tiers that extrapolate a window's growth and abandon it early (add_sequence): the multi-tile ones with
// compile-time capacities.  There a late overflow throws away the most work; the one-tile tiers' windows
// are so small that the bookkeeping alone costs more than it saves (measured on the pipeline shape mix).
template <bool kOneTile, int kTier>
constexpr bool kProjects = kTier >= 0 && !kOneTile;

struct GState {
    uint8_t* gbase;      // tiers L: this warp's arena in global memory
    ArenaLayout L;       // tiers with run-time capacities only
    uint32_t* fail_hist; // diagnostics: why windows were abandoned (may be null)
    const uint8_t* packed_end;   // end of the readable packed slab (bulk copies stay below it)
};


// Records why the window is being abandoned in this tier; always returns false.
__device__ __noinline__ bool give_up(const GState& st, int why) {
    if (lane_id() == 0 && st.fail_hist) atomicAdd(st.fail_hist + why, 1u);
    return false;
}

// Warps per window ("team"): 1 everywhere except T1m (tier 4: two), T1 (tier 5: four) and the team variant of
// the bound-driven tiers (kTier == -2: four).  One warp owns the window and runs every phase; all of them fill
// the DP matrix together, one 128-column tile at a time each (see team_fill).  These are the tiers whose
// shared-memory footprint (or window count) leaves most warp slots of an SM empty.
template <int kTier>
constexpr int kTeamOf = (kTier == 5 || kTier == -2) ? 4 : kTier == 4 ? 2 : 1;
// the arena (and workspace slot) index of this warp inside its CTA
template <int kTier>
__device__ __forceinline__ unsigned arena_slot() { return (threadIdx.x >> 5) / kTeamOf<kTier>; }

template <bool kSmem, int kTier>
__device__ __forceinline__ Graph bind_graph(const GState& st, const ArenaLayout& L) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t* base = kSmem ? (smem + arena_slot<kTier>() * L.total) : st.gbase;
    return bind_graph_at(base, L);
}

// kTier >= 0: a tier with compile-time capacities (every offset folds into an immediate and the
// layout never has to be fetched from the per-thread GState); kTier < 0: layout from st.L.
template <bool kSmem, int kTier>
__device__ __forceinline__ Graph make_graph(const GState& st) {
    if constexpr (kTier >= 0) {
        constexpr ArenaLayout L = arena_layout(fixed_caps(kTier));
        return bind_graph<kSmem, kTier>(st, L);
    } else {
        return bind_graph<kSmem, kTier>(st, st.L);
    }
}

template <bool kSmem, int kTier>
__device__ __forceinline__ WarpState* warp_state(const GState& st) {
    extern __shared__ __align__(16) uint8_t smem[];
    if constexpr (kTier >= 0) {
        constexpr ArenaLayout L = arena_layout(fixed_caps(kTier));
        return (WarpState*)((kSmem ? (smem + arena_slot<kTier>() * L.total) : st.gbase) + L.state);
    } else {
        return (WarpState*)((kSmem ? (smem + arena_slot<kTier>() * st.L.total) : st.gbase) + st.L.state);
    }
}

// The capacities a phase checks against: constants for the fixed tiers.
template <int kTier>
__device__ __forceinline__ Caps tier_caps(const Caps& dyn) {
    if constexpr (kTier >= 0) {
        constexpr Caps c = fixed_caps(kTier);
        return c;
    } else {
        return dyn;
    }
}


// ------------------------------------------------------------------------------------------
// Sequence decode (PackedSeq<2>/<4>::unpack, reference src/PackedSeq.cpp:231-262 with the bit
// layouts of :45,:48) straight into letter codes, plus the J/O markers of SHORT windows
// (reference include/Window.hpp:30-33, src/Window.cpp:98,105,116,127).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void decode2(const uint8_t* __restrict__ src, int len, uint8_t* dst) {
#pragma unroll 1
    for (int p = lane_id(); p < len; p += 32) dst[p] = (src[p >> 2] >> (6 - 2 * (p & 3))) & 3;
}
__device__ __forceinline__ void decode4(const uint8_t* __restrict__ src, int len, uint8_t* dst) {
#pragma unroll 1
    for (int p = lane_id(); p < len; p += 32) {
        int v = (src[p >> 1] >> ((p & 1) ? 0 : 4)) & 15;
        dst[p] = v > 4 ? 4 : v;
    }
}


// ------------------------------------------------------------------------------------------
// DP fill (reference sisd_alignment_engine.cpp:263-342, initialisation :158-159,197-211,
// 229-239).  Lane l owns columns [4l, 4l+4) of each 128-column tile as two s16x2 registers.
// ------------------------------------------------------------------------------------------


// Exclusive prefix max of the lane totals across the warp, radix 4: four dependent shuffle rounds
// (1, then 3 + 3 + 1 independent ones) instead of six.  shfl_up hands the lanes below the shift
// their own value back, which is harmless once the sequence has been shifted by one lane.  The
// totals travel packed in both halves of a register, so the result is the broadcast the row needs.
// `lane0_neg` is kNegInf2 on lane 0 and 0 elsewhere.
__device__ __forceinline__ uint32_t warp_excl_max(uint32_t tot2, uint32_t lane0_neg, uint32_t neg2) {
    uint32_t e = __shfl_up_sync(kFull, tot2, 1);
    e = lane0_neg ? neg2 : e;
    {
        const uint32_t a1 = __shfl_up_sync(kFull, e, 1), a2 = __shfl_up_sync(kFull, e, 2),
                       a3 = __shfl_up_sync(kFull, e, 3);
        e = __vmaxs2(__vimax3_s16x2(e, a1, a2), a3);
    }
    {
        const uint32_t b1 = __shfl_up_sync(kFull, e, 4), b2 = __shfl_up_sync(kFull, e, 8),
                       b3 = __shfl_up_sync(kFull, e, 12);
        e = __vmaxs2(__vimax3_s16x2(e, b1, b2), b3);
    }
    return __vmaxs2(e, __shfl_up_sync(kFull, e, 16));
}


// End cell (reference :276-288,328-340): best last-column score over the candidate rows
// (NW/ROV: nodes without out-edges, LOV: every node); strictly greater => lowest rank wins.
template <typename HT>
__device__ __forceinline__ EndCell end_cell(const Graph& g, const HT* __restrict__ H, int n, int cols,
                                            int len, int type) {
    const int lane = lane_id();
    // per lane: best score, its lowest row, and how many candidate rows reach it
    int best = INT_MIN, brow = 0x7fffffff, cnt = 0;
#pragma unroll 1
    for (int r = lane; r < n; r += 32) {
        const bool cand = (type == kLOV) || ((g.rowinfo[r] >> 27) & 1);
        if (cand) {
            const int v = (int)H[(unsigned)(r + 1) * (unsigned)cols + (unsigned)len];
            if (v > best) { best = v; brow = r + 1; cnt = 1; }
            else if (v == best) ++cnt;
        }
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        const int ob = __shfl_xor_sync(kFull, best, d);
        const int orow = __shfl_xor_sync(kFull, brow, d);
        const int ocnt = __shfl_xor_sync(kFull, cnt, d);
        if (ob > best) { best = ob; brow = orow; cnt = ocnt; }
        else if (ob == best) { brow = min(brow, orow); cnt += ocnt; }
    }
    EndCell ec;
    const bool any = brow != 0x7fffffff;
    ec.row = any ? brow : 0;
    ec.col = any ? len : 0;
    ec.score = any ? best : 0;
    // a second candidate with the same score?  (only then does the exact order matter)
    const int same = any ? cnt : 0;
    ec.tie = same > 1;
    return ec;
}

// The fill works one 128-column tile at a time, all rows of a tile before the next tile (cell (i, j)
// only needs columns <= j).  The last three rows of the tile stay in registers, each with its
// exclusive prefix max, which is exactly the value the lane to the left holds in its last column:
// rows whose predecessors all lie within three ranks (rowinfo bits 28-30; 97 % of the rows of the
// 30 x 120 shape) touch neither shared memory, nor the matrix, nor the shuffle unit before the scan.
// Tiles after the first one are seeded through lane 0: its "lane to the left" is column 128t-1 of
// the same row, which the previous tile's pass left in a per-row boundary array (rows are
// non-decreasing after the horizontal pass, so that one value is also the prefix max of everything
// to the left).
// (Unrolling the row loop by three to rotate the three register sets without moves removes 13 % of
// the kernel's instructions but no longer fits the L0 instruction cache: measured 10 % slower.)
struct RowRegs {
    uint32_t x[kNR];   // the lane's four columns
    uint32_t left;     // last column of the lane to the left, in both halves
};

struct DpConst {
    uint32_t g2, mm2, nn2, let4, neg2;
    uint32_t lane0;      // non-zero on lane 0
    uint32_t row0_left;  // `left` of the virtual row 0: -inf on lane 0 of the first tile, else 0
    uint32_t xinit0;     // initial value of the lane's first register (ROV pins column 0 to 0)
    unsigned stride;     // int16 elements per matrix row
};

// Rare rows with a predecessor further back (or none at all): predecessor rows come from the matrix.
// (inlined: a call inside the row loop costs 2 % even though it is almost never taken)
template <bool kSmem, bool kMulti>
__device__ __forceinline__ uint2 relax_far(typename Mem<kSmem>::addr_t prows, uint32_t info, int rk,
                                           const int16_t* __restrict__ Hl, const RowRegs& d1, const uint32_t (&pf)[kNR],
                                           const DpConst& c, const int16_t* __restrict__ bnd_prev) {
    typedef Mem<kSmem> M;
    uint32_t x[kNR] = {c.xinit0, c.neg2};
    const int np = (info >> 16) & 0xff;
    if (np == 0) {
        // no predecessor: virtual row 0 (reference :300-301)
        const uint32_t p[kNR] = {0u, 0u};
        relax(x, p, c.row0_left, pf, c.g2);
    } else {
        typename M::addr_t pa = prows + 2u * (info & 0xffffu);
#pragma unroll 1
        for (int k = 0; k < np; ++k, pa += 2) {
            const unsigned prow = M::ld16(pa);
            if (prow == (unsigned)rk) {
                relax(x, d1.x, d1.left, pf, c.g2);
            } else {
                const uint2 q = ldg64(Hl + prow * (kMulti ? c.stride : (unsigned)kTileCols));
                const uint32_t p[kNR] = {q.x, q.y};
                uint32_t left = __shfl_up_sync(kFull, q.y, 1);
                if (c.lane0) {
                    left = c.neg2;
                    if (kMulti && bnd_prev) left = bcast16(bnd_prev[prow]);
                }
                relax(x, p, left, pf, c.g2);
            }
        }
    }
    return make_uint2(x[0], x[1]);
}

// One DP row: d1/d2/d3 hold the rows 1/2/3 ranks back; the new row replaces d3.  `seed2` is what
// lane 0 feeds into the exclusive prefix max: -inf in the first tile, the row's boundary value after.
template <bool kSmem, bool kMulti>
__device__ __forceinline__ void dp_row(uint32_t info, int rk, const RowRegs& d1, const RowRegs& d2, RowRegs& d3,
                                       const DpConst& c, typename Mem<kSmem>::addr_t prows,
                                       const int16_t* __restrict__ Hl, int16_t*& Hrow, uint32_t seed2,
                                       const int16_t* __restrict__ bnd_prev, int16_t* __restrict__ bnd_next) {
    uint32_t pf[kNR];
    profile_regs(c.let4, (info >> 24) & 7u, c.mm2, c.nn2, pf);
    // first column: NW/LOV follow the vertical rule (done by relax with diag = -inf); ROV pins it to
    // 0 (reference :229-239): lane 0 starts its first column at 0, which every candidate
    // (-inf diagonal, 0 + g vertical, g <= 0) leaves in place
    uint32_t x[kNR] = {c.xinit0, c.neg2};
    const uint32_t near = info >> kRowNearShift;
    if (near == 1u) {
        relax(x, d1.x, d1.left, pf, c.g2);
    } else if (near) {
        if (near & 1u) relax(x, d1.x, d1.left, pf, c.g2);
        if (near & 2u) relax(x, d2.x, d2.left, pf, c.g2);
        if (kRowHist >= 3 && (near & 4u)) relax(x, d3.x, d3.left, pf, c.g2);
    } else {
        const uint2 q = relax_far<kSmem, kMulti>(prows, info, rk, Hl, d1, pf, c, bnd_prev);
        x[0] = q.x; x[1] = q.y;
    }
    const uint32_t cb = warp_excl_max(scan_inlane(x, c.neg2), c.lane0, seed2);
    d3.x[0] = __vmaxs2(x[0], cb);
    d3.x[1] = __vmaxs2(x[1], cb);
    d3.left = cb;   // == last column of the lane to the left (lane 0: the seed)
    Hrow += kMulti ? c.stride : (unsigned)kTileCols;
    stg64(Hrow, d3.x[0], d3.x[1]);
    if (kMulti && bnd_next && lane_id() == 31) bnd_next[rk + 1] = (int16_t)hi16(d3.x[1]);
}

// ------------------------------------------------------------------------------------------
// Teams: four warps fill one window's matrix together (T1, and the bound-driven tiers when a launch has
// fewer windows than warps to put them on).  Warp k of the team takes the tiles k, k + 4, ...; tile t may
// compute row r once tile t - 1 has left its boundary value of that row, so the warps run as a software
// pipeline, synchronised every kTeamBlock rows through progress counters in shared memory.  The owner (warp 0)
// posts the fill in the team's mailbox and runs every other phase alone; the helpers wait for the next fill.
// ------------------------------------------------------------------------------------------
constexpr int kTeamBlock = 16;        // rows between two synchronisations of neighbouring tiles
constexpr int kTeamMaxTiles = 32;     // a read with more tiles (> 4095 symbols) is filled by the owner alone
constexpr int kTeamsPerCta = 5;
struct TeamBox {
    uint32_t seq;                      // bumped by the owner for every command (helpers poll it)
    uint32_t cmd;                      // 1 = fill, 2 = exit
    uint32_t err;                      // a wait ran into its time limit (the window is abandoned)
    uint32_t done;                     // helpers that have finished a command (3 per command)
    int len, tiles, type, bnd_len;
    int m, n, g;
    int16_t* H;
    int16_t* bnd;
    uint32_t prog[kTeamMaxTiles];      // rows completed per tile
};
__device__ __forceinline__ uint32_t ld_volatile_shared(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_shared(uint32_t* p, uint32_t v) {
    asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}
// every lane waits until *p >= v (bounded: a wait that takes seconds sets err instead of hanging the device)
__device__ __forceinline__ bool team_wait_ge(const uint32_t* p, uint32_t v, uint32_t* err) {
    if (ld_volatile_shared(p) < v) {
        const long long t0 = clock64();
#pragma unroll 1
        while (ld_volatile_shared(p) < v) {
            __nanosleep(32);
            if (clock64() - t0 > (8ll << 30) || ld_volatile_shared(err)) { st_volatile_shared(err, 1u); return false; }
        }
    }
    __threadfence_block();
    return true;
}
// the whole warp publishes *p = v after everything it has written so far
__device__ __forceinline__ void team_publish(uint32_t* p, uint32_t v) {
    __threadfence_block();
    __syncwarp();
    if (lane_id() == 0) st_volatile_shared(p, v);
}

__shared__ TeamBox g_team_box[kTeamsPerCta];   // (only the team kernels reference it)

// `bnd`: one boundary array of bnd_len int16 per tile behind the matrix (multi-tile tiers only): entry r of
// tile t's array is the last column of row r of that tile.  Tiles t0, t0 + tstep, ... are filled; `box` is the
// team's mailbox (progress counters) when several warps share the fill.
template <bool kSmem, int kTier, bool kMulti, bool kTeamed>
__device__ __forceinline__ void dp_fill_tiles(const GState& st, const Graph& g, int16_t* __restrict__ H,
                                              int16_t* __restrict__ bnd, int bnd_len, int len, int tiles, int type,
                                              Scores sc, int t0, int tstep, TeamBox* box) {
    const int lane = lane_id();
    const int n = g.n_nodes;
    const int ntiles = kMulti ? tiles : 1;
    typedef Mem<kSmem> M;
    // loop constants are made opaque so that ptxas keeps them in registers instead of
    // re-deriving them in every row
    DpConst c;
    c.g2 = opaque(bcast16(sc.g));
    c.mm2 = opaque(bcast16(sc.m - sc.g));
    c.nn2 = opaque(bcast16(sc.n - sc.g));
    c.neg2 = opaque(kNegInf2);
    c.lane0 = opaque(lane == 0 ? 1u : 0u);
    c.stride = (unsigned)ntiles * kTileCols;
    const typename M::addr_t prows = M::addr(g.prows);
    const uint32_t n_pad = (uint32_t)(n + 1) & ~1u;   // rows the two-rows-per-trip loop really computes

#pragma unroll 1
    for (int t = t0; t < ntiles; t += tstep) {
        const unsigned toff = (unsigned)t * kTileCols;
        c.let4 = opaque(*reinterpret_cast<const uint32_t*>(g.colseq + toff + lane * 4));
        c.row0_left = opaque((lane == 0 && t == 0) ? kNegInf2 : 0u);
        c.xinit0 = opaque((type == kROV && lane == 0 && t == 0) ? (kNegInf2 & 0xffff0000u) : kNegInf2);
        const int16_t* Hl = opaque_ptr(H + toff + lane * 4);
        const int16_t* bnd_prev = (kMulti && t > 0) ? bnd + (size_t)(t - 1) * bnd_len : nullptr;
        int16_t* bnd_next = (kMulti && t + 1 < ntiles) ? bnd + (size_t)t * bnd_len : nullptr;
        typename M::addr_t ri = M::addr(g.rowinfo);
        if (kMulti && bnd_next && lane == 0) bnd_next[0] = 0;   // virtual row 0: H^[0][j] = 0

        // row 0: H^[0][j] = 0
        int16_t* Hrow = opaque_ptr(H + toff + lane * 4);
        stg64(Hrow, 0u, 0u);
        RowRegs A, B, C;
        A.x[0] = A.x[1] = 0u; A.left = c.row0_left;
        B = A; C = A;

        if (kTeamed && t > 0 && !team_wait_ge(&box->prog[t - 1], min((uint32_t)kTeamBlock + 2u, n_pad), &box->err)) return;
        uint32_t info = M::ld32(ri);
        uint32_t seed = (kMulti && bnd_prev) ? bcast16(bnd_prev[1]) : c.neg2;
#ifndef HYPO_DP_ROLLED
        // two rows per trip (measured +0.8 % over the rolled loop; three per trip overflows the L0
        // instruction cache) (a spare record / matrix row pads an odd row count): the register sets
        // are rotated once per pair
#pragma unroll 1
        for (int rk = 0; rk < n; rk += 2) {
            if (kTeamed && rk != 0 && (rk & (kTeamBlock - 1)) == 0) {
                // rows < rk of this tile are complete; the rows of the next block (and the seed read one row
                // ahead) need the previous tile's boundary values
                team_publish(&box->prog[t], (uint32_t)rk);
                if (t > 0 && !team_wait_ge(&box->prog[t - 1], min((uint32_t)(rk + kTeamBlock + 2), n_pad), &box->err)) return;
            }
            const uint32_t i0 = info, i1 = M::ld32(ri + 4), s0 = seed;
            uint32_t s1 = c.neg2;
            ri += 8;
            info = M::ld32(ri);
            if (kMulti && bnd_prev) { s1 = bcast16(bnd_prev[rk + 2]); seed = bcast16(bnd_prev[rk + 3]); }
#if HYPO_ROW_HIST == 2
            dp_row<kSmem, kMulti>(i0, rk, A, B, B, c, prows, Hl, Hrow, s0, bnd_prev, bnd_next);       // new row -> B
            dp_row<kSmem, kMulti>(i1, rk + 1, B, A, A, c, prows, Hl, Hrow, s1, bnd_prev, bnd_next);   // new row -> A
#else
            dp_row<kSmem, kMulti>(i0, rk, A, B, C, c, prows, Hl, Hrow, s0, bnd_prev, bnd_next);       // new row -> C
            dp_row<kSmem, kMulti>(i1, rk + 1, C, A, B, c, prows, Hl, Hrow, s1, bnd_prev, bnd_next);   // new row -> B
            const RowRegs r = A;
            A = B; B = C; C = r;
#endif
        }
#else
        static_assert(!kTeamed, "the team fill is written for the two-rows-per-trip loop");
#pragma unroll 1
        for (int rk = 0; rk < n; ++rk) {
            const uint32_t i0 = info, s0 = seed;
            ri += 4;
            info = M::ld32(ri);   // one row ahead (the array has spare entries)
            if (kMulti && bnd_prev) seed = bcast16(bnd_prev[rk + 2]);
            dp_row<kSmem, kMulti>(i0, rk, A, B, C, c, prows, Hl, Hrow, s0, bnd_prev, bnd_next);   // new row -> C
            const RowRegs r = C;
            C = B; B = A; A = r;
        }
#endif
        if (kTeamed) team_publish(&box->prog[t], n_pad);
        else __syncwarp();   // the boundary values of this tile are read by every lane in the next one
    }
}

template <bool kSmem, int kTier, bool kMulti>
__device__ __noinline__ EndCell dp_fill_row(const GState& st, int16_t* __restrict__ H, int16_t* __restrict__ bnd,
                                        int bnd_len, int len, int tiles, int type, Scores sc) {
    const Graph g = make_graph<kSmem, kTier>(st);
    dp_fill_tiles<kSmem, kTier, kMulti, false>(st, g, H, bnd, bnd_len, len, tiles, type, sc, 0, 1, nullptr);
    return end_cell<int16_t>(g, H, g.n_nodes, (kMulti ? tiles : 1) * kTileCols, len, type);
}
// Owner side of a team fill: post the command, fill the own tiles, wait for the helpers.  A row of -1 in the
// result means a wait ran into its time limit.
template <bool kSmem, int kTier>
__device__ __noinline__ EndCell team_fill(const GState& st, int16_t* __restrict__ H, int16_t* __restrict__ bnd,
                                          int bnd_len, int len, int tiles, int type, Scores sc) {
    constexpr int kTeam = kTeamOf<kTier>;
    TeamBox* const box = &g_team_box[arena_slot<kTier>()];
    const Graph g = make_graph<kSmem, kTier>(st);
    const int lane = lane_id();
#pragma unroll 1
    for (int t = lane; t < tiles; t += 32) box->prog[t] = 0;
    uint32_t seq = 0;
    if (lane == 0) {
        box->cmd = 1; box->len = len; box->tiles = tiles; box->type = type; box->bnd_len = bnd_len;
        box->m = sc.m; box->n = sc.n; box->g = sc.g; box->H = H; box->bnd = bnd;
        seq = box->seq + 1;
    }
    seq = __shfl_sync(kFull, seq, 0);
    __threadfence_block();
    __syncwarp();
    if (lane == 0) st_volatile_shared(&box->seq, seq);
    dp_fill_tiles<kSmem, kTier, true, true>(st, g, H, bnd, bnd_len, len, tiles, type, sc, 0, kTeam, box);
    // every helper reports back before the mailbox may change again (also after a failed wait)
    bool ok = team_wait_ge(&box->done, (uint32_t)(kTeam - 1) * seq, &box->err);
    __syncwarp();
    ok = ok && ld_volatile_shared(&box->err) == 0u;
    if (!ok) {
        EndCell bad;
        bad.row = -1; bad.col = 0; bad.score = 0; bad.tie = false;
        return bad;
    }
    return end_cell<int16_t>(g, H, g.n_nodes, tiles * kTileCols, len, type);
}

// Helper warps of a team: wait for the owner's commands.
template <bool kSmem, int kTier>
__device__ __noinline__ void team_helper(const GState& st, int role) {
    constexpr int kTeam = kTeamOf<kTier>;
    TeamBox* const box = &g_team_box[arena_slot<kTier>()];
    uint32_t last = 0;
#pragma unroll 1
    for (;;) {
        uint32_t s = ld_volatile_shared(&box->seq);
        if (s == last) {
            const long long t0 = clock64();
#pragma unroll 1
            unsigned nap = 64;   // (back off: an idle helper must not take issue slots from the owners)
#pragma unroll 1
            while ((s = ld_volatile_shared(&box->seq)) == last) {
                __nanosleep(nap);
                if (nap < 2048) nap *= 2;
                if (clock64() - t0 > (128ll << 30)) return;   // (a minute: the owner is gone)
            }
        }
        __threadfence_block();
        s = __shfl_sync(kFull, s, 0);
        last = s;
        const uint32_t cmd = ld_volatile_shared(&box->cmd);
        if (cmd == 2u) return;
        {
            const Scores sc = {box->m, box->n, box->g};
            const Graph g = make_graph<kSmem, kTier>(st);
            dp_fill_tiles<kSmem, kTier, true, true>(st, g, box->H, box->bnd, box->bnd_len, box->len, box->tiles, box->type,
                                                    sc, role, kTeam, box);
        }
        __threadfence_block();
        __syncwarp();
        if (lane_id() == 0) atomicAdd(&box->done, 1u);
    }
}

// Emitted here (not at its first use) so that the hottest loop of the compact tier sits at the front
// of the kernel's code: with 27 warps in different phases the placement of the row loop relative
// to the other per-read phases decides how well the instruction caches hold (measured: 5 %).
template __device__ EndCell dp_fill_row<true, 0, false>(const GState&, int16_t* __restrict__, int16_t* __restrict__, int,
                                                    int, int, int, Scores);

// ------------------------------------------------------------------------------------------
// 32-bit DP fill: the reference's H is int32 throughout (sisd_alignment_engine.cpp:263-342), so a
// window whose scores x size leave the 16-bit range (|H^| > kMaxH16) is never refused: the last
// tier fills such a read with plain 32-bit cells - one column per lane, 32 columns per step, the
// horizontal pass as a warp prefix max with a carry between steps.  Same g-normalised values
// (H^[i][j] = H[i][j] - j*g), same row-major layout (row stride `cols`), so end cell and traceback
// are the code above instantiated for int32_t.  Built for exactness, not speed: it only ever runs
// for windows no 16-bit tier can hold.
// ------------------------------------------------------------------------------------------
template <bool kSmem, int kTier>
__device__ __noinline__ EndCell dp_fill_wide(const GState& st, int32_t* __restrict__ H, int cols, int len,
                                             int type, Scores sc) {
    const Graph g = make_graph<kSmem, kTier>(st);
    const int lane = lane_id();
    const int n = g.n_nodes;
    const int mm = sc.m - sc.g, nn = sc.n - sc.g;
    constexpr int kNeg32 = -(1 << 30);
    const int used = ((len + 1 + 31) / 32) * 32;   // <= cols (a multiple of kTileCols)
#pragma unroll 1
    for (int j = lane; j < used; j += 32) H[j] = 0;   // row 0: H^[0][j] = 0
    __syncwarp();
#pragma unroll 1
    for (int rk = 0; rk < n; ++rk) {
        const uint32_t info = g.rowinfo[rk];
        const int np = (info >> 16) & 0xff;
        const int code = (info >> 24) & 7;
        const uint16_t* pr = g.prows + (info & 0xffffu);
        int32_t* Hi = H + (size_t)(rk + 1) * (size_t)cols;
        int carry = kNeg32;
#pragma unroll 1
        for (int j0 = 0; j0 < used; j0 += 32) {
            const int j = j0 + lane;
            const int s = ((int)g.colseq[j] == code) ? mm : nn;
            int x = kNeg32;
            int k = 0;
#pragma unroll 1
            do {   // no predecessor: the virtual row 0 (reference :300-301)
                const int32_t* Hp = H + (size_t)(np ? pr[k] : 0) * (size_t)cols;
                x = max(x, Hp[j] + sc.g);
                if (j >= 1) x = max(x, Hp[j - 1] + s);
            } while (++k < np);
            if (j == 0 && type == kROV) x = 0;   // reference :229-239
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int y = __shfl_up_sync(kFull, x, d);
                if (lane >= d) x = max(x, y);
            }
            x = max(x, carry);
            carry = __shfl_sync(kFull, x, 31);
            Hi[j] = x;
        }
        __syncwarp();   // later rows read this one across lanes
    }
    return end_cell<int32_t>(g, H, n, cols, len, type);
}

// ------------------------------------------------------------------------------------------
// Traceback (reference sisd_alignment_engine.cpp:344-437).
// Preference at every cell: diagonal via in-edge 0,1,..; vertical via in-edge 0,1,..; horizontal.
// Because "diagonal through the FIRST in-edge" is tested first, a run of such moves can be
// verified by up to 32 lanes at once: lane k reaches the row of step k through the jump pointers
// (fp4, then fp), tests the equality for its step, and the longest all-true prefix is committed; the
// number of lanes used adapts to how long the last run was.  Every other move is a general step:
// all candidates of the reference's preference list are fetched at once, one per lane, and the
// lowest matching lane wins.  The owning lanes record cur[pos] = aligned node (kNone for a read-only
// column).
// ------------------------------------------------------------------------------------------

template <bool kSmem, int kTier, typename HT>
__device__ __noinline__ AlnSpan traceback_dp(const GState& st, const HT* __restrict__ H, int cols,
                                          EndCell ec, int type, Scores sc, int max_steps) {
    const Graph g = make_graph<kSmem, kTier>(st);
    const int lane = lane_id();
    int i = ec.row, j = ec.col;
    AlnSpan span;
    span.first = -1; span.last = -1;
    const unsigned ucols = (unsigned)cols;
    int hij = ec.score;   // == H[i][j], already known from the end-cell search
    const int mm = sc.m - sc.g, nn = sc.n - sc.g;
    int steps = 0;
    int chunk = 32;   // lanes used by the next speculative run: 8, 16 or 32
    bool spec = true;  // false right after a run that was cut short: its last test already failed
#pragma unroll 1
    while ((type == kROV ? (i != 0 && j != 0) : (i != 0 || j != 0)) && steps++ < max_steps) {
        if (spec && i != 0 && j != 0) {
            // ---- speculative diagonal run through first predecessors: lane k takes step k, i.e.
            // needs row fp^k(i).  k = 4a + b: a jumps of four (fp4), then b single steps
            // (fp[0] = fp4[0] = 0 keeps the walk total).
            int my_r = i;
            {
                const int a = lane >> 2, b = lane & 3;
                const int na = (chunk >> 2) - 1;
#pragma unroll 1
                for (int s = 0; s < na; ++s) {
                    const int t = g.fp4[my_r];
                    if (a > s) my_r = t;
                }
#pragma unroll
                for (int s = 0; s < 3; ++s) {
                    const int t = g.fp[my_r];
                    if (b > s) my_r = t;
                }
            }
            const int my_rn = g.fp[my_r];
            const int jj = j - lane;
            bool ok = lane < chunk && my_r != 0 && jj >= 1;
            int hp = 0;
            if (ok) {
                const int hc = (int)H[(unsigned)my_r * ucols + (unsigned)jj];
                hp = (int)H[(unsigned)my_rn * ucols + (unsigned)(jj - 1)];
                const int s = (((g.rowinfo[my_r - 1] >> 24) & 7) == g.seq[jj - 1]) ? mm : nn;
                ok = hc == hp + s;
            }
            const unsigned mask = __ballot_sync(kFull, ok);
            const int nrun = __ffs(~mask) - 1;   // leading all-true lanes (mask == ~0 -> -1; only with chunk == 32)
            const int run = nrun < 0 ? 32 : nrun;
            if (run > 0) {
                if (lane < run) g.cur[jj - 1] = g.r2n[my_r - 1];
                if (span.last < 0) span.last = j - 1;
                span.first = j - run;
                i = __shfl_sync(kFull, my_rn, run - 1);
                hij = __shfl_sync(kFull, hp, run - 1);
                j -= run;
                steps += run - 1;
                // a run that used every lane was probably cut by the chunk, not by the alignment
                // a cut below the chunk size means the next cell fails this very test (or the
                // run reached row 0 / column 0): take the general step directly
#ifndef HYPO_TB_NOSKIP
                spec = run == chunk;
#endif
                chunk = run == chunk ? min(2 * chunk, 32) : (run < 8 ? 8 : 16);
                continue;
            }
            chunk = 8;
        }
        spec = true;
        // ---- one general step.  Candidates in the reference's preference order map to lanes:
        // lane k < 15: diagonal via in-edge k; lane 15 + k: vertical via in-edge k; lane 30:
        // horizontal.  All candidate cells are fetched at once and the lowest matching lane wins.
        int ni = i, nj = j, nh = hij;
        bool found = false;
        uint32_t info = 0;
        int ps = 0, deg = 0;
        if (i != 0) {
            info = g.rowinfo[i - 1];
            ps = info & 0xffff;
            deg = (info >> 16) & 0xff;
        }
        if (deg <= 15) {
            const int slots = deg == 0 ? 1 : deg;   // no in-edge: virtual row 0 (reference :300-301)
            int pi = i, h = 0;
            bool match = false;
            if (i != 0 && lane < 30) {
                const int k = lane < 15 ? lane : lane - 15;
                const bool diag = lane < 15;
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
            } else if (lane == 30 && j != 0) {
                h = (int)H[(unsigned)i * ucols + (unsigned)(j - 1)];
                match = hij == h;
            }
            const unsigned mm_ = __ballot_sync(kFull, match);
            if (mm_ != 0u) {
                const int w = __ffs(mm_) - 1;
                ni = __shfl_sync(kFull, pi, w);
                nh = __shfl_sync(kFull, h, w);
                nj = (w < 15 || w == 30) ? j - 1 : j;
                found = true;
            }
        } else {
            // very high in-degree: the serial walk, verbatim
            const int pe = ps + deg;
            if (j != 0) {
                const int s = (((info >> 24) & 7) == g.seq[j - 1]) ? mm : nn;
#pragma unroll 1
                for (int k = ps; k < pe; ++k) {
                    const int pi = g.prows[k];
                    const int h = (int)H[(unsigned)pi * ucols + (unsigned)(j - 1)];
                    if (hij == h + s) { ni = pi; nj = j - 1; nh = h; found = true; break; }
                }
            }
            if (!found) {
#pragma unroll 1
                for (int k = ps; k < pe; ++k) {
                    const int pi = g.prows[k];
                    const int h = (int)H[(unsigned)pi * ucols + (unsigned)j];
                    if (hij == h + sc.g) { ni = pi; nj = j; nh = h; found = true; break; }
                }
            }
            if (!found && j != 0) {
                const int h = (int)H[(unsigned)i * ucols + (unsigned)(j - 1)];
                if (hij == h) { ni = i; nj = j - 1; nh = h; found = true; }
            }
        }
        if (!found) break;   // impossible for a consistent H; never spin
        if (nj != j) {
            // pair (node or -1, j-1)
            if (lane == 0) g.cur[j - 1] = (ni != i) ? g.r2n[i - 1] : kNone;
            if (span.last < 0) span.last = j - 1;
            span.first = j - 1;
        }
        i = ni; j = nj; hij = nh;
    }
    __syncwarp();
    return span;
}

// ------------------------------------------------------------------------------------------
// Graph fusion (reference graph.cpp:154-291), warp-parallel over sequence positions.
// Returns false if a capacity was exceeded (window is abandoned and re-run in a larger tier).
// ------------------------------------------------------------------------------------------

template <bool kSmem, int kTier>
__device__ __noinline__ bool add_to_graph(GState& st, const Caps& caps_dyn, int len, AlnSpan span,
                                          uint16_t* path) {
    const Caps caps = tier_caps<kTier>(caps_dyn);
    Graph g = make_graph<kSmem, kTier>(st);
    const int lane = lane_id();
    const unsigned lt_mask = (1u << lane) - 1u;
    int first = span.first, last = span.last;
    if (first < 0) { first = len; last = len - 1; }   // empty alignment: whole read is a chain (:174-182)
    const int head_n = first, tail_n = len - 1 - last;
    const int base = g.n_nodes;
    if (base + head_n + tail_n > caps.ncap) return give_up(st, kFailNodes);

    // head chain [0, first) and tail chain (last, len): fresh nodes, allocated FIRST (:194-200)
#pragma unroll 1
    for (int p = lane; p < len; p += 32) {
        int id = -1;
        if (p < first) id = base + p;
        else if (p > last) id = base + head_n + (p - last - 1);
        if (id >= 0) {
            init_node(g, id, g.seq[p]);
            g.cur[p] = (uint16_t)id;
        }
    }
    int n_nodes = base + head_n + tail_n;
    int n_al = g.n_al;
    __syncwarp();

    // aligned part [first, last]: reuse / clique lookup / new node (:206-245)
#pragma unroll 1
    for (int p0 = first; p0 <= last; p0 += 32) {
        const int p = p0 + lane;
        const bool act = p <= last;
        int x = kNone, code = 0, res = -1;
        bool need_new = false;
        if (act) {
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
                        int a = g.al_pool[blk * g.als + k];
                        if ((g.ninfo[a] & 7) == code) { res = a; need_new = false; break; }
                    }
                }
            }
        }
        const unsigned newmask = __ballot_sync(kFull, need_new);
        const int n_new = __popc(newmask);
        if (n_nodes + n_new > caps.ncap) return give_up(st, kFailNodes);
        if (need_new) res = n_nodes + __popc(newmask & lt_mask);
        n_nodes += n_new;
        // aligned-list blocks: the new node needs one; so does x if it had none
        const bool link = need_new && x != kNone;
        const bool x_needs_blk = link && g.al_blk[x] == kNone;
        const unsigned m1 = __ballot_sync(kFull, link);
        const unsigned m2 = __ballot_sync(kFull, x_needs_blk);
        if (n_al + __popc(m1) + __popc(m2) > caps.acap) return give_up(st, kFailAligned);
        // cannot happen (clique letters are distinct, <= 7 letters); checked uniformly anyway
        if (__any_sync(kFull, link && g.al_cnt[x] + 1 > g.als)) return give_up(st, kFailClique);
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
        n_al += __popc(m1) + __popc(m2);
        if (act) g.cur[p] = (uint16_t)res;
        __syncwarp();
    }
    g.n_nodes = n_nodes;
    g.n_al = n_al;

    // edges (cur[p-1] -> cur[p]), weight 1+1 per traversal (:99-115,251-265,283-288).  Every
    // node of a sequence is distinct, so lanes touch disjoint in-lists / source nodes.
    int n_edges = g.n_edges;
#pragma unroll 1
    for (int p0 = 0; p0 < len; p0 += 32) {
        const int p = p0 + lane;
        bool need_edge = false, sat = false;
        int src = 0, dst = 0, tail = kNone;
        if (p < len) {
            dst = g.cur[p];
            if (path) path[p] = (uint16_t)dst;
            if (p >= 1) {
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
        }
        const unsigned em = __ballot_sync(kFull, need_edge);
        if (n_edges + __popc(em) > caps.ecap || __any_sync(kFull, sat)) return give_up(st, kFailEdges);
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
    if (lane == 0) {
        g.ws->n_nodes = n_nodes;
        g.ws->n_al = n_al;
        g.ws->n_edges = n_edges;
        g.ws->n_seq = g.n_seq + 1;
    }
    __syncwarp();
    return true;
}

// ------------------------------------------------------------------------------------------
// Topological sort (reference graph.cpp:293-353).  The rank order spoa's iterative DFS produces
// decides every tie (end cell, heaviest bundle, branch completion, MSA column ids), so where the
// order matters it is reproduced exactly: the reference's DFS, verbatim, on one lane.  Exact sorts
// are rare (order_update keeps a valid order between them), so the serial walk costs nothing on
// SHORT workloads and 7-10 % on LONG windows.
//
// mark bits: 0-1 = node mark (0 unmarked, 1 temporary, 2 permanent), bit 2 = "do not check
// aligned nodes" (check_aligned_nodes[id] == false).
// ------------------------------------------------------------------------------------------

template <bool kSmem, int kTier>
__device__ __noinline__ bool topo_sort(const GState& st, const Caps& caps_dyn) {
    const Caps caps = tier_caps<kTier>(caps_dyn);
    const Graph g = make_graph<kSmem, kTier>(st);
    const int lane = lane_id();
    const int n = g.n_nodes;
    __syncwarp();   // the sort's scratch aliases the row records other lanes may still be reading
#pragma unroll 1
    for (int i = lane; i < n; i += 32) g.mark[i] = 0;
    __syncwarp();
    int ok = 1;
    if (lane == 0) {
        int nr = 0;
#pragma unroll 1
        for (int id = 0; id < n && ok; ++id)
            if ((g.mark[id] & 3) != 2) ok = dfs_from(g, caps, id, nr) ? 1 : 0;
    }
    ok = __shfl_sync(kFull, ok, 0);
    if (!ok) return give_up(st, kFailStack);
    __syncwarp();
#pragma unroll 1
    for (int r = lane; r < n; r += 32) g.n2r[g.r2n[r]] = (uint16_t)r;
    __syncwarp();
    return true;
}

// ------------------------------------------------------------------------------------------
// Incremental ("cheap") order maintenance.  H[node][j] does not depend on which topological order
// the rows are processed in; spoa's exact DFS order only decides ties (end cell among equal
// candidates, heaviest bundle, MSA columns).  So between exact sorts the warp keeps a VALID
// topological order in which aligned cliques stay contiguous, by inserting the nodes the read just
// created, in path order:
//   * a new member of an existing clique goes to the end of that clique's block,
//   * any other new node goes immediately before the block of the next pre-existing column on the
//     read's path (or to the very end).
// Anchors are non-decreasing along the path, so new node i (path order) with anchor a_i gets rank
// a_i + i and an old node of rank r moves to r + #{i : a_i <= r}.  The exact order is re-derived
// (topo_sort) only when an alignment's end cell is tied and before the consensus is extracted.
// oracle/poa_oracle.c replays this scheme in checked mode and asserts it is a valid order.
// ------------------------------------------------------------------------------------------
template <bool kSmem, int kTier>
__device__ __noinline__ void order_update(const GState& st, int len, int nb) {
    const Graph g = make_graph<kSmem, kTier>(st);
    const int lane = lane_id();
    const int n = g.n_nodes;
    uint16_t* anch = g.anch;   // per position (the row records are dead here)
    uint16_t* newa = g.newa;   // per new node, path order

    // pass A (reverse): anchor of every position that holds a new node
    int carry = nb;   // block start of the next pre-existing column to the right (nb = none: the end)
#pragma unroll 1
    for (int p0 = ((len - 1) / 32) * 32; p0 >= 0; p0 -= 32) {
        const int p = p0 + lane;
        int colmin = 0x7fffffff;   // block start if column p is pre-existing (old node or new clique member)
        int mate_anchor = -1;
        bool is_new = false;
        if (p < len) {
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
        // exclusive suffix min within the chunk, then the carry from the chunks to the right
        int suf = colmin;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) suf = min(suf, __shfl_down_sync(kFull, suf, d));
        int excl = __shfl_down_sync(kFull, suf, 1);
        if (lane == 31) excl = 0x7fffffff;
        excl = min(excl, carry);
        if (is_new) anch[p] = (uint16_t)(mate_anchor >= 0 ? mate_anchor : excl);
        carry = min(carry, __shfl_sync(kFull, suf, 0));
    }
    __syncwarp();
    // pass B (forward): compact the new nodes in path order, give them their ranks
    int K = 0;
#pragma unroll 1
    for (int p0 = 0; p0 < len; p0 += 32) {
        const int p = p0 + lane;
        const int v = p < len ? (int)g.cur[p] : 0;
        const bool is_new = p < len && v >= nb;
        const unsigned m = __ballot_sync(kFull, is_new);
        if (is_new) {
            const int i = K + __popc(m & ((1u << lane) - 1u));
            const int a = anch[p];
            newa[i] = (uint16_t)a;
            g.n2r[v] = (uint16_t)(a + i);
        }
        K += __popc(m);
    }
    __syncwarp();
    // old nodes move right by the number of new nodes anchored at or before them
    if (K > 0) {
#pragma unroll 1
        for (int v = lane; v < nb; v += 32) {
            const int r = g.n2r[v];
            int lo = 0, hi = K;   // first index with newa[idx] > r
#pragma unroll 1
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if ((int)newa[mid] <= r) lo = mid + 1; else hi = mid;
            }
            g.n2r[v] = (uint16_t)(r + lo);
        }
        __syncwarp();
#pragma unroll 1
        for (int v = lane; v < n; v += 32) g.r2n[g.n2r[v]] = (uint16_t)v;
    }
    __syncwarp();
}

// Per-rank row records for the DP and the traceback: predecessor rows in in-edge order (CSR),
// letter code + sink flag, first predecessor row.
template <bool kSmem, int kTier>
__device__ __noinline__ void build_rows(const GState& st) {
    const Graph g = make_graph<kSmem, kTier>(st);
    const int lane = lane_id();
    const int n = g.n_nodes;
    int base = 0;
#pragma unroll 1
    for (int r0 = 0; r0 < n; r0 += 32) {
        const int r = r0 + lane;
        int v = 0, deg = 0, code = 0;
        if (r < n) {
            v = g.r2n[r];
            deg = g.in_deg[v];
            const int info = g.ninfo[v];
            code = (info & 7) | ((info & 8) ? 0 : 8);   // bit 3 = sink (no out-edges)
        }
        int off = deg;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int y = __shfl_up_sync(kFull, off, d);
            if (lane >= d) off += y;
        }
        const int total = __shfl_sync(kFull, off, 31);
        off = base + off - deg;
        if (r < n) {
            int k = off, first = 0;
            // near rows: every predecessor row is one of the three the DP still holds in registers
            uint32_t near = (deg == 0 && r == 0) ? 1u : 0u;
            bool far = deg == 0 && r != 0;
#pragma unroll 1
            for (int e = g.in_head[v]; e != kNone; e = g.e_next[e]) {
                const int prow = g.n2r[g.e_src[e]] + 1;
                if (k == off) first = prow;
                g.prows[k++] = (uint16_t)prow;
                const int dist = r + 1 - prow;
                if (dist >= 1 && dist <= kRowHist) near |= 1u << (dist - 1); else far = true;
            }
            if (far) near = 0u;
            g.fp[r + 1] = (uint16_t)first;
            g.rowinfo[r] = (uint32_t)off | ((uint32_t)deg << 16) | ((uint32_t)code << 24) | (near << kRowNearShift);
        }
        base += total;
    }
    // spare records behind the last row (the fill reads one record ahead)
    if (lane < 4) g.rowinfo[n + lane] = 1u << kRowNearShift;
    if (lane == 0) { g.fp[0] = 0; g.fp4[0] = 0; }
    __syncwarp();
    // jump pointers for the traceback's speculative runs
#pragma unroll 1
    for (int r = 1 + lane; r <= n; r += 32) g.fp4[r] = g.fp[g.fp[g.fp[g.fp[r]]]];
    __syncwarp();
}

// ------------------------------------------------------------------------------------------
// Heaviest bundle + branch completion (reference graph.cpp:610-705).  Serial on lane 0.
// Returns the consensus length; nodes in g.cons[0..len).
// ------------------------------------------------------------------------------------------

// With `exact_order` false the ranks are only SOME valid clique-contiguous topological order.  The
// per-node part of the traversal (scores, chosen predecessor) only looks at in-edges in stored order
// and at predecessor scores, so it is the same under every valid order; the rank order itself decides
// (a) which node wins when several reach the maximal score and (b) everything branch_completion does.
// In either case the function returns -1 and the caller re-derives spoa's exact order first.
template <bool kSmem, int kTier>
__device__ __noinline__ int heaviest_bundle(const GState& st, bool exact_order) {
    const Graph g = make_graph<kSmem, kTier>(st);
    const int lane = lane_id();
    const int n = g.n_nodes;
    int len = 0;
    __syncwarp();
    // Phase 1, one node per lane: the heaviest in-edge.  Which of several equally heavy in-edges wins
    // depends on the predecessors' path scores, so such nodes are only marked (kTiePred) and resolved
    // in the serial pass; g.cons temporarily holds the winning weight.
    constexpr int kTiePred = 0xFFFE;
#pragma unroll 1
    for (int v = lane; v < n; v += 32) {
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
    if (lane == 0) {
        // Phase 2, serial in rank order: path scores (reference graph.cpp:616-634)
        int best = 0, ties = 0, sb = -1;   // sb == g.score[best]
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
            if (v == best) sb = sv;   // node 0 competes with its real score once it has been visited
            if (sb < sv) { best = v; sb = sv; ties = 1; }
            else if (sb == sv) ++ties;
        }
        if (!exact_order && (ties > 1 || (g.ninfo[best] & 8))) best = -1;
        int guard = 0;
#pragma unroll 1
        while (best >= 0 && (g.ninfo[best] & 8) && guard++ <= n) best = branch_completion(g, g.n2r[best]);
        // backtrack (reversed in place afterwards)
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
    len = __shfl_sync(kFull, len, 0);
    __syncwarp();
    return len;
}

// The path scores of the heaviest bundle are sums of edge weights (2 per traversal): the reference
// accumulates them in int64 (graph.cpp:610-634), this kernel in int32 - exact as long as
// nodes x 2 x sequences stays below 2^31, which holds for every window a 16-bit node id and a 16-bit
// edge weight admit except the absurd corner (> 32 K nodes AND > 16 K reads); that corner is refused.
__device__ __forceinline__ bool bundle_fits(const GState& st, const WarpState* ws) {
    if ((long long)ws->n_nodes * 2ll * (long long)ws->n_seq <= 0x7fffffffll) return true;
    return give_up(st, kFailRange);
}


#ifdef HYPO_TMA_STAGE
// ------------------------------------------------------------------------------------------
// TMA staging of read bytes (north_star: "PackedSeq 2-bit read-segment batches staged from HBM to shared
// memory via TMA").  While read k is aligned, the packed bytes of read k+1 travel from the slab into a
// per-warp landing buffer with ONE 1-D bulk copy (cp.async.bulk + mbarrier, issued by lane 0 right after
// read k has been decoded, so the whole alignment of read k hides the copy); read k+1 is then decoded
// from shared memory.  The copy covers the 16-byte aligned window around the read's bytes.
// ------------------------------------------------------------------------------------------
struct StageCtl {
    unsigned long long mbar;
    int32_t pending;     // arm index (low 31 bits) whose bytes are in flight / in the buffer, -1 none
    uint32_t phase;      // parity of the mbarrier phase the next wait completes
};
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void stage_init(StageCtl* c) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&c->mbar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    c->pending = -1;
    c->phase = 0;
}
__device__ __forceinline__ void stage_issue(StageCtl* c, uint8_t* buf, const uint8_t* src, uint32_t bytes) {
    const uint32_t mb = smem_u32(&c->mbar);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic reads of the buffer
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(buf)), "l"(src), "r"(bytes), "r"(mb) : "memory");
}
__device__ __forceinline__ void stage_wait(StageCtl* c, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "STAGE_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra STAGE_DONE;\n"
        "bra STAGE_WAIT;\n"
        "STAGE_DONE:\n"
        "}\n" ::"r"(smem_u32(&c->mbar)), "r"(parity) : "memory");
}
template <bool kSmem, int kTier>
__device__ __forceinline__ uint8_t* stage_buf(const GState& st, StageCtl** ctl, uint32_t* cap) {
    extern __shared__ __align__(16) uint8_t smem[];
    if constexpr (kSmem && kTier >= 0) {
        constexpr ArenaLayout L = arena_layout(fixed_caps(kTier));
        static_assert(kTeamOf<kTier> == 1, "TMA staging is not built for the team tiers");
        uint8_t* base = smem + (threadIdx.x >> 5) * L.total;
        *ctl = (StageCtl*)(base + L.stage + L.stage_cap);
        *cap = L.stage_cap;
        return base + L.stage;
    } else {
        *ctl = nullptr;
        *cap = 0;
        return nullptr;
    }
}
// A window starts: a copy the previous window left in flight (it was abandoned, or its last read was
// longer than the tier) must have landed before the buffer and the barrier are used again.
template <bool kSmem, int kTier>
__device__ __forceinline__ void stage_drain(const GState& st) {
    StageCtl* c; uint32_t cap;
    if (stage_buf<kSmem, kTier>(st, &c, &cap) == nullptr) return;
    __syncwarp();
    if (c->pending >= 0) {
        stage_wait(c, c->phase);
        __syncwarp();
        if (lane_id() == 0) { c->phase ^= 1u; c->pending = -1; }
    }
    __syncwarp();
}
#endif

// ------------------------------------------------------------------------------------------
// One sequence: decode, align, fuse, re-sort.
// ------------------------------------------------------------------------------------------
struct SeqSrc {
    const uint8_t* bytes;   // packed source (nullptr => ASCII consensus in `ascii`)
    const char* ascii;
    int len;                // bases without markers
    int nb;                 // 2 or 4 bits per base
    bool head, tail;        // J / O markers
    int type;
#ifdef HYPO_TMA_STAGE
    int64_t arm_idx;        // identity of this read (index in the arm table), -1 if it is not an arm
    int64_t next_idx;       // the read that follows in the window's order (-1: none): staged by TMA
    const uint8_t* next_bytes;
    int next_len;
#endif
};

template <bool kSmem, bool kOneTile, int kTier, bool kWide>
__device__ __noinline__ bool add_sequence(GState& st, const Caps& caps_dyn, int16_t* H, const SeqSrc& s,
                                          Scores sc, uint16_t* path) {
    const Caps caps = tier_caps<kTier>(caps_dyn);
    const Graph g = make_graph<kSmem, kTier>(st);
    const int lane = lane_id();
    const int len = s.len + (s.head ? 1 : 0) + (s.tail ? 1 : 0);
    if (len > caps.lcap) return give_up(st, kFailLen);
    uint8_t* dst = g.seq + (s.head ? 1 : 0);
#ifdef HYPO_TMA_STAGE
    StageCtl* sctl; uint32_t scap;
    uint8_t* const sbuf = stage_buf<kSmem, kTier>(st, &sctl, &scap);
    if (sbuf && s.bytes && s.nb == 2) {
        const int32_t id = (int32_t)(s.arm_idx & 0x7fffffff);
        if (s.arm_idx >= 0 && sctl->pending == id) {
            stage_wait(sctl, sctl->phase);   // the bytes were put on their way while the previous read was aligned
            __syncwarp();
            decode2(sbuf + ((uintptr_t)s.bytes & 15u), s.len, dst);
            __syncwarp();
            if (lane == 0) { sctl->phase ^= 1u; sctl->pending = -1; }
        } else {
            decode2(s.bytes, s.len, dst);
        }
        __syncwarp();
        // the read after this one: one bulk copy of the aligned window around its bytes
        if (s.next_idx >= 0 && lane == 0 && sctl->pending < 0) {
            const uintptr_t a0 = (uintptr_t)s.next_bytes & ~(uintptr_t)15;
            const uint32_t nbytes = (uint32_t)((((uintptr_t)s.next_bytes + (uint32_t)(s.next_len + 3) / 4 + 15) & ~(uintptr_t)15) - a0);
            if (nbytes <= scap && a0 + nbytes <= (uintptr_t)st.packed_end) {
                stage_issue(sctl, sbuf, (const uint8_t*)a0, nbytes);
                sctl->pending = (int32_t)(s.next_idx & 0x7fffffff);
            }
        }
    } else
#endif
    if (s.bytes) {
        if (s.nb == 2) decode2(s.bytes, s.len, dst); else decode4(s.bytes, s.len, dst);
    } else {
#pragma unroll 1
        for (int p = lane; p < s.len; p += 32) {
            const char c = s.ascii[p];
            dst[p] = c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : 4;
        }
    }
    if (lane == 0) {
        if (s.head) g.seq[0] = kCodeJ;
        if (s.tail) g.seq[len - 1] = kCodeO;
        g.colseq[0] = 7;   // column 0 and the padding columns match no letter
    }
    {
        const int ncols = (kOneTile ? 1 : (len + 1 + kTileCols - 1) / kTileCols) * kTileCols;
#pragma unroll 1
        for (int j = len + 1 + lane; j < ncols; j += 32) g.colseq[j] = 7;
    }
    __syncwarp();

    AlnSpan span;
    span.first = -1; span.last = -1;
    WarpState* const ws = g.ws;
    const int nodes_before = g.n_nodes, edges_before = g.n_edges;
    // A DAG on course to outgrow this tier is handed on now, not after most of its reads have been
    // aligned: growth per read since the second sequence, extrapolated over the reads to come.  The
    // projection travels with the window (Params::need) so that later tiers it would not fit either
    // pass it on without work.  Only where the window runs is affected, never its result.
    if constexpr (kProjects<kOneTile, kTier>) {
        if (g.n_seq == 2) {
            if (lane == 0) ws->base = (uint32_t)nodes_before | ((uint32_t)edges_before << 16);
        } else if (g.n_seq >= kProjectFrom) {
            const uint32_t b = ws->base;
            const int left = ws->n_total - g.n_seq, seen = g.n_seq - 2;
            const int dn = nodes_before - (int)(b & 0xffffu), de = edges_before - (int)(b >> 16);
            const int capn = caps.ncap + caps.ncap / 8, cape = caps.ecap + caps.ecap / 8;
            // (growth slows as the DAG fills - later reads' errors coincide with nodes that exist - so
            // only 3/4 of the linear extrapolation is held against the capacity)
            if (3 * dn * left > 4 * (capn - nodes_before) * seen || 3 * de * left > 4 * (cape - edges_before) * seen) {
                const int pn = nodes_before + 3 * dn * left / (4 * seen);
                const int pe = edges_before + 3 * de * left / (4 * seen);
                if (lane == 0) ws->need = (uint32_t)min(pn, 0xffff) | ((uint32_t)min(pe, 0xffff) << 16);
                __syncwarp();
                return give_up(st, kFailProjected);
            }
        }
    }
    if (nodes_before > 0) {   // reference sisd_alignment_engine.cpp:249-251
        const int tiles = kOneTile ? 1 : (len + 1 + kTileCols - 1) / kTileCols;
        const int cols = tiles * kTileCols;
        // 16-bit range guard: |H^| <= S*(rows+cols) and <= 2*S*cols
        const int S = max(max(abs(sc.m), abs(sc.n)), abs(sc.g));
        const bool narrow = !(S * (nodes_before + 1 + cols) > kMaxH16 || 2 * S * cols > kMaxH16);
        if (!kWide && !narrow) return give_up(st, kFailRange);
        // (the reference's matrix of this read: (nodes + 1) x (len + 1), sisd_alignment_engine.cpp:52-75)
        if (lane == 0) ws->cells += (unsigned long long)(nodes_before + 1) * (unsigned long long)(len + 1);
        unsigned lines;   // 128-byte lines of the matrix this read leaves behind
        if (!kWide || narrow) {
            // boundary arrays of the multi-tile fill live behind the matrix slot
            const int bnd_len = caps.ncap + 4;
            int16_t* bnd = H + (size_t)(caps.ncap + 4) * (size_t)(caps.tiles * kTileCols);
            // (teams: reads of one tile, or of more tiles than the mailbox has counters, are left to the owner)
            const bool team = kTeamOf<kTier> > 1 && tiles > 1 && tiles <= kTeamMaxTiles;
            EndCell ec;
            if constexpr (kTeamOf<kTier> > 1) {
                ec = team ? team_fill<kSmem, kTier>(st, H, bnd, bnd_len, len, tiles, s.type, sc)
                          : dp_fill_row<kSmem, kTier, !kOneTile>(st, H, bnd, bnd_len, len, tiles, s.type, sc);
                if (ec.row < 0) return give_up(st, kFailTeam);
            } else {
                ec = dp_fill_row<kSmem, kTier, !kOneTile>(st, H, bnd, bnd_len, len, tiles, s.type, sc);
            }
            if (ec.tie && !ws->exact) {
                // the reference breaks this tie by rank in ITS order: derive it and redo the fill
                if (!topo_sort<kSmem, kTier>(st, caps)) return false;
                if (lane == 0) ws->exact = 1;
                __syncwarp();
                build_rows<kSmem, kTier>(st);
                if constexpr (kTeamOf<kTier> > 1) {
                    ec = team ? team_fill<kSmem, kTier>(st, H, bnd, bnd_len, len, tiles, s.type, sc)
                              : dp_fill_row<kSmem, kTier, !kOneTile>(st, H, bnd, bnd_len, len, tiles, s.type, sc);
                    if (ec.row < 0) return give_up(st, kFailTeam);
                } else {
                    ec = dp_fill_row<kSmem, kTier, !kOneTile>(st, H, bnd, bnd_len, len, tiles, s.type, sc);
                }
            }
            span = traceback_dp<kSmem, kTier, int16_t>(st, H, cols, ec, s.type, sc, nodes_before + len + 4);
            lines = (unsigned)(nodes_before + 1) * (unsigned)cols / 64u;
        } else {
            // scores x size beyond int16: 32-bit cells, as the reference computes them (the slot of a
            // kWide launch is sized for them)
            int32_t* H32 = reinterpret_cast<int32_t*>(H);
            EndCell ec = dp_fill_wide<kSmem, kTier>(st, H32, cols, len, s.type, sc);
            if (ec.tie && !ws->exact) {
                if (!topo_sort<kSmem, kTier>(st, caps)) return false;
                if (lane == 0) ws->exact = 1;
                __syncwarp();
                build_rows<kSmem, kTier>(st);
                ec = dp_fill_wide<kSmem, kTier>(st, H32, cols, len, s.type, sc);
            }
            span = traceback_dp<kSmem, kTier, int32_t>(st, H32, cols, ec, s.type, sc, nodes_before + len + 4);
            lines = (unsigned)(nodes_before + 1) * (unsigned)cols / 32u;
        }
        // The matrix of this read is dead now.  Drop its lines from L2 instead of letting them be
        // written back: without this every DP row ends up in HBM (1 TB per million windows) just
        // to be overwritten by the next read.
        {
            char* hb = reinterpret_cast<char*>(H);
#pragma unroll 1
            for (unsigned l = lane; l < lines; l += 32)
                asm volatile("discard.global.L2 [%0], 128;" ::"l"(hb + (size_t)l * 128) : "memory");
        }
    }
    if (!add_to_graph<kSmem, kTier>(st, caps, len, span, path)) return false;
    // a read that only re-walks existing nodes and edges leaves the DAG's structure, hence every
    // order, unchanged
    const int nodes_after = ws->n_nodes;
    if (nodes_after == nodes_before && ws->n_edges == edges_before) {
        if (lane == 0) ws->clean = 1;
        __syncwarp();
        return true;
    }
    // new nodes must be ranked; new edges alone keep the current order valid (they follow it),
    // but either may change what spoa's DFS would produce
    if (nodes_after != nodes_before) order_update<kSmem, kTier>(st, len, nodes_before);
    if (lane == 0) { ws->exact = 0; ws->clean = 0; }
    __syncwarp();
    build_rows<kSmem, kTier>(st);
    return true;
}

// ------------------------------------------------------------------------------------------
// Repeated reads.  Most reads of a real window are identical (Illumina reads of a 100 bp window are error-free
// four times out of five).  If the sequence added last left the DAG's structure untouched (it only re-walked
// existing nodes and edges: WarpState::clean) and the next sequence has the same bases, length and alignment
// type, then its DP matrix - a function of the DAG's structure, the order and the sequence - is the same matrix,
// its traceback the same path and its fusion the same node path, which `cur` still holds: the reference would
// compute all of that again and arrive at exactly the weight increments below (graph.cpp:99-115,283-288).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool same_packed(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int nbytes) {
    bool diff = false;
#pragma unroll 1
    for (int i = lane_id(); i < nbytes; i += 32) diff |= a[i] != b[i];
    return !__any_sync(kFull, diff);
}
template <bool kSmem, int kTier>
__device__ __noinline__ void repeat_sequence(const GState& st, int len, uint16_t* path) {
    const Graph g = make_graph<kSmem, kTier>(st);
    const int lane = lane_id();
#pragma unroll 1
    for (int p = lane; p < len; p += 32) {
        const int dst = g.cur[p];
        if (path) path[p] = (uint16_t)dst;
        if (p >= 1) {
            const int src = g.cur[p - 1];
#pragma unroll 1
            for (int e = g.in_head[dst]; e != kNone; e = g.e_next[e])
                if (g.e_src[e] == src) { g.e_w[e] = (uint16_t)(g.e_w[e] + 2); break; }
        }
    }
    __syncwarp();   // (every lane has taken its snapshot of the counts)
    if (lane == 0) g.ws->n_seq = g.n_seq + 1;
    __syncwarp();
}

// ------------------------------------------------------------------------------------------
// Window driver (reference src/Window.cpp:44-254)
// ------------------------------------------------------------------------------------------
template <bool kSmem, bool kOneTile, int kTier, bool kWide>
__device__ __forceinline__ int run_short(GState& g, const Params& P, const Caps& caps, int16_t* H,
                                         const WinDesc& w, char* out) {
    const int lane = lane_id();
    const ArmDesc* a = P.arms + w.first_arm;
    const Scores sc = {P.sr_m, P.sr_n, P.sr_g};
    const int n_arms = w.n_internal + w.n_pre + w.n_suf;
    // arms_added (reference :90,106,117,128)
    bool added = false;
    int n_added = 0;
#pragma unroll 1
    for (int k = lane; k < n_arms; k += 32) {
        const ArmDesc d = a[k];
        if constexpr (kProjects<kOneTile, kTier>) n_added += d.len > 0;
        else added |= d.len > 0;
        // the reads' packed bytes are needed one read at a time: pull them into L2 now
        if (d.len) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.packed + d.off));
    }
    if constexpr (kProjects<kOneTile, kTier>) {
        n_added = __reduce_add_sync(kFull, n_added);
        added = n_added != 0;
    } else {
        added = __any_sync(kFull, added);
    }
    if (!added) return -1;   // caller copies the draft (:150-152)

    WarpState* const ws = warp_state<kSmem, kTier>(g);
    if (lane == 0) {
        ws->n_nodes = 0; ws->n_edges = 0; ws->n_al = 0; ws->n_seq = 0; ws->exact = 1; ws->clean = 0;
        if constexpr (kProjects<kOneTile, kTier>) { ws->n_total = n_added + (w.n_internal == 0); ws->base = 0; ws->need = 0; }
    }
    // the arm added last (repeat_sequence): its bytes, length and kind (0 internal, 1 prefix, 2 suffix)
    const uint8_t* memo = nullptr;
    int memo_len = -1, memo_kind = -1;
    auto repeated = [&](const SeqSrc& q, int kind) -> bool {
        const bool cand = memo != nullptr && ws->clean != 0 && memo_kind == kind && memo_len == q.len;
        if (cand && same_packed(memo, q.bytes, (q.len + 3) / 4)) {
            repeat_sequence<kSmem, kTier>(g, q.len + (q.head ? 1 : 0) + (q.tail ? 1 : 0), nullptr);
            return true;
        }
        memo = q.bytes; memo_len = q.len; memo_kind = kind;
        return false;
    };
    __syncwarp();
    SeqSrc s;
    s.ascii = nullptr;
#ifdef HYPO_TMA_STAGE
    stage_drain<kSmem, kTier>(g);
    s.arm_idx = -1; s.next_idx = -1; s.next_bytes = nullptr; s.next_len = 0;
#endif
    if (w.n_internal == 0) {   // draft as backbone only without internal arms (:95-101)
        s.bytes = P.packed + w.draft_off; s.len = w.draft_len; s.nb = 4;
        s.head = true; s.tail = true; s.type = kNW;
        if (!add_sequence<kSmem, kOneTile, kTier, kWide>(g, caps, H, s, sc, nullptr)) return -2;
    }
    s.nb = 2;
#ifdef HYPO_TMA_STAGE
    {
        // one loop over the reference's order - internal arms (:102-110), prefix arms last to first
        // (:112-121, kLOV), suffix arms (:123-132, kROV) - so that every read knows its successor
        const int ni = (int)w.n_internal, np = (int)w.n_pre;
        auto arm_at = [&](int j) { return j < ni ? j : j < ni + np ? ni + (ni + np - 1 - j) : j; };
        auto next_nonempty = [&](int j) { while (j < n_arms && a[arm_at(j)].len == 0) ++j; return j; };
        int j = next_nonempty(0);
#pragma unroll 1
        while (j < n_arms) {
            const int jn = next_nonempty(j + 1);
            const int k = arm_at(j);
            s.bytes = P.packed + a[k].off; s.len = a[k].len;
            s.head = k < ni + np; s.tail = k < ni || k >= ni + np;
            s.type = k < ni ? kNW : k < ni + np ? kLOV : kROV;
            s.arm_idx = (int64_t)(w.first_arm + k);
            s.next_idx = -1;
            if (jn < n_arms) {
                const int kn = arm_at(jn);
                s.next_idx = (int64_t)(w.first_arm + kn); s.next_bytes = P.packed + a[kn].off; s.next_len = a[kn].len;
            }
            if (!add_sequence<kSmem, kOneTile, kTier, kWide>(g, caps, H, s, sc, nullptr)) return -2;
            j = jn;
        }
    }
#else
#pragma unroll 1
    for (uint32_t k = 0; k < w.n_internal; ++k) {   // :102-110
        if (a[k].len == 0) continue;
        s.bytes = P.packed + a[k].off; s.len = a[k].len; s.head = true; s.tail = true; s.type = kNW;
        if (repeated(s, 0)) continue;
        if (!add_sequence<kSmem, kOneTile, kTier, kWide>(g, caps, H, s, sc, nullptr)) return -2;
    }
    const ArmDesc* pre = a + w.n_internal;
#pragma unroll 1
    for (int k = (int)w.n_pre - 1; k >= 0; --k) {   // :112-121, reverse order, kLOV
        if (pre[k].len == 0) continue;
        s.bytes = P.packed + pre[k].off; s.len = pre[k].len; s.head = true; s.tail = false; s.type = kLOV;
        if (repeated(s, 1)) continue;
        if (!add_sequence<kSmem, kOneTile, kTier, kWide>(g, caps, H, s, sc, nullptr)) return -2;
    }
    const ArmDesc* suf = pre + w.n_pre;
#pragma unroll 1
    for (uint32_t k = 0; k < w.n_suf; ++k) {   // :123-132, kROV
        if (suf[k].len == 0) continue;
        s.bytes = P.packed + suf[k].off; s.len = suf[k].len; s.head = false; s.tail = true; s.type = kROV;
        if (repeated(s, 2)) continue;
        if (!add_sequence<kSmem, kOneTile, kTier, kWide>(g, caps, H, s, sc, nullptr)) return -2;
    }
#endif
    // The heaviest bundle rarely depends on WHICH valid order the ranks are in; only then is spoa's
    // exact order derived first.
    if (!bundle_fits(g, ws)) return -2;
    int nc = heaviest_bundle<kSmem, kTier>(g, ws->exact != 0);
    if (nc < 0) {
        if (!topo_sort<kSmem, kTier>(g, caps)) return -2;
        if (lane == 0) ws->exact = 1;
        __syncwarp();
        nc = heaviest_bundle<kSmem, kTier>(g, true);
    }
    // set_marked_consensus: strip first and last character (reference include/Window.hpp:144)
    const int n = nc >= 2 ? nc - 2 : 0;
    const Graph v = make_graph<kSmem, kTier>(g);
#pragma unroll 1
    for (int p = lane; p < n; p += 32) out[p] = code_to_char(v.ninfo[v.cons[p + 1]] & 7);
    return n;
}

// LONG windows: two rounds with the lr scores, all kNW, support counts and
// curation (reference src/Window.cpp:156-254, graph.cpp:371-388,533-568).
template <bool kSmem, bool kOneTile, int kTier, bool kWide>
__device__ __forceinline__ int run_long(GState& g, const Params& P, const Caps& caps, int16_t* H,
                                        const WinDesc& w, char* out, uint16_t* paths, uint64_t p_slot) {
    const int lane = lane_id();
    const ArmDesc* a = P.arms + w.first_arm;
    const Scores sc = {P.lr_m, P.lr_n, P.lr_g};
    const int n_arms = w.n_internal + w.n_pre + w.n_suf;
    int n_added = 0;
#pragma unroll 1
    for (int k = lane; k < n_arms; k += 32) n_added += a[k].len > 0;
    n_added = __reduce_add_sync(kFull, n_added);
    if (n_added == 0) return -1;
    // (UINT)std::floor(_num_internal * _cThresh), float _cThresh = 0.4 (reference :28,245)
    const uint32_t thres = (uint32_t)floorf(__fmul_rn((float)w.n_internal, 0.4f));

    // path slot: [0, n_added+2) 32-bit start offsets (as two u16 each) - one per sequence that is
    // actually added (seq 0 + the non-empty arms) plus the end marker - then node ids.  (Zero-length
    // arms are never added, reference src/Window.cpp:182, so they take no slot.)
    const uint64_t phdr = 2ull * ((uint64_t)n_added + 2);
    if (phdr > p_slot) { give_up(g, kFailPaths); return -2; }
    uint32_t* pstart = reinterpret_cast<uint32_t*>(paths);
    uint16_t* pnodes = paths + phdr;
    const uint64_t pcap = p_slot - phdr;

    int n_cons = 0;
#pragma unroll 1
    for (int round = 0; round < 2; ++round) {
        WarpState* const ws = warp_state<kSmem, kTier>(g);
        __syncwarp();
        if (lane == 0) {
            ws->n_nodes = 0; ws->n_edges = 0; ws->n_al = 0; ws->n_seq = 0; ws->exact = 1; ws->clean = 0;
            ws->n_total = n_added + 1; ws->base = 0; ws->need = 0;
        }
        const uint8_t* memo = nullptr;   // the arm added last (repeat_sequence)
        int memo_len = -1;
        __syncwarp();
        uint32_t used = 0;
        SeqSrc s;
        s.head = false; s.tail = false; s.type = kNW;
#ifdef HYPO_TMA_STAGE
        stage_drain<kSmem, kTier>(g);
        s.arm_idx = -1; s.next_idx = -1; s.next_bytes = nullptr; s.next_len = 0;
#endif
        auto add = [&](const SeqSrc& q) -> bool {
            if (used + (uint32_t)q.len > pcap) return give_up(g, kFailPaths);
            if (lane == 0) pstart[ws->n_seq] = used;
            bool ok = true;
            const bool cand = q.nb == 2 && memo != nullptr && ws->clean != 0 && memo_len == q.len;
            if (cand && same_packed(memo, q.bytes, (q.len + 3) / 4)) {
                repeat_sequence<kSmem, kTier>(g, q.len, pnodes + used);
            } else {
                memo = q.nb == 2 ? q.bytes : nullptr;
                memo_len = q.len;
                ok = add_sequence<kSmem, kOneTile, kTier, kWide>(g, caps, H, q, sc, pnodes + used);
            }
            used += q.len;
            return ok;
        };
        if (round == 0) {   // :174-179
            s.bytes = P.packed + w.draft_off; s.ascii = nullptr; s.len = w.draft_len; s.nb = 4;
            if (!add(s)) return -2;
        } else if (n_cons > 0) {   // :167-172
            s.bytes = nullptr; s.ascii = out; s.len = n_cons; s.nb = 0;
            if (!add(s)) return -2;
        }
        s.ascii = nullptr; s.nb = 2;
#pragma unroll 1
        for (int k = 0; k < n_arms; ++k) {   // :180-207 (container order, engine stays kNW)
            if (a[k].len == 0) continue;
            s.bytes = P.packed + a[k].off; s.len = a[k].len;
#ifdef HYPO_TMA_STAGE
            s.arm_idx = (int64_t)(w.first_arm + k);
            s.next_idx = -1;
            for (int kn = k + 1; kn < n_arms; ++kn)
                if (a[kn].len) { s.next_idx = (int64_t)(w.first_arm + kn); s.next_bytes = P.packed + a[kn].off; s.next_len = a[kn].len; break; }
#endif
            if (!add(s)) return -2;
        }
        if (lane == 0) pstart[ws->n_seq] = used;
        __syncwarp();

        if (!ws->exact) {   // consensus and MSA columns need spoa's exact rank order
            if (!topo_sort<kSmem, kTier>(g, caps)) return -2;
            if (lane == 0) ws->exact = 1;
            __syncwarp();
        }
        if (!bundle_fits(g, ws)) return -2;
        const int nc = heaviest_bundle<kSmem, kTier>(g, true);
        const Graph gv = make_graph<kSmem, kTier>(g);
        // MSA column ids (graph.cpp:371-388) -> reuse n2r (free after the bundle)
        uint16_t* msa = gv.n2r;
        if (lane == 0) {
            int id = 0;
#pragma unroll 1
            for (int i = 0; i < gv.n_nodes; ++i) {
                const int v = gv.r2n[i];
                msa[v] = (uint16_t)id;
                const int cnt = gv.al_cnt[v];
#pragma unroll 1
                for (int k = 0; k < cnt; ++k) msa[gv.r2n[++i]] = (uint16_t)id;
                ++id;
            }
        }
        // support counts (graph.cpp:542-564) -> reuse score (free after the bundle)
        uint32_t* sup = reinterpret_cast<uint32_t*>(gv.score);
        __syncwarp();
#pragma unroll 1
        for (int c = lane; c < nc; c += 32) sup[c] = 0;
        __syncwarp();
#pragma unroll 1
        for (int q = lane; q < gv.n_seq; q += 32) {
            const uint32_t b = pstart[q], e = pstart[q + 1];
            int c = 0;
#pragma unroll 1
            for (uint32_t k = b; k < e; ++k) {
                const int v = pnodes[k];
                const int mv = msa[v];
#pragma unroll 1
                while (c < nc && msa[gv.cons[c]] < mv) ++c;
                if (c >= nc) break;
                if (msa[gv.cons[c]] == mv && (gv.ninfo[v] & 7) == (gv.ninfo[gv.cons[c]] & 7)) atomicAdd(&sup[c], 1u);
            }
        }
        __syncwarp();
        // curate (src/Window.cpp:239-254): ordered compaction
        int kept = 0;
#pragma unroll 1
        for (int c0 = 0; c0 < nc; c0 += 32) {
            const int c = c0 + lane;
            const bool keep = c < nc && sup[c] >= thres;
            const unsigned km = __ballot_sync(kFull, keep);
            if (keep) out[kept + __popc(km & ((1u << lane) - 1u))] = code_to_char(gv.ninfo[gv.cons[c]] & 7);
            kept += __popc(km);
        }
        n_cons = kept;
        __syncwarp();
        __threadfence_block();
    }
    return n_cons;
}

// kMinBlocks = 3: the compact tier (27 warps / SM, <= 72 registers); 2: every other tier.
// kWide: reads whose scores x size leave the 16-bit DP range are filled with 32-bit cells (last tier).
template <bool kSmem, bool kOneTile, bool kLong, int kMinBlocks, int kTier, bool kWide = false>
__global__ void __launch_bounds__(kTeamOf<kTier> == 4 ? 32 * 4 * kTeamsPerCta : 288, kTeamOf<kTier> == 4 ? 1 : kMinBlocks)
poa_kernel(const Params P) {
    constexpr int kTeam = kTeamOf<kTier>;
    const int lane = lane_id();
    const int warp_in_cta = threadIdx.x >> 5;
    const int warps_per_cta = blockDim.x >> 5;
    // (one workspace slot per window in flight: per warp, or per team of warps)
    const int gwarp = blockIdx.x * (warps_per_cta / kTeam) + warp_in_cta / kTeam;
    const Caps caps = tier_caps<kTier>(P.caps);
    GState g;
    uint32_t arena_bytes;
    if constexpr (kTier >= 0) {
        constexpr ArenaLayout L = arena_layout(fixed_caps(kTier));
        arena_bytes = L.total;
    } else {
        g.L = arena_layout(caps);
        arena_bytes = g.L.total;
    }
    (void)arena_bytes;
    g.gbase = kSmem ? nullptr : (P.gws + (size_t)gwarp * P.g_slot);
    g.fail_hist = P.fail_hist;
    g.packed_end = P.packed + P.packed_end;
#ifdef HYPO_TMA_STAGE
    {
        StageCtl* sctl; uint32_t scap;
        if (stage_buf<kSmem, kTier>(g, &sctl, &scap) && lane == 0) stage_init(sctl);
        __syncwarp();
    }
#endif
    int16_t* H = P.H + (size_t)gwarp * P.h_slot;
    uint16_t* paths = P.paths ? P.paths + (size_t)gwarp * P.p_slot : nullptr;

    if constexpr (kTeam > 1) {
        TeamBox* const box = &g_team_box[warp_in_cta / kTeam];
        if (warp_in_cta % kTeam == 0 && lane == 0) { box->seq = 0; box->cmd = 0; box->err = 0; box->done = 0; }
        __syncthreads();
        if (warp_in_cta % kTeam != 0) {
            team_helper<kSmem, kTier>(g, warp_in_cta % kTeam);
            return;
        }
    }
    const uint32_t n_work = __ldg(P.n_work);   // complete: only earlier launches append to this tier's list
#pragma unroll 1
    for (;;) {
        uint32_t wi = 0;
        if (lane == 0) wi = atomicAdd(P.queue, 1u);
        wi = __shfl_sync(kFull, wi, 0);
        if (wi >= n_work) break;
        const uint32_t widx = P.work[wi];
        const WinDesc w = P.win[widx];
        char* out = P.out + P.out_pos[widx];
        const uint32_t n = w.n_internal + w.n_pre + w.n_suf;
        int res;
        // projected size from a tier that abandoned the window (0 = none)
        uint32_t need = 0;
        bool pass_on = false;
        if constexpr (kProjects<kOneTile, kTier>) {
            if (P.need) need = P.need[widx];
            pass_on = (int)(need & 0xffffu) > caps.ncap + caps.ncap / 8 || (int)(need >> 16) > caps.ecap + caps.ecap / 8;
        }
        if (lane == 0) warp_state<kSmem, kTier>(g)->cells = 0;
        if (w.n_empty > n) {
            res = 0;   // reference src/Window.cpp:47-49
        } else if (kProjects<kOneTile, kTier> && pass_on) {
            give_up(g, kFailForwarded);
            res = -2;
        } else if (kTier == -2 && P.long_only && w.wtype == 0 && n >= 2) {
            give_up(g, kFailForwarded);   // (T2s only takes LONG windows unless it is forced)
            res = -2;
        } else if (n >= 2) {
            if (w.wtype == 0) res = run_short<kSmem, kOneTile, kTier, kWide>(g, P, caps, H, w, out);
            else if (kLong && paths) res = run_long<kSmem, kOneTile, kTier, kWide>(g, P, caps, H, w, out, paths, P.p_slot);
            else { give_up(g, kFailNoLong); res = -2; }
        } else {
            res = -1;
        }
        if (res == -1) {   // draft copy (reference :58-60,150-152,233-235)
            const uint8_t* src = P.packed + w.draft_off;
#pragma unroll 1
            for (int p = lane; p < (int)w.draft_len; p += 32) {
                int v = (src[p >> 1] >> ((p & 1) ? 0 : 4)) & 15;
                out[p] = code_to_char(v > 4 ? 4 : v);
            }
            res = (int)w.draft_len;
        }
        if (lane == 0) {
            if (res == -2) {
                const uint32_t k = atomicAdd(P.next_count, 1u);
                if (P.abandoned) atomicAdd(P.abandoned, 1u);
                P.next_list[k] = widx;
                if constexpr (kProjects<kOneTile, kTier>) {
                    // (run_short / run_long zero it when they start a window; non-zero = abandoned on projection)
                    const uint32_t proj = pass_on ? 0u : warp_state<kSmem, kTier>(g)->need;
                    if (P.need && proj != 0)
                        P.need[widx] = max(proj & 0xffffu, need & 0xffffu) | (max(proj >> 16, need >> 16) << 16);
                }
            } else {
                P.out_len[widx] = (uint32_t)res;
                if (P.cells) atomicAdd(P.cells, warp_state<kSmem, kTier>(g)->cells);
            }
        }
        __syncwarp();
    }
    if constexpr (kTeam > 1) {   // the queue is empty: release the helpers
        TeamBox* const box = &g_team_box[warp_in_cta / kTeam];
        uint32_t seq = 0;
        if (lane == 0) { box->cmd = 2; seq = box->seq + 1; }
        __threadfence_block();
        __syncwarp();
        if (lane == 0) st_volatile_shared(&box->seq, seq);
    }
}

}  // namespace

int team_size(int tier, bool team_variant) { return tier == 5 || (tier > 5 && team_variant) ? 4 : tier == 4 ? 2 : 1; }

// ------------------------------------------------------------------------------------------
// Host-side launcher
// ------------------------------------------------------------------------------------------
// `warps_per_block` counts windows in flight per block: warps, or teams of four warps (T1 always; the
// bound-driven tiers with `team`).
cudaError_t launch_poa(const Params& P, int tier, bool smem_graph, bool wide, bool team, int blocks,
                       int warps_per_block, size_t smem_bytes, cudaStream_t stream) {
    void (*k)(const Params) = nullptr;
    const int tsize = team_size(tier, team);
    if (tsize == 4 ? warps_per_block > kTeamsPerCta : warps_per_block * tsize * 32 > 288) return cudaErrorInvalidConfiguration;
    // one-tile tiers only ever run SHORT windows (the LONG driver is compiled out of them)
    switch (tier) {
        case 0: k = poa_kernel<true, true, false, 3, 0>; break;    // Tc
        case 1: k = poa_kernel<true, true, false, 2, 1>; break;    // T0
        case 2: k = poa_kernel<true, true, false, 2, 2>; break;    // Tw
        case 3: k = poa_kernel<true, false, true, 2, 3>; break;    // T0b
        case 4: k = poa_kernel<true, false, true, 2, 4>; break;    // T1m
        case 5: k = poa_kernel<true, false, true, 2, 5>; break;    // T1
        default:                                                   // bound-driven tiers, DAG in global memory
            if (smem_graph) {   // T2s: shared-memory DAG with run-time capacities, always a team
                if (!team || wide) return cudaErrorInvalidConfiguration;
                k = poa_kernel<true, false, true, 1, -2, false>;
            } else
            if (team) k = wide ? poa_kernel<false, false, true, 1, -2, true> : poa_kernel<false, false, true, 1, -2, false>;
            else k = wide ? poa_kernel<false, false, true, 2, -1, true> : poa_kernel<false, false, true, 2, -1, false>;
    }
    cudaError_t err = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (err != cudaSuccess) return err;
    k<<<blocks, warps_per_block * tsize * 32, smem_bytes, stream>>>(P);
    return cudaGetLastError();
}

}  // namespace hypo_b200
