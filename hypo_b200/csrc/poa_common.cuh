// poa_common.cuh — pieces shared by the per-window kernels (poa_kernel.cu: one warp per window;
// poa_group.cu: several windows per warp in lock-step): the typed view of a window's arena, the
// packed 16-bit DP primitives, and the serial helpers that restate the reference.
// Each translation unit gets its own copies (anonymous namespace, device code only).
#pragma once
#include "poa_kernel.cuh"

namespace hypo_b200 {

namespace {

constexpr unsigned kFull = 0xffffffffu;

// ------------------------------------------------------------------------------------------
// Per-warp state
// ------------------------------------------------------------------------------------------
struct Graph {
    uint8_t* ninfo;       // node -> letter code (bits 0-2: A C G T N J O) | has out-edge (bit 3)
    uint8_t* al_cnt;      // node -> number of aligned nodes
    uint8_t* in_deg;      // node -> in-degree
    uint16_t* in_head;    // node -> first in-edge (insertion order), kNone if none
    uint16_t* al_blk;     // node -> block in al_pool holding its aligned_nodes_ids_, kNone
    uint16_t* n2r;        // node -> rank
    uint16_t* r2n;        // rank -> node
    uint16_t* e_src;
    uint16_t* e_w;        // Edge::total_weight_ (2 per traversal)
    uint16_t* e_next;     // next in-edge of the same destination
    uint16_t* al_pool;
    uint32_t* rowinfo;    // rank -> prows offset | #preds << 16 | letter code << 24 | sink << 27
    uint16_t* prows;      // predecessor DP rows, in-edge order
    uint16_t* fp;         // DP row -> first predecessor row
    uint16_t* fp4;        // DP row -> fp applied four times
    uint16_t* stack;      // toposort scratch: DFS stack
    uint8_t* colseq;      // letter code of DP column j (colseq[0] and the padding columns hold 7)
    uint8_t* seq;         // current sequence, letter codes (= colseq + 1)
    uint16_t* cur;        // per sequence position: aligned node / resolved node
    uint8_t* mark;        // toposort scratch
    uint16_t* anch;       // order_update scratch
    uint16_t* newa;
    int32_t* score;       // epilogue
    uint16_t* pred;
    uint16_t* cons;
    int n_nodes, n_edges, n_al, n_seq;   // snapshot of *ws taken by make_graph
    int als;              // slots per aligned-list block
    struct WarpState* ws; // the live counts (in the arena)
};

// What persists per warp between the phases (kept in local memory; the phases are separate
// functions so that each gets its own register allocation and the code stays small): where the
// arena is, its layout, and the element counts.  Every phase rebuilds its typed view with
// make_graph<kSmem>, which lets the compiler see that shared-memory tiers address __shared__
// (LDS/STS with 32-bit addresses) instead of falling back to generic loads.
// Element counts of the window a warp is building.  They live in the arena (shared memory for the
// shared-memory tiers) so that the phases - separate functions - exchange them without going through
// per-thread local memory.
struct WarpState {
    int n_nodes, n_edges, n_al, n_seq;
    int exact;           // r2n/n2r currently hold spoa's exact DFS order (not just a valid one)
    int n_total;         // sequences this window (round) will add in all
    uint32_t base;       // nodes | edges << 16 after the second sequence (growth is measured from here)
    uint32_t need;       // projected final nodes | edges << 16 when the window was abandoned on projection
    unsigned long long cells;   // DP cells filled for this window so far: sum of (rows + 1) x (columns + 1)
    int clean;           // the sequence added last only re-walked existing nodes and edges (and `cur` still holds its
                         // node path): an identical sequence right after it aligns identically - see repeat_sequence
};
static_assert(sizeof(WarpState) <= 48, "the arena reserves 48 bytes for it");

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

struct Scores {
    int m, n, g;
};

// Query profile (reference sisd_alignment_engine.cpp:101-108), g-normalised and computed on the
// fly: pf[j] = (letter(row) == colseq[j] ? m : n) - g.  A lane keeps the letter codes of its four
// columns in one register (one byte each, 0..7).  XOR with the row's code replicated to four bytes
// leaves a zero byte exactly where they match; 0x80 - 16*t moves that into bit 7 of each byte
// (t <= 7, so no borrow crosses a byte); PRMT's sign-replicate mode widens bit 7 to a 16-bit mask
// per column, which selects between the two packed constants.
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
__device__ __forceinline__ void profile_regs(uint32_t let4, uint32_t code, uint32_t mm2, uint32_t nn2,
                                             uint32_t (&pf)[kNR]) {
    const uint32_t t = let4 ^ (code * 0x01010101u);
    const uint32_t z = 0x80808080u - 16u * t;
    const uint32_t m01 = prmt(z, 0u, 0x9988u);
    const uint32_t m23 = prmt(z, 0u, 0xBBAAu);
    pf[0] = (m01 & mm2) | (~m01 & nn2);
    pf[1] = (m23 & mm2) | (~m23 & nn2);
}

__device__ __forceinline__ uint32_t opaque(uint32_t v) {
    asm volatile("" : "+r"(v));
    return v;
}
template <typename T>
__device__ __forceinline__ T* opaque_ptr(T* p) {
    asm volatile("" : "+l"(p));
    return p;
}
// Explicit address-space loads for the DP's inner loop: for the shared-memory tiers a 32-bit shared
// address that ptxas cannot re-derive (it otherwise rebuilds the shared window base in every row).
template <bool kSmem>
struct Mem;
template <>
struct Mem<true> {
    typedef uint32_t addr_t;
    static __device__ __forceinline__ addr_t addr(const void* p) {
        return opaque((uint32_t)__cvta_generic_to_shared(p));
    }
    static __device__ __forceinline__ uint32_t ld32(addr_t a) {
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
        return v;
    }
    static __device__ __forceinline__ uint32_t ld16(addr_t a) {
        uint32_t v;
        asm volatile("{ .reg .u16 t; ld.shared.u16 t, [%1]; cvt.u32.u16 %0, t; }" : "=r"(v) : "r"(a));
        return v;
    }
};
template <>
struct Mem<false> {
    typedef const uint8_t* addr_t;
    static __device__ __forceinline__ addr_t addr(const void* p) { return opaque_ptr((const uint8_t*)p); }
    static __device__ __forceinline__ uint32_t ld32(addr_t a) { return *reinterpret_cast<const uint32_t*>(a); }
    static __device__ __forceinline__ uint32_t ld16(addr_t a) { return *reinterpret_cast<const uint16_t*>(a); }
};
// DP rows in global memory, 8 bytes per lane (explicit .global: the pointers are opaque to ptxas)
__device__ __forceinline__ void stg64(void* p, uint32_t a, uint32_t b) {
    asm volatile("st.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ uint2 ldg64(const void* p) {
    uint2 v;
    asm volatile("ld.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t bcast16(int v) { return (uint32_t)(v & 0xffff) * 0x10001u; }
__device__ __forceinline__ int hi16(uint32_t v) { return (int)v >> 16; }

// x = max(x, diag + prof, vert + g) for one predecessor row.
__device__ __forceinline__ void relax(uint32_t (&x)[kNR], const uint32_t (&p)[kNR], uint32_t left,
                                      const uint32_t (&pf)[kNR], uint32_t g2) {
    uint32_t prevreg = left;   // pred[c0-1] in the HIGH half
#pragma unroll
    for (int r = 0; r < kNR; ++r) {
        uint32_t d = __byte_perm(prevreg, p[r], 0x5432);   // (pred[j-1] for lo, for hi)
        x[r] = __viaddmax_s16x2(d, pf[r], x[r]);
        x[r] = __viaddmax_s16x2(p[r], g2, x[r]);
        prevreg = p[r];
    }
}

// Horizontal pass: H^[i][j] = max(H^[i][j], H^[i][j-1]) == inclusive prefix max over columns.
// In-lane over the 2*kNR columns; returns the lane total (its last column) in both halves.
// `neg2` is kNegInf2 held in a register.
__device__ __forceinline__ uint32_t scan_inlane(uint32_t (&x)[kNR], uint32_t neg2) {
    uint32_t runb = neg2;   // running max broadcast to both halves
#pragma unroll
    for (int r = 0; r < kNR; ++r) {
        uint32_t t = __byte_perm(x[r], neg2, 0x1054);          // (lo: -inf, hi: x.lo)
        x[r] = __vimax3_s16x2(x[r], t, runb);
        runb = __byte_perm(x[r], 0, 0x3232);                    // (x.hi, x.hi)
    }
    return runb;
}

struct EndCell {
    int row;   // 0 if no candidate (reference clamps max_i=-1 to 0)
    int col;
    int score; // H^ of that cell (0 for the clamped cell (0, 0))
    bool tie;  // two or more candidate rows share the best score (the rank order decides)
};

// rowinfo bits 28-30: every predecessor row lies 1, 2 or 3 rows back (bit d-1 set for distance d;
// rank 0 without predecessor counts as distance 1: the virtual row 0).  0 = take the general path.
constexpr int kRowNearShift = 28;
// How many previous rows the fill keeps in registers (2 or 3).  With two, a pair of rows per trip needs no
// register rotation at all (row i overwrites the registers of row i-2 after using them).
#ifndef HYPO_ROW_HIST
#define HYPO_ROW_HIST 3
#endif
constexpr int kRowHist = HYPO_ROW_HIST;

struct AlnSpan {
    int first, last;
};

__device__ __forceinline__ void init_node(const Graph& g, int id, int code) {
    g.ninfo[id] = (uint8_t)code;
    g.al_cnt[id] = 0;
    g.in_deg[id] = 0;
    g.in_head[id] = kNone;
    g.al_blk[id] = kNone;
}

// Serial DFS from one root (lane 0), verbatim the reference's inner loop.
__device__ __forceinline__ bool dfs_from(const Graph& g, const Caps& caps, int root, int& nr) {
    int sp = 0;
    g.stack[sp++] = (uint16_t)root;
#pragma unroll 1
    while (sp != 0) {
        const int v = g.stack[sp - 1];
        bool valid = true;
        const int mv = g.mark[v];
        if ((mv & 3) != 2) {
#pragma unroll 1
            for (int e = g.in_head[v]; e != kNone; e = g.e_next[e]) {
                const int s = g.e_src[e];
                if ((g.mark[s] & 3) != 2) {
                    if (sp >= caps.scap) return false;
                    g.stack[sp++] = (uint16_t)s;
                    valid = false;
                }
            }
            const bool check = (mv & 4) == 0;
            const int cnt = g.al_cnt[v];
            const int blk = g.al_blk[v];
            if (check) {
#pragma unroll 1
                for (int k = 0; k < cnt; ++k) {
                    const int a = g.al_pool[blk * g.als + k];
                    const int ma = g.mark[a];
                    if ((ma & 3) != 2) {
                        if (sp >= caps.scap) return false;
                        g.stack[sp++] = (uint16_t)a;
                        g.mark[a] = (uint8_t)(ma | 4);
                        valid = false;
                    }
                }
            }
            if (valid) {
                g.mark[v] = (uint8_t)((mv & 4) | 2);
                if (check) {
                    g.r2n[nr++] = (uint16_t)v;
#pragma unroll 1
                    for (int k = 0; k < cnt; ++k) g.r2n[nr++] = g.al_pool[blk * g.als + k];
                }
            } else {
                g.mark[v] = (uint8_t)((mv & 4) | 1);
            }
        }
        if (valid) --sp;
    }
    return true;
}

__device__ __forceinline__ int branch_completion(const Graph& g, int rank) {
    const int n = g.n_nodes;
    const int node = g.r2n[rank];
    // for every successor d of node: invalidate the other sources of d's in-edges
#pragma unroll 1
    for (int d = 0; d < n; ++d) {
        bool succ = false;
#pragma unroll 1
        for (int e = g.in_head[d]; e != kNone; e = g.e_next[e]) succ |= g.e_src[e] == node;
        if (!succ) continue;
#pragma unroll 1
        for (int o = g.in_head[d]; o != kNone; o = g.e_next[o])
            if (g.e_src[o] != node) g.score[g.e_src[o]] = -1;
    }
    int max_score = 0, max_id = 0;
#pragma unroll 1
    for (int r = rank + 1; r < n; ++r) {
        const int v = g.r2n[r];
        int sv = -1, pv = kNone;
#pragma unroll 1
        for (int e = g.in_head[v]; e != kNone; e = g.e_next[e]) {
            const int s = g.e_src[e];
            const int ss = g.score[s];
            if (ss == -1) continue;
            const int w = g.e_w[e];
            if (sv < w || (sv == w && g.score[pv] <= ss)) { sv = w; pv = s; }
        }
        if (pv != kNone) sv += g.score[pv];
        g.score[v] = sv;
        g.pred[v] = (uint16_t)pv;
        if (max_score < sv) { max_score = sv; max_id = v; }
    }
    return max_id;
}

__device__ __forceinline__ char code_to_char(int c) {
    return "ACGTNJO"[c];
}

// Typed view of the arena at `base` (shared memory or a global workspace slot).
__device__ __forceinline__ Graph bind_graph_at(uint8_t* base, const ArenaLayout& L) {
    Graph g;
    g.ws = (WarpState*)(base + L.state);
    g.ninfo = base + L.ninfo;
    g.al_cnt = base + L.al_cnt;
    g.in_deg = base + L.in_deg;
    g.in_head = (uint16_t*)(base + L.in_head);
    g.al_blk = (uint16_t*)(base + L.al_blk);
    g.n2r = (uint16_t*)(base + L.n2r);
    g.r2n = (uint16_t*)(base + L.r2n);
    g.e_src = (uint16_t*)(base + L.e_src);
    g.e_w = (uint16_t*)(base + L.e_w);
    g.e_next = (uint16_t*)(base + L.e_next);
    g.al_pool = (uint16_t*)(base + L.al_pool);
    g.rowinfo = (uint32_t*)(base + L.rowinfo);
    g.prows = (uint16_t*)(base + L.prows);
    g.fp = (uint16_t*)(base + L.fp);
    g.fp4 = (uint16_t*)(base + L.fp4);
    g.stack = (uint16_t*)(base + L.stack);
    g.colseq = base + L.colseq;
    g.seq = base + L.colseq + 1;
    g.cur = (uint16_t*)(base + L.cur);
    g.mark = base + L.mark;
    g.anch = (uint16_t*)(base + L.anch);
    g.newa = (uint16_t*)(base + L.newa);
    g.score = (int32_t*)(base + L.score);
    g.pred = (uint16_t*)(base + L.pred);
    g.cons = (uint16_t*)(base + L.cons);
    g.n_nodes = g.ws->n_nodes; g.n_edges = g.ws->n_edges; g.n_al = g.ws->n_al; g.n_seq = g.ws->n_seq;
    g.als = L.alslots;
    return g;
}

}  // namespace

}  // namespace hypo_b200
