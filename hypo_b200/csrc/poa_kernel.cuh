// poa_kernel.cuh — device side of the B200-native POA-consensus path.
//
// One warp owns one weak-region window at a time and runs the whole of
// hypo::Window::generate_consensus for it (reference src/Window.cpp:44-254):
// for every read segment ("arm") in the reference's order it
//   1. decodes the 2-bit PackedSeq bytes (the query profile is computed in registers, per row),
//   2. fills the sequence-to-graph DP (spoa SISD linear-gap semantics,
//      reference external/spoa/src/sisd_alignment_engine.cpp:263-342),
//   3. walks the equality-driven traceback (:344-437),
//   4. fuses the alignment into the DAG (reference external/spoa/src/graph.cpp:154-271),
//   5. keeps a valid clique-contiguous topological order up to date; spoa's exact DFS order
//      (:293-353) is re-derived only where the rank order can change the result,
// and finally extracts the heaviest-bundle consensus (:610-705), for LONG windows
// with the per-column support counts + curation (:533-568, src/Window.cpp:239-254).
//
// Data placement:
//   * the growing DAG (SoA, 16-bit indices) and the element counts live in shared memory (tiers S,
//     layouts are compile-time constants) or in a global workspace (tiers L, for windows that
//     overflow the shared-memory capacities);
//   * the DP matrix lives in global memory, filled one 128-column tile at a time; every warp keeps
//     the last three rows of the tile in registers, so rows whose predecessors lie within three
//     ranks never touch memory before their store; rows are written once, coalesced, 8 bytes per
//     lane, and dropped from L2 (discard.global.L2) once the read's traceback is done;
//   * cells are 16-bit, two per register, updated with the sm_100 DPX instructions
//     (VIADDMNMX.S16x2 / VIMNMX3.S16x2); the horizontal gap pass is a warp-shuffle
//     prefix-max (the matrix is stored g-normalised, H^[i][j] = H[i][j] - j*g, which
//     turns the max-plus scan into a plain prefix max and leaves every equality test of
//     the traceback unchanged).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hypo_b200 {

// ---- ABI structs (must match include/hypo_b200.h) --------------------------------------
struct WinDesc {
    uint64_t draft_off, first_arm;
    uint32_t draft_len, n_internal, n_pre, n_suf, n_empty, wtype;
};
struct ArmDesc {
    uint64_t off;
    uint32_t len, reserved;
};
static_assert(sizeof(WinDesc) == 40 && sizeof(ArmDesc) == 16, "ABI layout");

// ---- constants ---------------------------------------------------------------------------
constexpr int kNR = 2;                       // packed registers per lane per tile
constexpr int kTileCols = 32 * 2 * kNR;      // 128 DP columns per tile
constexpr uint16_t kNone = 0xFFFF;
constexpr int kNegInf = -30000;              // "minus infinity" that survives one int16 add
constexpr uint32_t kNegInf2 = 0x8AD08AD0u;   // (kNegInf, kNegInf) packed
constexpr int kAlSlotsMax = 6;               // clique <= 7 letters (J O A C G T N) => <= 6 peers
constexpr int kNumCodes = 7;                 // A C G T N J O
constexpr int kCodeJ = 5, kCodeO = 6;
constexpr int kMaxH16 = 29000;               // int16 safety bound for |H^|

enum AlignType { kNW = 0, kLOV = 1, kROV = 2 };

// Why a window was abandoned in a tier (diagnostic histogram, hypo_gpu_last_fail_hist).
enum FailReason {
    kFailLen = 1,        // a sequence is longer than the tier's columns
    kFailRange = 2,      // scores x size could leave the 16-bit DP range
    kFailNodes = 3,      // node capacity
    kFailAligned = 4,    // aligned-list blocks
    kFailClique = 5,     // a clique with more than 7 members (cannot happen for 7 letters)
    kFailEdges = 6,      // edge capacity or in-degree > 254
    kFailStack = 7,      // DFS stack of the exact topological sort
    kFailPaths = 8,      // LONG: per-sequence node paths
    kFailNoLong = 9,     // LONG window in a tier compiled without the LONG driver
    kFailProjected = 10, // node / edge growth per read extrapolates beyond the tier: handed on early
    kFailForwarded = 11, // an earlier tier's projection exceeds this tier as well: passed on without work
    kFailTeam = 12,      // a team fill ran into its time limit (never observed; the window is re-run in the next tier)
    kNumFailReasons = 16
};

// Capacities of one tier (one kernel launch).
struct Caps {
    int ncap;      // nodes
    int ecap;      // edges
    int acap;      // aligned-list blocks (one per node that is member of a clique)
    int scap;      // DFS stack entries
    int lcap;      // sequence length incl. markers
    int tiles;     // ceil((lcap+1)/kTileCols)
    int alslots;   // slots per aligned-list block: a clique of more than alslots+1 nodes overflows the
                   // tier (6 always suffices; the small tiers trade slots for blocks: cliques beyond
                   // A/C/G/T need N or a marker letter aligned to bases)
    int tilecols = kTileCols;   // DP columns per tile: 4 per lane of the group that owns the window
};

struct Params {
    const WinDesc* win;
    const ArmDesc* arms;
    const uint8_t* packed;
    const uint32_t* work;        // window ids to process in this launch (this tier's list)
    const uint32_t* n_work;      // device: entries in `work` (routed windows + those earlier launches handed on)
    uint32_t* queue;             // [0] = next work index (atomic)
    char* out;                   // consensus bytes
    const uint64_t* out_pos;     // where window w writes
    uint32_t* out_len;           // consensus length of window w
    uint32_t* next_list;         // windows that exceed this tier are appended to the successor tier's list ...
    uint32_t* next_count;        // ... whose (atomic) length this is; the tier lists never need the host in between
    uint32_t* abandoned;         // how many windows this tier handed on (feedback for the tier probes; may be null)
    uint32_t* fail_hist;         // [kNumFailReasons] why windows left a tier (diagnostics; may be null)
    unsigned long long* cells;   // sum of the DP cells (rows x columns of every fill) of the windows this
                                 // launch completed: the work counter behind the GCUPS figures (may be null)
    uint32_t* need;              // per window: projected nodes | edges << 16 left by a tier that abandoned it
                                 // on projection (0 = none; may be null)
    int16_t* H;                  // DP workspace, one slot per warp
    uint64_t h_slot;             // elements per slot
    uint8_t* gws;                // tiers L: graph workspace, one slot per warp
    uint64_t g_slot;             // bytes per slot
    uint16_t* paths;             // LONG: per-warp node paths, one slot per warp
    uint64_t p_slot;             // elements per slot
    Caps caps;
    int sr_m, sr_n, sr_g, lr_m, lr_n, lr_g;
    uint64_t packed_end;         // bytes of `packed` that may be read (bulk copies never reach beyond it)
    uint32_t long_only;          // non-zero: SHORT windows are passed on to the successor tier untried (T2s)
};

__host__ __device__ constexpr uint32_t align16(uint32_t x) { return (x + 15u) & ~15u; }

// Per-warp graph arena.  Node arrays are indexed by node id, row arrays by rank (+1 = DP row).
struct ArenaLayout {
    uint32_t state;     // WarpState: element counts of the window being built (40 bytes)
    // nodes
    uint32_t ninfo;     // u8  letter code (bits 0-2) | has-out-edge (bit 3)
    uint32_t al_cnt;    // u8  number of aligned nodes
    uint32_t in_deg;    // u8  in-degree (saturating; 255 => tier overflow)
    uint32_t in_head;   // u16 first in-edge (insertion order)
    uint32_t al_blk;    // u16 block of al_pool holding aligned_nodes_ids_
    uint32_t n2r;       // u16 node -> rank
    uint32_t r2n;       // u16 rank -> node
    // edges
    uint32_t e_src, e_w, e_next;   // u16 each
    uint32_t al_pool;   // u16 [acap][alslots]
    // rows (rebuilt by build_rows after every change of the order; dead while the order is being
    // changed and once the last read has been aligned, so the sort / order-update / epilogue
    // scratch aliases this region)
    uint32_t rowinfo;   // u32 [ncap+4] by rank: prows offset (bits 0-15) | #preds (16-23) | letter code (24-26) | sink (27) | fast (28)
    uint32_t prows;     // u16 [ecap] predecessor DP rows in in-edge order
    uint32_t fp;        // u16 [ncap+1] first predecessor row of each DP row (0 = virtual row 0)
    uint32_t fp4;       // u16 [ncap+1] fp applied four times (traceback jump pointers)
    uint32_t mark;      // u8  [ncap]   sort marks            (aliases rows)
    uint32_t stack;     // u16 [scap]   DFS stack             (aliases rows)
    uint32_t anch;      // u16 [lcap+1] order_update anchors  (aliases rows)
    uint32_t newa;      // u16 [lcap+1] order_update          (aliases rows)
    uint32_t score;     // i32 [ncap]   epilogue              (aliases rows)
    uint32_t pred;      // u16 [ncap]
    uint32_t cons;      // u16 [ncap]
    // per-read scratch
    uint32_t colseq;    // u8  [tiles*128] letter code of DP column j (= seq[j-1]); 7 = matches nothing
    uint32_t cur;       // u16 [lcap+1] per position: aligned / resolved node
    uint32_t stage;     // HYPO_TMA_STAGE: landing buffer of the bulk copy of the NEXT read's packed bytes
                        // (16-byte aligned window around them), then mbarrier (8), pending arm (4), phase (4)
    uint32_t stage_cap; // bytes of the landing buffer
    uint32_t total;
    int alslots;
};

struct LayoutCursor {
    uint32_t o;
    __host__ __device__ constexpr uint32_t take(uint32_t bytes) {
        const uint32_t r = o;
        o = (o + bytes + 15u) & ~15u;
        return r;
    }
};

__host__ __device__ constexpr ArenaLayout arena_layout(const Caps& c) {
    ArenaLayout L{};
    LayoutCursor k{0};
    L.state = k.take(48);
    L.ninfo = k.take(c.ncap);
    L.al_cnt = k.take(c.ncap);
    L.in_deg = k.take(c.ncap);
    L.in_head = k.take(2u * c.ncap);
    L.al_blk = k.take(2u * c.ncap);
    L.n2r = k.take(2u * c.ncap);
    L.r2n = k.take(2u * c.ncap);
    L.e_src = k.take(2u * c.ecap);
    L.e_w = k.take(2u * c.ecap);
    L.e_next = k.take(2u * c.ecap);
    L.al_pool = k.take(2u * (uint32_t)c.alslots * c.acap);
    L.alslots = c.alslots;
    const uint32_t rows0 = k.o;
    L.rowinfo = k.take(4u * (c.ncap + 4));
    L.prows = k.take(2u * c.ecap);
    L.fp = k.take(2u * (c.ncap + 1));
    L.fp4 = k.take(2u * (c.ncap + 1));
    uint32_t end = k.o;
    // topological-sort scratch over the (dead) rows
    k.o = rows0;
    L.mark = k.take(c.ncap);
    L.stack = k.take(2u * c.scap);
    if (k.o > end) end = k.o;
    // order-update scratch
    k.o = rows0;
    L.anch = k.take(2u * (c.lcap + 1));
    L.newa = k.take(2u * (c.lcap + 1));
    if (k.o > end) end = k.o;
    // epilogue scratch
    k.o = rows0;
    L.score = k.take(4u * c.ncap);
    L.pred = k.take(2u * c.ncap);
    L.cons = k.take(2u * c.ncap);
    if (k.o > end) end = k.o;
    k.o = end;
    L.colseq = k.take((uint32_t)c.tiles * (uint32_t)c.tilecols);
    L.cur = k.take(2u * (c.lcap + 1));
#ifdef HYPO_TMA_STAGE
    L.stage_cap = ((uint32_t)(c.lcap + 3) / 4 + 15u + 15u) & ~15u;   // ceil(lcap / 4) bytes at any alignment
    L.stage = k.take(L.stage_cap + 16u);
#else
    L.stage_cap = 0;
    L.stage = k.o;
#endif
    L.total = (k.o + 15u) & ~15u;
    return L;
}

// Capacities of the shared-memory tiers are compile-time constants (the kernels fold every arena
// offset into an immediate); tiers >= kNumFixedTiers take theirs from Params at run time.
// Growth projection (add_sequence): from this many sequences in the graph on.
constexpr int kProjectFrom = 6;
constexpr int kNumFixedTiers = 6;
// Group tiers (poa_group.cu): several windows per warp, a group of 8 / 16 lanes each, for the small
// windows the pipeline mostly produces (median draft length 9 bp).  They sit behind the
// bound-driven tiers in the table so that the tier numbers of round 1 keep their meaning.
constexpr int kTierQuad = 8;   // Tq: 8 lanes per window, <= 31 symbols
constexpr int kTierHalf = 9;   // Th: 16 lanes per window, <= 63 symbols
// T2s: the DAG of a large window in SHARED memory, capacities from an ESTIMATE of its size (not the worst-case
// bound, which no shared memory holds), teams of four warps; runs between T1 and the bound-driven tiers.
constexpr int kTierBig = 10;
__host__ __device__ constexpr bool is_group_tier(int tier) { return tier == kTierQuad || tier == kTierHalf; }
__host__ __device__ constexpr bool is_fixed_tier(int tier) { return (tier >= 0 && tier < kNumFixedTiers) || is_group_tier(tier); }
__host__ __device__ constexpr int group_lanes(int tier) { return tier == kTierQuad ? 8 : tier == kTierHalf ? 16 : 32; }
__host__ __device__ constexpr Caps fixed_caps(int tier) {
    // (scap, the DFS stack of the exact sort, aliases the row records and costs no extra memory)
    //                 ncap  ecap  acap  scap  lcap  tiles alslots
    return tier == 0 ? Caps{212, 328, 212, 640, 127, 1, 3}      // Tc : compact, 27 warps / SM
         : tier == 1 ? Caps{320, 576, 304, 1024, 127, 1, 4}     // T0 : one tile
         : tier == 2 ? Caps{512, 1024, 384, 2048, 127, 1, 6}    // Tw : one tile, many reads per window
         : tier == 3 ? Caps{384, 768, 384, 1536, 255, 2, 6}     // T0b: two tiles
         : tier == 4 ? Caps{640, 1152, 512, 2048, 511, 4, 4}    // T1m: four tiles (500-bp windows), 8 warps / SM
         : tier == kTierQuad ? Caps{56, 96, 40, 224, 31, 1, 3, 32}    // Tq : four windows per warp
         : tier == kTierHalf ? Caps{120, 192, 80, 448, 63, 1, 3, 64}   // Th : two windows per warp
                     : Caps{1024, 1920, 1024, 4096, 1023, 8, 4}; // T1 : eight tiles, medium DAG
}

}  // namespace hypo_b200
