// api.cu — the C ABI of include/hypo_b200.h on top of the sm_100a kernels.
//
// Tiering: every window first runs in the fastest tier whose static limits
// it satisfies; a window that overflows a capacity at run time is abandoned, appended to the
// tier's overflow list on the device and re-run from scratch in the next tier.  The last tier
// is sized from exact upper bounds, so nothing ever falls back to the CPU.
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/hypo_b200.h"
#include "poa_kernel.cuh"

namespace hypo_b200 {
cudaError_t launch_poa(const Params& P, int tier, bool smem_graph, bool one_tile, int blocks,
                       int warps_per_block, size_t smem_bytes, cudaStream_t stream);
}

using namespace hypo_b200;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CUDA_TRY(x)                                                                        \
    do {                                                                                   \
        cudaError_t e_ = (x);                                                              \
        if (e_ != cudaSuccess)                                                             \
            return fail(HYPO_E_CUDA, "%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_),  \
                        __FILE__, __LINE__);                                               \
    } while (0)

// grow-only device buffer
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

// Per-window static facts, computed on the device by classify_kernel.
struct WinStat {
    uint32_t max_len;    // longest input sequence incl. markers (routing)
    uint32_t bound_len;  // upper bound of any sequence incl. the LONG round-2 backbone
    uint32_t sum_len;    // sum of sequence lengths incl. markers (node upper bound)
    uint32_t n_seq;      // sequences incl. draft/backbone
    uint32_t sum_raw;    // LONG: draft + arms, for the path slot
};

struct TierMax {
    uint32_t max_len, bound_len, sum_len, n_seq, sum_raw, any_long, count;
};

struct Ctx {
    bool init = false;
    int device = 0;
    int sms = 0;
    int smem_optin = 0;
    int8_t scores[6];
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;          // host-buffer entry point: H2D of the batch's tail
    cudaEvent_t ev_head = nullptr, ev_tail = nullptr;
    uint64_t launches = 0;
    DevBuf win, arms, packed, out_scratch, out_pos, out_len, out_off, out_compact;
    DevBuf stats, lists, ctrl, H, gws, paths, cub_tmp;
    void* pinned_ctrl = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float poa_ms = 0.f;          // device time of the POA kernels of the last batch call
    uint32_t poa_launches = 0;
    uint32_t tier_windows[8] = {0};
    uint32_t fail_hist[kNumFailReasons] = {0};   // why windows left a tier in the last batch call
} g;

std::mutex g_mu;

// ---------------------------------------------------------------------------------------
// Small helper kernels
// ---------------------------------------------------------------------------------------

// One thread per window: static facts + per-window output bound.
__global__ void classify_kernel(const WinDesc* __restrict__ win, const ArmDesc* __restrict__ arms,
                                uint64_t n_win, uint64_t n_arms, uint64_t packed_bytes,
                                WinStat* __restrict__ st, uint64_t* __restrict__ bound,
                                uint32_t* __restrict__ bad) {
    uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (w >= n_win) return;
    const WinDesc d = win[w];
    const uint64_t n = (uint64_t)d.n_internal + d.n_pre + d.n_suf;
    bool ok = d.first_arm + n <= n_arms && d.wtype <= 1 &&
              d.draft_off + (d.draft_len + 1) / 2 <= packed_bytes;
    uint32_t max_len = 0, bound_len = 0, sum_len = 0, n_seq = 0, sum_raw = d.draft_len;
    const uint32_t mark = d.wtype == 0 ? 2 : 0;
    if (ok) {
        for (uint64_t k = 0; k < n; ++k) {
            const ArmDesc a = arms[d.first_arm + k];
            if (a.off + (a.len + 3) / 4 > packed_bytes || a.reserved != 0) { ok = false; break; }
            if (a.len == 0) continue;
            const uint32_t l = a.len + mark;   // upper bound (prefix/suffix arms carry 1 marker)
            max_len = max(max_len, l);
            sum_len += l;
            sum_raw += a.len;
            ++n_seq;
        }
    }
    // backbone / seq 0: draft (SHORT without internal arms, LONG round 1) or the previous
    // consensus (LONG round 2, never longer than the node count of round 1)
    const uint32_t dl = d.draft_len + mark;
    if (d.wtype == 1) {
        // round-2 backbone = curated round-1 consensus <= nodes of round 1 <= sum_len + draft
        const uint32_t b = sum_len + d.draft_len;
        max_len = max(max_len, d.draft_len);
        bound_len = b;
        sum_len += b;
        sum_raw += b;   // generous: path slot must hold the round-2 backbone too
        ++n_seq;
    } else if (d.n_internal == 0) {
        max_len = max(max_len, dl);
        sum_len += dl;
        ++n_seq;
    }
    if (!ok) atomicAdd(bad, 1u);
    WinStat s;
    s.max_len = max_len; s.bound_len = max(bound_len, max_len); s.sum_len = sum_len; s.n_seq = n_seq; s.sum_raw = sum_raw;
    st[w] = s;
    uint64_t b = (uint64_t)sum_len + 2;
    if (b < d.draft_len) b = d.draft_len;
    bound[w] = ok ? b : 0;
}

// Route windows to the first tier whose static limits they satisfy.
// lists: [tier][n_win] window ids; counts[tier]; tmax[tier] running maxima.
__global__ void route_kernel(const WinDesc* __restrict__ win, const WinStat* __restrict__ st,
                             uint64_t n_win, int n_tiers, const uint32_t* __restrict__ tier_lcap,
                             const uint32_t* __restrict__ tier_long_ok, const uint32_t* __restrict__ tier_est_cap,
                             uint32_t* __restrict__ lists, TierMax* __restrict__ tmax) {
    uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (w >= n_win) return;
    const WinStat s = st[w];
    const bool is_long = win[w].wtype == 1;
    int t = 0;
    // (sum_len of a LONG window counts the round-2 backbone bound as well: about twice the bases)
    const uint32_t est = s.max_len + (uint32_t)(((uint64_t)s.sum_len * (is_long ? 8u : 15u)) / 1000u);
    while (t < n_tiers - 1 &&
           (s.max_len > tier_lcap[t] || (is_long ? !(tier_long_ok[t] & 1u) : (tier_long_ok[t] & 2u) != 0) ||
            est > tier_est_cap[t]))
        ++t;
    const uint32_t k = atomicAdd(&tmax[t].count, 1u);
    lists[(uint64_t)t * n_win + k] = (uint32_t)w;
}

// Running maxima over an explicit work list (used to size the bound-driven tiers).
__global__ void listmax_kernel(const WinDesc* __restrict__ win, const WinStat* __restrict__ st,
                               const uint32_t* __restrict__ list, uint32_t n, TierMax* __restrict__ tm) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t w = list[i];
    const WinStat s = st[w];
    atomicMax(&tm->max_len, s.max_len);
    atomicMax(&tm->bound_len, s.bound_len);
    atomicMax(&tm->sum_len, s.sum_len);
    atomicMax(&tm->n_seq, s.n_seq);
    atomicMax(&tm->sum_raw, s.sum_raw);
    if (win[w].wtype == 1) atomicMax(&tm->any_long, 1u);
}

// One warp per window: gather scratch consensus into the compact, window-ordered output.
__global__ void gather_kernel(const char* __restrict__ scratch, const uint64_t* __restrict__ pos,
                              const uint32_t* __restrict__ len, const uint64_t* __restrict__ off,
                              char* __restrict__ dst, uint64_t n_win) {
    const uint64_t w = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    if (w >= n_win) return;
    const int lane = threadIdx.x & 31;
    const char* s = scratch + pos[w];
    char* d = dst + off[w];
    for (uint32_t i = lane; i < len[w]; i += 32) d[i] = s[i];
}

__global__ void widen_kernel(const uint32_t* __restrict__ len, uint64_t* __restrict__ len64, uint64_t n) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n) len64[i] = len[i];
}

// ---------------------------------------------------------------------------------------
// Tier table
// ---------------------------------------------------------------------------------------
struct Tier {
    bool smem_graph, one_tile, long_ok, from_bounds;
    int ncap, ecap, acap, scap, lcap;
    int warps_per_block, blocks_per_sm;
    uint32_t est_cap;   // static routing: windows whose estimated node count exceeds this start in a later tier
    int next;           // tier that re-runs the windows overflowing this one (skips tiers that cannot help)
};

// Tc : SHORT windows whose sequences fit one 128-column tile and whose DAG stays small (the
//      30 x 120 headline shape at ~1 % read error: 187 nodes on average, 209 at most in 3000
//      windows): DAG in 8.3 KB of shared memory, 27 warps/SM (3 CTAs x 9 warps, <= 72 registers),
//      the last three DP rows carried in registers.
// T0 : same, larger DAG capacities, 18 warps/SM.  Windows that overflow Tc at run time land here.
// Tw : one tile, DAG sized for windows with many reads (100-200 x 120 bp), 10 warps/SM.
// Static routing skips tiers a window is unlikely to fit: est = longest sequence + 1.5 % of all bases
// (every base of a ~1 %-error read opens a new node with about that probability); a wrong guess only
// costs the run-time overflow path.
// T0b: windows up to 255 columns (two tiles; SHORT and LONG), 12 warps/SM.
// T1m: windows up to 511 columns (four tiles: the 500-bp windows of CCS reads and LONG windows) with
//      up to 640 nodes, 8 warps/SM.
// T1 : anything up to 1023 columns (LONG windows included) with a medium DAG in shared memory, 5 warps/SM.
// T2/T3: DAG in global memory, capacities from the windows' exact upper bounds (T2 capped); 16 and 8
//      warps/SM - nothing but occupancy hides the latency of the DAG accesses (16 warps/SM run the
//      30 x 500 bp, 5 %-error windows 2.8x faster than 4), the DP workspace permitting (the grid
//      shrinks when the slots would exceed 24 GB).
// The capacities of the first kNumFixedTiers rows are compile-time constants of the kernels
// (poa_kernel.cuh: fixed_caps); they are repeated here as documentation and checked at start-up.
const Tier kTiers[] = {
    {true, true, false, false, 212, 328, 212, 640, 127, 9, 3, 201, 1},
    {true, true, false, false, 320, 576, 304, 1024, 127, 8, 2, 300, 2},
    {true, true, false, false, 512, 1024, 384, 2048, 127, 5, 2, 486, 6},
    {true, false, true, false, 384, 768, 384, 1536, 255, 6, 2, 364, 4},
    {true, false, true, false, 640, 1152, 512, 2048, 511, 4, 2, 608, 5},
    {true, false, true, false, 1024, 1920, 1024, 4096, 1023, 5, 1, 972, 6},
    {false, false, true, true, 8192, 16384, 2048, 8192, 4095, 4, 4, 0xffffffffu, 7},
    {false, false, true, true, 65534, 65534, 65534, 65534, 0x7ffffff0, 2, 4, 0xffffffffu, 8},
};
const int kNumTiers = sizeof(kTiers) / sizeof(kTiers[0]);
// Static routing sends only LONG windows to T1: at 5 warps/SM it is slower than T2 at 16 for SHORT windows
// (30 x 500 bp: 19 vs 36 Mbp/s), while a LONG window's bound-driven capacities in T2 (the round-2 backbone
// is bounded by the node count) blow up the DP workspace and with it shrink the grid (5.7 vs 2.0 Mbp/s).
// SHORT windows still reach T1 as the overflow successor of T1m.
const int kLongOnlyTier = 5;

int check_scores(const int8_t s[6]) {
    if (s[2] > 0 || s[5] > 0)
        return fail(HYPO_E_SCORES, "gap penalty must be non-positive (sr %d, lr %d)", s[2], s[5]);
    return HYPO_OK;
}

// Core: everything on the device.  d_out_pos/d_bound semantics: window w may write up to
// bound[w] bytes at d_out + d_out_pos[w].
int run_device(const WinDesc* d_win, uint64_t n_win, const ArmDesc* d_arms, uint64_t n_arms,
               const uint8_t* d_packed, uint64_t packed_bytes, char* d_out, const uint64_t* d_out_pos,
               uint32_t* d_out_len, const WinStat* d_stats, cudaStream_t stream, bool accumulate = false) {
    if (!accumulate) {
        g.poa_ms = 0.f;
        g.poa_launches = 0;
        memset(g.tier_windows, 0, sizeof(g.tier_windows));
        memset(g.fail_hist, 0, sizeof(g.fail_hist));
    }
    if (n_win == 0) return HYPO_OK;
    if (n_win > 0xfffffff0ull) return fail(HYPO_E_ARG, "too many windows in one batch");
    // control block: [0..kNumTiers) TierMax, then queue counters
    const size_t ctrl_bytes = sizeof(TierMax) * (kNumTiers + 1) + (64 + kNumFailReasons) * sizeof(uint32_t);
    CUDA_TRY(g.ctrl.reserve(ctrl_bytes));
    CUDA_TRY(g.lists.reserve(sizeof(uint32_t) * ((uint64_t)kNumTiers * n_win + (n_win + 1) * 2 + 64)));
    CUDA_TRY(cudaMemsetAsync(g.ctrl.p, 0, ctrl_bytes, stream));
    TierMax* d_tmax = (TierMax*)g.ctrl.p;
    uint32_t* d_queue = (uint32_t*)((char*)g.ctrl.p + sizeof(TierMax) * (kNumTiers + 1));
    uint32_t* d_lists = (uint32_t*)g.lists.p;
    // behind the tier lists: the overflow list of the running launch, then the per-window size projections
    uint32_t* d_over = d_lists + (uint64_t)kNumTiers * n_win;
    uint32_t* d_need = d_over + (n_win + 1);
    CUDA_TRY(cudaMemsetAsync(d_need, 0, sizeof(uint32_t) * n_win, stream));
    uint32_t* d_tier_lcap = d_queue + 16;
    uint32_t* d_tier_long = d_queue + 32;
    uint32_t* d_tier_seq = d_queue + 48;
    uint32_t* d_fail = d_queue + 64;

    uint32_t h_lcap[16] = {0}, h_long[16] = {0}, h_seq[16] = {0};
    for (int t = 0; t < kNumTiers; ++t) {
        h_lcap[t] = (uint32_t)kTiers[t].lcap;
        h_long[t] = (kTiers[t].long_ok ? 1u : 0u) | (t == kLongOnlyTier ? 2u : 0u);   // bit 0: LONG allowed, bit 1: SHORT not routed here
        h_seq[t] = kTiers[t].est_cap;
    }
    CUDA_TRY(cudaMemcpyAsync(d_tier_lcap, h_lcap, sizeof(h_lcap), cudaMemcpyHostToDevice, stream));
    CUDA_TRY(cudaMemcpyAsync(d_tier_long, h_long, sizeof(h_long), cudaMemcpyHostToDevice, stream));
    CUDA_TRY(cudaMemcpyAsync(d_tier_seq, h_seq, sizeof(h_seq), cudaMemcpyHostToDevice, stream));
    const int tb = 256;
    route_kernel<<<(unsigned)((n_win + tb - 1) / tb), tb, 0, stream>>>(d_win, d_stats, n_win, kNumTiers,
                                                                     d_tier_lcap, d_tier_long, d_tier_seq, d_lists,
                                                                     d_tmax);
    ++g.launches;
    CUDA_TRY(cudaGetLastError());

    TierMax* h_tmax = (TierMax*)g.pinned_ctrl;
    uint32_t* h_ovf_count = (uint32_t*)((char*)g.pinned_ctrl + 1024);
    CUDA_TRY(cudaMemcpyAsync(h_tmax, d_tmax, sizeof(TierMax) * kNumTiers, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    uint32_t routed[16];
    for (int t = 0; t < kNumTiers; ++t) routed[t] = h_tmax[t].count;

    // Windows that overflow tier t at run time are appended to the list of tier kTiers[t].next
    // (always a later one), behind its routed windows: every list holds each window at most once.
    uint32_t pend[16] = {0};
    for (int t = 0; t < kNumTiers; ++t) {
        const Tier& T = kTiers[t];
        uint32_t* d_work = d_lists + (uint64_t)t * n_win;
        const uint32_t n_work = routed[t] + pend[t];
        if (n_work == 0) continue;
        CUDA_TRY(cudaMemsetAsync(d_over, 0, sizeof(uint32_t), stream));

        Caps caps;
        caps.ncap = T.ncap; caps.ecap = T.ecap; caps.acap = T.acap; caps.scap = T.scap; caps.lcap = T.lcap;
        caps.alslots = t < kNumFixedTiers ? fixed_caps(t).alslots : kAlSlotsMax;
        bool need_paths = false;
        uint32_t sum_raw = 0, n_seq = 0;
        if (T.from_bounds || T.long_ok) {
            TierMax* d_tm = d_tmax + kNumTiers;
            CUDA_TRY(cudaMemsetAsync(d_tm, 0, sizeof(TierMax), stream));
            listmax_kernel<<<(n_work + tb - 1) / tb, tb, 0, stream>>>(d_win, d_stats, d_work, n_work, d_tm);
            ++g.launches;
            CUDA_TRY(cudaMemcpyAsync(h_tmax, d_tm, sizeof(TierMax), cudaMemcpyDeviceToHost, stream));
            CUDA_TRY(cudaStreamSynchronize(stream));
            sum_raw = h_tmax->sum_raw; n_seq = h_tmax->n_seq;
            need_paths = h_tmax->any_long != 0;
            if (T.from_bounds) {
                uint32_t lc = std::min<uint32_t>(std::max<uint32_t>(h_tmax->bound_len, 1), (uint32_t)T.lcap);
                // The only sequence that can be longer than the longest input is the round-2 backbone of a
                // LONG window: bounded by the node count of round 1, in practice about as long as the draft.
                // Sizing the columns (hence the DP slot, hence how many warps fit the workspace) for that
                // bound starves the grid, so every bound-driven tier but the last sizes them for twice the
                // longest input; a backbone beyond that leaves the tier with kFailLen and runs in the next.
                if (t + 1 < kNumTiers) lc = std::min<uint32_t>(lc, std::max<uint32_t>(2 * h_tmax->max_len + 64, 1023));
                caps.lcap = (int)lc;
                const uint32_t nb = std::max<uint32_t>(h_tmax->sum_len + 2, 64);
                caps.ncap = (int)std::min<uint32_t>(nb, (uint32_t)T.ncap);
                caps.ecap = (int)std::min<uint32_t>(nb + 64, (uint32_t)T.ecap);
                caps.acap = (int)std::min<uint32_t>(nb, (uint32_t)T.acap);
                caps.scap = (int)std::min<uint32_t>(2 * nb + 64, (uint32_t)T.scap);
            }
            if (n_seq > 32000)
                return fail(HYPO_E_CAPACITY, "window with %u sequences exceeds 16-bit edge weights", n_seq);
        }
        caps.tiles = T.one_tile ? 1 : (caps.lcap + 1 + kTileCols - 1) / kTileCols;
        const ArenaLayout L = arena_layout(caps);

        int wpb = T.warps_per_block;
        const int bps = T.blocks_per_sm;
        size_t smem = 0;
        if (T.smem_graph) {
            smem = (size_t)L.total * wpb;
            while (smem > (size_t)g.smem_optin && wpb > 1) { wpb /= 2; smem = (size_t)L.total * wpb; }
            if (smem > (size_t)g.smem_optin) return fail(HYPO_E_CAPACITY, "tier %d does not fit shared memory", t);
        }
        int blocks = g.sms * bps;
        uint64_t warps = (uint64_t)blocks * wpb;
        if (warps > n_work) { blocks = (int)((n_work + wpb - 1) / wpb); warps = (uint64_t)blocks * wpb; }
        // matrix rows (+ spare) and, behind them, the two boundary arrays of the multi-tile fill
        const uint64_t h_slot = ((uint64_t)(caps.ncap + 4) * caps.tiles * kTileCols + 2ull * (caps.ncap + 4) + 63) & ~63ull;
        // keep the DP workspace bounded: shrink the grid if the slots would exceed ~24 GB
        while (warps * h_slot * 2 > (24ull << 30) && blocks > 1) { blocks = (blocks + 1) / 2; warps = (uint64_t)blocks * wpb; }
        CUDA_TRY(g.H.reserve(warps * h_slot * sizeof(int16_t)));
        uint64_t g_slot = 0, p_slot = 0;
        if (!T.smem_graph) {
            g_slot = ((uint64_t)L.total + 255) & ~255ull;
            CUDA_TRY(g.gws.reserve(warps * g_slot));
        }
        if (need_paths) {
            p_slot = ((uint64_t)sum_raw + 2 * ((uint64_t)n_seq + 4) + 64 + 7) & ~7ull;
            CUDA_TRY(g.paths.reserve(warps * p_slot * sizeof(uint16_t)));
        }

        Params P;
        P.win = d_win; P.arms = d_arms; P.packed = d_packed;
        P.work = d_work; P.n_work = n_work;
        P.queue = d_queue + t;
        P.out = d_out; P.out_pos = d_out_pos; P.out_len = d_out_len;
        P.overflow = d_over;
        P.fail_hist = d_fail;
        P.need = d_need;
        P.H = (int16_t*)g.H.p; P.h_slot = h_slot;
        P.gws = (uint8_t*)g.gws.p; P.g_slot = g_slot;
        P.paths = need_paths ? (uint16_t*)g.paths.p : nullptr; P.p_slot = p_slot;
        P.caps = caps;
        P.sr_m = g.scores[0]; P.sr_n = g.scores[1]; P.sr_g = g.scores[2];
        P.lr_m = g.scores[3]; P.lr_n = g.scores[4]; P.lr_g = g.scores[5];
        CUDA_TRY(cudaEventRecord(g.ev0, stream));
        CUDA_TRY(launch_poa(P, t, T.smem_graph, T.one_tile, blocks, wpb, smem, stream));
        CUDA_TRY(cudaEventRecord(g.ev1, stream));
        ++g.launches;
        CUDA_TRY(cudaMemcpyAsync(h_ovf_count, d_over, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
        {
            float ms = 0.f;
            CUDA_TRY(cudaEventElapsedTime(&ms, g.ev0, g.ev1));
            g.poa_ms += ms;
            g.poa_launches += 1;
            g.tier_windows[t] += n_work;
        }
        const uint32_t over = *h_ovf_count;
        if (over != 0) {
            const int nxt = T.next;
            if (nxt >= kNumTiers)
                return fail(HYPO_E_CAPACITY, "%u window(s) exceed every device capacity tier (graph > 65534 nodes or "
                                             "scores outside the 16-bit DP range)", over);
            CUDA_TRY(cudaMemcpyAsync(d_lists + (uint64_t)nxt * n_win + routed[nxt] + pend[nxt], d_over + 1,
                                     sizeof(uint32_t) * over, cudaMemcpyDeviceToDevice, stream));
            pend[nxt] += over;
        }
    }
    uint32_t* h_fail = (uint32_t*)((char*)g.pinned_ctrl + 2560);
    CUDA_TRY(cudaMemcpyAsync(h_fail, d_fail, sizeof(g.fail_hist), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    for (int k = 0; k < kNumFailReasons; ++k) g.fail_hist[k] += h_fail[k];
    return HYPO_OK;
}

// Per-window static facts and output bounds of windows [0, n_win) of d_win into d_stats / d_bound.
// n_arms / packed_bytes are the limits the descriptors are validated against (`quiet`: a violation is
// reported as HYPO_E_ARG without a message: the pipelined path uses it to detect an input whose head
// windows reference data of the tail).
int prepare_stats(const WinDesc* d_win, uint64_t n_win, const ArmDesc* d_arms, uint64_t n_arms,
                  uint64_t packed_bytes, WinStat* d_stats, uint64_t* d_bound, cudaStream_t stream,
                  bool quiet = false) {
    uint32_t* d_bad = (uint32_t*)((char*)g.stats.p + g.stats.cap - 64);
    CUDA_TRY(cudaMemsetAsync(d_bad, 0, sizeof(uint32_t), stream));
    const int tb = 128;
    classify_kernel<<<(unsigned)((n_win + tb - 1) / tb), tb, 0, stream>>>(d_win, d_arms, n_win, n_arms, packed_bytes,
                                                                        d_stats, d_bound, d_bad);
    ++g.launches;
    CUDA_TRY(cudaGetLastError());
    uint32_t* h_bad = (uint32_t*)((char*)g.pinned_ctrl + 2048);
    CUDA_TRY(cudaMemcpyAsync(h_bad, d_bad, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    if (*h_bad) {
        if (quiet) return HYPO_E_ARG;
        return fail(HYPO_E_ARG, "%u window descriptor(s) reference arms/bytes out of range", *h_bad);
    }
    return HYPO_OK;
}

}  // namespace

extern "C" {

int hypo_gpu_abi_version(void) { return HYPO_B200_ABI_VERSION; }

const char* hypo_gpu_last_error(void) { return g_err.c_str(); }

uint64_t hypo_gpu_launch_count(void) { return g.launches; }

int hypo_gpu_init(const int8_t scores[6], int device) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_err.clear();
    if (!scores) return fail(HYPO_E_ARG, "scores == NULL");
    if (int rc = check_scores(scores)) return rc;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
        return fail(HYPO_E_CUDA, "no usable CUDA device (%s); this library has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= n_dev) return fail(HYPO_E_ARG, "device %d out of range (0..%d)", device, n_dev - 1);
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(HYPO_E_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                    prop.major, prop.minor);
    if (!g.stream) CUDA_TRY(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
    if (!g.copy_stream) CUDA_TRY(cudaStreamCreateWithFlags(&g.copy_stream, cudaStreamNonBlocking));
    if (!g.ev_head) CUDA_TRY(cudaEventCreateWithFlags(&g.ev_head, cudaEventDisableTiming));
    if (!g.ev_tail) CUDA_TRY(cudaEventCreateWithFlags(&g.ev_tail, cudaEventDisableTiming));
    if (!g.pinned_ctrl) CUDA_TRY(cudaHostAlloc(&g.pinned_ctrl, 4096, cudaHostAllocDefault));
    if (!g.ev0) CUDA_TRY(cudaEventCreate(&g.ev0));
    if (!g.ev1) CUDA_TRY(cudaEventCreate(&g.ev1));
    g.device = device;
    g.sms = prop.multiProcessorCount;
    g.smem_optin = (int)prop.sharedMemPerBlockOptin;
    for (int t = 0; t < kNumFixedTiers; ++t) {   // the table must agree with the kernels' constants
        const Caps c = fixed_caps(t);
        const Tier& T = kTiers[t];
        if (c.ncap != T.ncap || c.ecap != T.ecap || c.acap != T.acap || c.scap != T.scap || c.lcap != T.lcap ||
            T.from_bounds || !T.smem_graph)
            return fail(HYPO_E_ARG, "internal: tier table row %d disagrees with fixed_caps", t);
    }
    memcpy(g.scores, scores, 6);
    g.launches = 0;
    g.init = true;
    return HYPO_OK;
}

uint64_t hypo_gpu_out_bound(const HypoWindowDesc* win, uint64_t n_win, const HypoArmDesc* arms, uint64_t n_arms) {
    uint64_t total = 0;
    for (uint64_t w = 0; w < n_win; ++w) {
        uint64_t b = (uint64_t)win[w].draft_len + 2;
        const uint64_t n = (uint64_t)win[w].n_internal + win[w].n_pre + win[w].n_suf;
        for (uint64_t k = 0; k < n && win[w].first_arm + k < n_arms; ++k) b += (uint64_t)arms[win[w].first_arm + k].len + 2;
        total += std::max<uint64_t>(b, win[w].draft_len);
    }
    return total;
}

int hypo_gpu_consensus_batch_device(const HypoWindowDesc* d_win, uint64_t n_win, const HypoArmDesc* d_arms,
                                    uint64_t n_arms, const uint8_t* d_packed, uint64_t packed_bytes,
                                    char* d_out, const uint64_t* d_out_pos, uint32_t* d_out_len, void* stream) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_err.clear();
    if (!g.init) return fail(HYPO_E_NOT_INIT, "hypo_gpu_init has not been called");
    CUDA_TRY(cudaSetDevice(g.device));
    cudaStream_t s = stream ? (cudaStream_t)stream : g.stream;
    if (n_win == 0) return HYPO_OK;
    CUDA_TRY(g.out_off.reserve(sizeof(uint64_t) * (n_win + 1)));
    CUDA_TRY(g.stats.reserve(sizeof(WinStat) * n_win + 128));
    if (int rc = prepare_stats((const WinDesc*)d_win, n_win, (const ArmDesc*)d_arms, n_arms, packed_bytes,
                               (WinStat*)g.stats.p, (uint64_t*)g.out_off.p, s))
        return rc;
    return run_device((const WinDesc*)d_win, n_win, (const ArmDesc*)d_arms, n_arms, d_packed, packed_bytes, d_out,
                      d_out_pos, d_out_len, (const WinStat*)g.stats.p, s);
}

// Host-buffer entry point.  The input is copied in two parts on a second stream: a head of the windows
// (with the arms and packed bytes they reference) and the tail; the POA kernels of the head run while the
// tail is still crossing PCIe, which hides most of the H2D time behind compute.  This relies on the
// batch being laid out in window order (what hypo::WindowBatch and every sane packer produce); the head's
// descriptors are validated on the device against the head's limits, and an input that is not laid out
// that way silently takes the single-copy path instead.
int hypo_gpu_consensus_batch(const HypoWindowDesc* win, uint64_t n_win, const HypoArmDesc* arms, uint64_t n_arms,
                             const uint8_t* packed, uint64_t packed_bytes, char* out, uint64_t out_cap,
                             uint64_t* out_off) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_err.clear();
    if (!g.init) return fail(HYPO_E_NOT_INIT, "hypo_gpu_init has not been called");
    if (!out_off) return fail(HYPO_E_ARG, "out_off == NULL");
    if (n_win == 0) { out_off[0] = 0; return HYPO_OK; }
    if (!win || (!arms && n_arms) || (!packed && packed_bytes)) return fail(HYPO_E_ARG, "NULL input buffer");
    CUDA_TRY(cudaSetDevice(g.device));
    cudaStream_t s = g.stream;

    CUDA_TRY(g.win.reserve(sizeof(WinDesc) * n_win));
    CUDA_TRY(g.arms.reserve(sizeof(ArmDesc) * std::max<uint64_t>(n_arms, 1)));
    CUDA_TRY(g.packed.reserve(packed_bytes + 16));
    CUDA_TRY(g.out_pos.reserve(sizeof(uint64_t) * (n_win + 1)));
    CUDA_TRY(g.out_off.reserve(sizeof(uint64_t) * (n_win + 1)));
    CUDA_TRY(g.out_len.reserve(sizeof(uint32_t) * n_win));
    CUDA_TRY(g.stats.reserve(sizeof(WinStat) * n_win + 128));
    const WinDesc* d_win = (const WinDesc*)g.win.p;
    const ArmDesc* d_arms = (const ArmDesc*)g.arms.p;
    const uint8_t* d_packed = (const uint8_t*)g.packed.p;
    uint64_t* d_bound = (uint64_t*)g.out_off.p;   // reused as the compact offsets later
    uint64_t* d_pos = (uint64_t*)g.out_pos.p;
    WinStat* d_stats = (WinStat*)g.stats.p;
    uint64_t* h64 = (uint64_t*)((char*)g.pinned_ctrl + 3072);
    size_t tmp_bytes = 0;
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_bound, d_pos, n_win + 1, s));
    CUDA_TRY(g.cub_tmp.reserve(tmp_bytes + 256));
    tmp_bytes = g.cub_tmp.cap;

    // ---- split: head = first ~1/8 of the windows -------------------------------------------------
    uint64_t w_s = 0, a_s = 0, b_s = 0;
    bool piped = n_win >= 65536 && n_arms > 0 && packed_bytes > 0;
    if (piped) {
        w_s = std::max<uint64_t>(32768, n_win / 8);
        a_s = win[w_s].first_arm;
        piped = a_s <= n_arms;
        if (piped) {
            b_s = std::min<uint64_t>(win[w_s].draft_off, a_s < n_arms ? arms[a_s].off : packed_bytes);
            piped = b_s <= packed_bytes;
        }
    }
    // an upper bound of the windows' scratch need that does not require looking at the arms: a window
    // may write up to 2 * sum(len + 2) + 2 * draft_len + 4 bytes (classify_kernel; the factor 2 is the
    // LONG round-2 backbone), and every base occupies at least 2 bits of the slab
    const uint64_t scratch_cap = 8 * packed_bytes + 4 * n_arms + 8 * n_win + 64;

    // the caller's host buffers must not be touched after this function returns, on any path
    struct CopyGuard {
        cudaStream_t c;
        bool armed = false;
        ~CopyGuard() { if (armed) cudaStreamSynchronize(c); }
    } guard{g.copy_stream};
    bool copies_issued = false;
    if (piped) {
        guard.armed = true;
        CUDA_TRY(g.out_scratch.reserve(scratch_cap + 16));
        cudaStream_t c = g.copy_stream;
        copies_issued = true;
        CUDA_TRY(cudaMemcpyAsync(g.win.p, win, sizeof(WinDesc) * w_s, cudaMemcpyHostToDevice, c));
        if (a_s) CUDA_TRY(cudaMemcpyAsync(g.arms.p, arms, sizeof(ArmDesc) * a_s, cudaMemcpyHostToDevice, c));
        if (b_s) CUDA_TRY(cudaMemcpyAsync(g.packed.p, packed, b_s, cudaMemcpyHostToDevice, c));
        CUDA_TRY(cudaEventRecord(g.ev_head, c));
        CUDA_TRY(cudaMemcpyAsync((WinDesc*)g.win.p + w_s, win + w_s, sizeof(WinDesc) * (n_win - w_s), cudaMemcpyHostToDevice, c));
        if (n_arms > a_s)
            CUDA_TRY(cudaMemcpyAsync((ArmDesc*)g.arms.p + a_s, arms + a_s, sizeof(ArmDesc) * (n_arms - a_s), cudaMemcpyHostToDevice, c));
        if (packed_bytes > b_s)
            CUDA_TRY(cudaMemcpyAsync((uint8_t*)g.packed.p + b_s, packed + b_s, packed_bytes - b_s, cudaMemcpyHostToDevice, c));
        CUDA_TRY(cudaEventRecord(g.ev_tail, c));

        CUDA_TRY(cudaStreamWaitEvent(s, g.ev_head, 0));
        CUDA_TRY(cudaMemsetAsync(g.out_len.p, 0, sizeof(uint32_t) * n_win, s));
        // head descriptors must only reference what has arrived: limits a_s / b_s
        const int rc_head = prepare_stats(d_win, w_s, d_arms, a_s, b_s, d_stats, d_bound, s, /*quiet=*/true);
        if (rc_head == HYPO_E_ARG) {
            piped = false;   // not laid out in window order (or really invalid): single-copy path decides
            CUDA_TRY(cudaStreamWaitEvent(s, g.ev_tail, 0));
        } else if (rc_head != HYPO_OK) {
            return rc_head;
        }
    }

    if (piped) {
        // ---- head: positions, kernels (the tail is still being copied) ----------------------------
        CUDA_TRY(cudaMemsetAsync(d_bound + w_s, 0, sizeof(uint64_t), s));
        CUDA_TRY(cub::DeviceScan::ExclusiveSum(g.cub_tmp.p, tmp_bytes, d_bound, d_pos, w_s + 1, s));
        ++g.launches;
        CUDA_TRY(cudaMemcpyAsync(h64, d_pos + w_s, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        const uint64_t head_bytes = *h64;
        if (head_bytes > scratch_cap) return fail(HYPO_E_CAPACITY, "internal: scratch bound exceeded");
        if (int rc = run_device(d_win, w_s, d_arms, n_arms, d_packed, packed_bytes, (char*)g.out_scratch.p, d_pos,
                                (uint32_t*)g.out_len.p, d_stats, s))
            return rc;
        // ---- tail -----------------------------------------------------------------------------------
        CUDA_TRY(cudaStreamWaitEvent(s, g.ev_tail, 0));
        const uint64_t n_tail = n_win - w_s;
        if (int rc = prepare_stats(d_win + w_s, n_tail, d_arms, n_arms, packed_bytes, d_stats + w_s, d_bound + w_s, s))
            return rc;
        CUDA_TRY(cudaMemsetAsync(d_bound + n_win, 0, sizeof(uint64_t), s));
        CUDA_TRY(cub::DeviceScan::ExclusiveScan(g.cub_tmp.p, tmp_bytes, d_bound + w_s, d_pos + w_s, cuda::std::plus<>{},
                                                (uint64_t)head_bytes, n_tail + 1, s));
        ++g.launches;
        CUDA_TRY(cudaMemcpyAsync(h64, d_pos + n_win, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        if (*h64 > scratch_cap) return fail(HYPO_E_CAPACITY, "internal: scratch bound exceeded");
        if (int rc = run_device(d_win + w_s, n_tail, d_arms, n_arms, d_packed, packed_bytes, (char*)g.out_scratch.p,
                                d_pos + w_s, (uint32_t*)g.out_len.p + w_s, d_stats + w_s, s, /*accumulate=*/true))
            return rc;
    } else {
        // ---- single copy --------------------------------------------------------------------------
        if (copies_issued) {
            // the split was attempted and abandoned: everything is on its way on the copy stream
            CUDA_TRY(cudaStreamSynchronize(g.copy_stream));
        } else {
            CUDA_TRY(cudaMemcpyAsync(g.win.p, win, sizeof(WinDesc) * n_win, cudaMemcpyHostToDevice, s));
            if (n_arms) CUDA_TRY(cudaMemcpyAsync(g.arms.p, arms, sizeof(ArmDesc) * n_arms, cudaMemcpyHostToDevice, s));
            if (packed_bytes) CUDA_TRY(cudaMemcpyAsync(g.packed.p, packed, packed_bytes, cudaMemcpyHostToDevice, s));
        }
        CUDA_TRY(cudaMemsetAsync((char*)g.out_off.p + sizeof(uint64_t) * n_win, 0, sizeof(uint64_t), s));
        if (int rc = prepare_stats(d_win, n_win, d_arms, n_arms, packed_bytes, d_stats, d_bound, s)) return rc;
        CUDA_TRY(cub::DeviceScan::ExclusiveSum(g.cub_tmp.p, tmp_bytes, d_bound, d_pos, n_win + 1, s));
        ++g.launches;
        CUDA_TRY(cudaMemcpyAsync(h64, d_pos + n_win, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        CUDA_TRY(g.out_scratch.reserve(*h64 + 16));
        CUDA_TRY(cudaMemsetAsync(g.out_len.p, 0, sizeof(uint32_t) * n_win, s));
        if (int rc = run_device(d_win, n_win, d_arms, n_arms, d_packed, packed_bytes, (char*)g.out_scratch.p, d_pos,
                                (uint32_t*)g.out_len.p, d_stats, s))
            return rc;
    }

    // compact on the device: lengths -> offsets -> gather; then one D2H of the exact bytes
    uint64_t* d_len64 = d_bound;
    const int tb = 256;
    widen_kernel<<<(unsigned)((n_win + tb - 1) / tb), tb, 0, s>>>((const uint32_t*)g.out_len.p, d_len64, n_win);
    CUDA_TRY(cudaMemsetAsync(d_len64 + n_win, 0, sizeof(uint64_t), s));
    CUDA_TRY(g.lists.reserve(sizeof(uint64_t) * (n_win + 1)));
    uint64_t* d_off = (uint64_t*)g.lists.p;
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(g.cub_tmp.p, tmp_bytes, d_len64, d_off, n_win + 1, s));
    CUDA_TRY(cudaMemcpyAsync(h64, d_off + n_win, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    const uint64_t total = *h64;
    if (total > out_cap) return fail(HYPO_E_OUT_CAP, "output needs %llu bytes, out_cap is %llu", (unsigned long long)total,
                                     (unsigned long long)out_cap);
    CUDA_TRY(g.out_compact.reserve(total + 16));
    gather_kernel<<<(unsigned)((n_win * 32 + tb - 1) / tb), tb, 0, s>>>((const char*)g.out_scratch.p, (const uint64_t*)g.out_pos.p,
                                                                      (const uint32_t*)g.out_len.p, d_off,
                                                                      (char*)g.out_compact.p, n_win);
    g.launches += 3;
    CUDA_TRY(cudaGetLastError());
    if (total) CUDA_TRY(cudaMemcpyAsync(out, g.out_compact.p, total, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(out_off, d_off, sizeof(uint64_t) * (n_win + 1), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return HYPO_OK;
}

int hypo_gpu_compact_device(const char* d_scratch, const uint64_t* d_out_pos, const uint32_t* d_out_len,
                            uint64_t n_win, char* d_compact, uint64_t compact_cap, uint64_t* d_off,
                            uint64_t* total, void* stream) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_err.clear();
    if (!g.init) return fail(HYPO_E_NOT_INIT, "hypo_gpu_init has not been called");
    CUDA_TRY(cudaSetDevice(g.device));
    cudaStream_t s = stream ? (cudaStream_t)stream : g.stream;
    if (n_win == 0) { if (total) *total = 0; return HYPO_OK; }
    CUDA_TRY(g.out_off.reserve(sizeof(uint64_t) * (n_win + 1)));
    uint64_t* d_len64 = (uint64_t*)g.out_off.p;
    const int tb = 256;
    widen_kernel<<<(unsigned)((n_win + tb - 1) / tb), tb, 0, s>>>(d_out_len, d_len64, n_win);
    CUDA_TRY(cudaMemsetAsync(d_len64 + n_win, 0, sizeof(uint64_t), s));
    size_t tmp_bytes = 0;
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_len64, d_off, n_win + 1, s));
    CUDA_TRY(g.cub_tmp.reserve(tmp_bytes));
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(g.cub_tmp.p, tmp_bytes, d_len64, d_off, n_win + 1, s));
    uint64_t* h64 = (uint64_t*)((char*)g.pinned_ctrl + 3072);
    CUDA_TRY(cudaMemcpyAsync(h64, d_off + n_win, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    if (total) *total = *h64;
    if (*h64 > compact_cap) return fail(HYPO_E_OUT_CAP, "compact output needs %llu bytes, capacity is %llu",
                                        (unsigned long long)*h64, (unsigned long long)compact_cap);
    gather_kernel<<<(unsigned)((n_win * 32 + tb - 1) / tb), tb, 0, s>>>(d_scratch, d_out_pos, d_out_len, d_off, d_compact, n_win);
    g.launches += 3;
    CUDA_TRY(cudaGetLastError());
    return HYPO_OK;
}

int hypo_gpu_last_timing(float* poa_kernel_ms, uint32_t* poa_launches, uint32_t tier_windows[8]) {
    if (poa_kernel_ms) *poa_kernel_ms = g.poa_ms;
    if (poa_launches) *poa_launches = g.poa_launches;
    if (tier_windows) for (int t = 0; t < 8; ++t) tier_windows[t] = g.tier_windows[t];
    return HYPO_OK;
}

int hypo_gpu_last_fail_hist(uint32_t reasons[16]) {
    if (reasons) for (int k = 0; k < kNumFailReasons; ++k) reasons[k] = g.fail_hist[k];
    return HYPO_OK;
}

void hypo_gpu_shutdown(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g.init && !g.stream) return;
    cudaSetDevice(g.device);
    DevBuf* bufs[] = {&g.win, &g.arms, &g.packed, &g.out_scratch, &g.out_pos, &g.out_len, &g.out_off, &g.out_compact,
                      &g.stats, &g.lists, &g.ctrl, &g.H, &g.gws, &g.paths, &g.cub_tmp};
    for (DevBuf* b : bufs) b->release();
    if (g.pinned_ctrl) cudaFreeHost(g.pinned_ctrl);
    g.pinned_ctrl = nullptr;
    if (g.stream) cudaStreamDestroy(g.stream);
    g.stream = nullptr;
    if (g.copy_stream) cudaStreamDestroy(g.copy_stream);
    g.copy_stream = nullptr;
    if (g.ev_head) cudaEventDestroy(g.ev_head);
    if (g.ev_tail) cudaEventDestroy(g.ev_tail);
    g.ev_head = g.ev_tail = nullptr;
    if (g.ev0) cudaEventDestroy(g.ev0);
    if (g.ev1) cudaEventDestroy(g.ev1);
    g.ev0 = g.ev1 = nullptr;
    g.init = false;
}

}  // extern "C"
