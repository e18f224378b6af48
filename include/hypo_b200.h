/*
 * hypo_b200.h — C ABI of the B200-native POA-consensus hot path of HyPo.
 *
 * This is the drop-in boundary for ONE path of the reference:
 *     hypo::Window::generate_consensus(engine_idx)        reference src/Window.cpp:44-61
 *       -> generate_consensus_short / _long / curate      reference src/Window.cpp:87-254
 *       -> spoa::AlignmentEngine::align (SISD, linear)     reference external/spoa/src/sisd_alignment_engine.cpp:246-439
 *       -> spoa::Graph::add_alignment / topological_sort   reference external/spoa/src/graph.cpp:154-353
 *       -> spoa::Graph::generate_consensus{,_custom}       reference external/spoa/src/graph.cpp:467-476,533-568,610-705
 * as it is driven by the "Polish with arms" block of Hypo::polish
 * (reference src/Hypo.cpp:236-248).  Everything else in HyPo stays host code.
 *
 * Plain C, plain pointers and sizes; no C++/torch types cross this boundary.
 * All entry points return 0 on success and a non-zero HYPO_E_* code on
 * failure; hypo_gpu_last_error() then returns a human-readable message
 * (the reference convention is fprintf(stderr,"[Hypo::X] Error: ...")+exit(1),
 * which the host wrapper reproduces around these calls).
 *
 * There is NO CPU fallback behind this API: if no CUDA device is usable the
 * calls fail loudly.
 */
#ifndef HYPO_B200_H
#define HYPO_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HYPO_B200_ABI_VERSION 2
#define HYPO_MAX_DEVICES 16

/* Window types — hypo::WindowType, reference include/Window.hpp:35-38 */
#define HYPO_WINDOW_SHORT 0u
#define HYPO_WINDOW_LONG  1u

/* Error codes */
#define HYPO_OK            0
#define HYPO_E_NOT_INIT    1   /* hypo_gpu_init was not called                         */
#define HYPO_E_CUDA        2   /* a CUDA runtime call failed (no device, OOM, ...)     */
#define HYPO_E_ARG         3   /* malformed descriptors / offsets out of range         */
#define HYPO_E_OUT_CAP     4   /* output buffer too small (see hypo_gpu_out_bound)     */
#define HYPO_E_SCORES      5   /* gap penalty > 0 (spoa rejects it too,
                                  reference external/spoa/src/alignment_engine.cpp:42-49) */
#define HYPO_E_CAPACITY    6   /* a window exceeds every device capacity tier: a graph of
                                  more than 65534 nodes, a node with more than 254
                                  in-edges, or more than 32000 reads in one window.
                                  Scores x size never cause it: reads beyond the 16-bit DP
                                  range are computed with 32-bit cells like the reference
                                  (external/spoa/src/sisd_alignment_engine.cpp:263-342)  */

/*
 * One weak-region window == one hypo::Window (reference include/Window.hpp:122-135).
 *
 *  - draft: the PackedSeq<4> bytes of Window::_draft verbatim: 2 bases/byte,
 *    base i in the HIGH nibble of byte i>>1 when i is even, LOW nibble when odd;
 *    nibble codes A0 C1 G2 T3, anything >=4 unpacks to 'N'
 *    (reference src/PackedSeq.cpp:45-48, include/PackedSeq.hpp:54-72).
 *  - arms [first_arm, first_arm + n_internal + n_pre + n_suf) of the arm table,
 *    in container order: _internal_arms, then _pre_arms, then _suf_arms
 *    (reference include/Window.hpp:131-133).  The device applies the
 *    reverse-prefix-order rule of generate_consensus_short itself.
 *  - n_internal/n_pre/n_suf are both the counters _num_internal/_num_pre/_num_suf
 *    and the vector sizes (they are always equal in the reference,
 *    include/Window.hpp:66-118); zero-length arms are allowed and are skipped
 *    exactly like the reference skips them (src/Window.cpp:103,114,125,182).
 *  - n_empty is Window::_num_empty.
 */
typedef struct HypoWindowDesc {
    uint64_t draft_off;    /* byte offset of the packed draft inside `packed`      */
    uint64_t first_arm;    /* index of this window's first arm in the arm table    */
    uint32_t draft_len;    /* draft length in bases (Window::get_window_len)       */
    uint32_t n_internal;
    uint32_t n_pre;
    uint32_t n_suf;
    uint32_t n_empty;
    uint32_t wtype;        /* HYPO_WINDOW_SHORT / HYPO_WINDOW_LONG                 */
} HypoWindowDesc;          /* 40 bytes */

/*
 * One arm == one PackedSeq<2> (reference include/PackedSeq.hpp:152-154):
 * 4 bases/byte, base i in bits (6 - 2*(i&3)) of byte i>>2, codes A0 C1 G2 T3
 * (reference src/PackedSeq.cpp:45, include/globalDefs.hpp:158-178).
 */
typedef struct HypoArmDesc {
    uint64_t off;          /* byte offset of the packed arm inside `packed`        */
    uint32_t len;          /* arm length in bases (PackedSeq::get_seq_size)        */
    uint32_t reserved;     /* must be 0                                            */
} HypoArmDesc;             /* 16 bytes */

/*
 * Replaces hypo::Window::prepare_for_poa(const ScoreParams&, num_threads)
 * (reference src/Window.cpp:31-42).  scores = {sr_match, sr_mismatch, sr_gap,
 * lr_match, lr_mismatch, lr_gap} == hypo::ScoreParams (reference
 * include/globalDefs.hpp:58-66).
 *
 * hypo_gpu_init       drives ONE device (CUDA ordinal `device`); this is what a
 *                     one-process-per-GPU launcher (torchrun, MPI) calls.
 * hypo_gpu_init_multi drives devices 0 .. n_gpus-1 from this ONE host process, as the
 *                     reference is one process (src/Hypo.cpp:34-35, OpenMP):
 *                     hypo_gpu_consensus_batch then cuts every batch into n_gpus
 *                     contiguous window ranges of equal estimated cost, runs them side by
 *                     side (one host thread per device) and returns the consensus strings
 *                     in window order - byte-identical for every n_gpus.  n_gpus <= 0:
 *                     all visible devices.
 * Either may be called again: with the same devices only the scores change; with other
 * devices everything held on the old ones (buffers, streams, events) is released first.
 */
int hypo_gpu_init(const int8_t scores[6], int device);
int hypo_gpu_init_multi(const int8_t scores[6], int n_gpus);

/* Number of devices the library currently drives (0 before init / after shutdown). */
int hypo_gpu_device_count(void);

/*
 * Knobs for tests and measurements (the defaults are what production uses):
 *   "first_tier" 0..10 routing starts at this capacity tier (7 = the bound-driven last tier; 8 / 9 = the
 *                      group tiers Tq / Th, 10 = T2s, the estimate-driven shared-memory tier for large windows:
 *                      windows that do not fit them are routed as from tier 0)
 *   "big_tier"   0|1   large windows whose estimated DAG fits shared memory start in T2s instead of the
 *                      bound-driven tier T2 (default 1)
 *   "group_tiers" 0|1|2  small SHORT windows (<= 63 symbols) start in the group tiers, several windows per
 *                      warp: 0 never, 1 in batches of at least 131072 windows (default: a small batch is better
 *                      off in one launch of the compact tier), 2 always; only changes where windows run
 *   "group_sort" 0|1   tier lists are ordered on the device before they run: by size class for the group tiers
 *                      (the windows of a warp advance in lock-step), by estimated cost, largest first, for
 *                      the others (default 1)
 *   "teams"      0|1   the bound-driven tiers give every window a team of four warps when a launch has few
 *                      windows (default 1; T1m / T1 always run teams of two / four warps)
 *   "scap"       n     DFS-stack entries of the bound-driven tiers except the last (0 = from bounds)
 *   "probe"      0|1   a shared-memory tier whose list holds >= 16384 windows runs the first 4096 alone and,
 *                      if a quarter of them outgrow the tier, hands the rest of the list to the successor tier
 *                      untried (default 1; only changes where windows run)
 *   "gather"     0|2   multi-device result gather: 0 every device copies its bytes to the host
 *                      itself, 2 NCCL send/recv to device 0 over NVLink, then one copy
 * Returns HYPO_E_ARG for an unknown name or a value out of range.
 */
int hypo_gpu_set_option(const char* name, int64_t value);

/*
 * Per-window upper bound of the bytes window w may write while it is being computed (its
 * consensus is never longer, its intermediate LONG round-1 consensus may be):
 *     SHORT: sum(arm_len) + 2 * n_arms + draft_len + 2      (>= nodes of the marked graph)
 *     LONG : 2 * sum(arm_len) + draft_len + 2               (round-2 backbone <= nodes of round 1)
 * This is the size every d_out_pos slot of hypo_gpu_consensus_batch_device must have, and what
 * the host entry point reserves internally.  hypo_gpu_out_bound is the sum over the batch
 * (a safe size for `out` of hypo_gpu_consensus_batch).  Pure host arithmetic, no device needed.
 */
int hypo_gpu_window_bounds(const HypoWindowDesc* win, uint64_t n_win,
                           const HypoArmDesc* arms, uint64_t n_arms, uint64_t* bound);
uint64_t hypo_gpu_out_bound(const HypoWindowDesc* win, uint64_t n_win,
                            const HypoArmDesc* arms, uint64_t n_arms);

/*
 * Replaces the OpenMP loop over Contig::generate_consensus(w, tid) ->
 * Window::generate_consensus(tid) (reference src/Hypo.cpp:238-247,
 * include/Contig.hpp:119) for a whole batch of windows.
 *
 * All pointers are HOST pointers (pinned memory makes the copies faster but is
 * not required).  On success, the consensus of window w — byte-identical to
 * what Window::get_consensus() returns in the reference's default (SISD) build —
 * is out[out_off[w] .. out_off[w+1]) (ASCII ACGTN, no terminator).
 * out_off must hold n_win+1 entries.  Nothing is retained after return.
 */
int hypo_gpu_consensus_batch(const HypoWindowDesc* win, uint64_t n_win,
                             const HypoArmDesc* arms, uint64_t n_arms,
                             const uint8_t* packed, uint64_t packed_bytes,
                             char* out, uint64_t out_cap, uint64_t* out_off);

/*
 * Same computation with every buffer already resident in device memory on the FIRST device
 * the library drives (d_* are device pointers; `stream` is a cudaStream_t passed as void*,
 * NULL = the library's own stream).  The call launches its kernels on `stream` and BLOCKS the
 * calling thread until they are done (it synchronises the stream two to four times for tier
 * bookkeeping); results are valid on return.
 *   d_out_len[w]  = consensus length of window w
 *   d_out         = consensus bytes, window w at d_out + d_out_pos[w]; d_out_pos is
 *                   caller-provided (n_win entries) and slot w must hold the bytes
 *                   hypo_gpu_window_bounds reports for w (e.g. an exclusive scan of them).
 */
int hypo_gpu_consensus_batch_device(const HypoWindowDesc* d_win, uint64_t n_win,
                                    const HypoArmDesc* d_arms, uint64_t n_arms,
                                    const uint8_t* d_packed, uint64_t packed_bytes,
                                    char* d_out, const uint64_t* d_out_pos,
                                    uint32_t* d_out_len, void* stream);

/*
 * Packs the per-window results of hypo_gpu_consensus_batch_device into one contiguous,
 * window-ordered byte string on the device (what the final consensus gather ships):
 *   d_off[w]  = start of window w inside d_compact (n_win+1 entries, device)
 *   *total    = total bytes (host); fails with HYPO_E_OUT_CAP if > compact_cap.
 */
int hypo_gpu_compact_device(const char* d_scratch, const uint64_t* d_out_pos, const uint32_t* d_out_len,
                            uint64_t n_win, char* d_compact, uint64_t compact_cap, uint64_t* d_off,
                            uint64_t* total, void* stream);

/*
 * Measurement hook: device time (CUDA events on the launching stream) and launch count of the
 * POA kernels of the most recent batch call, and how many windows each capacity tier ran.
 */
int hypo_gpu_last_timing(float* poa_kernel_ms, uint32_t* poa_launches, uint32_t tier_windows[8]);

/*
 * The tier histogram of the most recent batch call for ALL capacity tiers: entries 0..7 as above,
 * 8 = Tq and 9 = Th, the group tiers that run several small SHORT windows per warp (4 x 8 lanes for
 * windows of <= 31 symbols, 2 x 16 lanes for <= 63 symbols), 10 = T2s (large windows, DAG in shared memory,
 * capacities from an estimate); entries beyond the last tier are 0.
 */
int hypo_gpu_last_tier_windows(uint32_t* tier_windows, int n);

/*
 * Diagnostic hook: how many times windows were abandoned in a capacity tier during the most recent
 * batch call, by reason (index: 1 sequence longer than the tier's columns, 2 16-bit DP range,
 * 3 node capacity, 4 aligned-list blocks, 5 clique size, 6 edge capacity / in-degree, 7 DFS stack,
 * 8 LONG path slot, 9 LONG window in a SHORT-only tier, 10 growth per read extrapolates beyond the
 * tier (abandoned early), 11 passed on without work because an earlier tier's projection exceeds this
 * tier too).  Every abandoned window is re-run in a larger tier; this only explains the tier histogram
 * of hypo_gpu_last_timing.
 */
int hypo_gpu_last_fail_hist(uint32_t reasons[16]);

/*
 * Arm extraction: the step that FEEDS the path.  Replaces, for short reads,
 *     Alignment::initialise_pos / copy_data        reference src/Alignment.cpp:513-576
 *     Alignment::find_short_arms / find_bp         reference src/Alignment.cpp:222-259,321-404
 *     Alignment::prepare_short_arm                 reference src/Alignment.cpp:406-509
 *     Alignment::add_arms + the pruning rules of Contig::fill_short_windows
 *                                                  reference src/Alignment.cpp:299-318, src/Contig.cpp:249-289
 * and the host batch packer: the reads' bases go from the BAM records straight into the packed slab of the
 * batch, on the device.  Inputs are what the reference holds at that point of Hypo::polish
 * (src/Hypo.cpp:196-232), all host pointers:
 *   contigs / regions : the region table of every contig (Contig::_reg_pos / _reg_type / _reg_info /
 *                       _anchor_kmers): regions of contig c are [first_region, first_region + n_regions),
 *                       sorted by start; a strong region carries the k-mers at its first and last k bases
 *                       (src/Contig.cpp:128-129), a minimiser region the minimiser (:596,614)
 *   drafts            : the contigs' PackedSeq<4> bytes (Contig::_pseq), contig c at draft_off
 *   alns / cigar / seqs : the alignments the reference keeps (src/Hypo.cpp:296-300), in file order:
 *                       core.pos, the BAM CIGAR words (len << 4 | op) and bam_get_seq() bytes verbatim
 *                       (two bases per byte, high nibble first, A1 C2 G4 T8); soft clips are cut off on the
 *                       device, a read with another base in its aligned part is dropped like the reference
 *                       drops it (PackedSeq<2> cannot hold it, src/Alignment.cpp:556-571)
 *   k                 : the solid k-mer length (anchors); the minimiser length is 10 (src/main.cpp:86)
 * Outputs (host): the batch exactly as WindowBatch::pack would lay it out for the windows the reference
 * ends up with - dropped windows are absent, prefix / suffix arms already cleared where the reference
 * clears them, arms in alignment order - plus win_region[w], the region (batch-wide index) window w polishes.
 * *n_win / *n_arms / *packed_bytes report the sizes; HYPO_E_OUT_CAP if a capacity is too small (sizes are
 * still reported).
 */
#define HYPO_REG_SR  0u
#define HYPO_REG_MSR 1u
#define HYPO_REG_SWS 2u
#define HYPO_REG_WS  3u
#define HYPO_REG_SW  4u
#define HYPO_REG_MWM 5u
#define HYPO_REG_WM  6u
#define HYPO_REG_MW  7u
#define HYPO_REG_SWM 8u
#define HYPO_REG_MWS 9u
#define HYPO_REG_OTHER 10u
typedef struct HypoRegionRec {
    uint64_t key0;         /* strong region: its first k-mer (2 bits per base, first base highest); minimiser region: the minimiser */
    uint64_t key1;         /* strong region: its last k-mer                                        */
    uint32_t start;        /* first base of the region on its contig                               */
    uint32_t type;         /* HYPO_REG_*                                                           */
} HypoRegionRec;           /* 24 bytes */
typedef struct HypoContigDesc {
    uint64_t first_region; /* index of the contig's first region in `regions`                      */
    uint64_t draft_off;    /* byte offset of the contig's PackedSeq<4> draft in `drafts`           */
    uint32_t n_regions;
    uint32_t len;          /* contig length in bases                                               */
} HypoContigDesc;          /* 24 bytes */
typedef struct HypoAlnDesc {
    uint64_t cigar_off;    /* index of the first CIGAR word of this alignment in `cigar`           */
    uint64_t seq_off;      /* byte offset of bam_get_seq() of this read in `seqs`                  */
    uint32_t contig;
    uint32_t pos;          /* core.pos (0-based)                                                   */
    uint32_t n_cigar;
    uint32_t l_qseq;
} HypoAlnDesc;             /* 32 bytes */
int hypo_gpu_extract_arms(const HypoContigDesc* contigs, uint64_t n_contigs,
                          const HypoRegionRec* regions, uint64_t n_regions,
                          const uint8_t* drafts, uint64_t draft_bytes,
                          const HypoAlnDesc* alns, uint64_t n_alns,
                          const uint32_t* cigar, uint64_t n_cigar,
                          const uint8_t* seqs, uint64_t seq_bytes, uint32_t k,
                          HypoWindowDesc* win, uint64_t win_cap, uint64_t* n_win, uint64_t* win_region,
                          HypoArmDesc* arms, uint64_t arm_cap, uint64_t* n_arms,
                          uint8_t* packed, uint64_t packed_cap, uint64_t* packed_bytes);

/*
 * The fused step: alignments in, polished contigs out - arm extraction, the POA consensus of every window
 * and the stitching of the contigs without the batch or the consensus strings ever visiting the host
 * (Hypo::polish from "Short arms computing" to "Writing results", reference src/Hypo.cpp:196-267, for short
 * reads).  Same inputs as hypo_gpu_extract_arms; out receives the polished contigs back to back, out_off
 * (n_contigs + 1 entries) their starts.  One driven device.
 */
int hypo_gpu_polish_alignments(const HypoContigDesc* contigs, uint64_t n_contigs,
                               const HypoRegionRec* regions, uint64_t n_regions,
                               const uint8_t* drafts, uint64_t draft_bytes,
                               const HypoAlnDesc* alns, uint64_t n_alns,
                               const uint32_t* cigar, uint64_t n_cigar,
                               const uint8_t* seqs, uint64_t seq_bytes, uint32_t k,
                               char* out, uint64_t out_cap, uint64_t* out_off);

/*
 * Support counting: the step before windowing.  Replaces the two OpenMP loops of
 * Hypo::polish over the alignments of a contig batch (reference src/Hypo.cpp:135-171):
 *     Alignment::update_solidkmers_support      reference src/Alignment.cpp:65-131
 *     Alignment::update_minimisers_support      reference src/Alignment.cpp:133-220
 * The reference increments 16-bit counters under a mutex per k-mer / per region (include/Contig.hpp:40-52,
 * 207-214); here one thread per alignment adds to 32-bit counters with device atomics (identical values as
 * long as no counter passes 65 535).  Alignments are given as for hypo_gpu_extract_arms (position, BAM CIGAR
 * words, bam_get_seq() bytes; soft clips are cut off, reads with a base other than A/C/G/T in the aligned
 * part are skipped like the reference drops them).  All pointers are host pointers.
 *
 * hypo_gpu_solid_kmer_support:
 *   contig_first_kmer[c] .. [c+1] : the solid k-mers of contig c (Contig::_solid_pos / _kmerinfo,
 *                                   src/Contig.cpp:40-74), sorted by position: solid_pos[] and kmer_id[]
 *                                   (the k-mer, 2 bits per base, first base highest)
 *   coverage[i] / support[i]      : out, per solid k-mer (KmerInfo::coverage / support)
 * hypo_gpu_minimiser_support (after Contig::prepare_for_division, src/Contig.cpp:75-186):
 *   contig_first_bound[c] .. [c+1]: the boundaries between strong regions and the regions in between of
 *                                   contig c (set bits of Contig::_reg_pos at that stage, 0 and the contig
 *                                   length included), bounds[]
 *   contig_even[c]                : Contig::_is_win_even (the regions with even index are the weak ones)
 *   region_first_mini[r] .. [r+1] : the minimisers of the region that starts at bounds[r] (none for strong
 *                                   regions): mini_pos[] absolute positions, mini_val[] the minimisers
 *                                   (MWMinimiserInfo::rel_pos accumulated / ::minimisers); one entry per
 *                                   bound plus the end
 *   coverage[i] / support[i]      : out, per minimiser (MWMinimiserInfo::coverage / support)
 */
int hypo_gpu_solid_kmer_support(const uint64_t* contig_first_kmer, uint64_t n_contigs,
                                const uint32_t* solid_pos, const uint64_t* kmer_id, uint64_t n_kmers,
                                const HypoAlnDesc* alns, uint64_t n_alns,
                                const uint32_t* cigar, uint64_t n_cigar,
                                const uint8_t* seqs, uint64_t seq_bytes, uint32_t k,
                                uint32_t* coverage, uint32_t* support);
int hypo_gpu_minimiser_support(const uint64_t* contig_first_bound, const uint8_t* contig_even, uint64_t n_contigs,
                               const uint32_t* bounds, uint64_t n_bounds,
                               const uint64_t* region_first_mini,
                               const uint32_t* mini_pos, const uint32_t* mini_val, uint64_t n_minis,
                               const HypoAlnDesc* alns, uint64_t n_alns,
                               const uint32_t* cigar, uint64_t n_cigar,
                               const uint8_t* seqs, uint64_t seq_bytes,
                               uint32_t* coverage, uint32_t* support);

/*
 * Output stitching: the step after the path.  Replaces the region loop of Contig::operator<<
 * (reference src/Contig.cpp:345-366): the polished contig is its regions in order - a strong region
 * (or a window nobody polished) is copied from the contig's PackedSeq<4> draft, a polished window is
 * replaced by its consensus.  One region per HypoRegionDesc:
 *   window == HYPO_REGION_DRAFT : bases [src, src + len) of the contig's draft (unpacked: nibble codes
 *                                 A0 C1 G2 T3, anything else 'N', reference include/PackedSeq.hpp:54-72)
 *   otherwise                   : the consensus of window `window` of the batch (src / len ignored)
 * regions of contig c are [contig_first_region[c], contig_first_region[c+1]); its draft starts at byte
 * draft_off[c] of `drafts`.  cons / cons_off: the consensus bytes and n_win+1 offsets as returned by
 * hypo_gpu_consensus_batch - or cons == NULL to stitch straight from the result of the most recent
 * hypo_gpu_consensus_batch call, which is still resident on the device (one driven device only).
 * out receives the contigs' polished sequences back to back, out_off[c] their starts (n_contigs+1).
 * All pointers are host pointers.
 */
#define HYPO_REGION_DRAFT 0xffffffffu
typedef struct HypoRegionDesc {
    uint64_t src;
    uint32_t len;
    uint32_t window;
} HypoRegionDesc;          /* 16 bytes */
int hypo_gpu_stitch(const HypoRegionDesc* regions, uint64_t n_regions,
                    const uint64_t* contig_first_region, uint64_t n_contigs,
                    const uint8_t* drafts, const uint64_t* draft_off, uint64_t draft_bytes,
                    const char* cons, const uint64_t* cons_off, uint64_t n_win,
                    char* out, uint64_t out_cap, uint64_t* out_off);

/*
 * Measurement hooks for the compute roofline: the DP cells of the most recent batch
 * call - sum over windows and reads of (nodes + 1) x (read length + 1), the size of the matrix the
 * reference fills for that read (external/spoa/src/sisd_alignment_engine.cpp:52-75) - and the rate the
 * device sustains for one instruction of the fill's inner loop, measured on the spot in 10^9 warp
 * instructions per second over the whole GPU:
 *   op 0 VIADDMNMX.S16x2 (__viaddmax_s16x2)   1 VIMNMX3.S16x2 (__vimax3_s16x2)   2 VIADDMNMX.S32
 *      3 SHFL.UP                              4 PRMT                             5 IMAD
 */
uint64_t hypo_gpu_last_cells(void);
int hypo_gpu_issue_rate(int op, double* g_warp_instr_per_s);

/* Windows the tier probes of the most recent batch call sent on without trying them in a tier. */
uint64_t hypo_gpu_last_rerouted(void);

/* Number of kernel launches issued by this library since hypo_gpu_init. */
uint64_t hypo_gpu_launch_count(void);

/* Thread-local message describing the last failure ("" if none). */
const char* hypo_gpu_last_error(void);

/*
 * Page-locked host memory for the batch buffers (copies from it run at full PCIe speed and
 * asynchronously).  Optional: every entry point accepts ordinary memory as well.  Valid until
 * hypo_gpu_host_free; independent of init / shutdown.
 */
void* hypo_gpu_host_alloc(uint64_t bytes);
void hypo_gpu_host_free(void* p);

/* Releases device memory and streams. */
void hypo_gpu_shutdown(void);

int hypo_gpu_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* HYPO_B200_H */
