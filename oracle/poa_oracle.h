/* TEST INFRASTRUCTURE ONLY — see poa_oracle.c. */
#ifndef POA_ORACLE_H
#define POA_ORACLE_H
#include <stdint.h>
#include "../include/hypo_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Same contract as hypo_gpu_consensus_batch, computed by the CPU restatement.
 * n_threads <= 0: all OpenMP threads.  seconds (nullable): wall time of the loop. */
int poa_oracle_consensus_batch(const int8_t scores[6], const HypoWindowDesc* win, uint64_t n_win,
                               const HypoArmDesc* arms, uint64_t n_arms, const uint8_t* packed,
                               uint64_t packed_bytes, char* out, uint64_t out_cap,
                               uint64_t* out_off, int n_threads, double* seconds);

/* Raw POA: sequences added in order (align_type[i]: 0 kNW, 1 kLOV, 2 kROV; NULL = all kNW),
 * consensus = heaviest bundle.  Mirrors hypo_ref_spoa_consensus in ref_driver.cpp. */
int poa_oracle_spoa_consensus(int8_t m, int8_t n, int8_t g, const char* seqs,
                              const uint64_t* seq_off, uint32_t n_seq, const uint8_t* align_type,
                              char* out, uint64_t out_cap, uint64_t* out_len);

/* Per-window statistics of the last graph (for DESIGN.md sizing / roofline accounting):
 * stats[0]=nodes, [1]=edges, [2]=sum of DP cells, [3]=max in-degree, [4]=max clique size. */
int poa_oracle_window_stats(const int8_t scores[6], const HypoWindowDesc* win,
                            const HypoArmDesc* arms, const uint8_t* packed, uint64_t stats[5]);

/* How the window's DAG grows: growth[3*i..] = nodes, edges, cumulative DP cells after the i-th sequence
 * (both rounds of a LONG window, one after the other); *n = entries written (at most cap).  Feeds
 * tools/projection_sim.py, the offline model of the kernels' growth projection. */
int poa_oracle_window_growth(const int8_t scores[6], const HypoWindowDesc* win, const HypoArmDesc* arms,
                             const uint8_t* packed, uint64_t* growth, uint32_t cap, uint32_t* n);

#ifdef __cplusplus
}
#endif
#endif
