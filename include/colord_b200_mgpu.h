/* colord_b200_mgpu.h — C-ABI of the multi-GPU exchanges (libcolord_b200_mgpu.so: NCCL over NVLink / NVSwitch).
 *
 * The reference has no multi-device path; the seam these calls sit in is the one between its stage 1a and stage 1b
 * (compression.cpp:432-464 -> :564): every GPU holds a shard of the reads (contiguous read ids, one clb_ctx per GPU, all in one
 * process) and has counted the k-mers of its shard with clb_append_reads.  The two calls below then make (1) the filtered k-mer set and
 * (2) the reference-read set GLOBAL, after which clb_graph_build / clb_encode / the stage-3 encoders of every context produce, for
 * the reads of its shard, exactly what a single GPU produces for the same reads (candidates, tuples; the streams are per shard).
 *
 * Both calls are COLLECTIVE: one host thread per rank calls them, each with its own rank, all with the same group.  An error on one
 * rank is reported on every rank.  Plain pointers and sizes; no exceptions cross the boundary.
 */
#ifndef COLORD_B200_MGPU_H
#define COLORD_B200_MGPU_H

#include "colord_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct clb_group clb_group;

/* ctxs[n]: the contexts of the ranks, devices[n]: their CUDA device ordinals (ncclCommInitAll over them). */
clb_status clb_group_create(clb_ctx* const* ctxs, const int32_t* devices, uint32_t n, clb_group** out);
void       clb_group_destroy(clb_group* group);
const char* clb_group_last_error(const clb_group* group, uint32_t rank);

/* (1) k-mer counts.  k-mers are owned by hash partition (clb_counts_size / clb_counts_export): one all-to-all moves every
 * (k-mer, count) pair to its owner, the owner thresholds its share (clb_counts_merge + clb_count_finalize), one all-gather hands every
 * rank the union of the survivors (clb_filter_import): the filter of the WHOLE input on every rank, as CKmerCounter + CKmerFilter
 * leave it (count_kmers.cpp:28-68, filter_kmers.cpp:45-83).  global_stats: the statistics of the whole input.  Replaces the
 * single-GPU clb_count_finalize. */
clb_status clb_group_exchange_counts(clb_group* group, uint32_t rank, clb_kmer_stats* global_stats);

/* (2) reference reads.  sampled_local[n_local]: the sampler's decisions (CRefReadsAccepter over GLOBAL read ids) for this rank's reads,
 * lengths_local[n_local]: their lengths.  One all-gather distributes every rank's reference reads (sampled and free of N); the ones
 * of the ranks before this one are appended as context reads (clb_append_context_reads): reads_sim_graph.cpp:374-395 sees the same
 * earlier reference reads as on one GPU.  After clb_group_exchange_counts, before clb_graph_build.  n_context: how many were appended. */
clb_status clb_group_exchange_reference_reads(clb_group* group, uint32_t rank, const uint8_t* sampled_local, const uint32_t* lengths_local,
                                              uint32_t n_local, uint32_t* n_context);

#ifdef __cplusplus
}
#endif
#endif
