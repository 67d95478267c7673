/* colord_b200.h — C-ABI of the B200 (sm_100a) hot path of a CoLoRd-compatible long-read compressor.
 *
 * The reference (refresh-bio/colord @ 25b2860) has no FFI layer: its compression stages are C++ classes
 * wired by runCompression (src/colord/compression.cpp:344).  Each entry point below replaces one of those
 * internal seams; the reference line that a binding would swap out is cited per function.  Plain pointers
 * and sizes only; the library owns all device memory; no exit()/exceptions cross the boundary — every
 * call returns a clb_status and clb_last_error() explains a failure.  There is NO CPU fallback: without a
 * CUDA device every compute entry point returns CLB_ERR_NO_DEVICE.
 *
 * Read representation at the boundary: ASCII bases 'A','C','G','T','N' concatenated without separators
 * (what the reference's FASTQ reader holds before to_read_t, in_reads.cpp:24-42) + uint64 offsets[n+1].
 * Any other byte is an error (the reference aborts on it, in_reads.cpp:31-35).
 */
#ifndef COLORD_B200_H
#define COLORD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct clb_ctx clb_ctx;

typedef enum {
	CLB_OK = 0,
	CLB_ERR_NO_DEVICE = 1,     /* no CUDA device / driver: the product has no CPU path                 */
	CLB_ERR_CUDA = 2,          /* a CUDA runtime call or kernel failed (text in clb_last_error)       */
	CLB_ERR_BAD_ARG = 3,
	CLB_ERR_BAD_SYMBOL = 4,    /* input holds a byte outside ACGTN                                      */
	CLB_ERR_STATE = 5,         /* calls made in the wrong order                                        */
	CLB_ERR_CAPACITY = 6       /* caller-provided output buffer too small (needed size is reported)   */
} clb_status;

/* Parameters of stage 1.  Field meaning = CCompressorParams (src/colord/params.h:48) */
typedef struct {
	uint32_t kmer_len;          /* -k; 15..32 (compression.cpp:57-94 picks 20..26)                     */
	uint32_t modulo;            /* -f filterHashModulo: keep k-mers with murmur64(kmer) % modulo == 0  */
	uint32_t min_count;         /* -L minKmerCount (KMC -ci)                                           */
	uint32_t max_count;         /* -H maxKmerCount (KMC -cs; also the cap of a k-mer's read list)     */
	uint32_t max_candidates;    /* -c maxCandidates                                                     */
	uint32_t is_hifi;           /* DataSource::PBHiFi: the graph also returns the shared k-mers        */
	uint64_t expected_bases;    /* sizing hint for the count table (0 = grow on demand)                */
	int32_t  device;            /* CUDA device ordinal                                                 */
} clb_params;

/* What CKmerCounter::GetNReads/GetTotKmers/GetNUniqueCounted (count_kmers.h:32-34) and
 * CKmerFilter::GetTotalKmers (kmer_filter.h:140) report. */
typedef struct {
	uint64_t n_reads;               /* #Total_reads                                                     */
	uint64_t tot_kmers;             /* "#Total no. of k-mers": occurrences passing the modulo filter   */
	uint64_t n_unique;              /* distinct passing k-mers                                          */
	uint64_t n_unique_counted;      /* survivors: min_count <= count                                    */
	uint64_t total_count_filtered;  /* sum of survivor counts saturated at max_count                   */
} clb_kmer_stats;

clb_status clb_create(const clb_params* params, clb_ctx** out);
void       clb_destroy(clb_ctx* ctx);
const char* clb_last_error(const clb_ctx* ctx);          /* ctx may be NULL: error of the failed create  */
/* Device memory is taken from the device's default stream-ordered memory pool and returned to it; what a finished job (or a
 * destroyed context) returned stays cached in the pool for the next one, so that a process compressing file after file does
 * not pay the driver for mapping and unmapping tens of GB per file.  This call hands the cached memory back to the system. */
clb_status clb_release_cached_memory(int device);
/* Run all of this context's work on an existing CUDA stream (cudaStream_t as void*); default: own stream. */
clb_status clb_set_stream(clb_ctx* ctx, void* cuda_stream);
clb_status clb_synchronize(clb_ctx* ctx);

/* ---- Stage 1a: k-mer counting ---------------------------------------------------------------------
 * Replaces the CKmerCounter ctor -> run_filtering_kmc (count_kmers.cpp:28-68, filtering_kmc.h:18).
 * clb_append_reads = one read pack (read_pack_t, utils.h:46): bases are 2-bit packed on device
 * (layout of CReferenceReads, reference_reads.h:35-72, widened to 64-bit words) and kept resident; the
 * canonical k-mers passing the modulo filter are counted.  `bases`/`offsets` are HOST pointers unless
 * on_device != 0 (then both are device pointers valid on the context's stream).  May be called many times. */
clb_status clb_append_reads(clb_ctx* ctx, const uint8_t* bases, const uint64_t* offsets, uint32_t n_reads, int on_device);
/* Raw (k-mer, count) table exchange for the multi-GPU path: each rank counts its shard of
 * reads; k-mers are owned by partition part = owner(kmer) in [0, n_parts) (same function on every rank);
 * rank r exports partition p to rank p (all-to-all), resets, merges what it received for its own
 * partition, finalizes (thresholds its share), and the survivors are all-gathered and imported with
 * clb_filter_import.  n_parts <= 1 exports everything.  Buffers are device pointers iff on_device. */
clb_status clb_counts_size(clb_ctx* ctx, uint32_t part, uint32_t n_parts, uint64_t* n_entries);
clb_status clb_counts_export(clb_ctx* ctx, uint32_t part, uint32_t n_parts, uint64_t* kmers, uint32_t* counts, uint64_t cap, uint64_t* n_entries, int on_device);
/* The same for all partitions in ONE pass over the table each (what the exchanges call): the entries per partition
 * (sizes: host array of n_parts, n_parts <= 64), then every partition written at kmers / counts [first[p] ..) (first: host array
 * of n_parts — the caller's prefix sums of sizes; kmers / counts: DEVICE buffers of cap entries; order inside a partition is free). */
clb_status clb_counts_sizes(clb_ctx* ctx, uint32_t n_parts, uint64_t* sizes);
clb_status clb_counts_export_all(clb_ctx* ctx, uint32_t n_parts, const uint64_t* first, uint64_t* kmers, uint32_t* counts, uint64_t cap);
clb_status clb_counts_reset(clb_ctx* ctx);
clb_status clb_counts_merge(clb_ctx* ctx, const uint64_t* kmers, const uint32_t* counts, uint64_t n, uint64_t n_reads_remote, int on_device);
/* Thresholds (kb_sorter.h:1011-1065) and builds the filtered-k-mer set; replaces the CKmerFilter ctor
 * (filter_kmers.cpp:45-83).  After this call no more reads can be appended. */
clb_status clb_count_finalize(clb_ctx* ctx, clb_kmer_stats* stats);
/* Listing of the filtered set = what CKMCFile::ReadNextKmer yields (filter_kmers.cpp:68), any order.
 * Device buffers iff on_device.  Returns CLB_ERR_CAPACITY (with *n set) if cap < *n. */
clb_status clb_filter_list(clb_ctx* ctx, uint64_t* kmers, uint32_t* counts, uint64_t cap, uint64_t* n, int on_device);
/* Replace the filtered set by a listed one (multi-GPU: the survivors gathered from all owner ranks) and,
 * if global_stats != NULL, the statistics by the all-reduced ones. */
clb_status clb_filter_import(clb_ctx* ctx, const uint64_t* kmers, const uint32_t* counts, uint64_t n, const clb_kmer_stats* global_stats, int on_device);
/* CKmerFilter::Possible && Check for a batch of canonical k-mers (kmer_filter.h:129-137); HOST buffers. */
clb_status clb_filter_check(clb_ctx* ctx, const uint64_t* kmers, uint64_t n, uint8_t* possible, uint8_t* present);

/* Sequences that are counted like reads but are not reads: the reference hands the reference genome (-G) to KMC as a second input
 * file (compression.cpp:408-430), so its k-mers count towards the thresholds and the statistics (n_reads excepted: the caller adds the
 * number of sequences, compression.cpp:445-448).  Before clb_count_finalize.  The genome's pseudo-reads then enter through
 * clb_append_context_reads: always-accepted reference reads in front of the input's reads, never encoded (reads_sim_graph.cpp:295-322). */
clb_status clb_count_sequences(clb_ctx* ctx, const uint8_t* bases, const uint64_t* offsets, uint32_t n_seqs, int on_device);

/* ---- Multi-GPU: the global reference-read set -----------------------------------------
 * Reads shard by id, but the candidates of a read are EARLIER reference reads of the whole input (reads_sim_graph.cpp:374-395),
 * most of which sit at the start of the file in sparse mode (ref_reads_accepter.h:51-57).  Rank r therefore receives the
 * reference reads of the shards before its own as context reads: they take the read ids 0 .. n-1 in front of the rank's own
 * reads, are inserted into the k-mers -> reads table like any reference read (same order, same cap: the table a read of the
 * shard sees is the one the single-GPU run builds), are never queried and never encoded: clb_encode, clb_dna_encode and
 * clb_qual_encode cover only the reads after them (pack_sizes / qualities are given for those), and their tuples / streams are
 * identical to what a single GPU produces for the same reads.  Call once, after clb_count_finalize (context reads are not
 * counted: their k-mers were counted on their own rank) and before clb_graph_build; reads holding N are not reference reads
 * and must not be sent.  bases / offsets as in clb_append_reads. */
clb_status clb_append_context_reads(clb_ctx* ctx, const uint8_t* bases, const uint64_t* offsets, uint32_t n_reads, int on_device);
/* has-N flag of every read in the store (what CInputReads reports per read, utils.h:46); HOST buffer of n_reads bytes. */
clb_status clb_reads_have_n(clb_ctx* ctx, uint8_t* flags);
/* ASCII bases of the listed reads (ids: HOST, store numbering) back to back, e.g. a rank's reference reads for the exchange;
 * `bases` is a device pointer iff on_device.  CLB_ERR_CAPACITY if cap is below the sum of the reads' lengths. */
clb_status clb_reads_export(clb_ctx* ctx, const uint32_t* read_ids, uint32_t n, uint8_t* bases, uint64_t cap, int on_device);

/* ---- Stage 1b: similarity graph -------------------------------------------------------------------
 * Replaces CReadsSimilarityGraph (reads_sim_graph.cpp:530; per pack :324 / HiFi :429) over ALL appended
 * reads at once.  is_reference[i] (HOST, one byte per appended read after the context reads, input order) = the value of acceptRefRead
 * the reference would compute before its hasN test, i.e. the CRefReadsAccepter decision (all ones for
 * -R all); reads holding N are excluded inside.  n_pseudo leading reads are reference-genome
 * pseudo-reads (inserted uncapped, never queried: reads_sim_graph.cpp:295-322). */
clb_status clb_graph_build(clb_ctx* ctx, const uint8_t* is_reference, uint32_t n_pseudo);
/* Accepted k-mers per read (reads_sim_graph.cpp:134-164), CSR; HOST buffers; offsets has n_reads+1. */
clb_status clb_graph_accepted_size(clb_ctx* ctx, uint64_t* total);
clb_status clb_graph_accepted(clb_ctx* ctx, uint64_t* offsets, uint64_t* kmers, uint64_t cap);
/* Candidate reference reads per read (CCompressElem::ref_reads, queues_data.h:23): cand[i*max_candidates+j]
 * for j < cand_n[i], ordered (shared k-mers desc, reference id asc).  HOST buffers. */
clb_status clb_graph_candidates(clb_ctx* ctx, uint32_t* cand, uint32_t* cand_n);
/* HiFi only: CCompressElem::common_kmers.  common_off[i*max_candidates+j] .. + common_n[...] index `kmers`. */
clb_status clb_graph_common_size(clb_ctx* ctx, uint64_t* total);
clb_status clb_graph_common(clb_ctx* ctx, uint64_t* common_off, uint32_t* common_n, uint64_t* kmers, uint64_t cap);

/* ---- Stage 2: edit scripts -----------------------------------------------------------------------
 * Batch form of CEncoder::GetEditDist (encoder.cpp:1255-1283): optimal unit-cost alignment of a part of the read being
 * encoded against a part of the (oriented) reference read with edlib's path (edit_script.h:272-413, libs/edlib) followed
 * by refactor_edit_script (edit_script.h:661).  seqs: symbols 0..3 (HOST); task i aligns seqs[ref_off[i] .. +ref_len[i])
 * with seqs[enc_off[i] .. +enc_len[i]); the byte after each part must be readable and is what follows the part in its read
 * (255 at a read's end) — the reference's canonicalisation looks at it.  kind: 0 left flank (frag 0), 1 right flank (last
 * frag), 2 between anchors.  Scripts (alphabet M D A C G T X Y Z) are written back to back; out_off has n+1 entries. */
clb_status clb_edit_scripts(clb_ctx* ctx, const uint8_t* seqs, uint64_t n_seq_bytes, const uint64_t* ref_off, const uint32_t* ref_len,
	const uint64_t* enc_off, const uint32_t* enc_len, const uint32_t* kind, uint64_t n, uint64_t* out_off, char* out, uint64_t cap);

/* ---- Stage 2: the per-read encoder -------------------------------------------------------------------
 * Replaces the N CEncoder threads (CEncoder::Encode, encoder.cpp:1672-1691; ctor arguments encoder.h:371-410) over ALL
 * appended reads: m-mer anchors against the candidates (encoder.cpp:1058-1111), edit scripts of the parts between anchors
 * (:1255-1283), edit-script / plain / alternative-read decisions (:1315-1346, utils.h:1060-1126) and the CompactES tuple
 * bytes (:1513-1575, utils.h:69-273).  Field meaning = CCompressorParams (params.h:48). */
typedef struct {
	uint32_t anchor_len;            /* -a                                                                   */
	uint32_t min_part_len_alt;      /* minPartLenToConsiderAltRead                                          */
	uint32_t max_recurence;         /* maxRecurence                                                         */
	uint32_t min_anchors;           /* minAnchors                                                           */
	double min_mmer_frac;           /* minFractionOfMmersInEncode                                           */
	double min_mmer_force;          /* minFractionOfMmersInEncodeToAlwaysEncode                             */
	double max_matches_mult;        /* maxMatchesMultiplier                                                 */
	double es_cost_mult;            /* editScriptCostMultiplier                                             */
} clb_encode_params;
/* pack_sizes[n_packs] (HOST): number of reads of every read pack, in input order — CEntropyEstimator is reset at every
 * pack boundary (encoder.cpp:1677), so the packs are part of the result.  NULL: packs are cut by the reference's rule
 * (a pack closes when its reads hold >= 4 MiB including one guard byte per read; in_reads.cpp:62-76, defs.h:45).
 * Requires clb_graph_build.  The result stays on the device (stage 3 consumes it there). */
clb_status clb_encode(clb_ctx* ctx, const clb_encode_params* params, const uint32_t* pack_sizes, uint32_t n_packs);
clb_status clb_encode_size(clb_ctx* ctx, uint64_t* total_bytes);
/* es_t of every read back to back (what CEncoder pushes to compressed_queue, encoder.cpp:1681-1687); es_off has
 * n_reads + 1 entries.  Buffers are device pointers iff on_device. */
clb_status clb_encode_get(clb_ctx* ctx, uint64_t* es_off, uint8_t* es, uint64_t cap, int on_device);
/* Parity tap: the candidates of every read after anchor search (Candidate, encoder.h:46), best first.  Call
 * clb_encode_keep_candidates(ctx, 1) before clb_encode.  cand_off[n_reads + 1] indexes `data` (uint32 words): per candidate
 * ref_id, shouldReverse, tot_anchor_len, n_anchors, then n_anchors * (len, pos_enc, pos_ref).  HOST buffers. */
clb_status clb_encode_keep_candidates(clb_ctx* ctx, int on);

/* The counters of the reference's `-v` report (stats_collector.h:28-75, logged in encoder.cpp:663-676, :1069-1190, :1445-1575),
 * summed over all reads: call clb_encode_stats_enable(ctx, 1) before clb_encode, read them after it.  Levels = recursion depth of
 * AddEncodedReadWithCandidates (level 0 = the main reference read). */
#define CLB_MAX_STAT_LEVELS 8
typedef struct clb_level_stats {
	uint64_t n_alternative_left_flank, n_alternative_in_between, n_alternative_right_flank;
	uint64_t n_plain_symbols, n_symb_coded_with_edit_script, n_edit_script_symbols;
	uint64_t n_substitution, n_match, n_insertion, n_deletion;
	uint64_t n_symb_anchors, n_anchors, n_left_flank_symb, n_right_flank_symb;
} clb_level_stats;
typedef struct clb_encode_stats {
	uint64_t n_not_enough_unique_mmers_in_enc_read, n_too_many_matches, n_too_low_anchors;
	uint64_t n_non_rev_choosen, n_rev_choosen;
	uint64_t n_plain_reads_tot, n_plain_symb, n_plain_reads_with_n_tot, n_plain_with_n_symb;
	uint32_t n_levels, pad;
	clb_level_stats level[CLB_MAX_STAT_LEVELS];
} clb_encode_stats;
clb_status clb_encode_stats_enable(clb_ctx* ctx, int on);
clb_status clb_encode_stats_get(clb_ctx* ctx, clb_encode_stats* out);
clb_status clb_encode_candidates_size(clb_ctx* ctx, uint64_t* n_words);
clb_status clb_encode_candidates(clb_ctx* ctx, uint64_t* cand_off, uint32_t* data, uint64_t cap_words);

/* ---- Stage 3: DNA / edit-script stream --------------------------------------------------------------
 * Replaces CEntrComprReads::Compress -> CDNACoder::Encode (entr_read.h:56-80, dna_coder.cpp:26-231) over the tuples clb_encode
 * left on the device.  The reference's event model is kept (which symbols a read's tuples turn into and the context of each one:
 * read flag, length, plain symbols, reference ids, reverse-complement flags, tuple types under the tuple / symbol histories +
 * reference base + indel drift, anchor / skip lengths, insertions, substitutions, alternative reads — dna_coder.cpp:440-1239),
 * and so is the arithmetic of its range coder (sub_rc.h:83-201); the adaptive models running through the whole file are
 * replaced by static per-context tables (two passes) and 64 independent coder lanes per read pack.  Native container "DB01";
 * decoder (rebuilds the reads): oracle/stage3_dna.c.  level = compressionLevel (1..3: history widths, dna_coder.cpp:1253-1280). */
clb_status clb_dna_encode(clb_ctx* ctx, uint32_t level, const uint32_t* pack_sizes, uint32_t n_packs);
/* header_bytes (may be NULL): the part of the container that holds the frequency tables */
clb_status clb_dna_size(clb_ctx* ctx, uint64_t* total_bytes, uint64_t* header_bytes);
clb_status clb_dna_get(clb_ctx* ctx, uint8_t* stream, uint64_t cap, int on_device);

/* ---- Stage 3: quality stream ------------------------------------------------------------------------
 * Replaces CEntrComprQuals::Compress -> CQualityCoder::Encode (entr_qual.h:100-126, quality_coder.cpp:560) for the "*-avg"
 * modes (ONT default 4-avg, HiFi default 5-avg, 2-avg): the reference's lossy transform and context model
 * (quality_coder_impl.cpp:130-310: bins by the forward thresholds, per-read per-bin means, one bin symbol per base under the
 * previous symbols + neighbouring bases [+ match / anchor flags from the tuples when level > 1]) are kept, so the qualities a
 * decoder reconstructs are the reference's (test/<name>.quan).  The reference's single adaptive range-coder chain is replaced
 * by static per-context tables + 64 interleaved rANS lanes per read pack (native container "QB01"; layout, CPU twin and
 * decoder in oracle/stage3_qual.c; SURVEY.md §7 explains why the reference's byte stream cannot be produced in lockstep). */
typedef struct {
	uint32_t n_bins;                /* 2, 4 or 5 (QualityComprMode Binary/Quad/QuinaryAverage, params.h:37)             */
	uint32_t thresholds[4];         /* qualityFwdThresholds (-T): bin = number of thresholds <= phred; n_bins - 1 used  */
	uint32_t level;                 /* compressionLevel: > 1 puts the match / anchor flags in the context (needs clb_encode) */
} clb_qual_params;
/* Qualities kept on the device as the input streams in: the ASCII quality bytes of the reads of the clb_append_reads call just made, in the
 * same order (the reader's quals_pack follows its reads pack: in_reads.cpp:103-111).  HOST buffer unless on_device; it may be reused
 * when the call returns.  The quality encoders below then take quals == NULL ("the resident ones"). */
clb_status clb_append_quals(clb_ctx* ctx, const uint8_t* quals, uint64_t n_bytes, int on_device);
/* quals: phred+33 bytes of all appended reads, read r at quals[offsets[r] .. offsets[r+1]) with the reads' lengths
 * (device pointers iff on_device).  pack_sizes as for clb_encode (NULL: the reference's pack rule).  Result stays on the device. */
clb_status clb_qual_encode(clb_ctx* ctx, const clb_qual_params* params, const uint8_t* quals, const uint64_t* offsets, int on_device,
	const uint32_t* pack_sizes, uint32_t n_packs);
/* Lossless mode, "-q org" (QualityComprMode::Original; CQualityCoder::encode_original, quality_coder_impl.cpp:78-128): one
 * 96-symbol phred value per base under [the two previous values quantised to 4 bits by the data source's table | bases i, i-1,
 * (i-2), i+1 | match / anchor flags at level > 1] — the reference's context model; static tables (rare contexts fall back to
 * the two previous values) + 64 range-coder lanes per pack instead of its adaptive chain.  Native container "QO01"; CPU twin
 * and decoder: oracle/stage3_qorg.c.  source: 0 ONT, 1 PacBio CLR, 2 PacBio HiFi (DataSource, params.h:30 — selects the
 * quantiser, quality_coder.cpp:272-505); level = compressionLevel 1..3.  Other arguments as clb_qual_encode; the result is
 * fetched with clb_qual_size / clb_qual_get (one quality stream per context). */
clb_status clb_qual_encode_original(clb_ctx* ctx, uint32_t source, uint32_t level, const uint8_t* quals, const uint64_t* offsets, int on_device,
	const uint32_t* pack_sizes, uint32_t n_packs);
clb_status clb_qual_size(clb_ctx* ctx, uint64_t* total_bytes);
clb_status clb_qual_get(clb_ctx* ctx, uint8_t* stream, uint64_t cap, int on_device);

/* ---- Stage 3: header (read id) stream ---------------------------------------------------------------
 * Replaces CEntrComprHeaders::Compress -> CIDCoder::Encode / compress_lossless (entr_header.cpp:23-46, id_coder.cpp:210-383;
 * HeaderComprMode::Original, the default).  The reference's event model is kept (tokens cut at every character outside
 * [0-9A-Za-z@], the same-shape flag against the previous header, per token same / same length / the characters that differ,
 * plain fallback) and so is the arithmetic of its range coder; its adaptive models are replaced by static per-context tables
 * (two passes) and 64 independent coder lanes per pack; the first header of a pack has no predecessor.  Native container
 * "HB01"; CPU twin and decoder: oracle/stage3_hdr.c.  Independent of the other stages (may run before any read is appended).
 * bytes: the headers back to back exactly as they are to be restored (no NUL bytes), header r at bytes[offsets[r] ..
 * offsets[r+1]) with offsets[0] = 0; plus_id[r] != 0: the read's '+' line repeats the header (qual_header_type::eq_read_header,
 * entr_header.cpp:34; NULL = never).  Device pointers iff on_device.  pack_sizes (HOST; NULL: packs of 4096 headers). */
clb_status clb_hdr_encode(clb_ctx* ctx, const uint8_t* bytes, const uint64_t* offsets, const uint8_t* plus_id, uint64_t n_headers, int on_device,
	const uint32_t* pack_sizes, uint32_t n_packs);
/* table_bytes (may be NULL): the part of the container that holds the frequency tables */
clb_status clb_hdr_size(clb_ctx* ctx, uint64_t* total_bytes, uint64_t* table_bytes);
clb_status clb_hdr_get(clb_ctx* ctx, uint8_t* stream, uint64_t cap, int on_device);

/* ---- Stage 3, compat streams: the reference's own byte streams ------------------------------------
 * The same three seams (CEntrComprReads / CEntrComprQuals / CEntrComprHeaders::Compress -> CDNACoder / CQualityCoder / CIDCoder::Encode:
 * entr_read.h:56-80, entr_qual.h:100-126, entr_header.cpp:23-46) with the reference's OWN output: the parts of the "dna", "qual" and
 * "header" streams of a reference archive, byte for byte (adaptive models rc.h, range coder sub_rc.h; one part per pack, coder
 * restarted per pack, models kept).  An archive holding these parts is read by the unmodified `colord decompress`.  Every quality
 * mode of the reference is covered: mode = params.h QualityComprMode (0 org, 1 5-avg, 2 4-avg, 3 2-avg, 4 5-fix, 5 4-fix, 6 2-fix,
 * 7 avg, 8 none), source = DataSource (0 ONT, 1 PBRaw, 2 PBHiFi; the lossless quantiser), thresholds = the forward thresholds of the
 * binned modes (arg_parse.cpp:28-84).  pack_sizes: reads (headers) per part; NULL = the reference's pack rule (headers: one part).
 * Meant for inputs up to a few Gbases per GPU (about 30 bytes of device memory per coded symbol while a stream is built);
 * CLB_ERR_CAPACITY beyond 2^32 symbols per stream.  Results stay on the device until clb_xstream_get. */
clb_status clb_xdna_encode(clb_ctx* ctx, uint32_t level, const uint32_t* pack_sizes, uint32_t n_packs);
clb_status clb_xqual_encode(clb_ctx* ctx, uint32_t mode, uint32_t source, uint32_t level, const uint32_t* thresholds, const uint8_t* quals, const uint64_t* offsets,
                            int on_device, const uint32_t* pack_sizes, uint32_t n_packs);
clb_status clb_xhdr_encode(clb_ctx* ctx, const uint8_t* bytes, const uint64_t* offsets, const uint8_t* plus_id, uint64_t n_headers, int on_device,
                           const uint32_t* pack_sizes, uint32_t n_packs);
/* The stored reference genome (CReferenceGenome::Store(archive), reference_genome.cpp:319-360): the sequences (HOST: ASCII ACGT back to back
 * + offsets[n + 1]) coded as plain reads by a DNA coder of their own at `level` (the reference passes 9: one symbol of history); one part,
 * fetched as stream 3 of clb_xstream_get.  One range coder over the whole genome is one serial chain: meant for genomes of megabases. */
clb_status clb_xplain_encode(clb_ctx* ctx, const uint8_t* bases, const uint64_t* offsets, uint32_t n_seqs, uint32_t level);
/* which: 0 dna, 1 qual, 2 header, 3 ref-genome.  The parts lie back to back in `bytes`; part_sizes[n_parts] (HOST) receives their lengths. */
clb_status clb_xstream_size(clb_ctx* ctx, uint32_t which, uint64_t* total_bytes, uint32_t* n_parts);
clb_status clb_xstream_get(clb_ctx* ctx, uint32_t which, uint8_t* bytes, uint64_t cap, uint64_t* part_sizes, int on_device);

/* ---- Reference-read store (CReferenceReads, reference_reads.h:27) ---------------------------------
 * Read i of the appended input in the reference's byte layout (4 bases/byte MSB first + trailer byte).
 * HOST buffer of (len+3)/4+1 bytes; used by parity tests and by a host-side decoder. */
clb_status clb_get_packed_read(clb_ctx* ctx, uint32_t read_id, uint8_t* out, uint64_t cap, uint64_t* n_bytes);

/* ---- Sparse reference sampler (CRefReadsAccepter, ref_reads_accepter.h:27-57) ----------------------
 * Host-side, serial by construction (default-seeded std::mt19937 stream is part of the format).
 * decisions[i] = ShouldAddToReference(i) for i in [0, n). */
void clb_sampler(uint32_t range, double exponent, uint32_t n_pseudo, uint32_t n, uint8_t* decisions);

/* ---- Page-locked host memory --------------------------------------------------------------------
 * Buffers a host hands to clb_append_reads / clb_append_quals are copied at the full host-to-device rate (and asynchronously) only
 * when they are page-locked; the streaming reader of the command line takes its piece buffers from here.  NULL on failure. */
void* clb_host_alloc(uint64_t bytes);
void  clb_host_free(void* p);

/* ---- Instrumentation -------------------------------------------------------------------------------
 * Number of kernels this context has launched so far (bench.py's gpu_launches). */
uint64_t clb_kernel_launches(const clb_ctx* ctx);
/* Optional per-kernel device timing with CUDA events on the context's stream (off by default; enabling
 * resets the accumulators).  Kernel classes: k_pack, k_count, k_tab_misc, k_finalize, k_accept, k_postings,
 * k_vote, k_common, k_misc, k_align, k_anchors, k_encode (task lists), k_decide, k_estimate, k_emit, k_qual, k_dna, k_hdr.  clb_profile_get synchronizes the stream. */
clb_status clb_profile_enable(clb_ctx* ctx, int on);
clb_status clb_profile_get(clb_ctx* ctx, const char* kernel, double* ms, uint64_t* launches);

#ifdef __cplusplus
}
#endif
#endif
