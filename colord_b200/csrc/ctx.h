// ctx.h — the library context: parameters, the resident read store and every stage's device state.
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include <atomic>
#include <mutex>
#include <unordered_set>
#include <algorithm>
#include <cuda_runtime.h>
#include "../../include/colord_b200.h"
#include "util.cuh"
#include "stage2.h"
#include "slab.h"

namespace clb {

// Device memory.  Measured on a B200 (gpurun_out/tools/alloc_bench.cu): cudaMalloc / cudaFree of a 40 GiB block take 4 / 14 ms,
// while the stream-ordered pool needs 0.4-0.8 s to map a fresh block of that size (and, near the capacity of the device,
// seconds when it has to trim fragmented cached blocks first).  So: blocks of 32 MiB and more come straight from cudaMalloc
// and go back with cudaFree; small scratch comes from the device's default stream-ordered pool, which keeps it cached
// (clb_release_cached_memory hands it back).
constexpr uint64_t DEV_BIG_BYTES = 32ull << 20;
struct BigPtrs { std::mutex m; std::unordered_set<void*> v; };
inline BigPtrs& big_ptrs() { static BigPtrs b; return b; }
inline int dev_current() { int d = 0; cudaGetDevice(&d); return d; }
inline cudaError_t dev_malloc(void** p, uint64_t bytes, cudaStream_t s)
{
	if (bytes < DEV_BIG_BYTES) return cudaMallocAsync(p, bytes ? bytes : 1, s);
	Slab& slab = job_slab(dev_current());
	if (slab.active()) {      // slab.h: large blocks are cut from the device's slab, the driver is not called
		const uint64_t at = slab.alloc(bytes);
		if (at) { *p = reinterpret_cast<void*>(at); return cudaSuccess; }
	}
	const cudaError_t e = cudaMalloc(p, bytes);
	if (e == cudaSuccess) { BigPtrs& b = big_ptrs(); std::lock_guard<std::mutex> g(b.m); b.v.insert(*p); }
	return e;
}
// what a single large block may ask for: the slab's largest free range, or the driver's free memory without a slab
inline uint64_t dev_mem_available()
{
	size_t free_b = 0, total_b = 0;
	if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); free_b = 0; }
	const Slab& slab = job_slab(dev_current());
	return slab.active() ? std::max<uint64_t>(slab.largest_free(), free_b) : (uint64_t)free_b;
}
inline bool dev_is_big(void* p) { BigPtrs& b = big_ptrs(); std::lock_guard<std::mutex> g(b.m); return b.v.erase(p) != 0; }
// semantics of cudaFree: everything the device was doing is finished before the memory is reused
inline bool slab_free(void* p)
{
	const uint64_t a = reinterpret_cast<uint64_t>(p);
	for (int d = 0; d < SLAB_MAX_DEVICES; ++d) if (job_slab(d).owns(a)) { cudaDeviceSynchronize(); job_slab(d).free(a); return true; }
	return false;
}
inline void dev_free(void* p, cudaStream_t s) { if (!p) return; if (slab_free(p)) return; if (dev_is_big(p)) { cudaFree(p); return; } cudaDeviceSynchronize(); cudaFreeAsync(p, s); }
// stream-ordered free of scratch (a large block is freed by cudaFree, which waits for the device by itself)
inline cudaError_t dev_free_async(void* p, cudaStream_t s) { if (!p) return cudaSuccess; if (slab_free(p)) return cudaSuccess; if (dev_is_big(p)) return cudaFree(p); return cudaFreeAsync(p, s); }

// A growable device array, grown geometrically on the ctx stream.
template <typename T>
struct DevBuf {
	T* p = nullptr;
	uint64_t cap = 0;          // elements
	cudaStream_t st = nullptr;
	cudaError_t reserve(uint64_t n, cudaStream_t s, bool keep, uint64_t used = 0)
	{
		st = s;
		if (n <= cap) return cudaSuccess;
		uint64_t ncap = cap ? cap : 1;
		while (ncap < n) ncap = ncap + ncap / 2 + 1024;
		T* q = nullptr;
		cudaError_t e = dev_malloc((void**)&q, ncap * sizeof(T), s);
		if (e != cudaSuccess) return e;
		if (keep && p && used) {
			e = cudaMemcpyAsync(q, p, used * sizeof(T), cudaMemcpyDeviceToDevice, s);
			if (e != cudaSuccess) { dev_free_async(q, s); return e; }
		}
		if (p) dev_free(p, s);
		p = q; cap = ncap;
		return cudaSuccess;
	}
	void release() { if (p) dev_free(p, st); p = nullptr; cap = 0; }
};

struct CountSlot { uint64_t key; uint32_t cnt; uint32_t pad; };   // 16 B: two slots per 32 B sector

// kernel classes for the optional per-kernel timing (clb_profile_*)
enum KernelId : int { K_PACK = 0, K_COUNT, K_TAB_MISC, K_FINALIZE, K_ACCEPT, K_POSTINGS, K_VOTE, K_COMMON, K_MISC, K_ALIGN, K_ANCHORS, K_ENCODE, K_DECIDE, K_ESTIMATE, K_EMIT, K_QUAL, K_DNA, K_HDR, K_N };
static const char* const kernel_names[K_N] = { "k_pack", "k_count", "k_tab_misc", "k_finalize", "k_accept", "k_postings", "k_vote", "k_common", "k_misc", "k_align", "k_anchors", "k_encode", "k_decide", "k_estimate", "k_emit", "k_qual", "k_dna", "k_hdr" };
struct ProfRec { int kid; cudaEvent_t a, b; };

} // namespace clb

struct clb_ctx {
	clb_params prm{};
	clb::ModTest mt{};
	cudaStream_t stream = nullptr;
	bool own_stream = false;
	std::string err;
	std::atomic<uint64_t> launches{0};
	int n_sm = 148;
	// per-kernel device timing (off by default)
	bool prof_on = false;
	std::vector<clb::ProfRec> prof_open;
	double prof_ms[clb::K_N] = {0};
	uint64_t prof_n[clb::K_N] = {0};
	// The quality and header streams of stage 3 do not depend on stage 2 (quality: at level 1) and have their own stream and
	// profiling list, so that a second host thread can code them while stage 2 runs — the reference runs its quality and header
	// coders in threads of their own, too (compression.cpp:654-689).
	cudaStream_t stream3 = nullptr;
	std::vector<clb::ProfRec> prof_open3;
	cudaStream_t copy_stream = nullptr;      // H2D staging of host input overlaps the kernels
	cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_consumed[2] = {nullptr, nullptr};

	// ---- resident read store (device) ----
	clb::DevBuf<uint64_t> pk;        // packed bases, 32 per word
	clb::DevBuf<uint32_t> nmask;     // bit p&31 of word p>>5: position p is N / padding
	clb::DevBuf<uint32_t> smask;     // bit set: position p is the first base of a read
	clb::DevBuf<uint64_t> rd_start;  // per read: first position in the device stream
	clb::DevBuf<uint32_t> rd_len;    // per read: length in bases
	uint64_t n_pos = 0;              // positions used in the stream (multiple of 128)
	uint64_t n_reads = 0;
	uint64_t n_context = 0;          // leading reads that are context only: reference reads of earlier shards (never queried, not encoded)
	uint64_t n_reads_remote = 0;     // reads counted on other ranks (merged tables)
	uint64_t n_bases = 0;
	std::vector<uint64_t> h_rd_start; // host mirrors (small: 12 B per read)
	std::vector<uint32_t> h_rd_len;
	clb::DevBuf<uint8_t> dq;            // qualities kept on the device as the input streams in (clb_append_quals), read order
	uint64_t dq_n = 0;
	clb::DevBuf<uint8_t> stage_in[2];   // double-buffered staging for host ASCII input
	clb::DevBuf<uint64_t> stage_off;

	// ---- stage 1a: count table ----
	clb::CountSlot* tab = nullptr;
	uint32_t tab_log2 = 0;
	unsigned long long* d_scal = nullptr;   // device scalars, see enum below
	bool finalized = false;
	clb_kmer_stats stats{};
	struct FillState { uint64_t used_known = 0, maybe_new = 0; } fill;   // host-side bound on the table fill

	// ---- filtered set (survivors) ----
	uint64_t* sv_keys = nullptr;     // open addressing, EMPTY64
	uint32_t* sv_ids = nullptr;      // dense id per slot
	uint32_t sv_log2 = 0;
	uint64_t n_surv = 0;
	uint64_t sum_true = 0;           // sum of unsaturated counts of survivors (bounds accepted k-mers)
	uint64_t* sv_kmer = nullptr;     // dense: k-mer, saturated count
	uint32_t* sv_count = nullptr;

	// ---- stage 1b ----
	bool graph_done = false;
	uint8_t* d_has_n = nullptr;
	std::vector<uint8_t> h_has_n, h_is_ref;
	std::vector<uint32_t> h_ref_before;
	uint32_t* d_ref_before = nullptr;
	uint8_t* d_is_ref = nullptr;
	uint32_t n_ref = 0;
	uint64_t* acc_start = nullptr;   // per read: start in acc_id
	uint32_t* acc_n = nullptr;
	uint32_t* acc_id = nullptr;      // arena of dense survivor ids, read order inside a read
	uint64_t acc_cap = 0, acc_total = 0;
	uint32_t* post_cnt = nullptr;    // per survivor: number of reference reads holding it (then min(.,H))
	uint64_t* post_off = nullptr;    // exclusive scan of the uncapped counts
	uint32_t* post = nullptr;        // reference ids
	uint64_t post_total = 0;
	uint32_t* cand = nullptr;        // n_reads * max_candidates
	uint32_t* cand_votes = nullptr;
	uint32_t* cand_n = nullptr;
	uint64_t* common_off = nullptr;  // HiFi
	uint64_t* common = nullptr;
	uint64_t common_total = 0;

	// ---- stage 2 ----
	bool enc_done = false;
	clb::DevBuf<uint8_t> es;         // CompactES bytes of all reads, input order
	uint64_t* es_off = nullptr;      // n_reads + 1
	uint64_t es_total = 0;
	uint32_t* d_ref_to_read = nullptr;
	clb::DevBuf<uint8_t> s2_arena;   // pair / anchor arena of the current batch
	clb::DevBuf<uint8_t> s2_store;   // anchors of the chosen candidates of all reads
	clb::DevBuf<uint8_t> s2_scratch; // alignment scratch of the current waves
	uint64_t s2_budget = 0;          // its size for this job (set by the first level from the free memory)
	// per-batch scratch of the anchor search, kept from batch to batch (measured: allocating and freeing these blocks in every
	// batch left 3.5 s of a 13 s step to the driver — single batches of 0.3-1.2 s instead of 0.07 s)
	clb::DevBuf<uint8_t> s2_segs; clb::DevBuf<uint32_t> s2_gtab, s2_gbloom;
	clb::DevBuf<clb::Node> s2_nodes; clb::DevBuf<clb::CandView> s2_cviews;   // batch state, kept across batches (no per-batch malloc)
	clb::DevBuf<clb::Task> s2_tasks; clb::DevBuf<char> s2_esbuf;
	cudaStream_t s2_streams[16] = {};   // alignment bins run concurrently
	cudaEvent_t s2_fork = nullptr, s2_join[16] = {};
	// ---- stage 3: quality stream ----
	bool qual_done = false;
	clb::DevBuf<uint8_t> qs;         // native quality container
	uint64_t qs_total = 0;
	// ---- stage 3: DNA / edit-script stream ----
	bool dna_done = false;
	clb::DevBuf<uint8_t> ds;         // native DNA container
	uint64_t ds_total = 0, ds_header = 0;
	// ---- stage 3: header stream ----
	bool hdr_done = false;
	clb::DevBuf<uint8_t> hs;         // native header container
	uint64_t hs_total = 0, hs_header = 0;
	// ---- stage 3, compat streams (stage3_exact.cu): the reference's own parts, back to back ----
	clb::DevBuf<uint8_t> xd, xq, xh, xg;                     // dna, qual, header, ref-genome
	std::vector<uint64_t> xd_parts, xq_parts, xh_parts, xg_parts;      // bytes of every part
	uint64_t xg_total = 0;
	std::vector<uint64_t> xd_packs, xh_packs;                // reads / headers of every part (the parts' metadata)
	uint64_t xd_total = 0, xq_total = 0, xh_total = 0;
	// debugging / parity taps: candidates after E4 of every read (filled when keep_candidates is set)
	bool keep_candidates = false;
	// the reference's -v counters (clb_encode_stats_enable): device array of ST_COUNT sums, filled by k_select / k_stats_*
	bool collect_stats = false;
	unsigned long long* d_stats = nullptr;
	clb_encode_stats h_stats{};
	std::vector<std::vector<uint32_t>> dbg_cand;   // per read, per candidate: ref_id, rev, tot, n_anchors, then n_anchors * (len, pos_enc, pos_ref)
};

namespace clb {

enum Scalar : int {
	SC_TOT_KMERS = 0, SC_N_UNIQUE, SC_N_SURV, SC_TOT_FILTERED, SC_SUM_TRUE, SC_OVERFLOW, SC_BAD_SYMBOL,
	SC_CURSOR, SC_CURSOR2, SC_TAB_USED, SC_COUNT
};

// error plumbing
clb_status fail(clb_ctx* c, clb_status st, const std::string& msg);
clb_status cuda_fail(clb_ctx* c, cudaError_t e, const char* what);
#define CLB_CUDA(ctx, call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return clb::cuda_fail(ctx, e__, #call); } while (0)
#define CLB_LAUNCH_CHECK(ctx, name) do { ++(ctx)->launches; cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return clb::cuda_fail(ctx, e__, name); } while (0)

// per-kernel timing: CLB_TIMED(ctx, K_COUNT, kernel<<<...>>>(...));
void prof_begin(clb_ctx* c, int kid);
void prof_end(clb_ctx* c);
void prof_resolve(clb_ctx* c);
#define CLB_TIMED(ctx, kid, ...) do { clb::prof_begin(ctx, kid); __VA_ARGS__; clb::prof_end(ctx); } while (0)
void prof_begin3(clb_ctx* c, int kid);      // the same on stream3 (kernel classes k_qual, k_hdr)
void prof_end3(clb_ctx* c);
#define CLB_TIMED3(ctx, kid, ...) do { clb::prof_begin3(ctx, kid); __VA_ARGS__; clb::prof_end3(ctx); } while (0)

// stage entry points implemented in the .cu files
clb_status s1a_init(clb_ctx* c);
clb_status s1a_append(clb_ctx* c, const uint8_t* bases, const uint64_t* offsets, uint32_t n_reads, int on_device, bool context = false, bool count_only = false);
clb_status s1b_reads_have_n(clb_ctx* c, uint8_t* flags);
clb_status s1b_reads_export(clb_ctx* c, const uint32_t* read_ids, uint32_t n, uint8_t* bases, uint64_t cap, int on_device);
clb_status s1a_counts_size(clb_ctx* c, uint32_t part, uint32_t n_parts, uint64_t* n);
clb_status s1a_counts_export(clb_ctx* c, uint32_t part, uint32_t n_parts, uint64_t* kmers, uint32_t* counts, uint64_t cap, uint64_t* n_out, int on_device);
clb_status s1a_counts_sizes(clb_ctx* c, uint32_t n_parts, uint64_t* sizes);
clb_status s1a_counts_export_all(clb_ctx* c, uint32_t n_parts, const uint64_t* first, uint64_t* kmers, uint32_t* counts, uint64_t cap_out);
clb_status s1a_counts_reset(clb_ctx* c);
clb_status s1a_filter_import(clb_ctx* c, const uint64_t* kmers, const uint32_t* counts, uint64_t n, const clb_kmer_stats* global_stats, int on_device);
clb_status s1a_counts_merge(clb_ctx* c, const uint64_t* kmers, const uint32_t* counts, uint64_t n, uint64_t n_reads_remote, int on_device);
clb_status s1a_finalize(clb_ctx* c, clb_kmer_stats* stats);
clb_status s1a_filter_check(clb_ctx* c, const uint64_t* kmers, uint64_t n, uint8_t* possible, uint8_t* present);
clb_status s1b_build(clb_ctx* c, const uint8_t* is_reference, uint32_t n_pseudo);
void s1_free(clb_ctx* c);
clb_status exclusive_scan(clb_ctx* c, const uint32_t* in, uint64_t n, uint64_t* out, uint64_t* total);   // stage1b.cu; synchronizes
clb_status s2_encode(clb_ctx* c, const clb_encode_params* prm, const uint32_t* pack_sizes, uint32_t n_packs);
void s2_free(clb_ctx* c);
clb_status s3_dna_encode(clb_ctx* c, uint32_t level, const uint32_t* pack_sizes, uint32_t n_packs);
clb_status s3_hdr_encode(clb_ctx* c, const uint8_t* bytes, const uint64_t* offsets, const uint8_t* plus_id, uint64_t n, int on_device,
	const uint32_t* pack_sizes, uint32_t n_packs);
clb_status s3_qual_encode(clb_ctx* c, const clb_qual_params* prm, const uint8_t* quals, const uint64_t* offsets, int on_device,
	const uint32_t* pack_sizes, uint32_t n_packs);
clb_status s3_qual_flags(clb_ctx* c, const uint64_t* d_qoff, uint32_t n, uint8_t* d_flags);
clb_status s3_qual_encode_original(clb_ctx* c, uint32_t source, uint32_t level, const uint8_t* quals, const uint64_t* offsets, int on_device,
	const uint32_t* pack_sizes, uint32_t n_packs);
// qualities of the non-context reads: the caller's (host or device) or, with quals == NULL, the resident ones of clb_append_quals.
// h_off[n + 1] receives the offsets on the host, d_q the device pointer of the first quality (staged through `tmp_alloc` if needed).
clb_status resolve_quals(clb_ctx* c, const uint8_t* quals, const uint64_t* offsets, int on_device, cudaStream_t s, std::vector<uint64_t>& h_off, bool& resident);
clb_status s3x_dna_encode(clb_ctx* c, uint32_t level, const uint32_t* pack_sizes, uint32_t n_packs);
clb_status s3x_qual_encode(clb_ctx* c, uint32_t mode, uint32_t source, uint32_t level, const uint32_t* thr, const uint8_t* quals, const uint64_t* offsets, int on_device,
	const uint32_t* pack_sizes, uint32_t n_packs);
clb_status s3x_plain_encode(clb_ctx* c, const uint8_t* bases, const uint64_t* offsets, uint32_t n_seqs, uint32_t level);
clb_status s3x_hdr_encode(clb_ctx* c, const uint8_t* bytes, const uint64_t* offsets, const uint8_t* plus_id, uint64_t n, int on_device,
	const uint32_t* pack_sizes, uint32_t n_packs);
clb_status s2_edit_scripts(clb_ctx* c, const uint8_t* seqs, uint64_t n_seq_bytes, const uint64_t* ref_off, const uint32_t* ref_len,
	const uint64_t* enc_off, const uint32_t* enc_len, const uint32_t* kind, uint64_t n, uint64_t* out_off, char* out, uint64_t cap);

} // namespace clb
