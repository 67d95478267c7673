// api.cu — the extern "C" boundary (include/colord_b200.h) over the stage implementations.
#include "ctx.h"
#include <cstdlib>
#include <cstring>
#include <random>
#include <cmath>

static thread_local std::string g_create_error;

namespace clb {
clb_status fail(clb_ctx* c, clb_status st, const std::string& msg) { if (c) c->err = msg; else g_create_error = msg; return st; }
clb_status cuda_fail(clb_ctx* c, cudaError_t e, const char* what)
{
	std::string m = std::string(what) + ": " + cudaGetErrorString(e);
	cudaGetLastError();
	return fail(c, (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? CLB_ERR_NO_DEVICE : CLB_ERR_CUDA, m);
}
clb_status s1a_counts_size(clb_ctx* c, uint32_t part, uint32_t n_parts, uint64_t* n);
clb_status s1a_counts_export(clb_ctx* c, uint32_t part, uint32_t n_parts, uint64_t* kmers, uint32_t* counts, uint64_t cap_out, uint64_t* n_out, int on_device);
clb_status s1a_counts_reset(clb_ctx* c);
clb_status s1a_filter_import(clb_ctx* c, const uint64_t* kmers, const uint32_t* counts, uint64_t n, const clb_kmer_stats* gs, int on_device);
}
using namespace clb;

// slab.h: the first context on a device takes the slab, the last one gives it back (unless CLB_SLAB_GB asked for a fixed one)
namespace {
struct SlabRefs { std::mutex m; int refs[clb::SLAB_MAX_DEVICES] = {}; bool fixed[clb::SLAB_MAX_DEVICES] = {}; };
SlabRefs& slab_refs() { static SlabRefs r; return r; }
void slab_acquire(int dev)
{
	SlabRefs& r = slab_refs();
	std::lock_guard<std::mutex> g(r.m);
	if (dev < 0 || dev >= clb::SLAB_MAX_DEVICES) return;
	if (r.refs[dev]++ != 0 || clb::job_slab(dev).active()) return;
	uint64_t want = 0;
	if (const char* gb = std::getenv("CLB_SLAB_GB")) { want = static_cast<uint64_t>(std::atof(gb) * 1073741824.0); r.fixed[dev] = true; }
	else {
		const char* rs = std::getenv("CLB_SLAB_RESERVE_GB");
		const uint64_t reserve = static_cast<uint64_t>((rs ? std::atof(rs) : 8.0) * 1073741824.0);
		size_t free_b = 0, total_b = 0;
		if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && free_b > reserve + (4ull << 30)) want = (free_b - reserve) & ~((1ull << 21) - 1);
		r.fixed[dev] = false;
	}
	if (!want) return;
	void* q = nullptr;
	if (cudaMalloc(&q, want) == cudaSuccess) clb::job_slab(dev).init(reinterpret_cast<uint64_t>(q), want);
	else cudaGetLastError();                        // not enough memory for the slab: the job runs on cudaMalloc
}
void slab_release(int dev)
{
	SlabRefs& r = slab_refs();
	std::lock_guard<std::mutex> g(r.m);
	if (dev < 0 || dev >= clb::SLAB_MAX_DEVICES || r.refs[dev] == 0) return;
	if (--r.refs[dev] != 0 || r.fixed[dev] || !clb::job_slab(dev).active()) return;
	void* q = reinterpret_cast<void*>(clb::job_slab(dev).base());
	if (clb::job_slab(dev).reset()) cudaFree(q);
}
}

extern "C" {

clb_status clb_create(const clb_params* p, clb_ctx** out)
{
	if (!p || !out) return fail(nullptr, CLB_ERR_BAD_ARG, "null argument");
	*out = nullptr;
	if (p->kmer_len < 8 || p->kmer_len > 32) return fail(nullptr, CLB_ERR_BAD_ARG, "kmer_len must be in [8, 32]");
	if (p->modulo == 0 || p->max_candidates == 0 || p->min_count == 0 || p->max_count < p->min_count) return fail(nullptr, CLB_ERR_BAD_ARG, "bad filter parameters");
	int n_dev = 0;
	cudaError_t e = cudaGetDeviceCount(&n_dev);
	if (e != cudaSuccess || n_dev == 0) { cudaGetLastError(); return fail(nullptr, CLB_ERR_NO_DEVICE, "no CUDA device: colord_b200 has no CPU path for the hot stages"); }
	if (p->device < 0 || p->device >= n_dev) return fail(nullptr, CLB_ERR_BAD_ARG, "device ordinal out of range");
	clb_ctx* c = new clb_ctx();
	c->prm = *p;
	c->mt = make_modtest(p->modulo);
	e = cudaSetDevice(p->device);
	if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
	if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream3, cudaStreamNonBlocking);
	if (e == cudaSuccess) { c->own_stream = true; e = cudaDeviceGetAttribute(&c->n_sm, cudaDevAttrMultiProcessorCount, p->device); }
	if (e != cudaSuccess) { clb_status st = cuda_fail(nullptr, e, "clb_create"); delete c; return st; }
	{	// stream-ordered scratch (cudaMallocAsync) is recycled inside the pool instead of going back to the driver after every sync
		cudaMemPool_t pool;
		if (cudaDeviceGetDefaultMemPool(&pool, p->device) == cudaSuccess) { unsigned long long thr = ~0ull; cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr); }
	}
	slab_acquire(p->device);
	clb_status st = s1a_init(c);
	if (st != CLB_OK) { g_create_error = c->err; clb_destroy(c); return st; }
	*out = c;
	return CLB_OK;
}

void clb_destroy(clb_ctx* c)
{
	if (!c) return;
	cudaSetDevice(c->prm.device);
	cudaStreamSynchronize(c->stream);
	s1_free(c);
	s2_free(c);
	c->qs.release(); c->ds.release(); c->hs.release(); c->xd.release(); c->xq.release(); c->xh.release(); c->xg.release(); c->dq.release();
	if (c->stream3) { cudaStreamSynchronize(c->stream3); cudaStreamDestroy(c->stream3); }
	if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
	const int dev = c->prm.device;
	delete c;
	slab_release(dev);
}

const char* clb_last_error(const clb_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

clb_status clb_set_stream(clb_ctx* c, void* stream)
{
	if (!c) return CLB_ERR_BAD_ARG;
	cudaStreamSynchronize(c->stream);
	if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
	c->stream = static_cast<cudaStream_t>(stream); c->own_stream = false;
	return CLB_OK;
}

clb_status clb_synchronize(clb_ctx* c)
{
	if (!c) return CLB_ERR_BAD_ARG;
	CLB_CUDA(c, cudaStreamSynchronize(c->stream));
	return CLB_OK;
}

// a host thread that enters the library for the first time binds the device's primary context (cudaSetDevice + the cudaFree(0)
// idiom): the stage-3 quality / header calls may come from a thread of their own
static thread_local int g_bound_device = -1;
static inline void bind_device(int dev) { cudaSetDevice(dev); if (g_bound_device != dev) { cudaFree(0); g_bound_device = dev; } }
#define CLB_ENTER(c) do { if (!(c)) return CLB_ERR_BAD_ARG; bind_device((c)->prm.device); } while (0)

clb_status clb_append_reads(clb_ctx* c, const uint8_t* bases, const uint64_t* offsets, uint32_t n_reads, int on_device)
{
	CLB_ENTER(c);
	if (n_reads && (!offsets || !bases)) return fail(c, CLB_ERR_BAD_ARG, "null bases/offsets");
	return s1a_append(c, bases, offsets, n_reads, on_device);
}
void* clb_host_alloc(uint64_t bytes)
{
	void* p = nullptr;
	if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
	return p;
}
void clb_host_free(void* p) { if (p) cudaFreeHost(p); }
clb_status clb_append_quals(clb_ctx* c, const uint8_t* quals, uint64_t n, int on_device)
{
	CLB_ENTER(c);
	if (n && !quals) return fail(c, CLB_ERR_BAD_ARG, "null qualities");
	// on the stage-3 stream (the quality coders' own): a second host thread may bring the qualities in while stages 1 and 2 run
	cudaStream_t qs = c->stream3;
	if (!c->dq.cap) { const uint64_t hint = c->prm.expected_bases + c->prm.expected_bases / 16 + 1024; CLB_CUDA(c, c->dq.reserve(std::max(hint, n) + 16, qs, false)); }
	else CLB_CUDA(c, c->dq.reserve(c->dq_n + n + 16, qs, true, c->dq_n));
	if (n) CLB_CUDA(c, cudaMemcpyAsync(c->dq.p + c->dq_n, quals, n, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, qs));
	CLB_CUDA(c, cudaStreamSynchronize(qs));             // the caller may reuse its buffer
	c->dq_n += n;
	return CLB_OK;
}
clb_status clb_counts_size(clb_ctx* c, uint32_t part, uint32_t n_parts, uint64_t* n) { CLB_ENTER(c); return s1a_counts_size(c, part, n_parts, n); }
clb_status clb_counts_export(clb_ctx* c, uint32_t part, uint32_t n_parts, uint64_t* kmers, uint32_t* counts, uint64_t cap, uint64_t* n, int on_device)
{ CLB_ENTER(c); return s1a_counts_export(c, part, n_parts, kmers, counts, cap, n, on_device); }
clb_status clb_counts_sizes(clb_ctx* c, uint32_t n_parts, uint64_t* sizes) { CLB_ENTER(c); return s1a_counts_sizes(c, n_parts, sizes); }
clb_status clb_counts_export_all(clb_ctx* c, uint32_t n_parts, const uint64_t* first, uint64_t* kmers, uint32_t* counts, uint64_t cap)
{ CLB_ENTER(c); return s1a_counts_export_all(c, n_parts, first, kmers, counts, cap); }
clb_status clb_counts_reset(clb_ctx* c) { CLB_ENTER(c); return s1a_counts_reset(c); }
clb_status clb_counts_merge(clb_ctx* c, const uint64_t* kmers, const uint32_t* counts, uint64_t n, uint64_t n_reads_remote, int on_device)
{ CLB_ENTER(c); return s1a_counts_merge(c, kmers, counts, n, n_reads_remote, on_device); }
clb_status clb_count_finalize(clb_ctx* c, clb_kmer_stats* stats) { CLB_ENTER(c); return s1a_finalize(c, stats); }

clb_status clb_filter_list(clb_ctx* c, uint64_t* kmers, uint32_t* counts, uint64_t cap, uint64_t* n, int on_device)
{
	CLB_ENTER(c);
	if (!c->finalized) return fail(c, CLB_ERR_STATE, "clb_filter_list before clb_count_finalize");
	if (n) *n = c->n_surv;
	if (cap < c->n_surv) return fail(c, CLB_ERR_CAPACITY, "clb_filter_list: buffer too small");
	if (c->n_surv == 0) return CLB_OK;
	const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
	CLB_CUDA(c, cudaMemcpyAsync(kmers, c->sv_kmer, sizeof(uint64_t) * c->n_surv, kind, c->stream));
	CLB_CUDA(c, cudaMemcpyAsync(counts, c->sv_count, sizeof(uint32_t) * c->n_surv, kind, c->stream));
	CLB_CUDA(c, cudaStreamSynchronize(c->stream));
	return CLB_OK;
}
clb_status clb_filter_import(clb_ctx* c, const uint64_t* kmers, const uint32_t* counts, uint64_t n, const clb_kmer_stats* gs, int on_device)
{ CLB_ENTER(c); return s1a_filter_import(c, kmers, counts, n, gs, on_device); }
clb_status clb_filter_check(clb_ctx* c, const uint64_t* kmers, uint64_t n, uint8_t* possible, uint8_t* present)
{ CLB_ENTER(c); return s1a_filter_check(c, kmers, n, possible, present); }

clb_status clb_append_context_reads(clb_ctx* c, const uint8_t* bases, const uint64_t* offsets, uint32_t n_reads, int on_device)
{
	CLB_ENTER(c);
	if (!offsets || (!bases && n_reads)) return fail(c, CLB_ERR_BAD_ARG, "null argument");
	return s1a_append(c, bases, offsets, n_reads, on_device, true);
}
clb_status clb_reads_have_n(clb_ctx* c, uint8_t* flags) { CLB_ENTER(c); if (!flags) return fail(c, CLB_ERR_BAD_ARG, "null argument"); return s1b_reads_have_n(c, flags); }
clb_status clb_reads_export(clb_ctx* c, const uint32_t* read_ids, uint32_t n, uint8_t* bases, uint64_t cap, int on_device)
{
	CLB_ENTER(c);
	if ((!read_ids || !bases) && n) return fail(c, CLB_ERR_BAD_ARG, "null argument");
	return s1b_reads_export(c, read_ids, n, bases, cap, on_device);
}
clb_status clb_graph_build(clb_ctx* c, const uint8_t* is_reference, uint32_t n_pseudo) { CLB_ENTER(c); return s1b_build(c, is_reference, n_pseudo); }

clb_status clb_graph_accepted_size(clb_ctx* c, uint64_t* total)
{
	CLB_ENTER(c);
	if (!c->graph_done) return fail(c, CLB_ERR_STATE, "graph not built");
	*total = c->acc_total;
	return CLB_OK;
}

clb_status clb_graph_accepted(clb_ctx* c, uint64_t* offsets, uint64_t* kmers, uint64_t cap)
{
	CLB_ENTER(c);
	if (!c->graph_done) return fail(c, CLB_ERR_STATE, "graph not built");
	if (cap < c->acc_total) return fail(c, CLB_ERR_CAPACITY, "clb_graph_accepted: buffer too small");
	const uint64_t n = c->n_reads;
	std::vector<uint64_t> st(n), svk(c->n_surv); std::vector<uint32_t> cn(n), ids(c->acc_total);
	CLB_CUDA(c, cudaMemcpyAsync(st.data(), c->acc_start, sizeof(uint64_t) * n, cudaMemcpyDeviceToHost, c->stream));
	CLB_CUDA(c, cudaMemcpyAsync(cn.data(), c->acc_n, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, c->stream));
	CLB_CUDA(c, cudaMemcpyAsync(ids.data(), c->acc_id, sizeof(uint32_t) * c->acc_total, cudaMemcpyDeviceToHost, c->stream));
	CLB_CUDA(c, cudaMemcpyAsync(svk.data(), c->sv_kmer, sizeof(uint64_t) * c->n_surv, cudaMemcpyDeviceToHost, c->stream));
	CLB_CUDA(c, cudaStreamSynchronize(c->stream));
	uint64_t o = 0;
	for (uint64_t i = 0; i < n; ++i) {          // the device arena is in completion order; the API is CSR in read order
		offsets[i] = o;
		for (uint32_t j = 0; j < cn[i]; ++j) kmers[o++] = svk[ids[st[i] + j]];
	}
	offsets[n] = o;
	return CLB_OK;
}

clb_status clb_graph_candidates(clb_ctx* c, uint32_t* cand, uint32_t* cand_n)
{
	CLB_ENTER(c);
	if (!c->graph_done) return fail(c, CLB_ERR_STATE, "graph not built");
	const uint64_t n = c->n_reads;
	if (!n) return CLB_OK;
	CLB_CUDA(c, cudaMemcpyAsync(cand, c->cand, sizeof(uint32_t) * n * c->prm.max_candidates, cudaMemcpyDeviceToHost, c->stream));
	CLB_CUDA(c, cudaMemcpyAsync(cand_n, c->cand_n, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, c->stream));
	CLB_CUDA(c, cudaStreamSynchronize(c->stream));
	return CLB_OK;
}

clb_status clb_graph_common_size(clb_ctx* c, uint64_t* total)
{
	CLB_ENTER(c);
	if (!c->graph_done || !c->prm.is_hifi) return fail(c, CLB_ERR_STATE, "HiFi graph not built");
	*total = c->common_total;
	return CLB_OK;
}

clb_status clb_graph_common(clb_ctx* c, uint64_t* common_off, uint32_t* common_n, uint64_t* kmers, uint64_t cap)
{
	CLB_ENTER(c);
	if (!c->graph_done || !c->prm.is_hifi) return fail(c, CLB_ERR_STATE, "HiFi graph not built");
	if (cap < c->common_total) return fail(c, CLB_ERR_CAPACITY, "clb_graph_common: buffer too small");
	const uint64_t n = c->n_reads * c->prm.max_candidates;
	if (!n) return CLB_OK;
	std::vector<uint32_t> cn(c->n_reads);
	CLB_CUDA(c, cudaMemcpyAsync(common_off, c->common_off, sizeof(uint64_t) * n, cudaMemcpyDeviceToHost, c->stream));
	CLB_CUDA(c, cudaMemcpyAsync(common_n, c->cand_votes, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, c->stream));
	CLB_CUDA(c, cudaMemcpyAsync(cn.data(), c->cand_n, sizeof(uint32_t) * c->n_reads, cudaMemcpyDeviceToHost, c->stream));
	if (c->common_total) CLB_CUDA(c, cudaMemcpyAsync(kmers, c->common, sizeof(uint64_t) * c->common_total, cudaMemcpyDeviceToHost, c->stream));
	CLB_CUDA(c, cudaStreamSynchronize(c->stream));
	for (uint64_t i = 0; i < c->n_reads; ++i)
		for (uint32_t j = cn[i]; j < c->prm.max_candidates; ++j) { common_n[i * c->prm.max_candidates + j] = 0; common_off[i * c->prm.max_candidates + j] = 0; }
	return CLB_OK;
}

clb_status clb_edit_scripts(clb_ctx* c, const uint8_t* seqs, uint64_t n_seq_bytes, const uint64_t* ref_off, const uint32_t* ref_len,
	const uint64_t* enc_off, const uint32_t* enc_len, const uint32_t* kind, uint64_t n, uint64_t* out_off, char* out, uint64_t cap)
{
	CLB_ENTER(c);
	if (n && (!seqs || !ref_off || !ref_len || !enc_off || !enc_len || !kind || !out_off || !out)) return fail(c, CLB_ERR_BAD_ARG, "null argument");
	return s2_edit_scripts(c, seqs, n_seq_bytes, ref_off, ref_len, enc_off, enc_len, kind, n, out_off, out, cap);
}

clb_status clb_encode(clb_ctx* c, const clb_encode_params* prm, const uint32_t* pack_sizes, uint32_t n_packs)
{
	CLB_ENTER(c);
	if (!prm) return fail(c, CLB_ERR_BAD_ARG, "null params");
	return s2_encode(c, prm, pack_sizes, n_packs);
}
clb_status clb_encode_size(clb_ctx* c, uint64_t* total)
{
	CLB_ENTER(c);
	if (!c->enc_done) return fail(c, CLB_ERR_STATE, "clb_encode has not run");
	*total = c->es_total;
	return CLB_OK;
}
clb_status clb_encode_get(clb_ctx* c, uint64_t* es_off, uint8_t* es, uint64_t cap, int on_device)
{
	CLB_ENTER(c);
	if (!c->enc_done) return fail(c, CLB_ERR_STATE, "clb_encode has not run");
	if (cap < c->es_total) return fail(c, CLB_ERR_CAPACITY, "clb_encode_get: buffer too small");
	const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
	CLB_CUDA(c, cudaMemcpyAsync(es_off, c->es_off, sizeof(uint64_t) * (c->n_reads + 1), kind, c->stream));
	if (c->es_total) CLB_CUDA(c, cudaMemcpyAsync(es, c->es.p, c->es_total, kind, c->stream));
	CLB_CUDA(c, cudaStreamSynchronize(c->stream));
	return CLB_OK;
}
clb_status clb_encode_keep_candidates(clb_ctx* c, int on) { if (!c) return CLB_ERR_BAD_ARG; c->keep_candidates = on != 0; return CLB_OK; }
clb_status clb_encode_stats_enable(clb_ctx* c, int on) { if (!c) return CLB_ERR_BAD_ARG; c->collect_stats = on != 0; return CLB_OK; }
clb_status clb_encode_stats_get(clb_ctx* c, clb_encode_stats* out)
{
	if (!c || !out) return CLB_ERR_BAD_ARG;
	if (!c->enc_done || !c->collect_stats) return fail(c, CLB_ERR_STATE, "no statistics: clb_encode_stats_enable before clb_encode");
	*out = c->h_stats;
	return CLB_OK;
}
clb_status clb_encode_candidates_size(clb_ctx* c, uint64_t* n_words)
{
	if (!c || !c->enc_done || !c->keep_candidates) return fail(c, CLB_ERR_STATE, "candidates were not kept");
	uint64_t t = 0; for (auto& v : c->dbg_cand) t += v.size();
	*n_words = t;
	return CLB_OK;
}
clb_status clb_encode_candidates(clb_ctx* c, uint64_t* cand_off, uint32_t* data, uint64_t cap_words)
{
	if (!c || !c->enc_done || !c->keep_candidates) return fail(c, CLB_ERR_STATE, "candidates were not kept");
	uint64_t t = 0;
	for (uint64_t i = 0; i < c->dbg_cand.size(); ++i) {
		cand_off[i] = t;
		if (t + c->dbg_cand[i].size() > cap_words) return fail(c, CLB_ERR_CAPACITY, "clb_encode_candidates: buffer too small");
		std::memcpy(data + t, c->dbg_cand[i].data(), sizeof(uint32_t) * c->dbg_cand[i].size());
		t += c->dbg_cand[i].size();
	}
	cand_off[c->dbg_cand.size()] = t;
	return CLB_OK;
}

clb_status clb_dna_encode(clb_ctx* c, uint32_t level, const uint32_t* pack_sizes, uint32_t n_packs) { CLB_ENTER(c); return s3_dna_encode(c, level, pack_sizes, n_packs); }
clb_status clb_dna_size(clb_ctx* c, uint64_t* total, uint64_t* header)
{
	CLB_ENTER(c);
	if (!c->dna_done) return fail(c, CLB_ERR_STATE, "clb_dna_encode has not run");
	*total = c->ds_total; if (header) *header = c->ds_header;
	return CLB_OK;
}
clb_status clb_dna_get(clb_ctx* c, uint8_t* stream, uint64_t cap, int on_device)
{
	CLB_ENTER(c);
	if (!c->dna_done) return fail(c, CLB_ERR_STATE, "clb_dna_encode has not run");
	if (cap < c->ds_total) return fail(c, CLB_ERR_CAPACITY, "clb_dna_get: buffer too small");
	CLB_CUDA(c, cudaMemcpyAsync(stream, c->ds.p, c->ds_total, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
	CLB_CUDA(c, cudaStreamSynchronize(c->stream));
	return CLB_OK;
}
clb_status clb_hdr_encode(clb_ctx* c, const uint8_t* bytes, const uint64_t* offsets, const uint8_t* plus_id, uint64_t n, int on_device, const uint32_t* pack_sizes, uint32_t n_packs)
{
	CLB_ENTER(c);
	if (!offsets || (!bytes && n)) return fail(c, CLB_ERR_BAD_ARG, "null argument");
	return s3_hdr_encode(c, bytes, offsets, plus_id, n, on_device, pack_sizes, n_packs);
}
clb_status clb_hdr_size(clb_ctx* c, uint64_t* total, uint64_t* header)
{
	CLB_ENTER(c);
	if (!c->hdr_done) return fail(c, CLB_ERR_STATE, "clb_hdr_encode has not run");
	*total = c->hs_total; if (header) *header = c->hs_header;
	return CLB_OK;
}
clb_status clb_hdr_get(clb_ctx* c, uint8_t* stream, uint64_t cap, int on_device)
{
	CLB_ENTER(c);
	if (!c->hdr_done) return fail(c, CLB_ERR_STATE, "clb_hdr_encode has not run");
	if (cap < c->hs_total) return fail(c, CLB_ERR_CAPACITY, "clb_hdr_get: buffer too small");
	CLB_CUDA(c, cudaMemcpyAsync(stream, c->hs.p, c->hs_total, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
	CLB_CUDA(c, cudaStreamSynchronize(c->stream));
	return CLB_OK;
}
clb_status clb_xdna_encode(clb_ctx* c, uint32_t level, const uint32_t* pack_sizes, uint32_t n_packs)
{
	CLB_ENTER(c);
	return s3x_dna_encode(c, level, pack_sizes, n_packs);
}
clb_status clb_xqual_encode(clb_ctx* c, uint32_t mode, uint32_t source, uint32_t level, const uint32_t* thr, const uint8_t* quals, const uint64_t* offsets, int on_device, const uint32_t* pack_sizes, uint32_t n_packs)
{
	CLB_ENTER(c);
	return s3x_qual_encode(c, mode, source, level, thr, quals, offsets, on_device, pack_sizes, n_packs);
}
clb_status clb_xhdr_encode(clb_ctx* c, const uint8_t* bytes, const uint64_t* offsets, const uint8_t* plus_id, uint64_t n, int on_device, const uint32_t* pack_sizes, uint32_t n_packs)
{
	CLB_ENTER(c);
	if (n && (!offsets || !bytes)) return fail(c, CLB_ERR_BAD_ARG, "null argument");
	return s3x_hdr_encode(c, bytes, offsets, plus_id, n, on_device, pack_sizes, n_packs);
}
clb_status clb_count_sequences(clb_ctx* c, const uint8_t* bases, const uint64_t* offsets, uint32_t n, int on_device)
{
	CLB_ENTER(c);
	if (n && (!offsets || !bases)) return fail(c, CLB_ERR_BAD_ARG, "null bases/offsets");
	return s1a_append(c, bases, offsets, n, on_device, false, true);
}
clb_status clb_xplain_encode(clb_ctx* c, const uint8_t* bases, const uint64_t* offsets, uint32_t n_seqs, uint32_t level)
{
	CLB_ENTER(c);
	return s3x_plain_encode(c, bases, offsets, n_seqs, level);
}
clb_status clb_xstream_size(clb_ctx* c, uint32_t which, uint64_t* total, uint32_t* n_parts)
{
	CLB_ENTER(c);
	if (which > 3 || !total || !n_parts) return fail(c, CLB_ERR_BAD_ARG, "clb_xstream_size: bad argument");
	const std::vector<uint64_t>& parts = which == 0 ? c->xd_parts : which == 1 ? c->xq_parts : which == 2 ? c->xh_parts : c->xg_parts;
	*total = which == 0 ? c->xd_total : which == 1 ? c->xq_total : which == 2 ? c->xh_total : c->xg_total;
	*n_parts = (uint32_t)parts.size();
	return CLB_OK;
}
clb_status clb_xstream_get(clb_ctx* c, uint32_t which, uint8_t* bytes, uint64_t cap, uint64_t* part_sizes, int on_device)
{
	CLB_ENTER(c);
	if (which > 3) return fail(c, CLB_ERR_BAD_ARG, "clb_xstream_get: bad stream");
	const std::vector<uint64_t>& parts = which == 0 ? c->xd_parts : which == 1 ? c->xq_parts : which == 2 ? c->xh_parts : c->xg_parts;
	const uint64_t total = which == 0 ? c->xd_total : which == 1 ? c->xq_total : which == 2 ? c->xh_total : c->xg_total;
	const uint8_t* src = which == 0 ? c->xd.p : which == 1 ? c->xq.p : which == 2 ? c->xh.p : c->xg.p;
	if (cap < total) return fail(c, CLB_ERR_CAPACITY, "clb_xstream_get: buffer too small");
	if (part_sizes) for (size_t i = 0; i < parts.size(); ++i) part_sizes[i] = parts[i];
	if (total) {
		CLB_CUDA(c, cudaMemcpyAsync(bytes, src, total, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
		CLB_CUDA(c, cudaStreamSynchronize(c->stream));
	}
	return CLB_OK;
}
clb_status clb_qual_encode(clb_ctx* c, const clb_qual_params* prm, const uint8_t* quals, const uint64_t* offsets, int on_device, const uint32_t* pack_sizes, uint32_t n_packs)
{
	CLB_ENTER(c);
	if (!prm || (quals && !offsets)) return fail(c, CLB_ERR_BAD_ARG, "null argument");
	return s3_qual_encode(c, prm, quals, offsets, on_device, pack_sizes, n_packs);
}
clb_status clb_qual_encode_original(clb_ctx* c, uint32_t source, uint32_t level, const uint8_t* quals, const uint64_t* offsets, int on_device, const uint32_t* pack_sizes, uint32_t n_packs)
{
	CLB_ENTER(c);
	if (quals && !offsets) return fail(c, CLB_ERR_BAD_ARG, "null argument");
	return s3_qual_encode_original(c, source, level, quals, offsets, on_device, pack_sizes, n_packs);
}
clb_status clb_qual_size(clb_ctx* c, uint64_t* total)
{
	CLB_ENTER(c);
	if (!c->qual_done) return fail(c, CLB_ERR_STATE, "clb_qual_encode has not run");
	*total = c->qs_total;
	return CLB_OK;
}
clb_status clb_qual_get(clb_ctx* c, uint8_t* stream, uint64_t cap, int on_device)
{
	CLB_ENTER(c);
	if (!c->qual_done) return fail(c, CLB_ERR_STATE, "clb_qual_encode has not run");
	if (cap < c->qs_total) return fail(c, CLB_ERR_CAPACITY, "clb_qual_get: buffer too small");
	CLB_CUDA(c, cudaMemcpyAsync(stream, c->qs.p, c->qs_total, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
	CLB_CUDA(c, cudaStreamSynchronize(c->stream));
	return CLB_OK;
}

clb_status clb_get_packed_read(clb_ctx* c, uint32_t read_id, uint8_t* out, uint64_t cap, uint64_t* n_bytes)
{
	CLB_ENTER(c);
	if (read_id >= c->n_reads) return fail(c, CLB_ERR_BAD_ARG, "read id out of range");
	const uint64_t start = c->h_rd_start[read_id], len = c->h_rd_len[read_id];
	const uint64_t need = (len + 3) / 4 + 1;
	if (n_bytes) *n_bytes = need;
	if (cap < need) return fail(c, CLB_ERR_CAPACITY, "clb_get_packed_read: buffer too small");
	const uint64_t w0 = start >> 5, w1 = len ? (start + len - 1) >> 5 : w0;
	std::vector<uint64_t> w(w1 - w0 + 2, 0);
	if (len) {
		CLB_CUDA(c, cudaMemcpyAsync(w.data(), c->pk.p + w0, sizeof(uint64_t) * (w1 - w0 + 1), cudaMemcpyDeviceToHost, c->stream));
		CLB_CUDA(c, cudaStreamSynchronize(c->stream));
	}
	// re-serialise into the reference's byte layout: 4 bases per byte MSB first + trailer (reference_reads.h:35-72)
	uint64_t nb = 0; uint8_t b = 0; uint32_t in_byte = 0;
	for (uint64_t i = 0; i < len; ++i) {
		const uint64_t p = start + i;
		const uint32_t x = (uint32_t)(w[(p >> 5) - w0] >> (62 - 2 * (p & 31))) & 3u;
		b = (uint8_t)((b << 2) | x);
		if (++in_byte == 4) { out[nb++] = b; b = 0; in_byte = 0; }
	}
	if (in_byte) out[nb++] = (uint8_t)(b << (2 * (4 - in_byte)));
	out[nb++] = (uint8_t)in_byte;
	return CLB_OK;
}

// CRefReadsAccepter (ref_reads_accepter.h:27-57): default-seeded mt19937 + uniform_real_distribution<double>,
// one draw per non-pseudo read.  Uses the same libstdc++ facilities so the stream is identical by construction.
void clb_sampler(uint32_t range, double exponent, uint32_t n_pseudo, uint32_t n, uint8_t* decisions)
{
	std::mt19937 mt;
	std::uniform_real_distribution<double> dist(0.0, 1.0);
	if (range == 0) range = 1;
	for (uint32_t i = 0; i < n; ++i) {
		if (i < n_pseudo) { decisions[i] = 1; continue; }
		const uint32_t range_no = (i - n_pseudo) / range;
		const double p = std::pow(1.0 / (range_no + 1), exponent);
		decisions[i] = dist(mt) <= p;
	}
}

clb_status clb_release_cached_memory(int device)
{
	cudaMemPool_t pool;
	if (cudaSetDevice(device) != cudaSuccess || cudaDeviceGetDefaultMemPool(&pool, device) != cudaSuccess) return CLB_ERR_NO_DEVICE;
	cudaDeviceSynchronize();
	if (job_slab(device).active()) { void* q = reinterpret_cast<void*>(job_slab(device).base()); if (job_slab(device).reset()) cudaFree(q); }      // only when no context holds blocks of it
	return cudaMemPoolTrimTo(pool, 0) == cudaSuccess ? CLB_OK : CLB_ERR_CUDA;
}

uint64_t clb_kernel_launches(const clb_ctx* c) { return c ? c->launches.load() : 0; }

clb_status clb_profile_enable(clb_ctx* c, int on)
{
	if (!c) return CLB_ERR_BAD_ARG;
	prof_resolve(c);
	c->prof_on = on != 0;
	for (int i = 0; i < K_N; ++i) { c->prof_ms[i] = 0; c->prof_n[i] = 0; }
	return CLB_OK;
}

clb_status clb_profile_get(clb_ctx* c, const char* kernel, double* ms, uint64_t* launches)
{
	if (!c || !kernel) return CLB_ERR_BAD_ARG;
	prof_resolve(c);
	for (int i = 0; i < K_N; ++i)
		if (std::strcmp(kernel, kernel_names[i]) == 0) { if (ms) *ms = c->prof_ms[i]; if (launches) *launches = c->prof_n[i]; return CLB_OK; }
	return fail(c, CLB_ERR_BAD_ARG, "unknown kernel class");
}

} // extern "C"
