// mgpu.cu — the two exchanges of the multi-GPU path over NCCL, in C++ behind the C-ABI of include/colord_b200_mgpu.h:
// libcolord_b200_mgpu.so (the single-GPU library has no NCCL dependency).  One process, one clb_ctx per GPU, one host thread per
// rank; every entry point below is COLLECTIVE: all ranks' threads call it, each with its own rank.
//   1. k-mer counts      k-mers are owned by hash partition: one all-to-all (grouped ncclSend / ncclRecv) moves every (k-mer, count)
//                        pair to its owner, owners threshold their share, one ncclAllGather hands every rank the union of
//                        survivors — the filtered set of the WHOLE input (the collective the north star names)
//   2. reference reads   the candidates of a read are earlier reference reads of the whole input: one ncclAllGather of every rank's
//                        reference reads; rank r keeps those of the ranks before it as context reads (clb_append_context_reads)
// Same steps as colord_b200/dist.py (which does them through torch.distributed for bench.py); here sizes travel through host memory
// of the one process instead of extra collectives, and the table is scanned once per partition.
#include "../../include/colord_b200.h"
#include "../../include/colord_b200_mgpu.h"
#include <nccl.h>
#include <cuda_runtime.h>
#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

struct clb_group {
	uint32_t n = 0;
	std::vector<clb_ctx*> ctx; std::vector<int> dev; std::vector<ncclComm_t> comm; std::vector<cudaStream_t> stream;
	std::vector<std::string> err;
	// a reusable barrier over the rank threads + host-side exchange areas
	std::mutex m; std::condition_variable cv; uint32_t waiting = 0; uint64_t generation = 0;
	std::vector<std::vector<uint64_t>> sizes;          // [rank][partition]
	std::vector<clb_kmer_stats> local; std::vector<uint64_t> filt; std::vector<uint64_t> ref_reads, ref_bases;
	std::vector<int> failed;
	void barrier()
	{
		std::unique_lock<std::mutex> lk(m);
		const uint64_t g = generation;
		if (++waiting == n) { waiting = 0; ++generation; cv.notify_all(); }
		else cv.wait(lk, [&] { return generation != g; });
	}
};

namespace {
struct DevMem {      // device allocations of one call, freed at its end
	std::vector<void*> v;
	~DevMem() { for (void* p : v) cudaFree(p); }
	template <typename T> cudaError_t get(T** p, uint64_t count) { const cudaError_t e = cudaMalloc((void**)p, std::max<uint64_t>(1, count) * sizeof(T)); if (e == cudaSuccess) v.push_back(*p); return e; }
};
clb_status gfail(clb_group* g, uint32_t rank, clb_status st, const std::string& msg) { g->err[rank] = msg; g->failed[rank] = 1; return st; }
}
#define G_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return gfail(g, rank, CLB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); } while (0)
#define G_NCCL(call) do { ncclResult_t r__ = (call); if (r__ != ncclSuccess) return gfail(g, rank, CLB_ERR_CUDA, std::string(#call) + ": " + ncclGetErrorString(r__)); } while (0)
#define G_CLB(call) do { clb_status s__ = (call); if (s__ != CLB_OK) return gfail(g, rank, s__, std::string(#call) + ": " + clb_last_error(g->ctx[rank])); } while (0)
// after a barrier: did any rank fail before it?  (every rank reaches every barrier, so nobody is left waiting)
#define G_CHECK_ALL() do { for (uint32_t r__ = 0; r__ < g->n; ++r__) if (g->failed[r__]) return g->failed[rank] ? CLB_ERR_STATE : gfail(g, rank, CLB_ERR_STATE, "another rank failed: " + g->err[r__]); } while (0)

extern "C" {

clb_status clb_group_create(clb_ctx* const* ctxs, const int32_t* devices, uint32_t n, clb_group** out)
{
	if (!ctxs || !devices || !out || n < 1) return CLB_ERR_BAD_ARG;
	clb_group* g = new clb_group();
	g->n = n; g->ctx.assign(ctxs, ctxs + n); g->dev.assign(devices, devices + n); g->comm.resize(n); g->stream.resize(n); g->err.resize(n); g->failed.assign(n, 0);
	g->sizes.assign(n, std::vector<uint64_t>(n, 0)); g->local.resize(n); g->filt.assign(n, 0); g->ref_reads.assign(n, 0); g->ref_bases.assign(n, 0);
	if (ncclCommInitAll(g->comm.data(), (int)n, g->dev.data()) != ncclSuccess) { delete g; return CLB_ERR_CUDA; }
	for (uint32_t r = 0; r < n; ++r) { cudaSetDevice(g->dev[r]); if (cudaStreamCreateWithFlags(&g->stream[r], cudaStreamNonBlocking) != cudaSuccess) { delete g; return CLB_ERR_CUDA; } }
	*out = g;
	return CLB_OK;
}
void clb_group_destroy(clb_group* g)
{
	if (!g) return;
	for (uint32_t r = 0; r < g->n; ++r) { cudaSetDevice(g->dev[r]); if (g->stream[r]) cudaStreamDestroy(g->stream[r]); if (g->comm[r]) ncclCommDestroy(g->comm[r]); }
	delete g;
}
const char* clb_group_last_error(const clb_group* g, uint32_t rank) { return g && rank < g->n ? g->err[rank].c_str() : "bad group / rank"; }

clb_status clb_group_exchange_counts(clb_group* g, uint32_t rank, clb_kmer_stats* global_stats)
{
	if (!g || rank >= g->n || !global_stats) return CLB_ERR_BAD_ARG;
	const uint32_t N = g->n; clb_ctx* c = g->ctx[rank];
	cudaSetDevice(g->dev[rank]);
	cudaStream_t s = g->stream[rank];
	DevMem mem;
	clb_status st = CLB_OK;
	// 1. my table's partitions
	auto sizes_step = [&]() -> clb_status { G_CLB(clb_counts_sizes(c, N, g->sizes[rank].data())); return CLB_OK; };      // one pass for all partitions
	st = sizes_step();
	g->barrier(); G_CHECK_ALL();
	uint64_t n_send = 0, n_recv = 0;
	for (uint32_t p = 0; p < N; ++p) { n_send += g->sizes[rank][p]; n_recv += g->sizes[p][rank]; }
	uint64_t *send_k = nullptr, *recv_k = nullptr; uint32_t *send_c = nullptr, *recv_c = nullptr;
	auto a2a_step = [&]() -> clb_status {
		G_CUDA(mem.get(&send_k, n_send)); G_CUDA(mem.get(&send_c, n_send)); G_CUDA(mem.get(&recv_k, n_recv)); G_CUDA(mem.get(&recv_c, n_recv));
		std::vector<uint64_t> first(N, 0);
		for (uint32_t p = 1; p < N; ++p) first[p] = first[p - 1] + g->sizes[rank][p - 1];
		if (n_send) G_CLB(clb_counts_export_all(c, N, first.data(), send_k, send_c, n_send));      // one more pass writes every partition at its offset
		G_CLB(clb_synchronize(c));
		// 2. all-to-all of the pairs: partition p of every rank goes to rank p
		G_NCCL(ncclGroupStart());
		uint64_t so = 0, ro = 0;
		for (uint32_t p = 0; p < N; ++p) {
			const uint64_t ns = g->sizes[rank][p], nr = g->sizes[p][rank];
			if (ns) { G_NCCL(ncclSend(send_k + so, ns, ncclUint64, (int)p, g->comm[rank], s)); G_NCCL(ncclSend(send_c + so, ns, ncclUint32, (int)p, g->comm[rank], s)); }
			if (nr) { G_NCCL(ncclRecv(recv_k + ro, nr, ncclUint64, (int)p, g->comm[rank], s)); G_NCCL(ncclRecv(recv_c + ro, nr, ncclUint32, (int)p, g->comm[rank], s)); }
			so += ns; ro += nr;
		}
		G_NCCL(ncclGroupEnd());
		G_CUDA(cudaStreamSynchronize(s));
		// 3. my table = what every rank counted for my partition; threshold it
		G_CLB(clb_counts_reset(c));
		G_CLB(clb_counts_merge(c, recv_k, recv_c, n_recv, 0, 1));
		G_CLB(clb_count_finalize(c, &g->local[rank]));
		uint64_t n_mine = 0;
		const clb_status fs = clb_filter_list(c, nullptr, nullptr, 0, &n_mine, 1);
		if (fs != CLB_OK && fs != CLB_ERR_CAPACITY) return gfail(g, rank, fs, std::string("clb_filter_list: ") + clb_last_error(c));
		g->filt[rank] = n_mine;
		return CLB_OK;
	};
	if (st == CLB_OK) st = a2a_step();
	g->barrier(); G_CHECK_ALL();
	// 4. statistics are sums over the owners (n_reads over the shards)
	clb_kmer_stats tot{};
	for (uint32_t r = 0; r < N; ++r) { tot.n_reads += g->local[r].n_reads; tot.tot_kmers += g->local[r].tot_kmers; tot.n_unique += g->local[r].n_unique; tot.n_unique_counted += g->local[r].n_unique_counted; tot.total_count_filtered += g->local[r].total_count_filtered; }
	// 5. all-gather of the survivors, padded to the largest share
	auto gather_step = [&]() -> clb_status {
		const uint64_t pad = std::max<uint64_t>(1, *std::max_element(g->filt.begin(), g->filt.end()));
		uint64_t total = 0; for (uint64_t x : g->filt) total += x;
		uint64_t *my_k = nullptr, *all_k = nullptr, *uni_k = nullptr; uint32_t *my_c = nullptr, *all_c = nullptr, *uni_c = nullptr;
		G_CUDA(mem.get(&my_k, pad)); G_CUDA(mem.get(&my_c, pad)); G_CUDA(mem.get(&all_k, pad * N)); G_CUDA(mem.get(&all_c, pad * N)); G_CUDA(mem.get(&uni_k, total)); G_CUDA(mem.get(&uni_c, total));
		G_CUDA(cudaMemsetAsync(my_k, 0, pad * 8, s)); G_CUDA(cudaMemsetAsync(my_c, 0, pad * 4, s));
		G_CUDA(cudaStreamSynchronize(s));
		uint64_t got = 0;
		if (g->filt[rank]) { G_CLB(clb_filter_list(c, my_k, my_c, g->filt[rank], &got, 1)); G_CLB(clb_synchronize(c)); }
		G_NCCL(ncclAllGather(my_k, all_k, pad, ncclUint64, g->comm[rank], s));
		G_NCCL(ncclAllGather(my_c, all_c, pad, ncclUint32, g->comm[rank], s));
		uint64_t at = 0;
		for (uint32_t r = 0; r < N; ++r) if (g->filt[r]) {
			G_CUDA(cudaMemcpyAsync(uni_k + at, all_k + r * pad, g->filt[r] * 8, cudaMemcpyDeviceToDevice, s));
			G_CUDA(cudaMemcpyAsync(uni_c + at, all_c + r * pad, g->filt[r] * 4, cudaMemcpyDeviceToDevice, s));
			at += g->filt[r];
		}
		G_CUDA(cudaStreamSynchronize(s));
		G_CLB(clb_filter_import(c, uni_k, uni_c, total, &tot, 1));
		G_CLB(clb_synchronize(c));
		return CLB_OK;
	};
	st = gather_step();
	g->barrier(); G_CHECK_ALL();
	*global_stats = tot;
	return st;
}

clb_status clb_group_exchange_reference_reads(clb_group* g, uint32_t rank, const uint8_t* sampled_local, const uint32_t* lengths_local, uint32_t n_local, uint32_t* n_context_out)
{
	if (!g || rank >= g->n || (n_local && (!sampled_local || !lengths_local))) return CLB_ERR_BAD_ARG;
	const uint32_t N = g->n; clb_ctx* c = g->ctx[rank];
	cudaSetDevice(g->dev[rank]);
	cudaStream_t s = g->stream[rank];
	DevMem mem;
	std::vector<uint32_t> ids; std::vector<uint64_t> lens;
	uint64_t n_bases = 0;
	auto pick_step = [&]() -> clb_status {
		std::vector<uint8_t> has_n(std::max<uint32_t>(1, n_local));
		if (n_local) G_CLB(clb_reads_have_n(c, has_n.data()));
		for (uint32_t i = 0; i < n_local; ++i) if (sampled_local[i] && !has_n[i]) { ids.push_back(i); lens.push_back(lengths_local[i]); n_bases += lengths_local[i]; }
		g->ref_reads[rank] = ids.size(); g->ref_bases[rank] = n_bases;
		return CLB_OK;
	};
	clb_status st = pick_step();
	g->barrier(); G_CHECK_ALL();
	uint32_t n_ctx = 0;
	auto gather_step = [&]() -> clb_status {
		const uint64_t pad_r = std::max<uint64_t>(1, *std::max_element(g->ref_reads.begin(), g->ref_reads.end()));
		const uint64_t pad_b = (std::max<uint64_t>(1, *std::max_element(g->ref_bases.begin(), g->ref_bases.end())) + 15) & ~15ull;
		uint64_t *my_lens = nullptr, *all_lens = nullptr; uint8_t *my_bases = nullptr, *all_bases = nullptr;
		G_CUDA(mem.get(&my_lens, pad_r)); G_CUDA(mem.get(&all_lens, pad_r * N)); G_CUDA(mem.get(&my_bases, pad_b)); G_CUDA(mem.get(&all_bases, pad_b * N));
		G_CUDA(cudaMemsetAsync(my_lens, 0, pad_r * 8, s));
		if (!lens.empty()) G_CUDA(cudaMemcpyAsync(my_lens, lens.data(), lens.size() * 8, cudaMemcpyHostToDevice, s));
		G_CUDA(cudaStreamSynchronize(s));
		if (n_bases) { G_CLB(clb_reads_export(c, ids.data(), (uint32_t)ids.size(), my_bases, pad_b, 1)); G_CLB(clb_synchronize(c)); }
		G_NCCL(ncclAllGather(my_lens, all_lens, pad_r, ncclUint64, g->comm[rank], s));
		G_NCCL(ncclAllGather(my_bases, all_bases, pad_b, ncclUint8, g->comm[rank], s));
		G_CUDA(cudaStreamSynchronize(s));
		uint64_t ctx_reads = 0, ctx_bases = 0;
		for (uint32_t r = 0; r < rank; ++r) { ctx_reads += g->ref_reads[r]; ctx_bases += g->ref_bases[r]; }
		n_ctx = (uint32_t)ctx_reads;
		if (!ctx_reads) return CLB_OK;
		// the reference reads of the ranks before mine, back to back, + their offsets
		std::vector<uint64_t> h_lens(pad_r * rank), off(ctx_reads + 1, 0);
		G_CUDA(cudaMemcpy(h_lens.data(), all_lens, pad_r * rank * 8, cudaMemcpyDeviceToHost));
		uint8_t* d_bases = nullptr; uint64_t* d_off = nullptr;
		G_CUDA(mem.get(&d_bases, ctx_bases + 16)); G_CUDA(mem.get(&d_off, ctx_reads + 1));
		uint64_t k = 0, at = 0;
		for (uint32_t r = 0; r < rank; ++r) {
			for (uint64_t i = 0; i < g->ref_reads[r]; ++i, ++k) off[k + 1] = off[k] + h_lens[r * pad_r + i];
			if (g->ref_bases[r]) G_CUDA(cudaMemcpyAsync(d_bases + at, all_bases + r * pad_b, g->ref_bases[r], cudaMemcpyDeviceToDevice, s));
			at += g->ref_bases[r];
		}
		if (off[ctx_reads] != ctx_bases) return gfail(g, rank, CLB_ERR_STATE, "reference-read lengths do not add up");
		G_CUDA(cudaMemcpyAsync(d_off, off.data(), (ctx_reads + 1) * 8, cudaMemcpyHostToDevice, s));
		G_CUDA(cudaStreamSynchronize(s));
		G_CLB(clb_append_context_reads(c, d_bases, d_off, (uint32_t)ctx_reads, 1));
		G_CLB(clb_synchronize(c));
		return CLB_OK;
	};
	if (st == CLB_OK) st = gather_step();
	g->barrier(); G_CHECK_ALL();
	if (n_context_out) *n_context_out = n_ctx;
	return st;
}

} // extern "C"
