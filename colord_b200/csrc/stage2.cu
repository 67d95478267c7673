// stage2.cu — stage 2 on device.  This file: the edit-script batch (rows E6/E7 of SURVEY.md §8): tasks are
// classified by the number of 64-row blocks into lane groups of 1/2/4/8/32 and launched class by class.
#include "ctx.h"
#include "align.cuh"
#include <algorithm>
#include <cstring>
#include <numeric>

namespace clb {

struct EsTask {
	unsigned long long ref_off, enc_off;     // into the staged symbol buffer (each part is followed by one more readable byte)
	unsigned long long out_off, scratch_off;
	uint32_t rl, el, kind, pad;
};

template <int GROUP>
__global__ void __launch_bounds__(ALIGN_THREADS) k_edit_scripts(const EsTask* __restrict__ tasks, const uint32_t* __restrict__ list, uint32_t n_list,
	const uint8_t* __restrict__ seqs, char* __restrict__ out, uint32_t* __restrict__ out_len, uint8_t* __restrict__ scratch)
{
	const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t slot = tid / GROUP;
	if (slot >= n_list) return;                 // whole groups leave together
	const uint32_t ti = list[slot];
	const EsTask t = tasks[ti];
	__shared__ uint64_t s_peq[4 * ALIGN_THREADS];
	Aligner<GROUP> A;
	A.peq = s_peq + threadIdx.x;
	A.gl = threadIdx.x & (GROUP - 1);
	const uint32_t lane = threadIdx.x & 31;
	A.gmask = GROUP == 32 ? 0xffffffffu : (((1u << GROUP) - 1u) << (lane & ~(uint32_t)(GROUP - 1)));
	A.scratch = scratch + t.scratch_off;
	long long q, tt;
	align_task_dims(t.rl, t.el, t.kind, &q, &tt);
	A.lay = align_scratch_layout(q, tt);
	const SeqView ref{seqs + t.ref_off, 1}, enc{seqs + t.enc_off, 1};
	const uint32_t n = edit_script_task<GROUP>(A, ref, t.rl, enc, t.el, t.kind, out + t.out_off);
	if (A.gl == 0) out_len[ti] = n;
}

static int group_of(long long q)
{
	const long long B = (q + 63) / 64;
	return B <= 1 ? 1 : B <= 2 ? 2 : B <= 4 ? 4 : B <= 8 ? 8 : 32;
}

// Runs all tasks (device arrays d_tasks[n] with ref/enc/out offsets filled; scratch offsets are assigned here).
// h_tasks is the host copy.  out_len[n] on device receives the script lengths.
clb_status run_edit_scripts(clb_ctx* c, std::vector<EsTask>& h_tasks, const uint8_t* d_seqs, char* d_out, uint32_t* d_out_len)
{
	cudaStream_t s = c->stream;
	const size_t n = h_tasks.size();
	if (!n) return CLB_OK;
	const unsigned long long budget = 6ull << 30;          // scratch bytes per wave
	std::vector<uint32_t> order(n);
	std::iota(order.begin(), order.end(), 0u);
	std::vector<long long> cost(n);
	std::vector<int> grp(n);
	std::vector<unsigned long long> need(n);
	for (size_t i = 0; i < n; ++i) {
		long long q, t; align_task_dims(h_tasks[i].rl, h_tasks[i].el, h_tasks[i].kind, &q, &t);
		grp[i] = group_of(q); cost[i] = q * t; need[i] = align_scratch_layout(q, t).total;
	}
	// similar tasks next to each other: by group, then by cost (lanes of a warp finish together)
	std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return grp[a] != grp[b] ? grp[a] < grp[b] : cost[a] < cost[b]; });
	EsTask* d_tasks = nullptr; uint32_t* d_list = nullptr; uint8_t* d_scratch = nullptr;
	unsigned long long scratch_cap = 0;
	CLB_CUDA(c, dev_malloc((void**)&d_tasks, sizeof(EsTask) * n, s));
	CLB_CUDA(c, dev_malloc((void**)&d_list, sizeof(uint32_t) * n, s));
	clb_status st = CLB_OK;
	size_t pos = 0;
	while (pos < n && st == CLB_OK) {
		// one wave: same group, scratch within the budget
		const int g = grp[order[pos]];
		size_t end = pos; unsigned long long used = 0;
		while (end < n && grp[order[end]] == g && (end == pos || used + need[order[end]] <= budget)) { h_tasks[order[end]].scratch_off = used; used += need[order[end]]; ++end; }
		if (used > scratch_cap) {
			if (d_scratch) { cudaStreamSynchronize(s); dev_free(d_scratch, s); }
			scratch_cap = std::max(used, std::min<unsigned long long>(budget, scratch_cap * 2));
			cudaError_t e = dev_malloc((void**)&d_scratch, scratch_cap, s);
			if (e != cudaSuccess) { st = cuda_fail(c, e, "edit-script scratch"); break; }
		}
		const uint32_t m = (uint32_t)(end - pos);
		// tasks of this wave (scratch offsets changed) and the list
		std::vector<EsTask> wave(m); std::vector<uint32_t> ids(m);
		for (uint32_t i = 0; i < m; ++i) { ids[i] = order[pos + i]; }
		cudaError_t e = cudaSuccess;
		for (uint32_t i = 0; i < m && e == cudaSuccess; ++i) { /* tasks are scattered by id: copy individually only when few; else bulk below */ }
		// bulk: upload the whole task array once per wave (simple; waves are few)
		e = cudaMemcpyAsync(d_tasks, h_tasks.data(), sizeof(EsTask) * n, cudaMemcpyHostToDevice, s);
		if (e == cudaSuccess) e = cudaMemcpyAsync(d_list, ids.data(), sizeof(uint32_t) * m, cudaMemcpyHostToDevice, s);
		if (e != cudaSuccess) { st = cuda_fail(c, e, "edit-script upload"); break; }
		const uint32_t threads = 128;
		const uint64_t total_threads = (uint64_t)m * g;
		const uint32_t blocks = (uint32_t)((total_threads + threads - 1) / threads);
		prof_begin(c, K_ALIGN);
		switch (g) {
		case 1: k_edit_scripts<1><<<blocks, threads, 0, s>>>(d_tasks, d_list, m, d_seqs, d_out, d_out_len, d_scratch); break;
		case 2: k_edit_scripts<2><<<blocks, threads, 0, s>>>(d_tasks, d_list, m, d_seqs, d_out, d_out_len, d_scratch); break;
		case 4: k_edit_scripts<4><<<blocks, threads, 0, s>>>(d_tasks, d_list, m, d_seqs, d_out, d_out_len, d_scratch); break;
		case 8: k_edit_scripts<8><<<blocks, threads, 0, s>>>(d_tasks, d_list, m, d_seqs, d_out, d_out_len, d_scratch); break;
		default: k_edit_scripts<32><<<blocks, threads, 0, s>>>(d_tasks, d_list, m, d_seqs, d_out, d_out_len, d_scratch); break;
		}
		prof_end(c);
		++c->launches;
		e = cudaGetLastError();
		if (e == cudaSuccess) e = cudaStreamSynchronize(s);       // ids/wave vectors die here; scratch is reused by the next wave
		if (e != cudaSuccess) { st = cuda_fail(c, e, "k_edit_scripts"); break; }
		pos = end;
	}
	dev_free(d_tasks, s); dev_free(d_list, s); if (d_scratch) dev_free(d_scratch, s);
	return st;
}

// Test / host entry: symbols 0..3 in `seqs` (host), per task a (ref_off, ref_len, enc_off, enc_len, kind); the byte after
// every part must be readable (it is what follows the part in its read; 255 at a read's end).
clb_status s2_edit_scripts(clb_ctx* c, const uint8_t* seqs, uint64_t n_seq_bytes, const uint64_t* ref_off, const uint32_t* ref_len,
	const uint64_t* enc_off, const uint32_t* enc_len, const uint32_t* kind, uint64_t n, uint64_t* out_off, char* out, uint64_t cap)
{
	std::vector<EsTask> tasks(n);
	unsigned long long o = 0;
	for (uint64_t i = 0; i < n; ++i) {
		if (kind[i] > 2 || ref_off[i] + ref_len[i] >= n_seq_bytes + 0 || enc_off[i] + enc_len[i] >= n_seq_bytes) return fail(c, CLB_ERR_BAD_ARG, "clb_edit_scripts: part (plus its following byte) outside seqs");
		tasks[i] = EsTask{ref_off[i], enc_off[i], o, 0, ref_len[i], enc_len[i], kind[i], 0};
		o += (unsigned long long)ref_len[i] + enc_len[i] + 2;
	}
	uint8_t* d_seqs = nullptr; char* d_out = nullptr; uint32_t* d_len = nullptr;
	CLB_CUDA(c, dev_malloc((void**)&d_seqs, n_seq_bytes + 16, c->stream));
	CLB_CUDA(c, dev_malloc((void**)&d_out, o + 16, c->stream));
	CLB_CUDA(c, dev_malloc((void**)&d_len, sizeof(uint32_t) * (n + 1), c->stream));
	CLB_CUDA(c, cudaMemcpyAsync(d_seqs, seqs, n_seq_bytes, cudaMemcpyHostToDevice, c->stream));
	clb_status st = run_edit_scripts(c, tasks, d_seqs, d_out, d_len);
	std::vector<uint32_t> len(n); std::vector<char> raw(o + 1);
	if (st == CLB_OK) {
		cudaError_t e = cudaMemcpyAsync(len.data(), d_len, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, c->stream);
		if (e == cudaSuccess) e = cudaMemcpyAsync(raw.data(), d_out, o, cudaMemcpyDeviceToHost, c->stream);
		if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
		if (e != cudaSuccess) st = cuda_fail(c, e, "clb_edit_scripts download");
	}
	dev_free(d_seqs, c->stream); dev_free(d_out, c->stream); dev_free(d_len, c->stream);
	if (st != CLB_OK) return st;
	uint64_t w = 0;
	for (uint64_t i = 0; i < n; ++i) {
		out_off[i] = w;
		if (w + len[i] > cap) return fail(c, CLB_ERR_CAPACITY, "clb_edit_scripts: output buffer too small");
		std::memcpy(out + w, raw.data() + tasks[i].out_off, len[i]);
		w += len[i];
	}
	out_off[n] = w;
	return CLB_OK;
}

} // namespace clb
