// stage3_dna.cu — DNA / edit-script stream on device.
//
// Kept from the reference: the event model of CDNACoder (dna_model.h: which symbols a read's tuples turn into and the context
// of each) and the arithmetic of its range coder (sub_rc.h:83-201: 64-bit low / range, carry-less renormalisation byte by
// byte, 8-byte flush).  Replaced: the adaptive models whose state runs through the whole file (entr_read.h:56-80), which make
// the reference's stream one serial chain.  Here pass 1 counts (family, context, symbol) triples of all reads with atomics,
// the host turns the counts into static 12-bit frequency tables (rare contexts of the two big families share a fallback
// table; metadata-sized work) and writes them into the container header, pass 2 codes every read pack with 64 independent
// range-coder lanes (lane l takes reads l, l+64, ... of its pack), one thread per lane.  Native container "DB01"; decoder:
// oracle/stage3_dna.c (it rebuilds the reads, which is the round-trip proof).
#include "ctx.h"
#include "dna_model.h"
#include "static_tables.h"
#include "range_sink.cuh"
#include <algorithm>
#include <cstring>
#include <vector>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <ctime>

namespace clb {

constexpr uint32_t DB_LANES = 64, DB_MIN_CTX = 64;

using HistSink = HistSinkT<DnaModel>;
using RangeSink = RangeSinkT<DnaModel>;

struct DArgs {
	DnaReads R; DnaModel M;
	const uint32_t* pack_first; uint32_t n_packs; uint32_t n_reads;
	uint32_t* hist; const uint32_t* tab;
};

CLB_D uint32_t lane_flag_ctx(const DnaReads& R, uint32_t r, uint32_t pack_start)
{
	uint32_t c = 0;
	for (int k = 4; k >= 1; --k) { const long long rr = (long long)r - (long long)DB_LANES * k; if (rr >= (long long)pack_start) c = ((c << 2) + read_flag_of(R, (uint32_t)rr)) & 0xff; }
	return c;
}

// pass 1: one thread per read
__global__ void __launch_bounds__(128) k_d_count(DArgs a)
{
	const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= a.n_reads) return;
	uint32_t lo = 0, hi = a.n_packs;              // pack of the read: last pack_first <= r
	while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (a.pack_first[mid] <= r) lo = mid; else hi = mid; }
	HistSink s{a.hist, &a.M};
	dna_walk(a.M, a.R, r, lane_flag_ctx(a.R, r, a.pack_first[lo]), s);
}

struct DEnc { uint32_t* lane_bytes; const uint64_t* dst_off; const uint64_t* pack_hdr_off; uint8_t* out; const uint32_t* lane_cap; uint32_t* overflow; };

// slot of a lane in the temp: the tuple bytes of its reads + 64 (a coded read is a fraction of its tuples; a lane that would not
// fit raises `overflow` and the container is written by the two-walk path instead)
__global__ void __launch_bounds__(64) k_d_lane_cap(DArgs a, uint32_t* __restrict__ cap)
{
	const uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
	if (li >= a.n_packs * DB_LANES) return;
	const uint32_t p = li / DB_LANES, l = li % DB_LANES;
	uint64_t b = 64;
	for (uint32_t r = a.pack_first[p] + l; r < a.pack_first[p + 1]; r += DB_LANES) b += a.R.es_off[a.R.first + r + 1] - a.R.es_off[a.R.first + r];
	cap[li] = (uint32_t)min(b, (uint64_t)0xFFFFFFF0u);
}
// lane slots -> their final place in the container, pack headers; one warp per lane
__global__ void __launch_bounds__(128) k_d_compact(DArgs a, DEnc e, const uint8_t* __restrict__ tmp, const uint64_t* __restrict__ slot_off)
{
	const uint32_t li = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, t = threadIdx.x & 31;
	if (li >= a.n_packs * DB_LANES) return;
	const uint32_t nb = e.lane_bytes[li];
	const uint8_t* src = tmp + slot_off[li]; uint8_t* dst = e.out + e.dst_off[li];
	for (uint32_t k = t; k < nb; k += 32) dst[k] = src[k];
	if (t == 0) {
		const uint32_t p = li / DB_LANES, l = li % DB_LANES;
		uint8_t* h = e.out + e.pack_hdr_off[p];
		h[4 + 4 * l] = (uint8_t)nb; h[5 + 4 * l] = (uint8_t)(nb >> 8); h[6 + 4 * l] = (uint8_t)(nb >> 16); h[7 + 4 * l] = (uint8_t)(nb >> 24);
		if (l == 0) { const uint32_t np = a.pack_first[p + 1] - a.pack_first[p]; h[0] = (uint8_t)np; h[1] = (uint8_t)(np >> 8); h[2] = (uint8_t)(np >> 16); h[3] = (uint8_t)(np >> 24); }
	}
}

// pass 2: one thread per (pack, lane).  MODE 0 sizes the lane streams (the coder's output length does not depend on where it
// is stored), MODE 1 writes them — and the pack headers — at their final place in the container, MODE 2 writes them into the
// lane's temp slot and reports the size (one walk instead of two; k_d_compact moves them).
template <int MODE>
__global__ void __launch_bounds__(64) k_d_encode(DArgs a, DEnc e)
{
	constexpr bool WRITE = MODE == 1;
	const uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
	if (li >= a.n_packs * DB_LANES) return;
	const uint32_t p = li / DB_LANES, l = li % DB_LANES;
	const uint32_t r0 = a.pack_first[p], r1 = a.pack_first[p + 1];
	RangeSink s{a.tab, &a.M, MODE == 0 ? nullptr : e.out + e.dst_off[li], 0, 0, 0};
	if (MODE == 2) s.cap = e.lane_cap[li];
	s.start();
	uint32_t fctx = 0;
	for (uint32_t r = r0 + l; r < r1; r += DB_LANES) {
		dna_walk(a.M, a.R, r, fctx, s);
		fctx = ((fctx << 2) + read_flag_of(a.R, r)) & 0xff;
	}
	s.end();
	if (MODE == 2 && s.n > s.cap) atomicExch(e.overflow, 1u);
	if (!WRITE) { e.lane_bytes[li] = (uint32_t)s.n; return; }
	uint8_t* h = e.out + e.pack_hdr_off[p];
	const uint32_t nb = (uint32_t)s.n;
	h[4 + 4 * l] = (uint8_t)nb; h[5 + 4 * l] = (uint8_t)(nb >> 8); h[6 + 4 * l] = (uint8_t)(nb >> 16); h[7 + 4 * l] = (uint8_t)(nb >> 24);
	if (l == 0) { const uint32_t np = r1 - r0; h[0] = (uint8_t)np; h[1] = (uint8_t)(np >> 8); h[2] = (uint8_t)(np >> 16); h[3] = (uint8_t)(np >> 24); }
}

// ------------------------------------------------------------------------------------------------ host side
struct DTrace {            // CLB_S2_TRACE=1: wall time of every phase (synchronising; debugging aid)
	bool on; cudaStream_t s; double t0;
	static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
	explicit DTrace(cudaStream_t st) : on(std::getenv("CLB_S2_TRACE") != nullptr), s(st), t0(now()) {}
	void mark(const char* w) { if (!on) return; cudaStreamSynchronize(s); const double t = now(); fprintf(stderr, "[s3d] %-28s %9.3f ms\n", w, t - t0); t0 = t; }
};
clb_status s3_dna_encode(clb_ctx* c, uint32_t level, const uint32_t* pack_sizes, uint32_t n_packs)
{
	cudaStream_t s = c->stream;
	DTrace tr(s);
	const uint64_t nc = c->n_context, n = c->n_reads - nc;      // context reads are not coded
	if (!c->enc_done) return fail(c, CLB_ERR_STATE, "clb_dna_encode needs the tuples (clb_encode)");
	if (c->dna_done) return fail(c, CLB_ERR_STATE, "clb_dna_encode called twice");
	if (level < 1 || level > 3) return fail(c, CLB_ERR_BAD_ARG, "clb_dna_encode: level must be 1, 2 or 3");
	if (c->prm.max_candidates > 32) return fail(c, CLB_ERR_BAD_ARG, "clb_dna_encode: max_candidates above 32 is not supported");
	const DnaModel M = make_dna_model(level, c->prm.max_candidates);
	std::vector<uint32_t> pack_first{0};
	if (pack_sizes) {
		uint64_t at = 0;
		for (uint32_t i = 0; i < n_packs; ++i) { at += pack_sizes[i]; if (pack_sizes[i]) pack_first.push_back((uint32_t)at); }
		if (at != n) return fail(c, CLB_ERR_BAD_ARG, "pack_sizes do not sum to the number of reads");
	} else {
		uint64_t bytes = 0;
		for (uint64_t i = 0; i < n; ++i) { bytes += (uint64_t)c->h_rd_len[nc + i] + 1; if (bytes >= (2u << 21)) { bytes = 0; pack_first.push_back((uint32_t)(i + 1)); } }
		if (pack_first.back() != n) pack_first.push_back((uint32_t)n);
	}
	const uint32_t np = (uint32_t)pack_first.size() - 1;
	struct Tmp { std::vector<void*> v; cudaStream_t s; ~Tmp() { for (void* p : v) dev_free_async(p, s); } } tmp{{}, s};
	auto dalloc = [&](void** p, uint64_t bytes, Tmp& t) { cudaError_t e = dev_malloc(p, bytes ? bytes : 1, s); if (e == cudaSuccess) t.v.push_back(*p); return e; };
	const uint64_t n_entries = M.base[F_COUNT];
	uint32_t* d_hist = nullptr; uint32_t* d_pack_first = nullptr;
	CLB_CUDA(c, dalloc((void**)&d_hist, sizeof(uint32_t) * n_entries, tmp));
	CLB_CUDA(c, dalloc((void**)&d_pack_first, sizeof(uint32_t) * (np + 1), tmp));
	CLB_CUDA(c, cudaMemsetAsync(d_hist, 0, sizeof(uint32_t) * n_entries, s));
	CLB_CUDA(c, cudaMemcpyAsync(d_pack_first, pack_first.data(), sizeof(uint32_t) * (np + 1), cudaMemcpyHostToDevice, s));
	DArgs a{};
	a.R = DnaReads{c->pk.p, c->rd_start.p, c->rd_len.p, c->d_ref_to_read, c->es.p, c->es_off, (uint32_t)nc};
	a.M = M; a.pack_first = d_pack_first; a.n_packs = np; a.n_reads = (uint32_t)n; a.hist = d_hist;
	if (n) { CLB_TIMED(c, K_DNA, (k_d_count<<<(uint32_t)((n + 127) / 128), 128, 0, s>>>(a))); CLB_LAUNCH_CHECK(c, "k_d_count"); }
	tr.mark("k_d_count");
	// ---- counts -> static tables + container header (host, metadata-sized) ----
	std::vector<uint32_t> hist(n_entries);
	CLB_CUDA(c, cudaMemcpyAsync(hist.data(), d_hist, sizeof(uint32_t) * n_entries, cudaMemcpyDeviceToHost, s));
	CLB_CUDA(c, cudaStreamSynchronize(s));
	std::vector<uint32_t> tab(n_entries, 0);
	std::vector<uint8_t> hdr;
	hdr.insert(hdr.end(), {'D', 'B', '0', '1'}); st_put(hdr, level); st_put(hdr, c->prm.max_candidates); st_put(hdr, (uint64_t)n); st_put(hdr, np); st_put(hdr, (uint32_t)nc);
	st_build_tables(M, F_COUNT, hist, tab, hdr, DB_MIN_CTX);
	if (std::getenv("CLB_S3_BITS"))      // where the bits go: events, cost under the static tables, empirical context entropy
		for (uint32_t f = 0; f < F_COUNT; ++f) {
			const uint32_t A = M.A[f]; const uint64_t n_ctx = 1ull << M.cbits[f];
			const uint32_t* h = hist.data() + M.base[f]; const uint32_t* tb = tab.data() + M.base[f];
			double ev = 0, cost = 0, ent = 0;
			for (uint64_t x = 0; x < n_ctx; ++x) {
				uint64_t t = 0; for (uint32_t k = 0; k < A; ++k) t += h[x * A + k];
				if (!t) continue;
				for (uint32_t k = 0; k < A; ++k) if (h[x * A + k]) { const double cn = h[x * A + k]; ev += cn; cost += cn * -std::log2((tb[x * A + k] & 0xffff) / 4096.0); ent += cn * -std::log2(cn / (double)t); }
			}
			fprintf(stderr, "[s3d] family %2u: %12.0f events, %12.0f bytes coded, %12.0f bytes empirical\n", f, ev, cost / 8, ent / 8);
		}
	uint32_t* d_tab = nullptr;
	CLB_CUDA(c, dalloc((void**)&d_tab, sizeof(uint32_t) * n_entries, tmp));
	CLB_CUDA(c, cudaMemcpyAsync(d_tab, tab.data(), sizeof(uint32_t) * n_entries, cudaMemcpyHostToDevice, s));
	a.tab = d_tab;
	tr.mark("tables (host)");
	// ---- pass 2: code every lane into a temp slot (one walk), lay the container out, compact; if the temp cannot be had or a
	// lane outgrows its slot: size every lane stream (walk), lay out, write (second walk) ----
	const uint32_t nl = np * DB_LANES;
	uint32_t* d_bytes = nullptr; uint64_t* d_dst = nullptr; uint64_t* d_phdr = nullptr; uint32_t* d_cap = nullptr; uint64_t* d_slot = nullptr; uint32_t* d_ovf = nullptr;
	CLB_CUDA(c, dalloc((void**)&d_bytes, sizeof(uint32_t) * nl, tmp)); CLB_CUDA(c, dalloc((void**)&d_dst, sizeof(uint64_t) * nl, tmp)); CLB_CUDA(c, dalloc((void**)&d_phdr, sizeof(uint64_t) * np, tmp));
	CLB_CUDA(c, dalloc((void**)&d_cap, sizeof(uint32_t) * (nl + 1), tmp)); CLB_CUDA(c, dalloc((void**)&d_slot, sizeof(uint64_t) * (nl + 1), tmp));
	d_ovf = d_cap + nl;
	DEnc e{d_bytes, d_dst, d_phdr, nullptr, d_cap, d_ovf};
	uint8_t* d_tmp = nullptr; bool one_walk = false;
	if (nl && !std::getenv("CLB_DNA_TWO_WALKS")) {
		CLB_CUDA(c, cudaMemsetAsync(d_ovf, 0, sizeof(uint32_t), s));
		CLB_TIMED(c, K_DNA, (k_d_lane_cap<<<(nl + 63) / 64, 64, 0, s>>>(a, d_cap))); CLB_LAUNCH_CHECK(c, "k_d_lane_cap");
		uint64_t tmp_bytes = 0;
		clb_status st = exclusive_scan(c, d_cap, nl, d_slot, &tmp_bytes); if (st != CLB_OK) return st;
		if (tmp_bytes + (8ull << 30) < dev_mem_available() && dev_malloc((void**)&d_tmp, tmp_bytes + 16, s) == cudaSuccess) {
			e.out = d_tmp; e.dst_off = d_slot;
			CLB_TIMED(c, K_DNA, (k_d_encode<2><<<(nl + 63) / 64, 64, 0, s>>>(a, e))); CLB_LAUNCH_CHECK(c, "k_d_encode<temp>");
			uint32_t ovf = 0;
			CLB_CUDA(c, cudaMemcpyAsync(&ovf, d_ovf, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
			CLB_CUDA(c, cudaStreamSynchronize(s));
			one_walk = ovf == 0;
			e.dst_off = d_dst;
		} else cudaGetLastError();
	}
	if (nl && !one_walk) { CLB_TIMED(c, K_DNA, (k_d_encode<0><<<(nl + 63) / 64, 64, 0, s>>>(a, e))); CLB_LAUNCH_CHECK(c, "k_d_encode<size>"); }
	std::vector<uint32_t> bytes(nl);
	CLB_CUDA(c, cudaMemcpyAsync(bytes.data(), d_bytes, sizeof(uint32_t) * nl, cudaMemcpyDeviceToHost, s));
	CLB_CUDA(c, cudaStreamSynchronize(s));
	tr.mark(one_walk ? "k_d_encode<temp>" : "k_d_encode<size>");
	uint64_t out_at = hdr.size();
	std::vector<uint64_t> dst(nl), phdr(np);
	for (uint32_t p = 0; p < np; ++p) {
		phdr[p] = out_at; out_at += 4 + 4 * DB_LANES;
		for (uint32_t l = 0; l < DB_LANES; ++l) { dst[(size_t)p * DB_LANES + l] = out_at; out_at += bytes[(size_t)p * DB_LANES + l]; }
	}
	CLB_CUDA(c, c->ds.reserve(out_at + 16, s, false));
	CLB_CUDA(c, cudaMemcpyAsync(c->ds.p, hdr.data(), hdr.size(), cudaMemcpyHostToDevice, s));
	CLB_CUDA(c, cudaMemcpyAsync(d_dst, dst.data(), sizeof(uint64_t) * nl, cudaMemcpyHostToDevice, s));
	CLB_CUDA(c, cudaMemcpyAsync(d_phdr, phdr.data(), sizeof(uint64_t) * np, cudaMemcpyHostToDevice, s));
	e.out = c->ds.p;
	if (nl && one_walk) { CLB_TIMED(c, K_DNA, (k_d_compact<<<(nl * 32 + 127) / 128, 128, 0, s>>>(a, e, d_tmp, d_slot))); CLB_LAUNCH_CHECK(c, "k_d_compact"); }
	else if (nl) { CLB_TIMED(c, K_DNA, (k_d_encode<1><<<(nl + 63) / 64, 64, 0, s>>>(a, e))); CLB_LAUNCH_CHECK(c, "k_d_encode<write>"); }
	CLB_CUDA(c, cudaStreamSynchronize(s));
	if (d_tmp) dev_free(d_tmp, s);
	tr.mark(one_walk ? "k_d_compact" : "k_d_encode<write>");
	c->ds_total = out_at;
	c->ds_header = hdr.size();
	c->dna_done = true;
	return CLB_OK;
}

} // namespace clb
