// stage3_qorg.cu — lossless quality stream ("-q org") on device.
//
// Kept from the reference: the context model of CQualityCoder::encode_original (qorg_model.h) and the arithmetic of its range
// coder (sub_rc.h:83-201).  Replaced: the adaptive 96-symbol models + single coder chain of CEntrComprQuals::Compress
// (entr_qual.h:100-126).  The context of a base depends only on the input (the two previous qualities, four bases, the tuple
// flags), so pass 1 counts (context, symbol) pairs of all bases in parallel (one CTA per read), the host turns the counts into
// static 12-bit tables (contexts seen rarely fall back to the table of the two previous symbols; metadata-sized), pass 2
// codes every read pack with 64 independent range-coder lanes (lane l takes reads l, l+64, ... of its pack), first sizing
// and then writing the lane streams at their final place.  Native container "QO01"; CPU twin + decoder: oracle/stage3_qorg.c.
#include "ctx.h"
#include "qorg_model.h"
#include "static_tables.h"
#include "range_sink.cuh"
#include <vector>
#include <cstring>

namespace clb {

constexpr uint32_t QO_LANES = 64, QO_MIN_CTX = 256;

struct QoArgs {
	QoReads R; QoModel M;
	const uint32_t* pack_first; uint32_t n_packs; uint32_t n_reads;
	uint32_t* hist; const uint32_t* tab; uint32_t* bad;
};

// pass 1: one CTA per read
__global__ void __launch_bounds__(256) k_qo_count(QoArgs a)
{
	const uint32_t r = blockIdx.x;
	const uint32_t n = a.R.rd_len[r]; const uint64_t rs = a.R.rd_start[r];
	const uint8_t* q = a.R.quals + a.R.qoff[r];
	const uint8_t* fl = a.R.flags ? a.R.flags + a.R.qoff[r] : nullptr;
	bool bad = false;
	for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) bad |= q[i] < 33u || q[i] > 128u;
	if (__syncthreads_or(bad)) { if (threadIdx.x == 0) atomicExch(a.bad, 1u); return; }      // outside phred+33 of 0..95
	for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
		atomicAdd(&a.hist[st_entry(a.M, 0, qorg_context(a.M, a.R, rs, n, q, fl, i), q[i] - 33u)], 1u);
}

struct QoEnc { uint32_t* lane_bytes; const uint64_t* dst_off; const uint64_t* pack_hdr_off; uint8_t* out; };

// pass 2: one thread per (pack, lane); WRITE = false sizes the lane streams, WRITE = true writes them and the pack headers
template <bool WRITE>
__global__ void __launch_bounds__(64) k_qo_encode(QoArgs a, QoEnc e)
{
	const uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
	if (li >= a.n_packs * QO_LANES) return;
	const uint32_t p = li / QO_LANES, l = li % QO_LANES;
	const uint32_t r0 = a.pack_first[p], r1 = a.pack_first[p + 1];
	RangeSinkT<QoModel> s{a.tab, &a.M, WRITE ? e.out + e.dst_off[li] : nullptr, 0, 0, 0};
	s.start();
	for (uint32_t r = r0 + l; r < r1; r += QO_LANES) {
		const uint32_t n = a.R.rd_len[r]; const uint64_t rs = a.R.rd_start[r];
		const uint8_t* q = a.R.quals + a.R.qoff[r];
		const uint8_t* fl = a.R.flags ? a.R.flags + a.R.qoff[r] : nullptr;
		for (uint32_t i = 0; i < n; ++i) s.put(0, qorg_context(a.M, a.R, rs, n, q, fl, i), q[i] - 33u);
	}
	s.end();
	if (!WRITE) { e.lane_bytes[li] = (uint32_t)s.n; return; }
	uint8_t* h = e.out + e.pack_hdr_off[p];
	const uint32_t nb = (uint32_t)s.n;
	h[4 + 4 * l] = (uint8_t)nb; h[5 + 4 * l] = (uint8_t)(nb >> 8); h[6 + 4 * l] = (uint8_t)(nb >> 16); h[7 + 4 * l] = (uint8_t)(nb >> 24);
	if (l == 0) { const uint32_t np = r1 - r0; h[0] = (uint8_t)np; h[1] = (uint8_t)(np >> 8); h[2] = (uint8_t)(np >> 16); h[3] = (uint8_t)(np >> 24); }
}

clb_status s3_qual_encode_original(clb_ctx* c, uint32_t source, uint32_t level, const uint8_t* quals, const uint64_t* offsets, int on_device,
	const uint32_t* pack_sizes, uint32_t n_packs)
{
	cudaStream_t s = c->stream3;
	const uint64_t nc = c->n_context, n = c->n_reads - nc;      // context reads carry no qualities
	if (!c->finalized) return fail(c, CLB_ERR_STATE, "clb_qual_encode_original before the reads are complete (clb_count_finalize)");
	if (source > 2 || level < 1 || level > 3) return fail(c, CLB_ERR_BAD_ARG, "clb_qual_encode_original: source 0..2, level 1..3");
	if (level > 1 && !c->enc_done) return fail(c, CLB_ERR_STATE, "clb_qual_encode_original at level > 1 needs the tuples (clb_encode) for the match / anchor flags");
	if (c->qual_done) return fail(c, CLB_ERR_STATE, "the quality stream was already coded");
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
	std::vector<uint64_t> h_off; bool resident = false;
	{ const clb_status st = resolve_quals(c, quals, offsets, on_device, s, h_off, resident); if (st != CLB_OK) return st; }
	if (resident) { quals = c->dq.p; on_device = 1; }
	const uint64_t tot = h_off[n] - h_off[0];
	struct Tmp { std::vector<void*> v; cudaStream_t s; ~Tmp() { for (void* p : v) dev_free_async(p, s); } } tmp{{}, s};
	auto dalloc = [&](void** p, uint64_t bytes) { cudaError_t e = dev_malloc(p, bytes ? bytes : 1, s); if (e == cudaSuccess) tmp.v.push_back(*p); return e; };
	const uint8_t* d_q = nullptr; uint64_t* d_qoff = nullptr;
	CLB_CUDA(c, dalloc((void**)&d_qoff, sizeof(uint64_t) * (n + 1)));
	{
		std::vector<uint64_t> rel(n + 1);
		for (uint64_t i = 0; i <= n; ++i) rel[i] = h_off[i] - h_off[0];
		CLB_CUDA(c, cudaMemcpyAsync(d_qoff, rel.data(), sizeof(uint64_t) * (n + 1), cudaMemcpyHostToDevice, s));
		CLB_CUDA(c, cudaStreamSynchronize(s));
	}
	if (on_device) d_q = quals + h_off[0];
	else { uint8_t* b = nullptr; CLB_CUDA(c, dalloc((void**)&b, tot + 16)); CLB_CUDA(c, cudaMemcpyAsync(b, quals + h_off[0], tot, cudaMemcpyHostToDevice, s)); d_q = b; }
	uint8_t* d_flags = nullptr;
	if (level > 1) {
		CLB_CUDA(c, dalloc((void**)&d_flags, tot + 16));
		CLB_CUDA(c, cudaMemsetAsync(d_flags, 0, tot + 16, s));
		clb_status st = s3_qual_flags(c, d_qoff, (uint32_t)n, d_flags); if (st != CLB_OK) return st;
	}
	QoArgs a{};
	a.M = make_qorg_model(source, level);
	a.R = QoReads{c->pk.p, c->rd_start.p + nc, c->rd_len.p + nc, d_q, d_qoff, d_flags};
	a.n_reads = (uint32_t)n; a.n_packs = np;
	const uint64_t n_entries = a.M.base[1];
	uint32_t* d_hist = nullptr; uint32_t* d_pack_first = nullptr;
	CLB_CUDA(c, dalloc((void**)&d_hist, sizeof(uint32_t) * (n_entries + 1)));
	CLB_CUDA(c, dalloc((void**)&d_pack_first, sizeof(uint32_t) * (np + 1)));
	CLB_CUDA(c, cudaMemsetAsync(d_hist, 0, sizeof(uint32_t) * (n_entries + 1), s));
	CLB_CUDA(c, cudaMemcpyAsync(d_pack_first, pack_first.data(), sizeof(uint32_t) * (np + 1), cudaMemcpyHostToDevice, s));
	a.hist = d_hist; a.bad = d_hist + n_entries; a.pack_first = d_pack_first;
	if (n) { CLB_TIMED3(c, K_QUAL, (k_qo_count<<<(uint32_t)n, 256, 0, s>>>(a))); CLB_LAUNCH_CHECK(c, "k_qo_count"); }
	std::vector<uint32_t> hist(n_entries + 1);
	CLB_CUDA(c, cudaMemcpyAsync(hist.data(), d_hist, sizeof(uint32_t) * (n_entries + 1), cudaMemcpyDeviceToHost, s));
	CLB_CUDA(c, cudaStreamSynchronize(s));
	if (hist[n_entries]) return fail(c, CLB_ERR_BAD_ARG, "clb_qual_encode_original: a quality byte is outside phred+33 of 0..95");
	std::vector<uint32_t> tab(n_entries, 0);
	std::vector<uint8_t> hdr;
	hdr.insert(hdr.end(), {'Q', 'O', '0', '1'}); st_put(hdr, source); st_put(hdr, level); st_put(hdr, (uint64_t)n); st_put(hdr, np);
	st_build_tables(a.M, 1, hist, tab, hdr, QO_MIN_CTX);
	uint32_t* d_tab = nullptr;
	CLB_CUDA(c, dalloc((void**)&d_tab, sizeof(uint32_t) * n_entries));
	CLB_CUDA(c, cudaMemcpyAsync(d_tab, tab.data(), sizeof(uint32_t) * n_entries, cudaMemcpyHostToDevice, s));
	a.tab = d_tab;
	const uint32_t nl = np * QO_LANES;
	uint32_t* d_bytes = nullptr; uint64_t* d_dst = nullptr; uint64_t* d_phdr = nullptr;
	CLB_CUDA(c, dalloc((void**)&d_bytes, sizeof(uint32_t) * nl)); CLB_CUDA(c, dalloc((void**)&d_dst, sizeof(uint64_t) * nl)); CLB_CUDA(c, dalloc((void**)&d_phdr, sizeof(uint64_t) * np));
	QoEnc e{d_bytes, d_dst, d_phdr, nullptr};
	if (nl) { CLB_TIMED3(c, K_QUAL, (k_qo_encode<false><<<(nl + 63) / 64, 64, 0, s>>>(a, e))); CLB_LAUNCH_CHECK(c, "k_qo_encode<size>"); }
	std::vector<uint32_t> lane_bytes(nl);
	CLB_CUDA(c, cudaMemcpyAsync(lane_bytes.data(), d_bytes, sizeof(uint32_t) * nl, cudaMemcpyDeviceToHost, s));
	CLB_CUDA(c, cudaStreamSynchronize(s));
	uint64_t out_at = hdr.size();
	std::vector<uint64_t> dst(nl), phdr(np);
	for (uint32_t p = 0; p < np; ++p) {
		phdr[p] = out_at; out_at += 4 + 4 * QO_LANES;
		for (uint32_t l = 0; l < QO_LANES; ++l) { dst[(size_t)p * QO_LANES + l] = out_at; out_at += lane_bytes[(size_t)p * QO_LANES + l]; }
	}
	CLB_CUDA(c, c->qs.reserve(out_at + 16, s, false));
	CLB_CUDA(c, cudaMemcpyAsync(c->qs.p, hdr.data(), hdr.size(), cudaMemcpyHostToDevice, s));
	CLB_CUDA(c, cudaMemcpyAsync(d_dst, dst.data(), sizeof(uint64_t) * nl, cudaMemcpyHostToDevice, s));
	CLB_CUDA(c, cudaMemcpyAsync(d_phdr, phdr.data(), sizeof(uint64_t) * np, cudaMemcpyHostToDevice, s));
	e.out = c->qs.p;
	if (nl) { CLB_TIMED3(c, K_QUAL, (k_qo_encode<true><<<(nl + 63) / 64, 64, 0, s>>>(a, e))); CLB_LAUNCH_CHECK(c, "k_qo_encode<write>"); }
	CLB_CUDA(c, cudaStreamSynchronize(s));
	c->qs_total = out_at;
	c->qual_done = true;
	return CLB_OK;
}

} // namespace clb
