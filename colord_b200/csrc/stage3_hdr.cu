// stage3_hdr.cu — header (read id) stream on device.
//
// Kept from the reference: the event model of CIDCoder::compress_lossless (hdr_model.h: tokens, the same-shape flag, per-token
// same / same-length / differing characters, plain fallback) and the arithmetic of its range coder (sub_rc.h:83-201).
// Replaced: the adaptive models + single coder chain of CEntrComprHeaders::Compress (entr_header.cpp:23-46).  Every event of a
// header depends only on the header itself and on its predecessor — both are input — so all headers are walked in parallel:
// pass 0 computes the same-shape flag of every header, pass 1 counts (family, context, symbol) triples, the host turns the
// counts into static 12-bit tables (metadata-sized), pass 2 codes every pack with 64 independent range-coder lanes (lane l takes
// headers l, l+64, ... of its pack), first sizing and then writing the lane streams at their final place.
// Native container "HB01"; CPU twin + decoder: oracle/stage3_hdr.c.
#include "ctx.h"
#include "hdr_model.h"
#include "static_tables.h"
#include "range_sink.cuh"
#include <vector>

namespace clb {

constexpr uint32_t HB_LANES = 64, HB_MIN_CTX = 64, HB_PACK = 4096;

struct HArgs {
	HdrInput H; HdrModel M;
	const uint64_t* pack_first; uint32_t n_packs; uint64_t n;
	uint8_t* flags; uint32_t* hist; const uint32_t* tab; uint32_t* bad;
};

CLB_D uint64_t pack_of(const HArgs& a, uint64_t r)
{
	uint32_t lo = 0, hi = a.n_packs;              // last pack_first <= r
	while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (a.pack_first[mid] <= r) lo = mid; else hi = mid; }
	return a.pack_first[lo];
}
CLB_D uint32_t flag_ctx_of(const uint8_t* flags, uint64_t r, uint64_t pack_start)
{
	uint32_t c = 0;
	for (uint64_t k = r - pack_start < 8 ? pack_start : r - 8; k < r; ++k) c = (c << 1) + flags[k];
	return c & 0xff;
}

// pass 0: one thread per header: the same-shape flag; a NUL byte inside a header is refused (0 is the coder's terminator)
__global__ void __launch_bounds__(128) k_h_flags(HArgs a)
{
	const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= a.n) return;
	const uint8_t* cur = a.H.bytes + a.H.off[r]; const uint32_t nc = (uint32_t)(a.H.off[r + 1] - a.H.off[r]);
	bool nul = false;
	for (uint32_t i = 0; i < nc; ++i) nul |= cur[i] == 0;
	if (nul) atomicExch(a.bad, 1u);
	const bool has_prev = r > pack_of(a, r);
	a.flags[r] = has_prev && hdr_same_shape(cur, nc, a.H.bytes + a.H.off[r - 1], (uint32_t)(a.H.off[r] - a.H.off[r - 1]));
}
// pass 1: one thread per header
__global__ void __launch_bounds__(128) k_h_count(HArgs a)
{
	const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= a.n) return;
	const uint64_t p0 = pack_of(a, r);
	HistSinkT<HdrModel> s{a.hist, &a.M};
	hdr_walk(a.H, r, r > p0, flag_ctx_of(a.flags, r, p0), s);
}

struct HEnc { uint32_t* lane_bytes; const uint64_t* dst_off; const uint64_t* pack_hdr_off; uint8_t* out; };

// pass 2: one thread per (pack, lane); WRITE = false sizes the lane streams, WRITE = true writes them and the pack headers
template <bool WRITE>
__global__ void __launch_bounds__(64) k_h_encode(HArgs a, HEnc e)
{
	const uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
	if (li >= a.n_packs * HB_LANES) return;
	const uint32_t p = li / HB_LANES, l = li % HB_LANES;
	const uint64_t r0 = a.pack_first[p], r1 = a.pack_first[p + 1];
	RangeSinkT<HdrModel> s{a.tab, &a.M, WRITE ? e.out + e.dst_off[li] : nullptr, 0, 0, 0};
	s.start();
	for (uint64_t r = r0 + l; r < r1; r += HB_LANES) hdr_walk(a.H, r, r > r0, flag_ctx_of(a.flags, r, r0), s);
	s.end();
	if (!WRITE) { e.lane_bytes[li] = (uint32_t)s.n; return; }
	uint8_t* h = e.out + e.pack_hdr_off[p];
	const uint32_t nb = (uint32_t)s.n;
	h[4 + 4 * l] = (uint8_t)nb; h[5 + 4 * l] = (uint8_t)(nb >> 8); h[6 + 4 * l] = (uint8_t)(nb >> 16); h[7 + 4 * l] = (uint8_t)(nb >> 24);
	if (l == 0) { const uint32_t np = (uint32_t)(r1 - r0); h[0] = (uint8_t)np; h[1] = (uint8_t)(np >> 8); h[2] = (uint8_t)(np >> 16); h[3] = (uint8_t)(np >> 24); }
}

clb_status s3_hdr_encode(clb_ctx* c, const uint8_t* bytes, const uint64_t* offsets, const uint8_t* plus_id, uint64_t n, int on_device,
	const uint32_t* pack_sizes, uint32_t n_packs)
{
	cudaStream_t s = c->stream3;
	if (c->hdr_done) return fail(c, CLB_ERR_STATE, "clb_hdr_encode called twice");
	std::vector<uint64_t> pack_first{0};
	if (pack_sizes) {
		uint64_t at = 0;
		for (uint32_t i = 0; i < n_packs; ++i) { at += pack_sizes[i]; if (pack_sizes[i]) pack_first.push_back(at); }
		if (at != n) return fail(c, CLB_ERR_BAD_ARG, "pack_sizes do not sum to the number of headers");
	} else {
		for (uint64_t at = HB_PACK; at < n; at += HB_PACK) pack_first.push_back(at);
		if (n) pack_first.push_back(n);
	}
	const uint32_t np = (uint32_t)pack_first.size() - 1;
	struct Tmp { std::vector<void*> v; cudaStream_t s; ~Tmp() { for (void* p : v) dev_free_async(p, s); } } tmp{{}, s};
	auto dalloc = [&](void** p, uint64_t nbytes) { cudaError_t e = dev_malloc(p, nbytes ? nbytes : 1, s); if (e == cudaSuccess) tmp.v.push_back(*p); return e; };
	HArgs a{};
	a.M = make_hdr_model(); a.n = n; a.n_packs = np;
	if (on_device) a.H = HdrInput{bytes, offsets, plus_id};
	else {
		const uint64_t total = n ? offsets[n] : 0;
		if (n && offsets[0] != 0) return fail(c, CLB_ERR_BAD_ARG, "clb_hdr_encode: offsets[0] must be 0");
		uint8_t* d_b = nullptr; uint64_t* d_o = nullptr; uint8_t* d_p = nullptr;
		CLB_CUDA(c, dalloc((void**)&d_b, total)); CLB_CUDA(c, dalloc((void**)&d_o, sizeof(uint64_t) * (n + 1)));
		CLB_CUDA(c, cudaMemcpyAsync(d_b, bytes, total, cudaMemcpyHostToDevice, s));
		CLB_CUDA(c, cudaMemcpyAsync(d_o, offsets, sizeof(uint64_t) * (n + 1), cudaMemcpyHostToDevice, s));
		if (plus_id) { CLB_CUDA(c, dalloc((void**)&d_p, n)); CLB_CUDA(c, cudaMemcpyAsync(d_p, plus_id, n, cudaMemcpyHostToDevice, s)); }
		a.H = HdrInput{d_b, d_o, d_p};
	}
	const uint64_t n_entries = a.M.base[H_COUNT];
	uint64_t* d_pack_first = nullptr; uint32_t* d_hist = nullptr; uint32_t* d_bad = nullptr;
	CLB_CUDA(c, dalloc((void**)&d_pack_first, sizeof(uint64_t) * (np + 1)));
	CLB_CUDA(c, dalloc((void**)&a.flags, n));
	CLB_CUDA(c, dalloc((void**)&d_hist, sizeof(uint32_t) * (n_entries + 1)));
	d_bad = d_hist + n_entries;
	CLB_CUDA(c, cudaMemsetAsync(d_hist, 0, sizeof(uint32_t) * (n_entries + 1), s));
	CLB_CUDA(c, cudaMemcpyAsync(d_pack_first, pack_first.data(), sizeof(uint64_t) * (np + 1), cudaMemcpyHostToDevice, s));
	a.pack_first = d_pack_first; a.hist = d_hist; a.bad = d_bad;
	const uint32_t nblk = (uint32_t)((n + 127) / 128);
	if (n) {
		CLB_TIMED3(c, K_HDR, (k_h_flags<<<nblk, 128, 0, s>>>(a))); CLB_LAUNCH_CHECK(c, "k_h_flags");
		CLB_TIMED3(c, K_HDR, (k_h_count<<<nblk, 128, 0, s>>>(a))); CLB_LAUNCH_CHECK(c, "k_h_count");
	}
	// ---- counts -> static tables + container header (host, metadata-sized) ----
	std::vector<uint32_t> hist(n_entries + 1);
	CLB_CUDA(c, cudaMemcpyAsync(hist.data(), d_hist, sizeof(uint32_t) * (n_entries + 1), cudaMemcpyDeviceToHost, s));
	CLB_CUDA(c, cudaStreamSynchronize(s));
	if (hist[n_entries]) return fail(c, CLB_ERR_BAD_SYMBOL, "clb_hdr_encode: a header holds a NUL byte");
	std::vector<uint32_t> tab(n_entries, 0);
	std::vector<uint8_t> hdr;
	hdr.insert(hdr.end(), {'H', 'B', '0', '1'}); st_put(hdr, (uint64_t)n); st_put(hdr, np);
	st_build_tables(a.M, H_COUNT, hist, tab, hdr, HB_MIN_CTX);
	uint32_t* d_tab = nullptr;
	CLB_CUDA(c, dalloc((void**)&d_tab, sizeof(uint32_t) * n_entries));
	CLB_CUDA(c, cudaMemcpyAsync(d_tab, tab.data(), sizeof(uint32_t) * n_entries, cudaMemcpyHostToDevice, s));
	a.tab = d_tab;
	// ---- pass 2: size every lane stream, lay the container out, write ----
	const uint32_t nl = np * HB_LANES;
	uint32_t* d_bytes = nullptr; uint64_t* d_dst = nullptr; uint64_t* d_phdr = nullptr;
	CLB_CUDA(c, dalloc((void**)&d_bytes, sizeof(uint32_t) * nl)); CLB_CUDA(c, dalloc((void**)&d_dst, sizeof(uint64_t) * nl)); CLB_CUDA(c, dalloc((void**)&d_phdr, sizeof(uint64_t) * np));
	HEnc e{d_bytes, d_dst, d_phdr, nullptr};
	if (nl) { CLB_TIMED3(c, K_HDR, (k_h_encode<false><<<(nl + 63) / 64, 64, 0, s>>>(a, e))); CLB_LAUNCH_CHECK(c, "k_h_encode<size>"); }
	std::vector<uint32_t> lane_bytes(nl);
	CLB_CUDA(c, cudaMemcpyAsync(lane_bytes.data(), d_bytes, sizeof(uint32_t) * nl, cudaMemcpyDeviceToHost, s));
	CLB_CUDA(c, cudaStreamSynchronize(s));
	uint64_t out_at = hdr.size();
	std::vector<uint64_t> dst(nl), phdr(np);
	for (uint32_t p = 0; p < np; ++p) {
		phdr[p] = out_at; out_at += 4 + 4 * HB_LANES;
		for (uint32_t l = 0; l < HB_LANES; ++l) { dst[(size_t)p * HB_LANES + l] = out_at; out_at += lane_bytes[(size_t)p * HB_LANES + l]; }
	}
	CLB_CUDA(c, c->hs.reserve(out_at + 16, s, false));
	CLB_CUDA(c, cudaMemcpyAsync(c->hs.p, hdr.data(), hdr.size(), cudaMemcpyHostToDevice, s));
	CLB_CUDA(c, cudaMemcpyAsync(d_dst, dst.data(), sizeof(uint64_t) * nl, cudaMemcpyHostToDevice, s));
	CLB_CUDA(c, cudaMemcpyAsync(d_phdr, phdr.data(), sizeof(uint64_t) * np, cudaMemcpyHostToDevice, s));
	e.out = c->hs.p;
	if (nl) { CLB_TIMED3(c, K_HDR, (k_h_encode<true><<<(nl + 63) / 64, 64, 0, s>>>(a, e))); CLB_LAUNCH_CHECK(c, "k_h_encode<write>"); }
	CLB_CUDA(c, cudaStreamSynchronize(s));
	c->hs_total = out_at;
	c->hs_header = hdr.size();
	c->hdr_done = true;
	return CLB_OK;
}

} // namespace clb
