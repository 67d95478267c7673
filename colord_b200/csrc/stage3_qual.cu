// stage3_qual.cu — quality stream on device, the "*-avg" modes.
//
// What is kept from the reference (src/colord): the lossy transform and the context model —
//   quality_coder.cpp:250-270      phred -> bin by the forward thresholds
//   quality_coder_impl.cpp:191-249 per-read per-bin means coded as (uint32)(mean * 256) (:821-835), then one bin symbol per
//                                  base under [previous 3 (6) symbols] + [bases i-2 .. i+1] + [match / anchor flags, level > 1]
//   quality_coder_impl.cpp:25-76   the flags come from the read's tuples
// so the decoder's error-diffusion reconstruction (:559-601) prints exactly what the reference prints (test/*.quan).
// What is replaced: the reference codes the symbols with ONE adaptive range coder whose models persist over the whole file
// (entr_qual.h:100-126) — a serial chain.  Here the models are static: pass 1 counts (context, symbol) pairs of all reads with
// atomics, the host turns the 2^17..2^19-entry count table into 12-bit frequency tables (contexts seen fewer than 32 times
// share a fallback table; metadata-sized work), pass 2 codes every read pack as 4 streams of 32-way INTERLEAVED rANS (31-bit
// states, 16-bit renormalisation): one warp per stream, symbol k of a read belongs to state k mod 32 = lane k mod 32, so a warp
// codes 32 symbols per step in lockstep; the lanes that renormalise find their place in the word stream with one ballot + popc
// (container "QB02").  Container layout and its CPU twin + decoder: oracle/stage3_qual.c (the bytes must be identical).
// Round 1's "QB01" (64 single-state lanes per pack, the one state carried by all 32 threads of a warp: ~19 warp instructions per
// symbol, 0.9 s of k_q_encode per 25 Gbases) is still read by the decoders.
#include "ctx.h"
#include <algorithm>
#include <cstring>
#include <vector>
#include <cstdio>
#include <cstdlib>
#include <ctime>

namespace clb {

constexpr uint32_t QB_LANES = 4 /* streams per pack */, QB_STATES = 32, QB_PROB_BITS = 12, QB_M = 1u << QB_PROB_BITS, QB_L = 1u << 15, QB_MIN_CTX = 32;

struct QP { uint32_t nb, level, bps, cb, cbits; uint32_t thr[4]; };

struct QArgs {
	const uint64_t* pk; const uint64_t* rd_start; const uint32_t* rd_len;
	const uint8_t* quals; const uint64_t* qoff;         // quality bytes of read r: quals[qoff[r] .. + rd_len[r])
	const uint8_t* flags;                               // level > 1: per base 0 / 1 (match) / 2 (anchor), same layout as quals
	uint32_t n_reads; QP P;
	uint32_t* avg16;                                    // n_reads x 5
	uint32_t* hist; uint32_t* mhist;                    // 2^cbits x nb, nb x 128
	const uint4* tab;                                   // per (context, symbol): reciprocal, bias, complement | shift << 16, frequency
	const uint4* mtab;                                  // same for the means' high byte
};

CLB_D uint32_t q_bin(const QP& P, uint32_t phred) { uint32_t b = 0; while (b + 1 < P.nb && phred >= P.thr[b]) ++b; return b; }

// context of position i of a read (quality_coder_impl.cpp:222-240 with :528-537 unrolled)
CLB_D uint32_t q_context(const QArgs& a, uint64_t rs, uint32_t n, const uint8_t* q, const uint8_t* fl, uint32_t i)
{
	const QP& P = a.P;
	uint32_t c = 0;
	const uint32_t n_prev = P.cb / P.bps;
	for (uint32_t k = n_prev; k >= 1; --k) c = (c << P.bps) | (i >= k ? q_bin(P, q[i - k] - 33u) : ((1u << P.bps) - 1));
	uint32_t dna = 0;
	for (int d = -2; d <= 1; ++d) { const long long j = (long long)i + d; dna = (dna << 2) | ((j >= 0 && j < (long long)n) ? base_at(a.pk, rs + (uint64_t)j) : 0u); }
	c |= dna << P.cb;
	if (P.level > 1) { c |= (uint32_t)(fl[i] == 1) << (P.cb + 8); c |= (uint32_t)(fl[i] == 2) << (P.cb + 9); }
	return c;
}

clb_status resolve_quals(clb_ctx* c, const uint8_t* quals, const uint64_t* offsets, int on_device, cudaStream_t s, std::vector<uint64_t>& h_off, bool& resident)
{
	const uint64_t nc = c->n_context, n = c->n_reads - nc;
	h_off.assign(n + 1, 0);
	resident = quals == nullptr;
	if (resident) {
		for (uint64_t i = 0; i < n; ++i) h_off[i + 1] = h_off[i] + c->h_rd_len[nc + i];
		if (h_off[n] != c->dq_n) return fail(c, CLB_ERR_STATE, "no qualities given and the resident ones (clb_append_quals) do not cover the reads");
		return CLB_OK;
	}
	if (!offsets) return fail(c, CLB_ERR_BAD_ARG, "qualities without offsets");
	if (on_device) { CLB_CUDA(c, cudaMemcpyAsync(h_off.data(), offsets, sizeof(uint64_t) * (n + 1), cudaMemcpyDeviceToHost, s)); CLB_CUDA(c, cudaStreamSynchronize(s)); }
	else std::memcpy(h_off.data(), offsets, sizeof(uint64_t) * (n + 1));
	for (uint64_t i = 0; i < n; ++i) if (h_off[i + 1] - h_off[i] != c->h_rd_len[nc + i]) return fail(c, CLB_ERR_BAD_ARG, "quality lengths differ from the read lengths");
	return CLB_OK;
}

// per-base flags from the tuples (quality_coder_impl.cpp:25-76); one thread per read
__global__ void __launch_bounds__(128) k_q_flags(const uint8_t* __restrict__ es, const uint64_t* __restrict__ es_off, const uint64_t* __restrict__ qoff,
	const uint32_t* __restrict__ rd_len, uint32_t n_reads, uint8_t* __restrict__ flags)
{
	const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= n_reads) return;
	const uint8_t* t = es + es_off[r]; const uint64_t tn = es_off[r + 1] - es_off[r];
	uint8_t* fl = flags + qoff[r]; const uint32_t n = rd_len[r];
	if (!tn) return;
	const uint32_t t0 = t[0] >> 4;
	if (t0 == 9 || t0 == 11) return;                    // plain reads: no flags (the buffer is zeroed)
	uint64_t p = 5; uint32_t at = 0;
	while (p < tn) {
		const uint32_t ty = t[p] >> 4;
		if (ty == 4) { const uint32_t len = ((uint32_t)(t[p] & 15) << 24) | ((uint32_t)t[p + 1] << 16) | ((uint32_t)t[p + 2] << 8) | t[p + 3]; for (uint32_t k = 0; k < len && at < n; ++k) fl[at++] = 2; p += 4; }
		else if (ty == 5) p += 4;
		else if (ty == 6) p += 5;
		else { if (ty == 2) { if (at < n) fl[at] = 1; ++at; } else if (ty == 0 || ty == 3) ++at; p += 1; }
	}
}

// the flags of the reads after the context reads (shared with the lossless mode, stage3_qorg.cu); flags must be zeroed
clb_status s3_qual_flags(clb_ctx* c, const uint64_t* d_qoff, uint32_t n, uint8_t* d_flags)
{
	const uint64_t nc = c->n_context;
	if (n) { CLB_TIMED3(c, K_QUAL, (k_q_flags<<<(n + 127) / 128, 128, 0, c->stream3>>>(c->es.p, c->es_off + nc, d_qoff, c->rd_len.p + nc, n, d_flags))); CLB_LAUNCH_CHECK(c, "k_q_flags"); }
	return CLB_OK;
}

// pass 1: per-read means + (context, symbol) counts; one CTA per read
__global__ void __launch_bounds__(128) k_q_count(QArgs a)
{
	__shared__ unsigned long long s_sum[5]; __shared__ uint32_t s_cnt[5];
	const uint32_t r = blockIdx.x;
	const uint32_t n = a.rd_len[r]; const uint64_t rs = a.rd_start[r];
	const uint8_t* q = a.quals + a.qoff[r];
	const uint8_t* fl = a.flags ? a.flags + a.qoff[r] : nullptr;
	if (threadIdx.x < 5) { s_sum[threadIdx.x] = 0; s_cnt[threadIdx.x] = 0; }
	__syncthreads();
	unsigned long long sum[5] = {0, 0, 0, 0, 0}; uint32_t cnt[5] = {0, 0, 0, 0, 0};
	for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
		const uint32_t ph = q[i] - 33u, b = q_bin(a.P, ph);
#pragma unroll
		for (int k = 0; k < 5; ++k) if (b == (uint32_t)k) { sum[k] += ph; ++cnt[k]; }
		atomicAdd(&a.hist[(size_t)q_context(a, rs, n, q, fl, i) * a.P.nb + b], 1u);
	}
#pragma unroll
	for (int k = 0; k < 5; ++k) {
		for (int d = 16; d; d >>= 1) { sum[k] += __shfl_xor_sync(0xffffffffu, sum[k], d); cnt[k] += __shfl_xor_sync(0xffffffffu, cnt[k], d); }
		if ((threadIdx.x & 31) == 0 && cnt[k]) { atomicAdd(&s_sum[k], sum[k]); atomicAdd(&s_cnt[k], cnt[k]); }
	}
	__syncthreads();
	if (threadIdx.x < a.P.nb) {
		// quality_coder_impl.cpp:205-217: mean in double, then (uint32)(mean * 256); sums of integers are exact in either order
		const double avg = s_cnt[threadIdx.x] ? __ddiv_rn((double)s_sum[threadIdx.x], (double)s_cnt[threadIdx.x]) : 0.0;
		const uint32_t v = (uint32_t)__dmul_rn(avg, 256.0);
		a.avg16[(size_t)r * 5 + threadIdx.x] = v;
		atomicAdd(&a.mhist[threadIdx.x * 128 + ((v >> 8) & 127)], 1u);
	}
}

// Encoder symbol with the division replaced by a multiplication with a 32-bit reciprocal (exact for states below 2^31;
// the construction is the one of the public-domain ryg_rans RansEncSymbolInit): x' = x + bias + (x / f) * (M - f).
CLB_HD uint4 rans_symbol(uint32_t start, uint32_t freq)
{
	uint4 s;
	if (freq < 2) { s.x = ~0u; s.y = start + QB_M - 1; s.z = (QB_M - freq) | (0u << 16); }
	else {
		uint32_t shift = 0; while (freq > (1u << shift)) ++shift;
		s.x = (uint32_t)(((1ull << (shift + 31)) + freq - 1) / freq);
		s.y = start; s.z = (QB_M - freq) | ((shift - 1) << 16);
	}
	s.w = freq;
	return s;
}
struct QEnc {
	const uint32_t* pack_first; uint32_t pack_lo, n_packs;     // packs of this chunk
	const uint64_t* lane_off;                                   // first temp word of every lane of the chunk
	uint16_t* tmp; uint32_t* lane_words; uint32_t* lane_state;
};

// pass 2: one WARP per (pack, stream); stream s of a pack takes its reads s, s + 4, ...  Reads last to first, symbols last to first
// (rANS decodes in the opposite order).  The symbols of a read are its 2 x n_bins mean bytes, then one bin symbol per base; symbol k
// belongs to state k mod 32.  A step codes the 32 symbols k = 32 g + lane: every lane fetches its own table entry (quality bytes and
// packed bases are read side by side), renormalises (the 16-bit words of a step go out in ascending lane order in DECODING order,
// i.e. descending here: position = number of renormalising lanes above mine, one ballot + popc) and takes its rANS step with a
// reciprocal multiply.
__global__ void __launch_bounds__(128) k_q_encode(QArgs a, QEnc e)
{
	const uint32_t li = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, t = threadIdx.x & 31;
	if (li >= e.n_packs * QB_LANES) return;
	const unsigned FULL = 0xffffffffu;
	const uint32_t p = e.pack_lo + li / QB_LANES, l = li % QB_LANES;
	const uint32_t r0 = e.pack_first[p], r1 = e.pack_first[p + 1];
	uint16_t* w = e.tmp + e.lane_off[li]; uint32_t nw = 0;
	uint32_t x = QB_L;
	const QP& P = a.P;
	const uint32_t above = ~((2u << t) - 1u);                // lanes above mine (t = 31: none)
	if (r0 + l < r1) {
		const uint32_t last = r0 + l + ((r1 - 1 - (r0 + l)) / QB_LANES) * QB_LANES;
		for (long long r = last; r >= (long long)(r0 + l); r -= QB_LANES) {
			const uint32_t n = a.rd_len[r]; const uint64_t rs = a.rd_start[r];
			const uint8_t* q = a.quals + a.qoff[r];
			const uint8_t* fl = a.flags ? a.flags + a.qoff[r] : nullptr;
			const uint32_t ns = 2 * P.nb, T = ns + n;
			// the table entry of a symbol depends on the input alone: the entries of the NEXT step are fetched before the coder's
			// arithmetic of this one, so that their way from the L2 lies beside it instead of in front of it
			auto fetch = [&](long long g, bool& valid) -> uint4 {
				const uint32_t k = (uint32_t)(g << 5) + t;
				valid = k < T;
				uint4 fc = make_uint4(0, 0, 0, 1);
				if (valid) {
					if (k < ns) {                                    // bin k / 2: high byte of mean * 256 under the bin's table, then the low byte (uniform)
						const uint32_t b = k >> 1, v = a.avg16[(size_t)r * 5 + b], a1 = (v >> 8) & 127, a2 = v & 0xff;
						fc = (k & 1) ? rans_symbol(a2 * (QB_M >> 8), QB_M >> 8) : a.mtab[b * 128 + a1];
					} else {
						const uint32_t j = k - ns;
						fc = a.tab[(size_t)q_context(a, rs, n, q, fl, j) * P.nb + q_bin(P, q[j] - 33u)];
					}
				}
				return fc;
			};
			bool valid, valid_next = false;
			uint4 fc = fetch((long long)((T - 1) >> 5), valid), fc_next = fc;
			for (long long g = (long long)((T - 1) >> 5); g >= 0; --g) {
				if (g > 0) fc_next = fetch(g - 1, valid_next);
				const bool emit = valid && x >= (fc.w << 19);       // f <= 2^12: no overflow
				const uint32_t m = __ballot_sync(FULL, emit);
				if (emit) { w[nw + __popc(m & above)] = (uint16_t)x; x >>= 16; }
				nw += __popc(m);
				if (valid) { const uint32_t qq = __umulhi(x, fc.x) >> (fc.z >> 16); x = x + fc.y + qq * (fc.z & 0xffffu); }
				fc = fc_next; valid = valid_next;
			}
		}
	}
	if (t == 0) e.lane_words[li] = nw;
	e.lane_state[(size_t)li * QB_STATES + t] = x;
}

// streams into the final container: the 32 states, then the words in decoding order (last written first); one warp per stream
__global__ void __launch_bounds__(128) k_q_gather(QEnc e, const uint64_t* __restrict__ dst_off, const uint64_t* __restrict__ pack_hdr_off, uint8_t* __restrict__ out)
{
	const uint32_t li = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (li >= e.n_packs * QB_LANES) return;
	const uint32_t nw = e.lane_words[li];
	uint8_t* d = out + dst_off[li];
	const uint16_t* w = e.tmp + e.lane_off[li];
	{
		const uint32_t x = e.lane_state[(size_t)li * QB_STATES + lane];
		uint8_t* ds = d + 4 * lane;
		ds[0] = (uint8_t)x; ds[1] = (uint8_t)(x >> 8); ds[2] = (uint8_t)(x >> 16); ds[3] = (uint8_t)(x >> 24);
	}
	if (lane == 0) {      // size field in the pack header, and the pack's read count
		const uint32_t p = li / QB_LANES, l = li % QB_LANES;
		uint8_t* h = out + pack_hdr_off[p];
		const uint32_t bytes = 4 * QB_STATES + 2 * nw;
		h[4 + 4 * l] = (uint8_t)bytes; h[5 + 4 * l] = (uint8_t)(bytes >> 8); h[6 + 4 * l] = (uint8_t)(bytes >> 16); h[7 + 4 * l] = (uint8_t)(bytes >> 24);
		if (l == 0) { const uint32_t np = e.pack_first[e.pack_lo + p + 1] - e.pack_first[e.pack_lo + p]; h[0] = (uint8_t)np; h[1] = (uint8_t)(np >> 8); h[2] = (uint8_t)(np >> 16); h[3] = (uint8_t)(np >> 24); }
	}
	for (uint32_t k = lane; k < nw; k += 32) { const uint16_t v = w[nw - 1 - k]; d[4 * QB_STATES + 2 * k] = (uint8_t)v; d[4 * QB_STATES + 1 + 2 * k] = (uint8_t)(v >> 8); }
}

// ------------------------------------------------------------------------------------------------ host side
struct QTrace {            // CLB_S2_TRACE=1: wall time of every phase (synchronising; debugging aid)
	bool on; cudaStream_t s; double t0;
	static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
	explicit QTrace(cudaStream_t st) : on(std::getenv("CLB_S2_TRACE") != nullptr), s(st), t0(now()) {}
	void mark(const char* w) { if (!on) return; cudaStreamSynchronize(s); const double t = now(); fprintf(stderr, "[s3q] %-28s %9.3f ms\n", w, t - t0); t0 = t; }
};
static void normalise(const uint32_t* cnt, uint32_t n, uint16_t* f)
{
	uint64_t tot = 0; uint32_t best = 0;
	for (uint32_t i = 0; i < n; ++i) { tot += cnt[i]; if (cnt[i] > cnt[best]) best = i; }
	if (!tot) { for (uint32_t i = 0; i < n; ++i) f[i] = 0; return; }
	uint32_t sum = 0;
	for (uint32_t i = 0; i < n; ++i) { uint32_t v = (uint32_t)(((uint64_t)cnt[i] << QB_PROB_BITS) / tot); if (cnt[i] && !v) v = 1; f[i] = (uint16_t)v; sum += v; }
	f[best] = (uint16_t)(f[best] + QB_M - sum);
}
template <typename T> static void put(std::vector<uint8_t>& o, const T& v) { const uint8_t* p = reinterpret_cast<const uint8_t*>(&v); o.insert(o.end(), p, p + sizeof(T)); }

clb_status s3_qual_encode(clb_ctx* c, const clb_qual_params* prm, const uint8_t* quals, const uint64_t* offsets, int on_device,
	const uint32_t* pack_sizes, uint32_t n_packs)
{
	cudaStream_t s = c->stream3;
	QTrace tr(s);
	const uint64_t nc = c->n_context, n = c->n_reads - nc;      // context reads carry no qualities: the stream covers the reads after them
	if (!c->finalized) return fail(c, CLB_ERR_STATE, "clb_qual_encode before the reads are complete (clb_count_finalize)");
	if (prm->n_bins != 2 && prm->n_bins != 4 && prm->n_bins != 5) return fail(c, CLB_ERR_BAD_ARG, "clb_qual_encode: the *-avg modes have 2, 4 or 5 bins");
	if (prm->level > 1 && !c->enc_done) return fail(c, CLB_ERR_STATE, "clb_qual_encode at level > 1 needs the tuples (clb_encode) for the match / anchor flags");
	if (c->qual_done) return fail(c, CLB_ERR_STATE, "clb_qual_encode called twice");
	QP P{}; P.nb = prm->n_bins; P.level = prm->level; P.bps = P.nb == 2 ? 2 : 3; P.cb = P.bps * (P.nb == 2 ? 6 : 3); P.cbits = P.cb + 8 + (P.level > 1 ? 2 : 0);
	for (int i = 0; i < 4; ++i) P.thr[i] = prm->thresholds[i];
	const uint64_t n_ctx = 1ull << P.cbits;
	// packs (same rule as stage 2)
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
	// qualities on the device
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
	if (P.level > 1) {
		CLB_CUDA(c, dalloc((void**)&d_flags, tot + 16));
		CLB_CUDA(c, cudaMemsetAsync(d_flags, 0, tot + 16, s));
		if (n) { CLB_TIMED3(c, K_QUAL, (k_q_flags<<<(uint32_t)((n + 127) / 128), 128, 0, s>>>(c->es.p, c->es_off + nc, d_qoff, c->rd_len.p + nc, (uint32_t)n, d_flags))); CLB_LAUNCH_CHECK(c, "k_q_flags"); }
	}
	QArgs a{};
	a.pk = c->pk.p; a.rd_start = c->rd_start.p + nc; a.rd_len = c->rd_len.p + nc; a.quals = d_q; a.qoff = d_qoff; a.flags = d_flags; a.n_reads = (uint32_t)n; a.P = P;
	uint32_t* d_avg = nullptr; uint32_t* d_hist = nullptr; uint32_t* d_mhist = nullptr; uint4* d_tab = nullptr; uint4* d_mtab = nullptr;
	CLB_CUDA(c, dalloc((void**)&d_avg, sizeof(uint32_t) * 5 * (n + 1)));
	CLB_CUDA(c, dalloc((void**)&d_hist, sizeof(uint32_t) * n_ctx * P.nb)); CLB_CUDA(c, dalloc((void**)&d_mhist, sizeof(uint32_t) * 5 * 128));
	CLB_CUDA(c, dalloc((void**)&d_tab, sizeof(uint4) * n_ctx * P.nb)); CLB_CUDA(c, dalloc((void**)&d_mtab, sizeof(uint4) * 5 * 128));
	CLB_CUDA(c, cudaMemsetAsync(d_hist, 0, sizeof(uint32_t) * n_ctx * P.nb, s)); CLB_CUDA(c, cudaMemsetAsync(d_mhist, 0, sizeof(uint32_t) * 5 * 128, s));
	a.avg16 = d_avg; a.hist = d_hist; a.mhist = d_mhist;
	tr.mark("setup");
	if (n) { CLB_TIMED3(c, K_QUAL, (k_q_count<<<(uint32_t)n, 128, 0, s>>>(a))); CLB_LAUNCH_CHECK(c, "k_q_count"); }
	tr.mark("k_q_count");
	// ---- count table -> frequency tables + the container's header (metadata-sized, on the host) ----
	std::vector<uint32_t> hist(n_ctx * P.nb), mh(5 * 128);
	CLB_CUDA(c, cudaMemcpyAsync(hist.data(), d_hist, sizeof(uint32_t) * hist.size(), cudaMemcpyDeviceToHost, s));
	CLB_CUDA(c, cudaMemcpyAsync(mh.data(), d_mhist, sizeof(uint32_t) * mh.size(), cudaMemcpyDeviceToHost, s));
	CLB_CUDA(c, cudaStreamSynchronize(s));
	const uint32_t nb = P.nb, n_fb = 1u << P.cb;
	std::vector<uint32_t> fbh((size_t)n_fb * nb, 0); std::vector<uint8_t> dense(n_ctx, 0);
	for (uint64_t x = 0; x < n_ctx; ++x) {
		uint64_t t = 0; for (uint32_t k = 0; k < nb; ++k) t += hist[x * nb + k];
		if (t >= QB_MIN_CTX) dense[x] = 1; else for (uint32_t k = 0; k < nb; ++k) fbh[(x & (n_fb - 1)) * nb + k] += hist[x * nb + k];
	}
	std::vector<uint16_t> freq(n_ctx * nb), fbf((size_t)n_fb * nb), mf(5 * 128, 0);
	for (uint32_t x = 0; x < n_fb; ++x) normalise(&fbh[(size_t)x * nb], nb, &fbf[(size_t)x * nb]);
	std::vector<uint4> tab(n_ctx * nb), mtab(5 * 128, make_uint4(0, 0, 0, 1));
	for (uint64_t x = 0; x < n_ctx; ++x) {
		if (dense[x]) normalise(&hist[x * nb], nb, &freq[x * nb]); else std::memcpy(&freq[x * nb], &fbf[(x & (n_fb - 1)) * nb], 2 * nb);
		uint32_t acc = 0; for (uint32_t k = 0; k < nb; ++k) { tab[x * nb + k] = freq[x * nb + k] ? rans_symbol(acc, freq[x * nb + k]) : make_uint4(0, 0, 0, 1); acc += freq[x * nb + k]; }
	}
	for (uint32_t b = 0; b < nb; ++b) { normalise(&mh[b * 128], 128, &mf[b * 128]); uint32_t acc = 0; for (uint32_t k = 0; k < 128; ++k) { mtab[b * 128 + k] = mf[b * 128 + k] ? rans_symbol(acc, mf[b * 128 + k]) : make_uint4(0, 0, 0, 1); acc += mf[b * 128 + k]; } }
	std::vector<uint8_t> hdr;
	hdr.insert(hdr.end(), {'Q', 'B', '0', '2'}); put(hdr, nb); put(hdr, P.level); for (int i = 0; i < 4; ++i) put(hdr, P.thr[i]); put(hdr, (uint64_t)n); put(hdr, np); put(hdr, P.cbits);
	for (uint32_t i = 0; i < nb * 128; ++i) put(hdr, mf[i]);
	for (uint32_t x = 0; x < n_fb; ++x) for (uint32_t k = 0; k + 1 < nb; ++k) put(hdr, fbf[(size_t)x * nb + k]);
	{
		uint32_t nd = 0; for (uint64_t x = 0; x < n_ctx; ++x) nd += dense[x];
		put(hdr, nd);
		uint64_t prev = 0;
		for (uint64_t x = 0; x < n_ctx; ++x) if (dense[x]) {
			uint64_t gap = x - prev; prev = x;
			do { uint8_t by = (uint8_t)(gap & 127); gap >>= 7; if (gap) by |= 128; hdr.push_back(by); } while (gap);
			for (uint32_t k = 0; k + 1 < nb; ++k) put(hdr, freq[x * nb + k]);
		}
	}
	CLB_CUDA(c, cudaMemcpyAsync(d_tab, tab.data(), sizeof(uint4) * tab.size(), cudaMemcpyHostToDevice, s));
	CLB_CUDA(c, cudaMemcpyAsync(d_mtab, mtab.data(), sizeof(uint4) * mtab.size(), cudaMemcpyHostToDevice, s));
	a.tab = d_tab; a.mtab = d_mtab;
	tr.mark("tables (host)");
	// ---- pass 2 in chunks of packs (the temp holds one 16-bit word per symbol at worst) ----
	uint32_t* d_pack_first = nullptr;
	CLB_CUDA(c, dalloc((void**)&d_pack_first, sizeof(uint32_t) * (np + 1)));
	CLB_CUDA(c, cudaMemcpyAsync(d_pack_first, pack_first.data(), sizeof(uint32_t) * (np + 1), cudaMemcpyHostToDevice, s));
	CLB_CUDA(c, c->qs.reserve(hdr.size() + tot / 3 + (uint64_t)np * (4 + 4 * QB_LANES + 4 * QB_LANES * QB_STATES) + 1024, s, false));
	CLB_CUDA(c, cudaMemcpyAsync(c->qs.p, hdr.data(), hdr.size(), cudaMemcpyHostToDevice, s));
	uint64_t out_at = hdr.size();
	uint64_t chunk_syms = 1ull << 31;       // more symbols per chunk = more warps in flight; the temp is 2 bytes per symbol
	{ const uint64_t avail = dev_mem_available(); while (chunk_syms < (1ull << 33) && 4 * chunk_syms + (16ull << 30) < avail) chunk_syms <<= 1; }
	uint16_t* d_tmp = nullptr; uint64_t tmp_cap = 0;      // one temp for all chunks (grown if a later chunk is larger)
	for (uint32_t p0 = 0; p0 < np;) {
		uint32_t p1 = p0; uint64_t syms = 0;
		while (p1 < np && (p1 == p0 || syms + (h_off[pack_first[p1 + 1]] - h_off[pack_first[p1]]) <= chunk_syms)) { syms += h_off[pack_first[p1 + 1]] - h_off[pack_first[p1]] + 2ull * nb * (pack_first[p1 + 1] - pack_first[p1]); ++p1; }
		const uint32_t cp = p1 - p0, nl = cp * QB_LANES;
		std::vector<uint64_t> lane_off(nl + 1, 0);
		for (uint32_t p = p0; p < p1; ++p)
			for (uint32_t r = pack_first[p]; r < pack_first[p + 1]; ++r) lane_off[(size_t)(p - p0) * QB_LANES + (r - pack_first[p]) % QB_LANES + 1] += c->h_rd_len[nc + r] + 2ull * nb;
		for (uint32_t i = 0; i < nl; ++i) lane_off[i + 1] += lane_off[i];
		uint64_t* d_lane_off = nullptr; uint32_t* d_words = nullptr; uint32_t* d_state = nullptr; uint64_t* d_dst = nullptr; uint64_t* d_phdr = nullptr;
		Tmp ct{{}, s};
		auto calloc_ = [&](void** q, uint64_t bytes) { cudaError_t e = dev_malloc(q, bytes ? bytes : 1, s); if (e == cudaSuccess) ct.v.push_back(*q); return e; };
		CLB_CUDA(c, calloc_((void**)&d_lane_off, sizeof(uint64_t) * (nl + 1))); if (lane_off[nl] + 8 > tmp_cap) { tmp_cap = lane_off[nl] + 8; CLB_CUDA(c, dalloc((void**)&d_tmp, sizeof(uint16_t) * tmp_cap)); }
		CLB_CUDA(c, calloc_((void**)&d_words, sizeof(uint32_t) * nl)); CLB_CUDA(c, calloc_((void**)&d_state, sizeof(uint32_t) * nl * QB_STATES));
		CLB_CUDA(c, calloc_((void**)&d_dst, sizeof(uint64_t) * nl)); CLB_CUDA(c, calloc_((void**)&d_phdr, sizeof(uint64_t) * cp));
		CLB_CUDA(c, cudaMemcpyAsync(d_lane_off, lane_off.data(), sizeof(uint64_t) * (nl + 1), cudaMemcpyHostToDevice, s));
		QEnc e{d_pack_first, p0, cp, d_lane_off, d_tmp, d_words, d_state};
		CLB_TIMED3(c, K_QUAL, (k_q_encode<<<(nl * 32 + 127) / 128, 128, 0, s>>>(a, e)));
		CLB_LAUNCH_CHECK(c, "k_q_encode");
		tr.mark("k_q_encode");
		std::vector<uint32_t> words(nl);
		CLB_CUDA(c, cudaMemcpyAsync(words.data(), d_words, sizeof(uint32_t) * nl, cudaMemcpyDeviceToHost, s));
		CLB_CUDA(c, cudaStreamSynchronize(s));
		std::vector<uint64_t> dst(nl), phdr(cp);
		for (uint32_t p = 0; p < cp; ++p) {
			phdr[p] = out_at; out_at += 4 + 4 * QB_LANES;
			for (uint32_t l = 0; l < QB_LANES; ++l) { dst[(size_t)p * QB_LANES + l] = out_at; out_at += 4ull * QB_STATES + 2ull * words[(size_t)p * QB_LANES + l]; }
		}
		CLB_CUDA(c, c->qs.reserve(out_at + 16, s, true, phdr[0]));
		CLB_CUDA(c, cudaMemcpyAsync(d_dst, dst.data(), sizeof(uint64_t) * nl, cudaMemcpyHostToDevice, s));
		CLB_CUDA(c, cudaMemcpyAsync(d_phdr, phdr.data(), sizeof(uint64_t) * cp, cudaMemcpyHostToDevice, s));
		CLB_TIMED3(c, K_QUAL, (k_q_gather<<<(nl * 32 + 127) / 128, 128, 0, s>>>(e, d_dst, d_phdr, c->qs.p)));
		CLB_LAUNCH_CHECK(c, "k_q_gather");
		CLB_CUDA(c, cudaStreamSynchronize(s));
		tr.mark("k_q_gather");
		p0 = p1;
	}
	c->qs_total = out_at;
	c->qual_done = true;
	return CLB_OK;
}

} // namespace clb
