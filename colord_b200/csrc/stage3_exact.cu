// stage3_exact.cu — the reference's OWN stage-3 byte streams produced on the device ("compat" streams):
// the parts of the "dna", "qual" and "header" streams of a reference archive, byte for byte, so that `colord decompress` of the
// unmodified reference reads an archive written here and the archive size is the reference's by construction.
//
// The reference codes each stream with ONE serial chain (entr_read.h:56-80, entr_qual.h:100-126, entr_header.cpp:23-46): adaptive
// frequency models looked up by context (rc.h:34-221, :487-740; context_hm.h) feeding one 64-bit range coder (sub_rc.h:72-211) that
// is restarted for every read pack while the models live on.  That chain is taken apart along the two dependencies it really has:
//   1. events      which (family, context, symbol, excluded symbols) a read / header turns into depends only on the input — all
//                  reads are walked in parallel and write their events at their place in stream order (dna_model.h with EXACT,
//                  the quality contexts of quality_coder_impl.cpp:78-450, the header events of id_coder.cpp:210-383);
//   2. models      the state of a context's model when an event meets it depends only on the EARLIER EVENTS OF THE SAME CONTEXT:
//                  a stable radix sort by (family, context) brings every context's events together in stream order, one thread per
//                  context replays its model (counts, +adder, halving at max_total; Encode / EncodeExcluding rc.h:780-803,
//                  :861-893) and leaves (frequency, cumulative frequency, total) at each event;
//   3. range coder restarts per pack, so the packs are independent: one warp per pack runs sub_rc.h:83-201 over its events'
//                  triples (the only serial loop left: one pack = 4 MiB of bases), parts are then laid out back to back.
// Memory is ~30 bytes per event for the duration of a stream, so this path is meant for inputs up to a few Gbases per GPU; the
// native containers (stage3_dna.cu, stage3_qual.cu, ...) stay the path for the whole-genome scale.
// CPU restatement pinned byte for byte on the stock binary's archives: oracle/stage3_exact.c (tests/test_oracle_exact.py); the
// device bytes are compared with it and with the stock binary's parts in tests/test_gpu_exact.py.
#include "ctx.h"
#include "dna_model.h"
#include "hdr_model.h"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <cstring>
#include <vector>

namespace clb {

constexpr uint32_t X_CTX_BITS = 44, X_KEY_BITS = 48;      // key = family << 44 | context
constexpr uint32_t X_MAX_FAM = 16;
struct XFam { uint32_t n_sym, max_total, adder; };
struct XFams { XFam f[X_MAX_FAM]; };

// ---- event sinks: sizing pass and writing pass of a walk -------------------------------------------------------------
struct XCountSink {
	uint32_t n = 0;
	CLB_D void put(uint32_t, uint64_t, uint32_t) { ++n; }
	CLB_D void putx(uint32_t, uint64_t, uint32_t, uint32_t) { ++n; }
};
struct XWriteSink {
	uint64_t* key; uint16_t* info; uint64_t at; uint32_t* bad;
	CLB_D void putx(uint32_t f, uint64_t ctx, uint32_t sym, uint32_t excl)
	{
		if (ctx >> X_CTX_BITS) atomicExch(bad, 1u);
		key[at] = ((uint64_t)f << X_CTX_BITS) | (ctx & ((1ull << X_CTX_BITS) - 1));
		info[at] = (uint16_t)(sym | (excl << 8));
		++at;
	}
	CLB_D void put(uint32_t f, uint64_t ctx, uint32_t sym) { putx(f, ctx, sym, 0); }
};

// ---- 2. models: one thread per context -------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_x_iota(uint32_t* __restrict__ v, uint64_t n)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) v[i] = (uint32_t)i;
}
// What the model leaves at an event for the range coder: x = frequency | cumulative frequency << 21 | total << 42 (total < max_total
// + adder <= 2^20 + 64) and y = floor((2^64 - 1) / total), with which the coder's `range / total` is one multiply-high and a fix-up.
CLB_D ulonglong2 x_triple(uint32_t freq, uint32_t cum, uint32_t tot) { return make_ulonglong2((uint64_t)freq | ((uint64_t)cum << 21) | ((uint64_t)tot << 42), ~0ull / tot); }

__global__ void __launch_bounds__(256) k_x_heads(const uint64_t* __restrict__ key, uint64_t n, uint8_t* __restrict__ flag)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) flag[i] = i == 0 || key[i - 1] != key[i];
}
// One WARP per context (heads[s] = first sorted position of context s; grid-stride over the contexts).  Counts, +adder, halving at
// max_total, Encode / EncodeExcluding: rc.h:122-221, :780-803, :861-893.
//   alphabets up to 8 symbols (every family with exclusions is one): 32 events per step — lane i takes event i, the counts it meets
//     are the step's base counts + adder x (events of that symbol in the lanes before it: one ballot per symbol); a step ends where the
//     total would reach max_total, the halving is applied between steps
//   larger alphabets: events one by one, the counters spread over the lanes (lane l holds symbols l, l + 32, ...), cumulative frequency by
//     a warp reduction
__global__ void __launch_bounds__(128) k_x_model(const uint64_t* __restrict__ key, const uint32_t* __restrict__ idx, const uint16_t* __restrict__ info,
	uint64_t n, const uint32_t* __restrict__ heads, const uint32_t* __restrict__ n_heads_p, XFams fams, ulonglong2* __restrict__ triple)
{
	const uint32_t lane = threadIdx.x & 31, n_heads = *n_heads_p;
	const uint32_t lt = (1u << lane) - 1;
	for (uint64_t sg = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; sg < n_heads; sg += ((uint64_t)gridDim.x * blockDim.x) >> 5) {
		const uint64_t i0 = heads[sg], i1 = sg + 1 < n_heads ? heads[sg + 1] : n;
		const XFam F = fams.f[key[i0] >> X_CTX_BITS];
		const uint32_t A = F.n_sym;
		if (A <= 8) {
			uint32_t cb[8];
#pragma unroll
			for (int s = 0; s < 8; ++s) cb[s] = 1;
			uint32_t tb = A;
			for (uint64_t i = i0; i < i1;) {
				const bool valid = i + lane < i1;
				uint32_t e = 0, sym = 0, excl = 0;
				if (valid) { e = idx[i + lane]; const uint32_t in = info[e]; sym = in & 0xff; excl = in >> 8; }
				uint32_t n_step = __popc(__ballot_sync(0xffffffffu, valid));
				n_step = min(n_step, (F.max_total - tb + F.adder - 1) / F.adder);      // the n_step-th event's update is the first to reach max_total
				const bool mine = lane < n_step;
				uint32_t cum = 0, tot = tb + F.adder * lane, freq = 0;
#pragma unroll
				for (int s = 0; s < 8; ++s) {
					const uint32_t m = __ballot_sync(0xffffffffu, mine && sym == (uint32_t)s);
					const uint32_t c = cb[s] + F.adder * __popc(m & lt);
					const bool ex = (excl >> s) & 1;
					if ((uint32_t)s < sym && !ex) cum += c;
					if (ex && (uint32_t)s < A) tot -= c;
					if ((uint32_t)s == sym) freq = c;
					cb[s] += F.adder * __popc(m);
				}
				if (mine) triple[e] = x_triple(freq, cum, tot);
				tb += F.adder * n_step;
				while (tb >= F.max_total) {
					uint32_t t = 0;
#pragma unroll
					for (int s = 0; s < 8; ++s) { cb[s] = (cb[s] + 1) >> 1; if ((uint32_t)s < A) t += cb[s]; }
					tb = t;
				}
				i += n_step;
			}
		} else {
			uint32_t c[8];                                    // symbol lane + 32 j
#pragma unroll
			for (int j = 0; j < 8; ++j) c[j] = lane + 32 * j < A ? 1u : 0u;
			uint32_t total = A;
			for (uint64_t i = i0; i < i1; ++i) {
				const uint32_t e = idx[i], sym = info[e] & 0xff;
				uint32_t part = 0, mine = 0;
#pragma unroll
				for (int j = 0; j < 8; ++j) { const uint32_t s = lane + 32 * j; if (s < sym) part += c[j]; if (s == sym) mine = c[j]; }
				const uint32_t cum = __reduce_add_sync(0xffffffffu, part);
				const uint32_t freq = __shfl_sync(0xffffffffu, mine, sym & 31);
				if (lane == 0) triple[e] = x_triple(freq, cum, total);
#pragma unroll
				for (int j = 0; j < 8; ++j) if (lane + 32 * j == sym) c[j] += F.adder;
				total += F.adder;
				while (total >= F.max_total) {
					uint32_t t = 0;
#pragma unroll
					for (int j = 0; j < 8; ++j) { c[j] = lane + 32 * j < A ? (c[j] + 1) >> 1 : 0u; t += c[j]; }
					total = __reduce_add_sync(0xffffffffu, t);
				}
			}
		}
	}
}

// ---- 3. range coder: one WARP per pack (sub_rc.h:72-211) --------------------------------------------------------------
// The coder's state is one serial chain per pack.  A lone thread walking its pack's triples pays a DRAM round trip for every other
// event (measured: ~1000 cycles per event, profiles/r02e_compat_launches.csv), so the warp fetches 32 events at a time with one
// coalesced load (the next 32 already on their way), every lane then follows the same chain — event i's triple comes from lane i by
// shuffle, the state is identical in all lanes — and lane 0 stores the bytes.
__global__ void __launch_bounds__(128) k_x_code(const ulonglong2* __restrict__ triple, const uint64_t* __restrict__ pack_ev, uint32_t n_packs,
	uint8_t* __restrict__ tmp, const uint64_t* __restrict__ tmp_off, uint64_t* __restrict__ part_bytes)
{
	const uint32_t p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (p >= n_packs) return;
	uint8_t* out = tmp + tmp_off[p];
	uint64_t n = 0;
	unsigned long long low = 0, range = 0xff00000000000000ULL;
	const uint64_t e0 = pack_ev[p], e1 = pack_ev[p + 1];
	ulonglong2 nxt = e0 + lane < e1 ? triple[e0 + lane] : make_ulonglong2(0, 1);
	for (uint64_t e = e0; e < e1; e += 32) {
		const ulonglong2 cur = nxt;
		if (e + 32 + lane < e1) nxt = triple[e + 32 + lane];
		const uint32_t m = (uint32_t)min((uint64_t)32, e1 - e);
		for (uint32_t i = 0; i < m; ++i) {
			const unsigned long long x = __shfl_sync(0xffffffffu, cur.x, i), inv = __shfl_sync(0xffffffffu, cur.y, i);
			const uint32_t freq = (uint32_t)(x & 0x1fffff), cum = (uint32_t)((x >> 21) & 0x1fffff), tot = (uint32_t)(x >> 42);
			unsigned long long q = __umul64hi(range, inv);       // range / tot: inv = floor((2^64 - 1) / tot) gives the quotient or one less
			if (range - q * tot >= tot) ++q;
			range = q;
			low += range * cum;
			range *= freq;
			for (int k = 0; k < 8 && range <= 0x00ffffffffffffULL; ++k) {      // UNROLL_FREQUENCY_CODING: at most 8 bytes per symbol
				if ((low ^ (low + range)) & 0xff00000000000000ULL) { const unsigned long long y = low; range = (y | 0x00ffffffffffffULL) - y; }
				if (lane == 0) out[n] = (uint8_t)(low >> 56);
				++n;
				low <<= 8; range <<= 8;
			}
		}
	}
	for (int i = 0; i < 8; ++i) { if (lane == 0) out[n] = (uint8_t)(low >> 56); ++n; low <<= 8; }      // End (sub_rc.h:203-210)
	if (lane == 0) part_bytes[p] = n;
}
__global__ void __launch_bounds__(256) k_x_compact(const uint8_t* __restrict__ tmp, const uint64_t* __restrict__ tmp_off, const uint64_t* __restrict__ dst_off,
	const uint64_t* __restrict__ part_bytes, uint8_t* __restrict__ out)
{
	const uint32_t p = blockIdx.x;
	const uint8_t* src = tmp + tmp_off[p]; uint8_t* dst = out + dst_off[p];
	for (uint64_t i = threadIdx.x; i < part_bytes[p]; i += blockDim.x) dst[i] = src[i];
}

struct XTmp { std::vector<void*> v; cudaStream_t s; ~XTmp() { for (void* p : v) dev_free_async(p, s); } };

// events in stream order -> the parts of the stream.  pack_ev[n_packs + 1]: first event of every pack.
static clb_status x_code_stream(clb_ctx* c, cudaStream_t s, int kid, const XFams& fams, uint64_t* d_key, uint16_t* d_info, uint64_t n_ev,
	const std::vector<uint64_t>& pack_ev, clb::DevBuf<uint8_t>& out, std::vector<uint64_t>& part_sizes, uint64_t& total)
{
	const uint32_t np = (uint32_t)pack_ev.size() - 1;
	if (n_ev >= 0xfffffff0ull) return fail(c, CLB_ERR_CAPACITY, "compat streams: more than 2^32 events in one stream (use the native containers for inputs of this size)");
	XTmp tmp{{}, s};
	auto dalloc = [&](void** p, uint64_t bytes) { cudaError_t e = dev_malloc(p, bytes ? bytes : 1, s); if (e == cudaSuccess) tmp.v.push_back(*p); return e; };
	auto timed_begin = [&]() { if (s == c->stream3) prof_begin3(c, kid); else prof_begin(c, kid); };
	auto timed_end = [&]() { if (s == c->stream3) prof_end3(c); else prof_end(c); };
	ulonglong2* d_triple = nullptr;
	CLB_CUDA(c, dalloc((void**)&d_triple, sizeof(ulonglong2) * (n_ev + 1)));
	if (n_ev) {
		uint64_t* d_key2 = nullptr; uint32_t* d_idx = nullptr; uint32_t* d_idx2 = nullptr; void* d_sort = nullptr; size_t sort_bytes = 0;
		CLB_CUDA(c, dalloc((void**)&d_key2, sizeof(uint64_t) * n_ev)); CLB_CUDA(c, dalloc((void**)&d_idx, sizeof(uint32_t) * n_ev)); CLB_CUDA(c, dalloc((void**)&d_idx2, sizeof(uint32_t) * n_ev));
		timed_begin(); k_x_iota<<<(uint32_t)((n_ev + 255) / 256), 256, 0, s>>>(d_idx, n_ev); timed_end(); CLB_LAUNCH_CHECK(c, "k_x_iota");
		// stable LSD radix sort of (key, event index) pairs (CUB, part of the CUDA toolkit): plumbing between the walk and the models
		CLB_CUDA(c, cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, d_key, d_key2, d_idx, d_idx2, (unsigned long long)n_ev, 0, (int)X_KEY_BITS, s));
		CLB_CUDA(c, dalloc(&d_sort, sort_bytes));
		timed_begin();
		cudaError_t e = cub::DeviceRadixSort::SortPairs(d_sort, sort_bytes, d_key, d_key2, d_idx, d_idx2, (unsigned long long)n_ev, 0, (int)X_KEY_BITS, s);
		timed_end();
		if (e != cudaSuccess) return cuda_fail(c, e, "cub::DeviceRadixSort::SortPairs");
		++c->launches;
		// the first sorted position of every context (flag + CUB select: plumbing), then one warp per context
		uint8_t* d_flag = nullptr; uint32_t* d_heads = nullptr; uint32_t* d_n_heads = nullptr; void* d_sel = nullptr; size_t sel_bytes = 0;
		CLB_CUDA(c, dalloc((void**)&d_flag, n_ev)); CLB_CUDA(c, dalloc((void**)&d_heads, sizeof(uint32_t) * n_ev)); CLB_CUDA(c, dalloc((void**)&d_n_heads, 4));
		timed_begin(); k_x_heads<<<(uint32_t)((n_ev + 255) / 256), 256, 0, s>>>(d_key2, n_ev, d_flag); timed_end(); CLB_LAUNCH_CHECK(c, "k_x_heads");
		cub::CountingInputIterator<uint32_t> first(0);
		CLB_CUDA(c, cub::DeviceSelect::Flagged(nullptr, sel_bytes, first, d_flag, d_heads, d_n_heads, (unsigned long long)n_ev, s));
		CLB_CUDA(c, dalloc(&d_sel, sel_bytes));
		timed_begin();
		e = cub::DeviceSelect::Flagged(d_sel, sel_bytes, first, d_flag, d_heads, d_n_heads, (unsigned long long)n_ev, s);
		timed_end();
		if (e != cudaSuccess) return cuda_fail(c, e, "cub::DeviceSelect::Flagged");
		++c->launches;
		timed_begin(); k_x_model<<<c->n_sm * 16, 128, 0, s>>>(d_key2, d_idx2, d_info, n_ev, d_heads, d_n_heads, fams, d_triple); timed_end(); CLB_LAUNCH_CHECK(c, "k_x_model");
	}
	// temp slot of a pack: at most 21 bits leave the coder per event (frequency >= 1 of a total < 2^21), + the 8-byte flush
	std::vector<uint64_t> tmp_off(np + 1, 0);
	for (uint32_t p = 0; p < np; ++p) tmp_off[p + 1] = tmp_off[p] + 3 * (pack_ev[p + 1] - pack_ev[p]) + 16;
	uint64_t* d_pack_ev = nullptr; uint64_t* d_tmp_off = nullptr; uint64_t* d_bytes = nullptr; uint64_t* d_dst = nullptr; uint8_t* d_tmp = nullptr;
	CLB_CUDA(c, dalloc((void**)&d_pack_ev, sizeof(uint64_t) * (np + 1))); CLB_CUDA(c, dalloc((void**)&d_tmp_off, sizeof(uint64_t) * (np + 1)));
	CLB_CUDA(c, dalloc((void**)&d_bytes, sizeof(uint64_t) * (np + 1))); CLB_CUDA(c, dalloc((void**)&d_dst, sizeof(uint64_t) * (np + 1)));
	CLB_CUDA(c, dalloc((void**)&d_tmp, tmp_off[np] + 16));
	CLB_CUDA(c, cudaMemcpyAsync(d_pack_ev, pack_ev.data(), sizeof(uint64_t) * (np + 1), cudaMemcpyHostToDevice, s));
	CLB_CUDA(c, cudaMemcpyAsync(d_tmp_off, tmp_off.data(), sizeof(uint64_t) * (np + 1), cudaMemcpyHostToDevice, s));
	part_sizes.assign(np, 0);
	total = 0;
	if (!np) return CLB_OK;
	timed_begin(); k_x_code<<<(np + 3) / 4, 128, 0, s>>>(d_triple, d_pack_ev, np, d_tmp, d_tmp_off, d_bytes); timed_end(); CLB_LAUNCH_CHECK(c, "k_x_code");
	CLB_CUDA(c, cudaMemcpyAsync(part_sizes.data(), d_bytes, sizeof(uint64_t) * np, cudaMemcpyDeviceToHost, s));
	CLB_CUDA(c, cudaStreamSynchronize(s));
	std::vector<uint64_t> dst(np + 1, 0);
	for (uint32_t p = 0; p < np; ++p) dst[p + 1] = dst[p] + part_sizes[p];
	total = dst[np];
	CLB_CUDA(c, out.reserve(total + 16, s, false));
	CLB_CUDA(c, cudaMemcpyAsync(d_dst, dst.data(), sizeof(uint64_t) * (np + 1), cudaMemcpyHostToDevice, s));
	timed_begin(); k_x_compact<<<np, 256, 0, s>>>(d_tmp, d_tmp_off, d_dst, d_bytes, out.p); timed_end(); CLB_LAUNCH_CHECK(c, "k_x_compact");
	CLB_CUDA(c, cudaStreamSynchronize(s));
	return CLB_OK;
}

// pack boundaries as read indices: pack_sizes given, or the reference's pack rule (in_reads.cpp:62-76) over the read lengths
static clb_status x_packs(clb_ctx* c, const uint32_t* pack_sizes, uint32_t n_packs, uint64_t n, std::vector<uint64_t>& pack_first)
{
	pack_first.assign(1, 0);
	if (pack_sizes) {
		uint64_t at = 0;
		for (uint32_t i = 0; i < n_packs; ++i) { at += pack_sizes[i]; pack_first.push_back(at); }      // empty packs are parts too (8-byte flush)
		if (at != n) return fail(c, CLB_ERR_BAD_ARG, "pack_sizes do not sum to the number of reads");
	} else {
		uint64_t bytes = 0;
		for (uint64_t i = 0; i < n; ++i) { bytes += (uint64_t)c->h_rd_len[c->n_context + i] + 1; if (bytes >= (2u << 21)) { bytes = 0; pack_first.push_back(i + 1); } }
		if (pack_first.back() != n) pack_first.push_back(n);
	}
	return CLB_OK;
}

// ================================================================================================ DNA stream
struct XDArgs { DnaReads R; DnaModel M; uint32_t n_reads; uint32_t* n_ev; const uint64_t* ev_off; uint64_t* key; uint16_t* info; uint32_t* bad; };

CLB_D uint32_t x_flag_ctx(const DnaReads& R, uint32_t r)      // the last four read flags, whatever the pack (dna_coder.cpp:459-462)
{
	uint32_t ctx = 0;
	for (int k = 4; k >= 1; --k) if (r >= (uint32_t)k) ctx = ((ctx << 2) + read_flag_of(R, r - (uint32_t)k)) & 0xff;
	return ctx;
}
template <bool WRITE>
__global__ void __launch_bounds__(128) k_xd_events(XDArgs a)
{
	const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= a.n_reads) return;
	if (WRITE) { XWriteSink s{a.key, a.info, a.ev_off[r], a.bad}; dna_walk<true>(a.M, a.R, r, x_flag_ctx(a.R, r), s); }
	else { XCountSink s; dna_walk<true>(a.M, a.R, r, x_flag_ctx(a.R, r), s); a.n_ev[r] = s.n; }
}

clb_status s3x_dna_encode(clb_ctx* c, uint32_t level, const uint32_t* pack_sizes, uint32_t n_packs)
{
	cudaStream_t s = c->stream;
	if (!c->enc_done) return fail(c, CLB_ERR_STATE, "clb_xdna_encode before clb_encode");
	if (level < 1 || level > 3) return fail(c, CLB_ERR_BAD_ARG, "clb_xdna_encode: level must be 1, 2 or 3");
	if (c->prm.max_candidates > 32) return fail(c, CLB_ERR_BAD_ARG, "clb_xdna_encode: at most 32 candidates");
	// context reads (the pseudo-reads of a reference genome) are reference reads in front of the input's reads: not coded, but they
	// count in the read ids — the reference starts its coder at start_read_id = n_ref_genome_pseudo_reads (compression.cpp:641)
	const uint64_t nc = c->n_context, n = c->n_reads - nc;
	std::vector<uint64_t> pack_first;
	{ const clb_status st = x_packs(c, pack_sizes, n_packs, n, pack_first); if (st != CLB_OK) return st; }
	const uint32_t np = (uint32_t)pack_first.size() - 1;
	XTmp tmp{{}, s};
	auto dalloc = [&](void** p, uint64_t bytes) { cudaError_t e = dev_malloc(p, bytes ? bytes : 1, s); if (e == cudaSuccess) tmp.v.push_back(*p); return e; };
	XDArgs a{};
	a.M = make_dna_model(level, c->prm.max_candidates);
	a.R = DnaReads{c->pk.p, c->rd_start.p, c->rd_len.p, c->d_ref_to_read, c->es.p, c->es_off, (uint32_t)nc};
	a.n_reads = (uint32_t)n;
	uint64_t* d_ev_off = nullptr;
	CLB_CUDA(c, dalloc((void**)&a.n_ev, sizeof(uint32_t) * (n + 1))); CLB_CUDA(c, dalloc((void**)&d_ev_off, sizeof(uint64_t) * (n + 1))); CLB_CUDA(c, dalloc((void**)&a.bad, 4));
	CLB_CUDA(c, cudaMemsetAsync(a.bad, 0, 4, s));
	const uint32_t nblk = (uint32_t)((n + 127) / 128);
	uint64_t n_ev = 0;
	if (n) {
		CLB_TIMED(c, K_DNA, (k_xd_events<false><<<nblk, 128, 0, s>>>(a))); CLB_LAUNCH_CHECK(c, "k_xd_events<count>");
		const clb_status st = exclusive_scan(c, a.n_ev, n, d_ev_off, &n_ev);
		if (st != CLB_OK) return st;
	}
	std::vector<uint64_t> ev_off(n + 1, 0);
	if (n) CLB_CUDA(c, cudaMemcpyAsync(ev_off.data(), d_ev_off, sizeof(uint64_t) * n, cudaMemcpyDeviceToHost, s));
	CLB_CUDA(c, cudaStreamSynchronize(s));
	ev_off[n] = n_ev;
	CLB_CUDA(c, dalloc((void**)&a.key, sizeof(uint64_t) * (n_ev + 1))); CLB_CUDA(c, dalloc((void**)&a.info, sizeof(uint16_t) * (n_ev + 1)));
	a.ev_off = d_ev_off;
	if (n) { CLB_TIMED(c, K_DNA, (k_xd_events<true><<<nblk, 128, 0, s>>>(a))); CLB_LAUNCH_CHECK(c, "k_xd_events<write>"); }
	uint32_t bad = 0;
	CLB_CUDA(c, cudaMemcpyAsync(&bad, a.bad, 4, cudaMemcpyDeviceToHost, s));
	CLB_CUDA(c, cudaStreamSynchronize(s));
	if (bad) return fail(c, CLB_ERR_CAPACITY, "clb_xdna_encode: a context does not fit the event key");
	// model parameters: dna_coder.h:48-60, dna_coder.cpp:1316-1336
	XFams F{};
	const XFam f0[F_COUNT] = {{3, 1u << 15, 1}, {32, 1u << 18, 8}, {256, 1u << 18, 8}, {4, 1u << 10, 1}, {5, 1u << 10, 1}, {256, 1u << 13, 1}, {2, 1u << 15, 1},
		{8, 1u << 15, 1}, {24, 1u << 15, 1}, {256, 1u << 15, 1}, {256, 1u << 15, 1}, {2, 1u << 15, 1}, {c->prm.max_candidates, 1u << 13, 1}};
	for (uint32_t f = 0; f < F_COUNT; ++f) F.f[f] = f0[f];
	std::vector<uint64_t> pack_ev(np + 1);
	for (uint32_t p = 0; p <= np; ++p) pack_ev[p] = ev_off[pack_first[p]];
	c->xd_packs.assign(np, 0);
	for (uint32_t p = 0; p < np; ++p) c->xd_packs[p] = pack_first[p + 1] - pack_first[p];
	return x_code_stream(c, s, K_DNA, F, a.key, a.info, n_ev, pack_ev, c->xd, c->xd_parts, c->xd_total);
}

// ================================================================================================ plain sequences (the stored reference genome)
// CReferenceGenome::Store(archive) (reference_genome.cpp:319-360): every sequence of the genome goes through a CDNACoder of its own as a
// plain read (start_plain + one plain tuple per base) at "level 9", which Init maps to one symbol of history (dna_coder.cpp:1275-1280);
// one part for the whole genome, metadata = number of sequences.
struct XPArgs { const uint8_t* bases; const uint64_t* off; uint32_t n_seqs; uint32_t n_s; const uint64_t* ev_off; uint64_t* key; uint16_t* info; uint32_t* bad; };
CLB_D uint32_t xp_code(uint8_t ch) { return ch == 'C' ? 1u : ch == 'G' ? 2u : ch == 'T' ? 3u : 0u; }
template <class Sink>
CLB_D void xp_head(uint64_t len64, Sink& sink)      // read flag + encode_read_len (dna_coder.cpp:440-463, :1004-1056)
{
	sink.put(F_FLAG, 0, 0);
	uint32_t len = (uint32_t)len64;
	const uint32_t nbits = ilog2_bits(len);
	sink.put(F_LENBITS, 0, nbits);
	if (nbits >= 2) {
		uint64_t ctx = (uint64_t)nbits << 3;
		len -= 1u << (nbits - 1);
		uint32_t prefix = len, suffix = 0;
		if (nbits > 9) { prefix = len >> (nbits - 9); suffix = len - (prefix << (nbits - 9)); }
		sink.put(F_LENDATA, ctx, prefix);
		if (nbits > 9) { ctx += 4; for (int nb = (int)nbits - 9; nb > 0; nb -= 8) { sink.put(F_LENDATA, ctx, suffix & 0xff); suffix >>= 8; ++ctx; } }
	}
}
__global__ void __launch_bounds__(128) k_xp_heads(XPArgs a)
{
	const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
	if (q >= a.n_seqs) return;
	XWriteSink s{a.key, a.info, a.ev_off[q], a.bad};
	xp_head(a.off[q + 1] - a.off[q], s);
}
__global__ void __launch_bounds__(256) k_xp_symbols(XPArgs a, uint64_t n_bases)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_bases) return;
	uint32_t lo = 0, hi = a.n_seqs;                       // sequence of base i: last off <= i
	while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (a.off[mid] <= i) lo = mid; else hi = mid; }
	const uint64_t at = i - a.off[lo];
	uint64_t ctx = 0;
	for (uint32_t k = a.n_s; k >= 1; --k) ctx = (ctx << 2) | (at >= k ? xp_code(a.bases[i - k]) : 3u);
	const uint64_t e = a.ev_off[lo + 1] - (a.off[lo + 1] - a.off[lo]) + at;      // the head events come first
	a.key[e] = ((uint64_t)F_SYM << X_CTX_BITS) | (ctx << 2);
	a.info[e] = (uint16_t)xp_code(a.bases[i]);
}

clb_status s3x_plain_encode(clb_ctx* c, const uint8_t* bases, const uint64_t* offsets, uint32_t n_seqs, uint32_t level)
{
	cudaStream_t s = c->stream;
	if (!n_seqs || !bases || !offsets) return fail(c, CLB_ERR_BAD_ARG, "clb_xplain_encode: no sequences");
	const uint64_t n_bases = offsets[n_seqs] - offsets[0];
	if (offsets[0] != 0) return fail(c, CLB_ERR_BAD_ARG, "clb_xplain_encode: offsets[0] must be 0");
	XTmp tmp{{}, s};
	auto dalloc = [&](void** p, uint64_t bytes) { cudaError_t e = dev_malloc(p, bytes ? bytes : 1, s); if (e == cudaSuccess) tmp.v.push_back(*p); return e; };
	XPArgs a{};
	a.n_seqs = n_seqs; a.n_s = level >= 3 ? (level == 3 ? 8u : 1u) : level == 2 ? 7u : level == 1 ? 5u : 1u;      // dna_coder.cpp:1253-1280: anything but 1, 2, 3 keeps one symbol
	std::vector<uint64_t> ev_off(n_seqs + 1, 0);
	for (uint32_t q = 0; q < n_seqs; ++q) {
		const uint64_t len = offsets[q + 1] - offsets[q];
		if (len >> 32) return fail(c, CLB_ERR_BAD_ARG, "clb_xplain_encode: a sequence of 4 Gbases or more");
		uint32_t nbits = 0; for (uint64_t x = len; x; x >>= 1) ++nbits;
		ev_off[q + 1] = ev_off[q] + 2 + (nbits >= 2 ? 1 + (nbits > 9 ? (nbits - 9 + 7) / 8 : 0) : 0) + len;
	}
	const uint64_t n_ev = ev_off[n_seqs];
	uint8_t* d_b = nullptr; uint64_t* d_o = nullptr; uint64_t* d_e = nullptr;
	CLB_CUDA(c, dalloc((void**)&d_b, n_bases + 16)); CLB_CUDA(c, dalloc((void**)&d_o, sizeof(uint64_t) * (n_seqs + 1))); CLB_CUDA(c, dalloc((void**)&d_e, sizeof(uint64_t) * (n_seqs + 1)));
	CLB_CUDA(c, dalloc((void**)&a.key, sizeof(uint64_t) * (n_ev + 1))); CLB_CUDA(c, dalloc((void**)&a.info, sizeof(uint16_t) * (n_ev + 1))); CLB_CUDA(c, dalloc((void**)&a.bad, 4));
	CLB_CUDA(c, cudaMemcpyAsync(d_b, bases, n_bases, cudaMemcpyHostToDevice, s));
	CLB_CUDA(c, cudaMemcpyAsync(d_o, offsets, sizeof(uint64_t) * (n_seqs + 1), cudaMemcpyHostToDevice, s));
	CLB_CUDA(c, cudaMemcpyAsync(d_e, ev_off.data(), sizeof(uint64_t) * (n_seqs + 1), cudaMemcpyHostToDevice, s));
	CLB_CUDA(c, cudaMemsetAsync(a.bad, 0, 4, s));
	a.bases = d_b; a.off = d_o; a.ev_off = d_e;
	CLB_TIMED(c, K_DNA, (k_xp_heads<<<(n_seqs + 127) / 128, 128, 0, s>>>(a))); CLB_LAUNCH_CHECK(c, "k_xp_heads");
	if (n_bases) { CLB_TIMED(c, K_DNA, (k_xp_symbols<<<(uint32_t)((n_bases + 255) / 256), 256, 0, s>>>(a, n_bases))); CLB_LAUNCH_CHECK(c, "k_xp_symbols"); }
	CLB_CUDA(c, cudaStreamSynchronize(s));
	XFams F{};
	const XFam f0[F_COUNT] = {{3, 1u << 15, 1}, {32, 1u << 18, 8}, {256, 1u << 18, 8}, {4, 1u << 10, 1}, {5, 1u << 10, 1}, {256, 1u << 13, 1}, {2, 1u << 15, 1},
		{8, 1u << 15, 1}, {24, 1u << 15, 1}, {256, 1u << 15, 1}, {256, 1u << 15, 1}, {2, 1u << 15, 1}, {2, 1u << 13, 1}};
	for (uint32_t f = 0; f < F_COUNT; ++f) F.f[f] = f0[f];
	const std::vector<uint64_t> pack_ev{0, n_ev};
	return x_code_stream(c, s, K_DNA, F, a.key, a.info, n_ev, pack_ev, c->xg, c->xg_parts, c->xg_total);
}

// ================================================================================================ quality stream
// mode: params.h QualityComprMode — 0 original, 1 / 2 / 3 quinary / quad / binary average, 4 / 5 / 6 quinary / quad / binary threshold,
// 7 average, 8 none.  source: 0 ONT, 1 PacBio CLR, 2 PacBio HiFi (the lossless mode's quantiser).
enum { XQ_SYM = 0, XQ_BYTE = 1 };
struct XQArgs {
	const uint64_t* pk; const uint64_t* rd_start; const uint32_t* rd_len; const uint8_t* quals; const uint64_t* qoff; const uint8_t* flags;
	const uint64_t* ev_off; uint64_t* key; uint16_t* info; uint32_t* bad;
	uint32_t mode, level, n_bins, bps, ncs;
	uint8_t map[96], quant[96];
};

__global__ void __launch_bounds__(128) k_xq_events(XQArgs a)
{
	__shared__ uint32_t st[128];
	const uint32_t r = blockIdx.x, n = a.rd_len[r];
	const uint64_t rs = a.rd_start[r];
	const uint8_t* q = a.quals + a.qoff[r];
	const uint8_t* fl = a.flags ? a.flags + a.qoff[r] : nullptr;
	uint64_t at = a.ev_off[r];
	auto emit = [&](uint64_t e, uint32_t f, uint64_t ctx, uint32_t sym) { a.key[e] = ((uint64_t)f << X_CTX_BITS) | ctx; a.info[e] = (uint16_t)sym; };
	// AVG (quality_coder_impl.cpp:821-834): (uint32)(x * 256) as two bytes, the second under the first
	auto emit_avg = [&](uint64_t e, uint64_t ctx, double x) { const uint32_t v = (uint32_t)__dmul_rn(x, 256.0); emit(e, XQ_BYTE, ctx, (v >> 8) & 0xff); emit(e + 1, XQ_BYTE, (uint64_t)((v >> 8) & 0xff) + 0x100ull, v & 0xff); };
	const bool averages = a.mode >= 1 && a.mode <= 3;
	if (averages || a.mode == 7) {
		st[threadIdx.x] = 0;
		__syncthreads();
		for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(&st[q[i] & 127], 1u);
		__syncthreads();
		if (threadIdx.x == 0) {
			if (a.mode == 7) {      // encode_average impl:441-450: sum of (q - 33) in read order = the same integer in any order
				double sum = 0.0;
				for (uint32_t v = 33; v < 128; ++v) sum += (double)(v - 33u) * st[v];
				emit_avg(at, 0ull, __ddiv_rn(sum, (double)n));
			} else {                 // encode_*_average impl:130-310
				double sum[5] = {0, 0, 0, 0, 0}; uint32_t cnt[5] = {0, 0, 0, 0, 0};
				for (uint32_t v = 33; v < 128; ++v) { const uint32_t b = a.map[v - 33 < 96 ? v - 33 : 95]; sum[b] += (double)(v - 33u) * st[v]; cnt[b] += st[v]; }
				uint64_t ctx_p = 0;
				for (uint32_t b = 0; b < a.n_bins; ++b) {
					const double avg = cnt[b] ? __ddiv_rn(sum[b], (double)cnt[b]) : 0.0;
					emit_avg(at + 2 * b, (1ull << 30) + ((uint64_t)b << 24) + (ctx_p << 16), avg);
					ctx_p = (uint64_t)avg;
				}
			}
		}
		at += a.mode == 7 ? 2 : 2 * a.n_bins;
		if (a.mode == 7) return;
	}
	const uint32_t cbits = a.bps * a.ncs, ones = (1u << a.bps) - 1;
	for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
		const uint32_t qv = q[i] - 33u;
		if (qv >= 96) { atomicExch(a.bad, 1u); continue; }
		uint64_t ctx = 0;                                    // the last ncs coded symbols (their quantised values in the lossless mode), all ones before the read
		for (uint32_t k = a.ncs; k >= 1; --k) {
			uint32_t h = ones;
			if (i >= k) { const uint32_t pv = q[i - k] - 33u; const uint32_t ps = a.map[pv < 96 ? pv : 95]; h = a.mode == 0 ? a.quant[ps] : ps; }
			ctx = (ctx << a.bps) | h;
		}
		auto vs = [&](uint32_t j) -> uint64_t { return base_at(a.pk, rs + j); };
		uint32_t sh = cbits;
		if (averages) {                                      // bases i-2 .. i+1 as one byte (impl:222-236)
			uint64_t dna = 0;
			for (int d = -2; d <= 1; ++d) { const long long j = (long long)i + d; dna = (dna << 2) | ((j >= 0 && j < (long long)n) ? vs((uint32_t)j) : 0ull); }
			ctx += dna << sh; sh += 8;
		} else {                                             // encode_original impl:78-128, encode_*_threshold impl:312-438
			ctx += vs(i) << sh; sh += 2;
			if (i > 0) ctx += vs(i - 1) << sh;
			sh += 2;
			if (a.mode != 0 || a.level == 3) { if (i > 1) ctx += vs(i - 2) << sh; sh += 2; }
			else { if (i > 1) ctx += (uint64_t)(vs(i - 2) == vs(i - 1)) << sh; sh += 1; }
			if (i + 1 < n) ctx += vs(i + 1) << sh;
			sh += 2;
		}
		if (a.level > 1) { ctx += (uint64_t)(fl[i] == 1) << sh; ++sh; ctx += (uint64_t)(fl[i] == 2) << sh; }
		emit(at + i, XQ_SYM, ctx, a.map[qv]);
	}
}

// lossless quantisers (quality_coder.cpp:276-338 ONT, :356-420 PacBio CLR, :441-505 PacBio HiFi)
static void xq_quantize(uint32_t source, uint32_t level, uint8_t* q)
{
	std::memset(q, 0, 96);
	auto fill = [&](int a, int b, int v) { for (int i = a; i < b; ++i) q[i] = (uint8_t)v; };
	if (source == 0) {
		if (level >= 3) { static const int e[] = {0, 1, 2, 4, 7, 11, 16, 22, 29, 37, 46, 56, 67, 79, 90, 96}; for (int k = 0; k < 15; ++k) fill(e[k], e[k + 1], k); }
		else { static const int e[] = {0, 1, 2, 5, 10, 15, 20, 25, 35, 50, 70, 96}; for (int k = 0; k < 11; ++k) fill(e[k], e[k + 1], k); }
	} else {
		const int s = source == 2 ? 1 : 0;
		if (level >= 3) { static const int e[] = {1, 10, 20, 30, 39, 45, 51, 57, 63, 69, 75, 81, 87, 93}; q[0] = (uint8_t)s; for (int k = 0; k < 13; ++k) fill(e[k], e[k + 1], k + 1 + s); q[93] = s ? 0 : 14; }
		else { static const int e[] = {1, 15, 29, 41, 53, 63, 72, 80, 87, 93}; q[0] = (uint8_t)s; for (int k = 0; k < 9; ++k) fill(e[k], e[k + 1], k + 1 + s); q[93] = s ? 0 : 10; }
	}
}

clb_status s3x_qual_encode(clb_ctx* c, uint32_t mode, uint32_t source, uint32_t level, const uint32_t* thr, const uint8_t* quals, const uint64_t* offsets, int on_device,
	const uint32_t* pack_sizes, uint32_t n_packs)
{
	cudaStream_t s = c->stream3;
	const uint64_t nc = c->n_context, n = c->n_reads - nc;      // context reads carry no qualities
	if (!c->finalized) return fail(c, CLB_ERR_STATE, "clb_xqual_encode before the reads are complete (clb_count_finalize)");
	if (mode > 8 || source > 2 || level < 1 || level > 3) return fail(c, CLB_ERR_BAD_ARG, "clb_xqual_encode: bad mode / source / level");
	if (level > 1 && mode != 8 && !c->enc_done) return fail(c, CLB_ERR_STATE, "clb_xqual_encode at level > 1 needs the tuples (clb_encode) for the match / anchor flags");
	std::vector<uint64_t> pack_first;
	{ const clb_status st = x_packs(c, pack_sizes, n_packs, n, pack_first); if (st != CLB_OK) return st; }
	const uint32_t np = (uint32_t)pack_first.size() - 1;
	XQArgs a{};
	a.mode = mode; a.level = level;
	a.n_bins = (mode == 1 || mode == 4) ? 5 : (mode == 2 || mode == 5) ? 4 : (mode == 3 || mode == 6) ? 2 : 0;
	if (mode == 0) { a.bps = 4; a.ncs = 2; } else if (mode == 7) { a.bps = 8; a.ncs = 2; } else if (a.n_bins == 2) { a.bps = 2; a.ncs = 6; } else { a.bps = 3; a.ncs = 3; }      // quality_coder.cpp:59-262
	if (mode == 0) { for (int i = 0; i < 96; ++i) a.map[i] = (uint8_t)i; xq_quantize(source, level, a.quant); }
	else if (a.n_bins) {      // adjust_quality_map_symbols (quality_coder.cpp:264-283)
		if (!thr) return fail(c, CLB_ERR_BAD_ARG, "clb_xqual_encode: the binned modes need their thresholds");
		for (uint32_t b = 0; b + 2 < a.n_bins; ++b) if (thr[b] > thr[b + 1]) return fail(c, CLB_ERR_BAD_ARG, "clb_xqual_encode: thresholds must not decrease");
		std::memset(a.map, 0, 96);
		for (uint32_t bin = 1; bin + 1 < a.n_bins; ++bin) for (uint32_t i = thr[bin - 1]; i < thr[bin] && i < 96; ++i) a.map[i] = (uint8_t)bin;
		for (uint32_t i = thr[a.n_bins - 2]; i < 96; ++i) a.map[i] = (uint8_t)(a.n_bins - 1);
	}
	XFams F{};
	F.f[XQ_SYM] = XFam{mode == 0 ? 96u : a.n_bins ? a.n_bins : 2u, mode == 0 ? 1u << 20 : 1u << 18, mode == 0 ? 32u : 8u};      // quality_coder.h:35-39
	F.f[XQ_BYTE] = XFam{256, 1u << 18, 8};
	XTmp tmp{{}, s};
	auto dalloc = [&](void** p, uint64_t bytes) { cudaError_t e = dev_malloc(p, bytes ? bytes : 1, s); if (e == cudaSuccess) tmp.v.push_back(*p); return e; };
	// events per read and their place in the stream
	const uint32_t extra = (mode >= 1 && mode <= 3) ? 2 * a.n_bins : mode == 7 ? 2 : 0;
	std::vector<uint64_t> ev_off(n + 1, 0);
	for (uint64_t i = 0; i < n; ++i) ev_off[i + 1] = ev_off[i] + (mode == 8 ? 0 : extra + (mode == 7 ? 0 : c->h_rd_len[nc + i]));
	const uint64_t n_ev = ev_off[n];
	std::vector<uint64_t> pack_ev(np + 1);
	for (uint32_t p = 0; p <= np; ++p) pack_ev[p] = ev_off[pack_first[p]];
	uint64_t* d_key = nullptr; uint16_t* d_info = nullptr;
	CLB_CUDA(c, dalloc((void**)&d_key, sizeof(uint64_t) * (n_ev + 1))); CLB_CUDA(c, dalloc((void**)&d_info, sizeof(uint16_t) * (n_ev + 1)));
	if (mode != 8 && n) {
		std::vector<uint64_t> h_off; bool resident = false;
		{ const clb_status st = resolve_quals(c, quals, offsets, on_device, s, h_off, resident); if (st != CLB_OK) return st; }
		if (resident) { quals = c->dq.p; on_device = 1; }
		const uint64_t tot = h_off[n] - h_off[0];
		uint64_t* d_qoff = nullptr; uint64_t* d_ev_off = nullptr;
		CLB_CUDA(c, dalloc((void**)&d_qoff, sizeof(uint64_t) * (n + 1))); CLB_CUDA(c, dalloc((void**)&d_ev_off, sizeof(uint64_t) * (n + 1))); CLB_CUDA(c, dalloc((void**)&a.bad, 4));
		std::vector<uint64_t> rel(n + 1);
		for (uint64_t i = 0; i <= n; ++i) rel[i] = h_off[i] - h_off[0];
		CLB_CUDA(c, cudaMemcpyAsync(d_qoff, rel.data(), sizeof(uint64_t) * (n + 1), cudaMemcpyHostToDevice, s));
		CLB_CUDA(c, cudaMemcpyAsync(d_ev_off, ev_off.data(), sizeof(uint64_t) * (n + 1), cudaMemcpyHostToDevice, s));
		CLB_CUDA(c, cudaMemsetAsync(a.bad, 0, 4, s));
		const uint8_t* d_q = nullptr;
		if (on_device) d_q = quals + h_off[0];
		else { uint8_t* b = nullptr; CLB_CUDA(c, dalloc((void**)&b, tot + 16)); CLB_CUDA(c, cudaMemcpyAsync(b, quals + h_off[0], tot, cudaMemcpyHostToDevice, s)); d_q = b; }
		uint8_t* d_flags = nullptr;
		if (level > 1) {
			CLB_CUDA(c, dalloc((void**)&d_flags, tot + 16));
			CLB_CUDA(c, cudaMemsetAsync(d_flags, 0, tot + 16, s));
			const clb_status st = s3_qual_flags(c, d_qoff, (uint32_t)n, d_flags);
			if (st != CLB_OK) return st;
		}
		a.pk = c->pk.p; a.rd_start = c->rd_start.p + nc; a.rd_len = c->rd_len.p + nc; a.quals = d_q; a.qoff = d_qoff; a.flags = d_flags; a.ev_off = d_ev_off; a.key = d_key; a.info = d_info;
		CLB_TIMED3(c, K_QUAL, (k_xq_events<<<(uint32_t)n, 128, 0, s>>>(a))); CLB_LAUNCH_CHECK(c, "k_xq_events");
		uint32_t bad = 0;
		CLB_CUDA(c, cudaMemcpyAsync(&bad, a.bad, 4, cudaMemcpyDeviceToHost, s));
		CLB_CUDA(c, cudaStreamSynchronize(s));      // also: the host vectors above are not needed by the device any more
		if (bad) return fail(c, CLB_ERR_BAD_SYMBOL, "clb_xqual_encode: a quality value outside '!' .. '~' + 2");
	}
	return x_code_stream(c, s, K_QUAL, F, d_key, d_info, n_ev, pack_ev, c->xq, c->xq_parts, c->xq_total);
}

// ================================================================================================ header stream
// id_coder.cpp:210-383 (compress_lossless) with its adaptive models (id_coder.h:50-59); the previous header is the previous header of
// the FILE (Restart at a pack boundary only clears the flag history, id_coder.cpp:80-91).
enum { XH_PLUS = 0, XH_FLAGS, XH_SAME, XH_SAMELEN, XH_LITERAL, XH_PLAIN, XH_COUNT };
struct XHArgs { HdrInput H; uint64_t n; const uint64_t* pack_first; uint32_t n_packs; uint8_t* flags; uint32_t* n_ev; const uint64_t* ev_off; uint64_t* key; uint16_t* info; uint32_t* bad; };

CLB_D uint64_t xh_pack_of(const XHArgs& a, uint64_t r)
{
	uint32_t lo = 0, hi = a.n_packs;
	while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (a.pack_first[mid] <= r) lo = mid; else hi = mid; }
	return a.pack_first[lo];
}
// the same-shape flag of every header against its predecessor in the file (header 0: against nothing -> 0)
__global__ void __launch_bounds__(128) k_xh_flags(XHArgs a)
{
	const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= a.n) return;
	const uint8_t* cur = a.H.bytes + a.H.off[r]; const uint32_t nc = (uint32_t)(a.H.off[r + 1] - a.H.off[r]);
	bool nul = false;
	for (uint32_t i = 0; i < nc; ++i) nul |= cur[i] == 0 || cur[i] >= 128;      // 0 is the terminator; the plain model has 128 symbols
	if (nul) atomicExch(a.bad, 1u);
	a.flags[r] = r > 0 && hdr_same_shape(cur, nc, a.H.bytes + a.H.off[r - 1], (uint32_t)(a.H.off[r] - a.H.off[r - 1]));
}
// contexts are identifiers: the reference's sums (flags + position) + (1 << 32) + (token << 40) [+ (1 << 60)] are re-packed into 44
// bits — (flags + position) 28 | token 15 | "against the previous header" 1 — keeping the reference's own aliasing of flags + position
template <class Sink>
CLB_D void xh_walk(const XHArgs& a, uint64_t r, Sink& sink, uint32_t* bad)
{
	const HdrInput& H = a.H;
	const uint8_t* cur = H.bytes + H.off[r]; const uint32_t nc = (uint32_t)(H.off[r + 1] - H.off[r]);
	const uint8_t* prv = r ? H.bytes + H.off[r - 1] : cur; const uint32_t np = r ? (uint32_t)(H.off[r] - H.off[r - 1]) : 0;
	const uint64_t p0 = xh_pack_of(a, r);
	uint64_t ctx_flags = 0;                                  // the flags since the pack's first header (at most 8)
	for (uint64_t k = r - p0 < 8 ? p0 : r - 8; k < r; ++k) ctx_flags = ((ctx_flags << 1) + a.flags[k]) & 0xff;
	sink.put(XH_PLUS, 0, H.plus ? (H.plus[r] != 0) : 0u);
	const uint32_t flag = a.flags[r];
	sink.put(XH_FLAGS, ctx_flags, flag);
	if (!flag) {
		if (nc >= (1u << 20)) { atomicExch(bad, 1u); return; }
		for (uint32_t j = 0; j < nc; ++j) sink.put(XH_PLAIN, j, cur[j]);
		sink.put(XH_PLAIN, nc, 0);
		return;
	}
	ctx_flags = ((ctx_flags << 1) + 1) & 0xff;             // the literal contexts see the flag just coded (id_coder.cpp:221)
	uint32_t i = 0, j = 0;
	for (uint32_t t = 0;; ++t) {
		uint32_t ie = i, je = j;
		while (ie < nc && hdr_is_literal(cur[ie])) ++ie;
		while (je < np && hdr_is_literal(prv[je])) ++je;
		const uint32_t lc = ie - i, lp = je - j;
		if (t >= (1u << 15) || lc >= (1u << 20) - 1) { atomicExch(bad, 1u); return; }
		bool same = lc == lp;
		if (same) for (uint32_t k = 0; k < lc; ++k) if (cur[i + k] != prv[j + k]) { same = false; break; }
		sink.put(XH_SAME, t, same);
		if (!same) {
			sink.put(XH_SAMELEN, t, lc == lp);
			const uint64_t base = (uint64_t)t << 28;
			if (lc == lp) for (uint32_t k = 0; k < lc; ++k) { const uint32_t ch = cur[i + k]; sink.put(XH_LITERAL, base | (ctx_flags + k) | (1ull << 43), ch == prv[j + k] ? 0u : ch); }
			else {
				for (uint32_t k = 0; k < lc; ++k) sink.put(XH_LITERAL, base | (ctx_flags + k), cur[i + k]);
				sink.put(XH_LITERAL, base | (ctx_flags + lc), 0);
			}
		}
		if (ie == nc) break;
		i = ie + 1; j = je + 1;
	}
}
template <bool WRITE>
__global__ void __launch_bounds__(128) k_xh_events(XHArgs a)
{
	const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= a.n) return;
	if (WRITE) { XWriteSink s{a.key, a.info, a.ev_off[r], a.bad}; xh_walk(a, r, s, a.bad); }
	else { XCountSink s; xh_walk(a, r, s, a.bad); a.n_ev[r] = s.n; }
}

clb_status s3x_hdr_encode(clb_ctx* c, const uint8_t* bytes, const uint64_t* offsets, const uint8_t* plus_id, uint64_t n, int on_device,
	const uint32_t* pack_sizes, uint32_t n_packs)
{
	cudaStream_t s = c->stream3;
	if (n >= 0xffffffffull) return fail(c, CLB_ERR_BAD_ARG, "clb_xhdr_encode: too many headers");
	std::vector<uint64_t> pack_first{0};
	if (pack_sizes) {
		uint64_t at = 0;
		for (uint32_t i = 0; i < n_packs; ++i) { at += pack_sizes[i]; pack_first.push_back(at); }
		if (at != n) return fail(c, CLB_ERR_BAD_ARG, "pack_sizes do not sum to the number of headers");
	} else if (n) pack_first.push_back(n);
	const uint32_t np = (uint32_t)pack_first.size() - 1;
	XTmp tmp{{}, s};
	auto dalloc = [&](void** p, uint64_t nbytes) { cudaError_t e = dev_malloc(p, nbytes ? nbytes : 1, s); if (e == cudaSuccess) tmp.v.push_back(*p); return e; };
	XHArgs a{};
	a.n = n; a.n_packs = np;
	if (on_device) a.H = HdrInput{bytes, offsets, plus_id};
	else {
		const uint64_t total = n ? offsets[n] : 0;
		if (n && offsets[0] != 0) return fail(c, CLB_ERR_BAD_ARG, "clb_xhdr_encode: offsets[0] must be 0");
		uint8_t* d_b = nullptr; uint64_t* d_o = nullptr; uint8_t* d_p = nullptr;
		CLB_CUDA(c, dalloc((void**)&d_b, total)); CLB_CUDA(c, dalloc((void**)&d_o, sizeof(uint64_t) * (n + 1)));
		if (n) { CLB_CUDA(c, cudaMemcpyAsync(d_b, bytes, total, cudaMemcpyHostToDevice, s)); CLB_CUDA(c, cudaMemcpyAsync(d_o, offsets, sizeof(uint64_t) * (n + 1), cudaMemcpyHostToDevice, s)); }
		if (plus_id && n) { CLB_CUDA(c, dalloc((void**)&d_p, n)); CLB_CUDA(c, cudaMemcpyAsync(d_p, plus_id, n, cudaMemcpyHostToDevice, s)); }
		a.H = HdrInput{d_b, d_o, d_p};
	}
	uint64_t* d_pack_first = nullptr; uint64_t* d_ev_off = nullptr;
	CLB_CUDA(c, dalloc((void**)&d_pack_first, sizeof(uint64_t) * (np + 1))); CLB_CUDA(c, dalloc((void**)&a.flags, n + 1));
	CLB_CUDA(c, dalloc((void**)&a.n_ev, sizeof(uint32_t) * (n + 1))); CLB_CUDA(c, dalloc((void**)&d_ev_off, sizeof(uint64_t) * (n + 1))); CLB_CUDA(c, dalloc((void**)&a.bad, 4));
	CLB_CUDA(c, cudaMemsetAsync(a.bad, 0, 4, s));
	CLB_CUDA(c, cudaMemcpyAsync(d_pack_first, pack_first.data(), sizeof(uint64_t) * (np + 1), cudaMemcpyHostToDevice, s));
	a.pack_first = d_pack_first;
	const uint32_t nblk = (uint32_t)((n + 127) / 128);
	std::vector<uint32_t> n_ev_h(n);
	if (n) {
		CLB_TIMED3(c, K_HDR, (k_xh_flags<<<nblk, 128, 0, s>>>(a))); CLB_LAUNCH_CHECK(c, "k_xh_flags");
		CLB_TIMED3(c, K_HDR, (k_xh_events<false><<<nblk, 128, 0, s>>>(a))); CLB_LAUNCH_CHECK(c, "k_xh_events<count>");
		CLB_CUDA(c, cudaMemcpyAsync(n_ev_h.data(), a.n_ev, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, s));
	}
	uint32_t bad = 0;
	CLB_CUDA(c, cudaMemcpyAsync(&bad, a.bad, 4, cudaMemcpyDeviceToHost, s));
	CLB_CUDA(c, cudaStreamSynchronize(s));
	if (bad) return fail(c, CLB_ERR_BAD_SYMBOL, "clb_xhdr_encode: a header holds a NUL / non-ASCII byte or is longer than the compat stream's contexts allow");
	std::vector<uint64_t> ev_off(n + 1, 0);      // the scan runs on the host: stream3 may run beside stage 2, whose scan scratch lives on the other stream
	for (uint64_t i = 0; i < n; ++i) ev_off[i + 1] = ev_off[i] + n_ev_h[i];
	const uint64_t n_ev = ev_off[n];
	CLB_CUDA(c, cudaMemcpyAsync(d_ev_off, ev_off.data(), sizeof(uint64_t) * (n + 1), cudaMemcpyHostToDevice, s));
	CLB_CUDA(c, dalloc((void**)&a.key, sizeof(uint64_t) * (n_ev + 1))); CLB_CUDA(c, dalloc((void**)&a.info, sizeof(uint16_t) * (n_ev + 1)));
	a.ev_off = d_ev_off;
	if (n) { CLB_TIMED3(c, K_HDR, (k_xh_events<true><<<nblk, 128, 0, s>>>(a))); CLB_LAUNCH_CHECK(c, "k_xh_events<write>"); }
	CLB_CUDA(c, cudaStreamSynchronize(s));
	XFams F{};
	const XFam f0[XH_COUNT] = {{2, 1u << 15, 1}, {2, 1u << 15, 1}, {2, 1u << 15, 1}, {2, 1u << 15, 1}, {256, 1u << 20, 64}, {128, 1u << 19, 32}};      // id_coder.h:50-59
	for (uint32_t f = 0; f < XH_COUNT; ++f) F.f[f] = f0[f];
	std::vector<uint64_t> pack_ev(np + 1);
	for (uint32_t p = 0; p <= np; ++p) pack_ev[p] = ev_off[pack_first[p]];
	c->xh_packs.assign(np, 0);
	for (uint32_t p = 0; p < np; ++p) c->xh_packs[p] = pack_first[p + 1] - pack_first[p];
	return x_code_stream(c, s, K_HDR, F, a.key, a.info, n_ev, pack_ev, c->xh, c->xh_parts, c->xh_total);
}

} // namespace clb
