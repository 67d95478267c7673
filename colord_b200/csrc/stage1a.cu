// stage1a.cu — ingest (ASCII -> 2-bit resident read store), canonical k-mer scan with the
// murmur64 % f filter, the open-addressed count table, thresholding and the filtered-k-mer set.
//
// What it replaces in the reference (semantics only; the disk-bin / super-k-mer / radix-sort machinery
// of filtering-KMC is not reproduced — one hash-count table in HBM takes its place):
//   to_read_t                      src/colord/in_reads.cpp:24-42
//   CKmerWalker::NextKmer          src/colord/in_reads.h:59-73
//   checkModuloHash / hash_mm      src/filtering-KMC/hash_filter.h:8-78, src/colord/filter_kmers.cpp:24-32
//   count thresholds + statistics  src/filtering-KMC/kb_sorter.h:1011-1065, kmc.h:1471-1479
//   CKmerFilter / CCompactedKmers  src/colord/kmer_filter.h:30-199, filter_kmers.cpp:45-83
#include "ctx.h"
#include <ctime>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>

namespace clb {

// ------------------------------------------------------------------------------------------------
// Ingest: 32 ASCII bases per thread -> one packed word + one N-mask word.
// HBM traffic per base: 1 B read + 0.25 B + 0.125 B written.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pack(const uint8_t* __restrict__ bases, uint64_t n_bases, uint64_t n_words,
	uint64_t* __restrict__ pk, uint32_t* __restrict__ nmask, int aligned16,
	unsigned long long* __restrict__ scal)
{
	uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	bool bad = false;
	for (; w < n_words; w += stride) {
		const uint64_t p0 = w << 5;
		uint32_t raw[8];
		if (p0 + 32 <= n_bases && aligned16) {
			const uint4* src = reinterpret_cast<const uint4*>(bases + p0);
			uint4 a = __ldg(src), b = __ldg(src + 1);
			raw[0] = a.x; raw[1] = a.y; raw[2] = a.z; raw[3] = a.w;
			raw[4] = b.x; raw[5] = b.y; raw[6] = b.z; raw[7] = b.w;
		} else {
#pragma unroll
			for (int i = 0; i < 8; ++i) {
				uint32_t v = 0;
#pragma unroll
				for (int j = 0; j < 4; ++j) {
					uint64_t p = p0 + 4 * i + j;
					uint32_t c = p < n_bases ? bases[p] : 0u;      // 0 = padding, becomes an N position
					v |= c << (8 * j);
				}
				raw[i] = v;
			}
		}
		uint64_t word = 0; uint32_t nm = 0;
#pragma unroll
		for (int i = 0; i < 8; ++i) {
#pragma unroll
			for (int j = 0; j < 4; ++j) {
				const uint32_t c = (raw[i] >> (8 * j)) & 0xFF;
				const int idx = 4 * i + j;
				// 'A' 0x41 'C' 0x43 'G' 0x47 'T' 0x54: (c>>1)&3 = 0,1,3,2 ; x ^ (x>>1) = 0,1,2,3
				uint32_t x = (c >> 1) & 3; x ^= x >> 1;
				const bool acgt = (c == 'A') | (c == 'C') | (c == 'G') | (c == 'T');
				const bool pad = (p0 + idx >= n_bases);
				if (!acgt) { x = 0; nm |= 1u << idx; if (c != 'N' && !pad) bad = true; }
				word |= (uint64_t)x << (62 - 2 * idx);
			}
		}
		pk[w] = word; nmask[w] = nm;
	}
	if (bad) atomicOr(&scal[SC_BAD_SYMBOL], 1ULL);
}

__global__ void k_mark_starts(const uint64_t* __restrict__ offsets, uint32_t n_reads, uint64_t pos0, uint64_t read0,
	uint32_t* __restrict__ smask, uint64_t* __restrict__ rd_start, uint32_t* __restrict__ rd_len)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_reads) return;
	const uint64_t s = offsets[i] - offsets[0], e = offsets[i + 1] - offsets[0];
	const uint64_t p = pos0 + s;
	rd_start[read0 + i] = p;
	rd_len[read0 + i] = (uint32_t)(e - s);
	if (e > s) atomicOr(&smask[p >> 5], 1u << (p & 31));
}

// ------------------------------------------------------------------------------------------------
// Count table.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t MAX_PROBE = 1u << 14;

CLB_D bool tab_add(CountSlot* __restrict__ tab, uint32_t log2cap, uint64_t kmer, uint64_t h, uint32_t add, uint32_t& n_new)
{
	const uint64_t mask = (1ULL << log2cap) - 1;
	uint64_t s = slot_of(h, log2cap);
	for (uint32_t probe = 0; probe < MAX_PROBE; ++probe) {
		unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long*>(&tab[s].key), EMPTY64, kmer);
		if (old == EMPTY64 || old == kmer) {
			atomicAdd(&tab[s].cnt, add);
			n_new += (old == EMPTY64);
			return true;
		}
		s = (s + 1) & mask;
	}
	return false;
}

__global__ void k_tab_clear(CountSlot* tab, uint64_t cap)
{
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (; i < cap; i += stride) { uint4 v; v.x = v.y = 0xFFFFFFFFu; v.z = 0; v.w = 0; reinterpret_cast<uint4*>(tab)[i] = v; }
}

// Scan kernel.  One thread owns the 32 k-mers ENDING in one packed word (halo = previous word); the
// canonical k-mer is rolled base by base (in_reads.h:59-73), hashed, and the ~1/f survivors are pushed to a
// per-warp shared-memory queue with a ballot; whenever 32 are queued the whole warp inserts them in
// lockstep (32 independent atomics in flight, duplicates inside the batch merged by match.any).
// Algorithmic HBM bytes per base: 0.25 (packed) + 0.25 (two masks) + 16/f (one 8 B key + count RMW per
// passing k-mer).
constexpr int COUNT_THREADS = 256;

template <bool COUNT_ONLY>
__global__ void __launch_bounds__(COUNT_THREADS) k_count(const uint64_t* __restrict__ pk, const uint32_t* __restrict__ nmask,
	const uint32_t* __restrict__ smask, uint64_t w_first, uint64_t w_begin, uint64_t w_end, uint32_t k, ModTest mt,
	CountSlot* __restrict__ tab, uint32_t log2cap, unsigned long long* __restrict__ scal)
{
	__shared__ uint64_t q[COUNT_THREADS / 32][64];
	const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint32_t lt = (1u << lane) - 1;
	uint64_t* myq = q[wid];
	uint32_t qn = 0, n_new = 0, n_pass = 0;
	bool ovf = false;

	const uint64_t kmask = k == 32 ? ~0ULL : ((1ULL << (2 * k)) - 1);
	const uint32_t rsh = 2 * (k - 1);
	const uint64_t win_n = ((k == 64 ? 0 : (1ULL << k)) - 1) << (33 - k);   // bits of the k positions ending at bit 32
	const uint64_t win_s = win_n & (win_n - 1);                               // same without the first position

	const uint64_t n_words = w_end - w_begin;
	const uint64_t warp_stride = (uint64_t)gridDim.x * COUNT_THREADS;
	for (uint64_t base = (uint64_t)blockIdx.x * COUNT_THREADS + (wid << 5); base < n_words; base += warp_stride) {
		const uint64_t i = base + lane;
		const bool live = i < n_words;
		const uint64_t w = w_begin + (live ? i : 0);
		uint64_t cur = 0, prev = 0, N64 = ~0ULL, S64 = 0;
		if (live) {
			cur = pk[w];
			const uint32_t nm = nmask[w], sm = smask[w];
			uint32_t pnm = 0xFFFFFFFFu, psm = 0;
			if (w > w_first) { prev = pk[w - 1]; pnm = nmask[w - 1]; psm = smask[w - 1]; }   // no halo across appends
			N64 = ((uint64_t)nm << 32) | pnm;
			S64 = ((uint64_t)sm << 32) | psm;
		}
		uint64_t fw = prev & (kmask >> 2);
		uint64_t rc = revcomp(fw, k);
#pragma unroll 4
		for (int j = 0; j < 32; ++j) {
			const uint64_t b = (cur >> (62 - 2 * j)) & 3;
			fw = ((fw << 2) | b) & kmask;
			rc = (rc >> 2) | ((3 - b) << rsh);
			const uint64_t can = fw < rc ? fw : rc;
			const bool ok = ((N64 & (win_n << j)) | (S64 & (win_s << j))) == 0;
			const bool pass = ok && divisible(murmur64(can), mt);
			const uint32_t bal = __ballot_sync(0xffffffffu, pass);
			if (bal) {
				if (COUNT_ONLY) { n_pass += pass; continue; }
				if (pass) { myq[qn + __popc(bal & lt)] = can; ++n_pass; }
				qn += __popc(bal);
				if (qn >= 32) {
					__syncwarp();
					qn -= 32;
					const uint64_t x = myq[qn + lane];
					const uint32_t same = __match_any_sync(0xffffffffu, x);
					if ((uint32_t)(__ffs(same) - 1) == lane)
						ovf |= !tab_add(tab, log2cap, x, murmur64(x), __popc(same), n_new);
					__syncwarp();
				}
			}
		}
	}
	if (!COUNT_ONLY) {
		__syncwarp();
		const uint32_t act = __ballot_sync(0xffffffffu, lane < qn);
		if (lane < qn) {
			const uint64_t x = myq[lane];
			const uint32_t same = __match_any_sync(act, x);
			if ((uint32_t)(__ffs(same) - 1) == lane)
				ovf |= !tab_add(tab, log2cap, x, murmur64(x), __popc(same), n_new);
		}
	}
	// block totals -> one atomic each
	__shared__ uint32_t red[2][COUNT_THREADS / 32];
#pragma unroll
	for (int d = 16; d; d >>= 1) { n_new += __shfl_xor_sync(0xffffffffu, n_new, d); n_pass += __shfl_xor_sync(0xffffffffu, n_pass, d); }
	if (lane == 0) { red[0][wid] = n_new; red[1][wid] = n_pass; }
	__syncthreads();
	if (threadIdx.x == 0) {
		uint32_t a = 0, b = 0;
		for (int i = 0; i < COUNT_THREADS / 32; ++i) { a += red[0][i]; b += red[1][i]; }
		if (a) atomicAdd(&scal[SC_TAB_USED], (unsigned long long)a);
		if (b) atomicAdd(&scal[SC_TOT_KMERS], (unsigned long long)b);
	}
	if (ovf) atomicOr(&scal[SC_OVERFLOW], 1ULL);
}

// Re-insert (k-mer, count) pairs: table growth and the multi-GPU merge.
__global__ void k_tab_reinsert(const CountSlot* __restrict__ src, uint64_t n_src, CountSlot* __restrict__ dst, uint32_t log2cap,
	unsigned long long* __restrict__ scal)
{
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	uint32_t n_new = 0; bool ovf = false;
	for (; i < n_src; i += stride) {
		const CountSlot s = src[i];
		if (s.key != EMPTY64) ovf |= !tab_add(dst, log2cap, s.key, murmur64(s.key), s.cnt, n_new);
	}
	if (n_new) atomicAdd(&scal[SC_TAB_USED], (unsigned long long)n_new);
	if (ovf) atomicOr(&scal[SC_OVERFLOW], 1ULL);
}

__global__ void k_tab_merge(const uint64_t* __restrict__ kmers, const uint32_t* __restrict__ counts, uint64_t n,
	CountSlot* __restrict__ dst, uint32_t log2cap, unsigned long long* __restrict__ scal)
{
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	uint32_t n_new = 0; bool ovf = false;
	for (; i < n; i += stride) ovf |= !tab_add(dst, log2cap, kmers[i], murmur64(kmers[i]), counts[i], n_new);
	if (n_new) atomicAdd(&scal[SC_TAB_USED], (unsigned long long)n_new);
	if (ovf) atomicOr(&scal[SC_OVERFLOW], 1ULL);
}

// Owner partition of a k-mer for the multi-GPU exchange (any rank computes the same value).
CLB_HD uint32_t owner_of(uint64_t kmer, uint32_t n_parts)
{
	return n_parts <= 1 ? 0u : (uint32_t)(((murmur64(kmer) >> 32) * (uint64_t)n_parts) >> 32);
}

// Export the occupied slots of partition `part` (compaction with a warp-aggregated cursor).
template <bool WRITE>
__global__ void k_tab_export(const CountSlot* __restrict__ tab, uint64_t cap, uint32_t part, uint32_t n_parts,
	uint64_t* __restrict__ kmers, uint32_t* __restrict__ counts, uint64_t out_cap, unsigned long long* __restrict__ cursor)
{
	const uint32_t lane = threadIdx.x & 31;
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	const uint64_t n_iter = (cap + stride - 1) / stride;
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	for (uint64_t it = 0; it < n_iter; ++it, i += stride) {
		CountSlot s; s.key = EMPTY64; s.cnt = 0;
		if (i < cap) s = tab[i];
		const bool take = s.key != EMPTY64 && owner_of(s.key, n_parts) == part;
		const uint32_t bal = __ballot_sync(0xffffffffu, take);
		if (!bal) continue;
		unsigned long long base = 0;
		if (lane == 0) base = atomicAdd(cursor, (unsigned long long)__popc(bal));
		base = __shfl_sync(0xffffffffu, base, 0);
		if (WRITE && take) {
			const uint64_t o = base + __popc(bal & ((1u << lane) - 1));
			if (o < out_cap) { kmers[o] = s.key; counts[o] = s.cnt; }
		}
	}
}

// All partitions in ONE pass over the table (the multi-GPU exchange asked for 2 N scans of it before: N sizes, N exports, each
// paying one global atomic per warp on a single cursor).  A CTA takes tiles of 256 x EXPORT_PER_THREAD slots: the occupied slots are
// ranked per partition in shared memory (one atomic per warp and partition present: match.any), the tile's share of every
// partition is reserved with one global atomic per partition, then the entries are written at first[partition] + their rank.
// WRITE == false only counts (cursor[p] = entries of partition p).
constexpr int EXPORT_PER_THREAD = 8;
constexpr uint32_t EXPORT_MAX_PARTS = 64;
template <bool WRITE>
__global__ void __launch_bounds__(256) k_tab_export_all(const CountSlot* __restrict__ tab, uint64_t cap, uint32_t n_parts,
	uint64_t* __restrict__ kmers, uint32_t* __restrict__ counts, uint64_t out_cap, unsigned long long* __restrict__ cursor, const unsigned long long* __restrict__ first)
{
	__shared__ uint32_t s_cnt[EXPORT_MAX_PARTS];
	__shared__ unsigned long long s_base[EXPORT_MAX_PARTS];
	const uint32_t lane = threadIdx.x & 31;
	const uint64_t tile_slots = 256ull * EXPORT_PER_THREAD, n_tiles = (cap + tile_slots - 1) / tile_slots;
	for (uint64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
		if (threadIdx.x < n_parts) s_cnt[threadIdx.x] = 0;
		__syncthreads();
		CountSlot sl[EXPORT_PER_THREAD]; uint32_t own[EXPORT_PER_THREAD], rank[EXPORT_PER_THREAD];
#pragma unroll
		for (int k = 0; k < EXPORT_PER_THREAD; ++k) {
			const uint64_t i = t * tile_slots + (uint64_t)k * 256 + threadIdx.x;
			sl[k].key = EMPTY64; sl[k].cnt = 0;
			if (i < cap) sl[k] = tab[i];
			const bool take = sl[k].key != EMPTY64;
			own[k] = take ? owner_of(sl[k].key, n_parts) : 0xFFFFFFFFu;
			const uint32_t peers = __match_any_sync(0xffffffffu, own[k]);      // lanes of the warp with the same partition (or none)
			uint32_t base = 0;
			if (take && lane == (uint32_t)__ffs((int)peers) - 1) base = atomicAdd(&s_cnt[own[k]], (uint32_t)__popc(peers));
			base = __shfl_sync(0xffffffffu, base, __ffs((int)peers) - 1);
			rank[k] = base + __popc(peers & ((1u << lane) - 1));
		}
		__syncthreads();
		if (threadIdx.x < n_parts) s_base[threadIdx.x] = s_cnt[threadIdx.x] ? atomicAdd(&cursor[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]) : 0ull;
		__syncthreads();
		if (WRITE) {
#pragma unroll
			for (int k = 0; k < EXPORT_PER_THREAD; ++k) if (own[k] != 0xFFFFFFFFu) {
				const uint64_t o = first[own[k]] + s_base[own[k]] + rank[k];
				if (o < out_cap) { kmers[o] = sl[k].key; counts[o] = sl[k].cnt; }
			}
		}
		__syncthreads();
	}
}

// Thresholding statistics (kb_sorter.h:1011-1065): a k-mer survives iff min_count <= count <= 1e9; its
// stored count saturates at max_count.
__global__ void __launch_bounds__(256) k_tab_stats(const CountSlot* __restrict__ tab, uint64_t cap, uint32_t min_count, uint32_t max_count,
	unsigned long long* __restrict__ scal)
{
	unsigned long long v[5] = {0, 0, 0, 0, 0};   // tot, unique, surv, tot_filtered, sum_true
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (; i < cap; i += stride) {
		const CountSlot s = tab[i];
		if (s.key == EMPTY64) continue;
		v[0] += s.cnt; v[1] += 1;
		if (s.cnt >= min_count && s.cnt <= 1000000000u) { v[2] += 1; v[3] += min(s.cnt, max_count); v[4] += s.cnt; }
	}
	__shared__ unsigned long long red[5][8];
	const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
	for (int q = 0; q < 5; ++q) {
#pragma unroll
		for (int d = 16; d; d >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], d);
		if (lane == 0) red[q][wid] = v[q];
	}
	__syncthreads();
	if (threadIdx.x < 5) {
		unsigned long long s = 0;
		for (int w = 0; w < 8; ++w) s += red[threadIdx.x][w];
		const int dst[5] = {SC_TOT_KMERS, SC_N_UNIQUE, SC_N_SURV, SC_TOT_FILTERED, SC_SUM_TRUE};
		if (s) atomicAdd(&scal[dst[threadIdx.x]], s);
	}
}

CLB_D void sv_insert(uint64_t* __restrict__ keys, uint32_t* __restrict__ ids, uint32_t log2cap, uint64_t kmer, uint32_t id)
{
	const uint64_t mask = (1ULL << log2cap) - 1;
	uint64_t s = slot_of(murmur64(kmer), log2cap);
	for (;;) {
		unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long*>(&keys[s]), EMPTY64, kmer);
		if (old == EMPTY64) { ids[s] = id; return; }
		if (old == kmer) return;                 // cannot happen for distinct keys; harmless
		s = (s + 1) & mask;
	}
}

// Survivors -> dense arrays + the filtered set (open addressing, 64-bit key -> dense id).
__global__ void k_build_survivors(const CountSlot* __restrict__ tab, uint64_t cap, uint32_t min_count, uint32_t max_count,
	uint64_t* __restrict__ sv_kmer, uint32_t* __restrict__ sv_count, uint64_t* __restrict__ sv_keys, uint32_t* __restrict__ sv_ids,
	uint32_t sv_log2, unsigned long long* __restrict__ cursor)
{
	const uint32_t lane = threadIdx.x & 31;
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	const uint64_t n_iter = (cap + stride - 1) / stride;
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	for (uint64_t it = 0; it < n_iter; ++it, i += stride) {
		CountSlot s; s.key = EMPTY64; s.cnt = 0;
		if (i < cap) s = tab[i];
		const bool take = s.key != EMPTY64 && s.cnt >= min_count && s.cnt <= 1000000000u;
		const uint32_t bal = __ballot_sync(0xffffffffu, take);
		if (!bal) continue;
		unsigned long long base = 0;
		if (lane == 0) base = atomicAdd(cursor, (unsigned long long)__popc(bal));
		base = __shfl_sync(0xffffffffu, base, 0);
		if (take) {
			const uint32_t id = (uint32_t)(base + __popc(bal & ((1u << lane) - 1)));
			sv_kmer[id] = s.key; sv_count[id] = min(s.cnt, max_count);
			sv_insert(sv_keys, sv_ids, sv_log2, s.key, id);
		}
	}
}

// Import a listed filtered set (multi-GPU: survivors gathered from the owner ranks).
__global__ void k_sv_import(const uint64_t* __restrict__ sv_kmer, uint64_t n, uint64_t* __restrict__ sv_keys, uint32_t* __restrict__ sv_ids, uint32_t sv_log2)
{
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (; i < n; i += stride) sv_insert(sv_keys, sv_ids, sv_log2, sv_kmer[i], (uint32_t)i);
}

__global__ void k_filter_check(const uint64_t* __restrict__ kmers, uint64_t n, ModTest mt, const uint64_t* __restrict__ sv_keys,
	uint32_t sv_log2, uint8_t* __restrict__ possible, uint8_t* __restrict__ present)
{
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint64_t x = kmers[i], h = murmur64(x);
	possible[i] = divisible(h, mt);
	const uint64_t mask = (1ULL << sv_log2) - 1;
	uint64_t s = slot_of(h, sv_log2);
	uint8_t f = 0;
	for (;;) { const uint64_t kx = sv_keys[s]; if (kx == x) { f = 1; break; } if (kx == EMPTY64) break; s = (s + 1) & mask; }
	present[i] = f;
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
static inline uint32_t grid_for(uint64_t n_items, uint32_t threads, int n_sm, uint32_t per_sm)
{
	uint64_t g = (n_items + threads - 1) / threads;
	uint64_t cap = (uint64_t)n_sm * per_sm;
	if (g > cap) g = cap;             // persistent-style: a multiple of the SM count, grid-stride inside
	if (g == 0) g = 1;
	return (uint32_t)g;
}

static void prof_resolve_list(clb_ctx* c, std::vector<ProfRec>& open, cudaStream_t st);
void prof_begin(clb_ctx* c, int kid)
{
	if (!c->prof_on) return;
	ProfRec r; r.kid = kid;
	cudaEventCreate(&r.a); cudaEventCreate(&r.b);
	cudaEventRecord(r.a, c->stream);
	c->prof_open.push_back(r);
}
void prof_end(clb_ctx* c)
{
	if (!c->prof_on || c->prof_open.empty()) return;
	cudaEventRecord(c->prof_open.back().b, c->stream);
	if (c->prof_open.size() > 4096) prof_resolve_list(c, c->prof_open, c->stream);
}
static void prof_resolve_list(clb_ctx* c, std::vector<ProfRec>& open, cudaStream_t st)
{
	if (open.empty()) return;
	cudaStreamSynchronize(st);
	for (auto& r : open) {
		float ms = 0;
		if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) { c->prof_ms[r.kid] += ms; c->prof_n[r.kid] += 1; }
		cudaEventDestroy(r.a); cudaEventDestroy(r.b);
	}
	open.clear();
}
void prof_resolve(clb_ctx* c) { prof_resolve_list(c, c->prof_open, c->stream); prof_resolve_list(c, c->prof_open3, c->stream3); }
void prof_begin3(clb_ctx* c, int kid)
{
	if (!c->prof_on) return;
	ProfRec r; r.kid = kid;
	cudaEventCreate(&r.a); cudaEventCreate(&r.b);
	cudaEventRecord(r.a, c->stream3);
	c->prof_open3.push_back(r);
}
void prof_end3(clb_ctx* c)
{
	if (!c->prof_on || c->prof_open3.empty()) return;
	cudaEventRecord(c->prof_open3.back().b, c->stream3);
	if (c->prof_open3.size() > 4096) prof_resolve_list(c, c->prof_open3, c->stream3);
}

static clb_status read_scalars(clb_ctx* c, unsigned long long* out)
{
	CLB_CUDA(c, cudaMemcpyAsync(out, c->d_scal, sizeof(unsigned long long) * SC_COUNT, cudaMemcpyDeviceToHost, c->stream));
	CLB_CUDA(c, cudaStreamSynchronize(c->stream));
	return CLB_OK;
}

static clb_status tab_alloc(clb_ctx* c, uint32_t log2cap, CountSlot** out)
{
	CountSlot* t = nullptr;
	const uint64_t cap = 1ULL << log2cap;
	CLB_CUDA(c, dev_malloc((void**)&t, cap * sizeof(CountSlot), c->stream));
	CLB_TIMED(c, K_TAB_MISC, (k_tab_clear<<<grid_for(cap, 256, c->n_sm, 16), 256, 0, c->stream>>>(t, cap)));
	CLB_LAUNCH_CHECK(c, "k_tab_clear");
	*out = t;
	return CLB_OK;
}

static uint32_t log2_for(uint64_t n_keys)
{
	// load factor <= 0.6
	uint64_t need = n_keys + n_keys * 2 / 3 + 1024;
	uint32_t l = 10;
	while ((1ULL << l) < need) ++l;
	return l;
}

clb_status s1a_init(clb_ctx* c)
{
	CLB_CUDA(c, dev_malloc((void**)&c->d_scal, sizeof(unsigned long long) * SC_COUNT, c->stream));
	CLB_CUDA(c, cudaMemsetAsync(c->d_scal, 0, sizeof(unsigned long long) * SC_COUNT, c->stream));
	// distinct keys <= passing occurrences ~ bases / f (the table grows by re-insertion if the hint was low)
	const uint64_t exp_keys = c->prm.expected_bases / c->prm.modulo + c->prm.expected_bases / (32 * (uint64_t)c->prm.modulo);
	c->tab_log2 = log2_for(exp_keys);
	return tab_alloc(c, c->tab_log2, &c->tab);
}


static clb_status tab_grow(clb_ctx* c, uint32_t new_log2)
{
	CountSlot* nt = nullptr;
	clb_status st = tab_alloc(c, new_log2, &nt);
	if (st != CLB_OK) return st;
	CLB_CUDA(c, cudaMemsetAsync(&c->d_scal[SC_TAB_USED], 0, sizeof(unsigned long long), c->stream));
	const uint64_t cap = 1ULL << c->tab_log2;
	CLB_TIMED(c, K_TAB_MISC, (k_tab_reinsert<<<grid_for(cap, 256, c->n_sm, 16), 256, 0, c->stream>>>(c->tab, cap, nt, new_log2, c->d_scal)));
	CLB_LAUNCH_CHECK(c, "k_tab_reinsert");
	CLB_CUDA(c, cudaStreamSynchronize(c->stream));
	dev_free(c->tab, c->stream);
	c->tab = nt; c->tab_log2 = new_log2;
	return CLB_OK;
}

// Make sure `incoming` more distinct keys fit under the load-factor limit.  `used_known` is refreshed from
// the device only when the pessimistic host-side bound says the table might be getting full.
static clb_status tab_ensure(clb_ctx* c, uint64_t incoming)
{
	clb_ctx::FillState& f = c->fill;
	const auto limit = [&]() { return (uint64_t)((1ULL << c->tab_log2) * 0.6); };
	if (f.used_known + f.maybe_new + incoming <= limit()) { f.maybe_new += incoming; return CLB_OK; }
	unsigned long long sc[SC_COUNT];
	clb_status st = read_scalars(c, sc);
	if (st != CLB_OK) return st;
	f.used_known = sc[SC_TAB_USED]; f.maybe_new = 0;
	if (f.used_known + incoming > limit()) {
		uint32_t nl = c->tab_log2 + 1;
		while ((uint64_t)((1ULL << nl) * 0.6) < f.used_known + 2 * incoming) ++nl;
		st = tab_grow(c, nl);
		if (st != CLB_OK) return st;
	}
	f.maybe_new += incoming;
	return CLB_OK;
}

// One append = one read pack.  The bases are processed in chunks of 64 Mi positions: host input is staged
// through two device buffers on a copy stream so the H2D of chunk i+1 overlaps k_pack/k_count of chunk i.
constexpr uint64_t APPEND_CHUNK = 1ULL << 26;      // positions; multiple of 128

// context = true (clb_append_context_reads): the reads are packed into the store but not counted, and they take the read
// ids 0 .. n_reads-1 in front of the reads appended so far (multi-GPU: reference reads of earlier shards)
clb_status s1a_append(clb_ctx* c, const uint8_t* bases, const uint64_t* offsets, uint32_t n_reads, int on_device, bool context, bool count_only)
{
	if (!context && c->finalized) return fail(c, CLB_ERR_STATE, "clb_append_reads after clb_count_finalize");
	if (context && (!c->finalized || c->graph_done || c->n_context)) return fail(c, CLB_ERR_STATE, "clb_append_context_reads: once, after clb_count_finalize and before clb_graph_build");
	if (n_reads == 0) return CLB_OK;
	cudaStream_t s = c->stream;
	struct ATrace {      // CLB_S2_TRACE=1: wall time of the append (synchronising; debugging aid)
		bool on; cudaStream_t s; timespec t0; uint64_t bytes = 0; int dev;
		ATrace(cudaStream_t st, int d) : on(std::getenv("CLB_S2_TRACE") != nullptr), s(st), dev(d) { clock_gettime(CLOCK_MONOTONIC, &t0); }
		~ATrace() { if (!on) return; cudaStreamSynchronize(s); timespec t1; clock_gettime(CLOCK_MONOTONIC, &t1);
			const double ms = (t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6;
			fprintf(stderr, "[s1] append (%s)               %9.3f ms  %.2f GB/s\n", dev ? "device input" : "host input", ms, bytes / ms / 1e6); }
	} atrace(s, on_device);
	// offsets are needed on the host too (read bookkeeping is tiny: 12 B per read)
	std::vector<uint64_t> h_off(n_reads + 1);
	const uint64_t* d_off = nullptr;
	if (on_device) {
		CLB_CUDA(c, cudaMemcpyAsync(h_off.data(), offsets, sizeof(uint64_t) * (n_reads + 1), cudaMemcpyDeviceToHost, s));
		CLB_CUDA(c, cudaStreamSynchronize(s));
		d_off = offsets;
	} else {
		std::memcpy(h_off.data(), offsets, sizeof(uint64_t) * (n_reads + 1));
	}
	for (uint32_t i = 0; i < n_reads; ++i)
		if (h_off[i + 1] < h_off[i] || h_off[i + 1] - h_off[i] > 0xFFFFFFFFull) return fail(c, CLB_ERR_BAD_ARG, "offsets must be non-decreasing, reads < 4 Gbases");
	const uint64_t nb = h_off[n_reads] - h_off[0];
	atrace.bytes = nb;
	const uint8_t* src = bases + h_off[0];
	// the append occupies a multiple of 128 positions, padding is N-masked
	const uint64_t pos0 = c->n_pos;
	const uint64_t n_words = ((nb + 127) / 128) * 4;
	const uint64_t w0 = pos0 >> 5;
	const uint64_t hint_words = (c->prm.expected_bases + c->prm.expected_bases / 16) / 32 + 1024;
	const uint64_t want_words = std::max(w0 + n_words, w0 == 0 ? hint_words : 0);
	CLB_CUDA(c, c->pk.reserve(want_words + 8, s, true, w0));      // stage 2 reads windows across a word boundary and stages 16-byte aligned tiles: a few words of slack
	CLB_CUDA(c, c->nmask.reserve(want_words, s, true, w0));
	CLB_CUDA(c, c->smask.reserve(want_words, s, true, w0));
	if (!context) {
		CLB_CUDA(c, c->rd_start.reserve(c->n_reads + n_reads, s, true, c->n_reads));
		CLB_CUDA(c, c->rd_len.reserve(c->n_reads + n_reads, s, true, c->n_reads));
	} else {      // shift the existing reads up by n_reads ids
		DevBuf<uint64_t> ns; DevBuf<uint32_t> nl;
		CLB_CUDA(c, ns.reserve(c->n_reads + n_reads, s, false)); CLB_CUDA(c, nl.reserve(c->n_reads + n_reads, s, false));
		if (c->n_reads) {
			CLB_CUDA(c, cudaMemcpyAsync(ns.p + n_reads, c->rd_start.p, sizeof(uint64_t) * c->n_reads, cudaMemcpyDeviceToDevice, s));
			CLB_CUDA(c, cudaMemcpyAsync(nl.p + n_reads, c->rd_len.p, sizeof(uint32_t) * c->n_reads, cudaMemcpyDeviceToDevice, s));
			CLB_CUDA(c, cudaStreamSynchronize(s));
		}
		c->rd_start.release(); c->rd_len.release();
		c->rd_start = ns; c->rd_len = nl;
	}
	if (!on_device) {
		CLB_CUDA(c, c->stage_off.reserve(n_reads + 1, s, false));
		CLB_CUDA(c, cudaMemcpyAsync(c->stage_off.p, offsets, sizeof(uint64_t) * (n_reads + 1), cudaMemcpyHostToDevice, s));
		d_off = c->stage_off.p;
		const uint64_t stage_bytes = std::min<uint64_t>(nb, APPEND_CHUNK) + 64;
		for (int b = 0; b < 2; ++b) {
			CLB_CUDA(c, c->stage_in[b].reserve(stage_bytes, s, false));
			if (!c->ev_copied[b]) {
				CLB_CUDA(c, cudaEventCreateWithFlags(&c->ev_copied[b], cudaEventDisableTiming));
				CLB_CUDA(c, cudaEventCreateWithFlags(&c->ev_consumed[b], cudaEventDisableTiming));
			}
		}
		if (!c->copy_stream) CLB_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
		// the copy stream must not run ahead of whatever the compute stream did to the staging buffers before
		CLB_CUDA(c, cudaEventRecord(c->ev_consumed[0], s));
		CLB_CUDA(c, cudaEventRecord(c->ev_consumed[1], s));
	}

	// read-start mask of the whole append first (needs only the offsets), then chunk by chunk: copy, pack, count
	CLB_CUDA(c, cudaMemsetAsync(c->smask.p + w0, 0, sizeof(uint32_t) * n_words, s));
	k_mark_starts<<<(n_reads + 255) / 256, 256, 0, s>>>(d_off, n_reads, pos0, context ? 0 : c->n_reads, c->smask.p, c->rd_start.p, c->rd_len.p);
	CLB_LAUNCH_CHECK(c, "k_mark_starts");
	uint32_t ci = 0;
	for (uint64_t p = 0; p < n_words * 32; p += APPEND_CHUNK, ++ci) {
		const uint64_t pe = std::min(n_words * 32, p + APPEND_CHUNK);      // chunk = positions [p, pe) of this append
		const uint64_t cb = p < nb ? std::min(nb, pe) - p : 0;              // real bases in the chunk
		const uint64_t cw = (pe - p) >> 5, wb = w0 + (p >> 5);
		const uint8_t* d_src = src + p;
		if (!on_device) {
			const int b = ci & 1;
			CLB_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->ev_consumed[b], 0));
			if (cb) CLB_CUDA(c, cudaMemcpyAsync(c->stage_in[b].p, src + p, cb, cudaMemcpyHostToDevice, c->copy_stream));
			CLB_CUDA(c, cudaEventRecord(c->ev_copied[b], c->copy_stream));
			CLB_CUDA(c, cudaStreamWaitEvent(s, c->ev_copied[b], 0));
			d_src = c->stage_in[b].p;
		}
		CLB_TIMED(c, K_PACK, (k_pack<<<grid_for(cw, 256, c->n_sm, 8), 256, 0, s>>>(d_src, cb, cw, c->pk.p + wb, c->nmask.p + wb,
			(reinterpret_cast<uintptr_t>(d_src) & 15) == 0, c->d_scal)));
		CLB_LAUNCH_CHECK(c, "k_pack");
		if (!on_device) CLB_CUDA(c, cudaEventRecord(c->ev_consumed[ci & 1], s));
		if (context) continue;
		clb_status st = tab_ensure(c, (pe - p) / c->prm.modulo + (pe - p) / (4 * (uint64_t)c->prm.modulo) + 4096);
		if (st != CLB_OK) return st;
		CLB_TIMED(c, K_COUNT, (k_count<false><<<grid_for(cw, COUNT_THREADS, c->n_sm, 8), COUNT_THREADS, 0, s>>>(c->pk.p, c->nmask.p, c->smask.p,
			w0, wb, wb + cw, c->prm.kmer_len, c->mt, c->tab, c->tab_log2, c->d_scal)));
		CLB_LAUNCH_CHECK(c, "k_count");
	}
	if (count_only) {      // the sequences were counted and are forgotten: the next append takes their place in the store
		CLB_CUDA(c, cudaStreamSynchronize(s));
		unsigned long long sc[SC_COUNT];
		clb_status st = read_scalars(c, sc); if (st != CLB_OK) return st;
		if (sc[SC_BAD_SYMBOL]) return fail(c, CLB_ERR_BAD_SYMBOL, "input holds a symbol outside ACGTN");
		return CLB_OK;
	}
	if (context) {
		std::vector<uint64_t> hs(n_reads); std::vector<uint32_t> hl(n_reads);
		for (uint32_t i = 0; i < n_reads; ++i) { hs[i] = pos0 + (h_off[i] - h_off[0]); hl[i] = (uint32_t)(h_off[i + 1] - h_off[i]); }
		c->h_rd_start.insert(c->h_rd_start.begin(), hs.begin(), hs.end());
		c->h_rd_len.insert(c->h_rd_len.begin(), hl.begin(), hl.end());
		c->n_context = n_reads;
		unsigned long long sc[SC_COUNT];
		clb_status st = read_scalars(c, sc); if (st != CLB_OK) return st;
		if (sc[SC_BAD_SYMBOL]) return fail(c, CLB_ERR_BAD_SYMBOL, "input holds a symbol outside ACGTN");
	} else
	for (uint32_t i = 0; i < n_reads; ++i) {
		c->h_rd_start.push_back(pos0 + (h_off[i] - h_off[0]));
		c->h_rd_len.push_back((uint32_t)(h_off[i + 1] - h_off[i]));
	}
	c->n_pos = pos0 + n_words * 32;
	c->n_reads += n_reads;
	c->n_bases += nb;
	if (!on_device) CLB_CUDA(c, cudaStreamSynchronize(s));   // the caller may reuse its host buffers
	return CLB_OK;
}

clb_status s1a_counts_size(clb_ctx* c, uint32_t part, uint32_t n_parts, uint64_t* n)
{
	if (c->finalized) return fail(c, CLB_ERR_STATE, "count table already finalized");
	CLB_CUDA(c, cudaMemsetAsync(&c->d_scal[SC_CURSOR], 0, sizeof(unsigned long long), c->stream));
	const uint64_t cap = 1ULL << c->tab_log2;
	k_tab_export<false><<<grid_for(cap, 256, c->n_sm, 16), 256, 0, c->stream>>>(c->tab, cap, part, n_parts, nullptr, nullptr, 0, &c->d_scal[SC_CURSOR]);
	CLB_LAUNCH_CHECK(c, "k_tab_export<size>");
	unsigned long long sc[SC_COUNT];
	clb_status st = read_scalars(c, sc);
	if (st != CLB_OK) return st;
	if (sc[SC_OVERFLOW]) return fail(c, CLB_ERR_CUDA, "count table overflow (probe limit hit)");
	*n = sc[SC_CURSOR];
	return CLB_OK;
}

clb_status s1a_counts_export(clb_ctx* c, uint32_t part, uint32_t n_parts, uint64_t* kmers, uint32_t* counts, uint64_t cap_out, uint64_t* n_out, int on_device)
{
	if (c->finalized) return fail(c, CLB_ERR_STATE, "count table already finalized");
	uint64_t* dk = kmers; uint32_t* dc = counts;
	if (!on_device) {
		CLB_CUDA(c, dev_malloc((void**)&dk, sizeof(uint64_t) * (cap_out + 1), c->stream));
		CLB_CUDA(c, dev_malloc((void**)&dc, sizeof(uint32_t) * (cap_out + 1), c->stream));
	}
	CLB_CUDA(c, cudaMemsetAsync(&c->d_scal[SC_CURSOR], 0, sizeof(unsigned long long), c->stream));
	const uint64_t cap = 1ULL << c->tab_log2;
	k_tab_export<true><<<grid_for(cap, 256, c->n_sm, 16), 256, 0, c->stream>>>(c->tab, cap, part, n_parts, dk, dc, cap_out, &c->d_scal[SC_CURSOR]);
	CLB_LAUNCH_CHECK(c, "k_tab_export");
	unsigned long long sc[SC_COUNT];
	clb_status st = read_scalars(c, sc);
	if (st == CLB_OK && !on_device) {
		const uint64_t n = std::min<uint64_t>(sc[SC_CURSOR], cap_out);
		cudaMemcpy(kmers, dk, sizeof(uint64_t) * n, cudaMemcpyDeviceToHost);
		cudaMemcpy(counts, dc, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost);
	}
	if (!on_device) { dev_free(dk, c->stream); dev_free(dc, c->stream); }
	if (st != CLB_OK) return st;
	*n_out = sc[SC_CURSOR];
	if (sc[SC_CURSOR] > cap_out) return fail(c, CLB_ERR_CAPACITY, "clb_counts_export: buffer too small");
	return CLB_OK;
}

// sizes != nullptr: entries per partition (host); first != nullptr: write every partition at its offset (host array of n_parts)
static clb_status counts_all(clb_ctx* c, uint32_t n_parts, uint64_t* sizes, const uint64_t* first, uint64_t* kmers, uint32_t* counts, uint64_t cap_out)
{
	if (c->finalized) return fail(c, CLB_ERR_STATE, "count table already finalized");
	if (n_parts == 0 || n_parts > EXPORT_MAX_PARTS) return fail(c, CLB_ERR_BAD_ARG, "the count table is exported to 1 .. 64 partitions");
	unsigned long long* d = nullptr;      // cursor[n_parts], first[n_parts]
	CLB_CUDA(c, dev_malloc((void**)&d, sizeof(unsigned long long) * 2 * EXPORT_MAX_PARTS, c->stream));
	cudaError_t e = cudaMemsetAsync(d, 0, sizeof(unsigned long long) * 2 * EXPORT_MAX_PARTS, c->stream);
	if (e == cudaSuccess && first) e = cudaMemcpyAsync(d + EXPORT_MAX_PARTS, first, sizeof(uint64_t) * n_parts, cudaMemcpyHostToDevice, c->stream);
	const uint64_t cap = 1ULL << c->tab_log2;
	const uint64_t n_tiles = (cap + 256ull * EXPORT_PER_THREAD - 1) / (256ull * EXPORT_PER_THREAD);
	const uint32_t grid = (uint32_t)std::min<uint64_t>(n_tiles, (uint64_t)c->n_sm * 8);
	if (e == cudaSuccess) {
		if (first) k_tab_export_all<true><<<grid, 256, 0, c->stream>>>(c->tab, cap, n_parts, kmers, counts, cap_out, d, d + EXPORT_MAX_PARTS);
		else k_tab_export_all<false><<<grid, 256, 0, c->stream>>>(c->tab, cap, n_parts, nullptr, nullptr, 0, d, d + EXPORT_MAX_PARTS);
		++c->launches;
		e = cudaGetLastError();
	}
	unsigned long long h[EXPORT_MAX_PARTS] = {};
	if (e == cudaSuccess) e = cudaMemcpyAsync(h, d, sizeof(unsigned long long) * n_parts, cudaMemcpyDeviceToHost, c->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
	dev_free_async(d, c->stream);
	if (e != cudaSuccess) return cuda_fail(c, e, "k_tab_export_all");
	uint64_t total = 0;
	for (uint32_t p = 0; p < n_parts; ++p) { if (sizes) sizes[p] = h[p]; total += h[p]; }
	if (first && total > cap_out) return fail(c, CLB_ERR_CAPACITY, "clb_counts_export_all: buffer too small");
	return CLB_OK;
}
clb_status s1a_counts_sizes(clb_ctx* c, uint32_t n_parts, uint64_t* sizes)
{
	if (!sizes) return fail(c, CLB_ERR_BAD_ARG, "null argument");
	return counts_all(c, n_parts, sizes, nullptr, nullptr, nullptr, 0);
}
clb_status s1a_counts_export_all(clb_ctx* c, uint32_t n_parts, const uint64_t* first, uint64_t* kmers, uint32_t* counts, uint64_t cap_out)
{
	if (!first || !kmers || !counts) return fail(c, CLB_ERR_BAD_ARG, "null argument");
	return counts_all(c, n_parts, nullptr, first, kmers, counts, cap_out);
}

clb_status s1a_counts_reset(clb_ctx* c)
{
	if (c->finalized) return fail(c, CLB_ERR_STATE, "count table already finalized");
	const uint64_t cap = 1ULL << c->tab_log2;
	k_tab_clear<<<grid_for(cap, 256, c->n_sm, 16), 256, 0, c->stream>>>(c->tab, cap);
	CLB_LAUNCH_CHECK(c, "k_tab_clear");
	CLB_CUDA(c, cudaMemsetAsync(&c->d_scal[SC_TAB_USED], 0, sizeof(unsigned long long), c->stream));
	c->fill = clb_ctx::FillState{};
	return CLB_OK;
}

clb_status s1a_counts_merge(clb_ctx* c, const uint64_t* kmers, const uint32_t* counts, uint64_t n, uint64_t n_reads_remote, int on_device)
{
	if (c->finalized) return fail(c, CLB_ERR_STATE, "count table already finalized");
	c->n_reads_remote += n_reads_remote;
	if (n == 0) return CLB_OK;
	const uint64_t* dk = kmers; const uint32_t* dc = counts;
	uint64_t* tk = nullptr; uint32_t* tc = nullptr;
	if (!on_device) {
		CLB_CUDA(c, dev_malloc((void**)&tk, sizeof(uint64_t) * n, c->stream));
		CLB_CUDA(c, dev_malloc((void**)&tc, sizeof(uint32_t) * n, c->stream));
		CLB_CUDA(c, cudaMemcpyAsync(tk, kmers, sizeof(uint64_t) * n, cudaMemcpyHostToDevice, c->stream));
		CLB_CUDA(c, cudaMemcpyAsync(tc, counts, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, c->stream));
		dk = tk; dc = tc;
	}
	clb_status st = tab_ensure(c, n);
	if (st == CLB_OK) {
		k_tab_merge<<<grid_for(n, 256, c->n_sm, 16), 256, 0, c->stream>>>(dk, dc, n, c->tab, c->tab_log2, c->d_scal);
		++c->launches;
		cudaError_t e = cudaGetLastError();
		if (e != cudaSuccess) st = cuda_fail(c, e, "k_tab_merge");
	}
	if (!on_device) { cudaStreamSynchronize(c->stream); dev_free(tk, c->stream); dev_free(tc, c->stream); }
	return st;
}

static clb_status sv_alloc(clb_ctx* c, uint64_t n_surv)
{
	c->n_surv = n_surv;
	c->sv_log2 = 10;
	while ((1ULL << c->sv_log2) < 2 * n_surv + 1024) ++c->sv_log2;
	const uint64_t sv_cap = 1ULL << c->sv_log2;
	CLB_CUDA(c, dev_malloc((void**)&c->sv_keys, sizeof(uint64_t) * sv_cap, c->stream));
	CLB_CUDA(c, dev_malloc((void**)&c->sv_ids, sizeof(uint32_t) * sv_cap, c->stream));
	CLB_CUDA(c, dev_malloc((void**)&c->sv_kmer, sizeof(uint64_t) * (n_surv + 1), c->stream));
	CLB_CUDA(c, dev_malloc((void**)&c->sv_count, sizeof(uint32_t) * (n_surv + 1), c->stream));
	CLB_CUDA(c, cudaMemsetAsync(c->sv_keys, 0xFF, sizeof(uint64_t) * sv_cap, c->stream));
	CLB_CUDA(c, cudaMemsetAsync(c->sv_ids, 0xFF, sizeof(uint32_t) * sv_cap, c->stream));
	return CLB_OK;
}

clb_status s1a_finalize(clb_ctx* c, clb_kmer_stats* stats)
{
	if (c->finalized) { if (stats) *stats = c->stats; return CLB_OK; }
	cudaStream_t s = c->stream;
	unsigned long long sc[SC_COUNT];
	clb_status st = read_scalars(c, sc);
	if (st != CLB_OK) return st;
	if (sc[SC_BAD_SYMBOL]) return fail(c, CLB_ERR_BAD_SYMBOL, "input holds a symbol outside ACGTN");
	if (sc[SC_OVERFLOW]) return fail(c, CLB_ERR_CUDA, "count table overflow (probe limit hit)");
	const uint64_t local_pass = sc[SC_TOT_KMERS];
	CLB_CUDA(c, cudaMemsetAsync(c->d_scal, 0, sizeof(unsigned long long) * SC_COUNT, s));
	const uint64_t cap = 1ULL << c->tab_log2;
	CLB_TIMED(c, K_FINALIZE, (k_tab_stats<<<grid_for(cap, 256, c->n_sm, 16), 256, 0, s>>>(c->tab, cap, c->prm.min_count, c->prm.max_count, c->d_scal)));
	CLB_LAUNCH_CHECK(c, "k_tab_stats");
	st = read_scalars(c, sc);
	if (st != CLB_OK) return st;
	clb_kmer_stats r{};
	r.n_reads = c->n_reads + c->n_reads_remote;
	r.tot_kmers = sc[SC_TOT_KMERS];
	r.n_unique = sc[SC_N_UNIQUE];
	r.n_unique_counted = sc[SC_N_SURV];
	r.total_count_filtered = sc[SC_TOT_FILTERED];
	if (r.n_unique_counted >= 0xFFFFFFF0ull) return fail(c, CLB_ERR_BAD_ARG, "more than 2^32 filtered k-mers");
	st = sv_alloc(c, r.n_unique_counted);
	if (st != CLB_OK) return st;
	CLB_TIMED(c, K_FINALIZE, (k_build_survivors<<<grid_for(cap, 256, c->n_sm, 16), 256, 0, s>>>(c->tab, cap, c->prm.min_count, c->prm.max_count,
		c->sv_kmer, c->sv_count, c->sv_keys, c->sv_ids, c->sv_log2, &c->d_scal[SC_CURSOR])));
	CLB_LAUNCH_CHECK(c, "k_build_survivors");
	CLB_CUDA(c, cudaStreamSynchronize(s));
	dev_free(c->tab, c->stream); c->tab = nullptr;
	c->sum_true = local_pass;           // local passing occurrences bound the accepted k-mers of local reads
	c->stats = r; c->finalized = true;
	if (stats) *stats = r;
	return CLB_OK;
}

// Replace the filtered set by a listed one (multi-GPU: every rank imports the gathered survivors).
clb_status s1a_filter_import(clb_ctx* c, const uint64_t* kmers, const uint32_t* counts, uint64_t n, const clb_kmer_stats* global_stats, int on_device)
{
	if (!c->finalized) return fail(c, CLB_ERR_STATE, "clb_filter_import before clb_count_finalize");
	if (c->graph_done) return fail(c, CLB_ERR_STATE, "clb_filter_import after clb_graph_build");
	dev_free(c->sv_keys, c->stream); dev_free(c->sv_ids, c->stream); dev_free(c->sv_kmer, c->stream); dev_free(c->sv_count, c->stream);
	c->sv_keys = nullptr; c->sv_ids = nullptr; c->sv_kmer = nullptr; c->sv_count = nullptr;
	clb_status st = sv_alloc(c, n);
	if (st != CLB_OK) return st;
	const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
	if (n) {
		CLB_CUDA(c, cudaMemcpyAsync(c->sv_kmer, kmers, sizeof(uint64_t) * n, kind, c->stream));
		CLB_CUDA(c, cudaMemcpyAsync(c->sv_count, counts, sizeof(uint32_t) * n, kind, c->stream));
		k_sv_import<<<grid_for(n, 256, c->n_sm, 16), 256, 0, c->stream>>>(c->sv_kmer, n, c->sv_keys, c->sv_ids, c->sv_log2);
		CLB_LAUNCH_CHECK(c, "k_sv_import");
	}
	CLB_CUDA(c, cudaStreamSynchronize(c->stream));
	if (global_stats) c->stats = *global_stats;
	return CLB_OK;
}

clb_status s1a_filter_check(clb_ctx* c, const uint64_t* kmers, uint64_t n, uint8_t* possible, uint8_t* present)
{
	if (!c->finalized) return fail(c, CLB_ERR_STATE, "clb_filter_check before clb_count_finalize");
	if (n == 0) return CLB_OK;
	uint64_t* dk = nullptr; uint8_t* dp = nullptr;
	CLB_CUDA(c, dev_malloc((void**)&dk, sizeof(uint64_t) * n, c->stream));
	CLB_CUDA(c, dev_malloc((void**)&dp, 2 * n, c->stream));
	CLB_CUDA(c, cudaMemcpyAsync(dk, kmers, sizeof(uint64_t) * n, cudaMemcpyHostToDevice, c->stream));
	k_filter_check<<<(uint32_t)((n + 255) / 256), 256, 0, c->stream>>>(dk, n, c->mt, c->sv_keys, c->sv_log2, dp, dp + n);
	++c->launches;
	cudaError_t e = cudaGetLastError();
	if (e == cudaSuccess) e = cudaMemcpyAsync(possible, dp, n, cudaMemcpyDeviceToHost, c->stream);
	if (e == cudaSuccess) e = cudaMemcpyAsync(present, dp + n, n, cudaMemcpyDeviceToHost, c->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
	dev_free(dk, c->stream); dev_free(dp, c->stream);
	if (e != cudaSuccess) return cuda_fail(c, e, "k_filter_check");
	return CLB_OK;
}

} // namespace clb
