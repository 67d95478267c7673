// stage1b.cu — accepted k-mers per read and the causal similarity graph, re-derived for a GPU.
//
// Reference semantics (src/colord/reads_sim_graph.cpp):
//   :134-164  per read: canonical k-mers in order, kept iff Possible && first occurrence in the read && Check
//   :324-427  serial loop in input order: for read i, every (k-mer -> earlier reference read) entry votes;
//             then, if read i is a reference, each of its k-mers whose list is shorter than maxKmerCount
//             gains (k-mer -> ref id of i); top max_candidates by (votes desc, id asc)
//   :429-528  HiFi: also the shared k-mers of every chosen candidate, in the read's k-mer order
//   :295-322  reference-genome pseudo-reads: inserted first, without the cap
//
// The serial loop is order-dependent only through two facts, both of which have a closed form:
//   (1) a k-mer's list ends up holding the pseudo-read entries plus the FIRST (by reference id) normal
//       reference reads that contain it, until the list is maxKmerCount long;
//   (2) read i sees exactly the entries whose reference id is smaller than the number of reference reads
//       that precede i in the input.
// So the table is built for all reads at once (count -> scan -> fill, lists longer than the cap are sorted
// and truncated) and every read votes independently against the finished table with an id limit.
#include "ctx.h"
#include <algorithm>
#include <numeric>

namespace clb {

// ------------------------------------------------------------------------------------------------
__global__ void k_read_flags(const uint32_t* __restrict__ nmask, const uint64_t* __restrict__ rd_start, const uint32_t* __restrict__ rd_len,
	uint32_t n_reads, uint8_t* __restrict__ has_n)
{
	const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (r >= n_reads) return;
	const uint64_t s = rd_start[r], e = s + rd_len[r];
	uint32_t any = 0;
	if (e > s) {
		const uint64_t w0 = s >> 5, w1 = (e - 1) >> 5;
		for (uint64_t w = w0 + lane; w <= w1; w += 32) {
			uint32_t m = nmask[w];
			if (w == w0) m &= ~0u << (s & 31);
			if (w == w1 && ((e & 31) != 0)) m &= (1u << (e & 31)) - 1;
			any |= m;
		}
	}
	any = __any_sync(0xffffffffu, any != 0);
	if (lane == 0) has_n[r] = (uint8_t)any;
}

// ------------------------------------------------------------------------------------------------
// Accepted k-mers.  One CTA per read.  Shared (or, for very long reads, global) scratch:
//   map   : open-addressed u64 (dense survivor id << 32 | first position in the read), atomicMin keeps the
//           first occurrence
//   bits  : one bit per read position that starts... ends a kept k-mer; a prefix popcount turns a position
//           into the k-mer's rank, so the list comes out in read order without a sort
// ------------------------------------------------------------------------------------------------
constexpr int ACC_THREADS = 128;
constexpr uint32_t PENDING = 0xFFFFFFFFu;

struct AccArgs {
	const uint64_t* pk; const uint32_t* nmask; const uint32_t* smask;
	const uint64_t* rd_start; const uint32_t* rd_len; const uint8_t* has_n;
	const uint64_t* sv_keys; const uint32_t* sv_ids; uint32_t sv_log2;
	uint32_t k; ModTest mt; uint32_t modulo;
	const uint32_t* list; uint32_t n_list;      // reads to process (nullptr = all)
	uint64_t* acc_start; uint32_t* acc_n; uint32_t* acc_id; uint64_t acc_cap;
	unsigned long long* scal;                    // SC_CURSOR = arena cursor, SC_CURSOR2 = #pending, SC_OVERFLOW
	// global scratch variant
	uint64_t* g_map; uint32_t* g_bits; const uint64_t* g_map_off; const uint64_t* g_bits_off;
};

CLB_D uint32_t sv_lookup(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ ids, uint32_t log2cap, uint64_t kmer, uint64_t h)
{
	const uint64_t mask = (1ULL << log2cap) - 1;
	uint64_t s = slot_of(h, log2cap);
	for (;;) {
		const uint64_t kx = keys[s];
		if (kx == kmer) return ids[s];
		if (kx == EMPTY64) return EMPTY32;
		s = (s + 1) & mask;
	}
}

// returns false if the map is too full (caller flags the read as pending for a larger scratch class)
CLB_D bool map_put(uint64_t* map, uint32_t cap_mask, uint32_t log2cap, uint32_t id, uint32_t pos, uint32_t max_probe)
{
	const unsigned long long val = ((unsigned long long)id << 32) | pos;
	uint32_t s = (id * 0x9E3779B1u) >> (32 - log2cap);
	for (uint32_t probe = 0; probe < max_probe; ++probe) {
		unsigned long long cur = map[s];
		if (cur == EMPTY64) {
			cur = atomicCAS(reinterpret_cast<unsigned long long*>(&map[s]), EMPTY64, val);
			if (cur == EMPTY64) return true;
		}
		if ((uint32_t)(cur >> 32) == id) { atomicMin(reinterpret_cast<unsigned long long*>(&map[s]), val); return true; }
		s = (s + 1) & cap_mask;
	}
	return false;
}

template <int MAP_SLOTS, int BM_WORDS, bool GLOBAL>
__global__ void __launch_bounds__(ACC_THREADS) k_accept(AccArgs a)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	__shared__ uint64_t q_kmer[ACC_THREADS / 32][64];
	__shared__ uint32_t q_pos[ACC_THREADS / 32][64];
	__shared__ uint32_t ws[33];
	__shared__ uint32_t s_fail;
	__shared__ unsigned long long s_base;

	const uint32_t bi = blockIdx.x;
	const uint32_t r = a.list ? a.list[bi] : bi;
	const uint32_t len = a.rd_len[r];
	const uint32_t k = a.k;
	if (a.has_n[r] || len < k) {
		if (threadIdx.x == 0) { a.acc_n[r] = 0; a.acc_start[r] = 0; }
		return;
	}
	const uint32_t bm_words = (len + 31) / 32;
	// scratch capacity for this read
	uint32_t want = 2 * (len / a.modulo) + 128;
	uint32_t log2cap = 7;
	uint64_t* map; uint32_t* bits; uint32_t* pref;
	if (GLOBAL) {
		const uint64_t mo = a.g_map_off[bi], mc = a.g_map_off[bi + 1] - mo;
		while ((1ULL << (log2cap + 1)) <= mc) ++log2cap;     // largest pow2 <= reserved
		map = a.g_map + mo;
		bits = a.g_bits + a.g_bits_off[bi];
		pref = bits + bm_words;
	} else {
		while ((1u << log2cap) < want && (1u << log2cap) < (uint32_t)MAP_SLOTS) ++log2cap;
		map = reinterpret_cast<uint64_t*>(smem_raw);
		bits = reinterpret_cast<uint32_t*>(smem_raw + sizeof(uint64_t) * MAP_SLOTS);
		pref = bits + BM_WORDS;
		if (bm_words > (uint32_t)BM_WORDS) {                 // host classification should prevent this
			if (threadIdx.x == 0) { a.acc_n[r] = PENDING; atomicAdd(&a.scal[SC_CURSOR2], 1ULL); }
			return;
		}
	}
	const uint32_t cap = 1u << log2cap, cap_mask = cap - 1;
	const uint32_t max_probe = GLOBAL ? cap : 128;
	for (uint32_t i = threadIdx.x; i < cap; i += ACC_THREADS) map[i] = EMPTY64;
	for (uint32_t i = threadIdx.x; i < bm_words; i += ACC_THREADS) bits[i] = 0;
	if (threadIdx.x == 0) s_fail = 0;
	__syncthreads();

	const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5, lt = (1u << lane) - 1;
	uint64_t* myqk = q_kmer[wid]; uint32_t* myqp = q_pos[wid];
	uint32_t qn = 0; bool failed = false;
	const uint64_t kmask = k == 32 ? ~0ULL : ((1ULL << (2 * k)) - 1);
	const uint32_t rsh = 2 * (k - 1);
	const uint64_t win_n = ((1ULL << k) - 1) << (33 - k);
	const uint64_t win_s = win_n & (win_n - 1);
	const uint64_t start = a.rd_start[r], end = start + len;          // [start, end)
	const uint64_t first_end = start + k - 1;                           // first position that ends a k-mer
	const uint64_t gw0 = first_end >> 5, gw1 = (end - 1) >> 5;
	const uint64_t n_words = gw1 - gw0 + 1;

	auto drain = [&](uint32_t n_take, uint32_t from) {
		// lanes < n_take look one queued k-mer up in the filtered set and record its first position
		if (lane < n_take) {
			const uint64_t x = myqk[from + lane];
			const uint32_t id = sv_lookup(a.sv_keys, a.sv_ids, a.sv_log2, x, murmur64(x));
			if (id != EMPTY32 && !map_put(map, cap_mask, log2cap, id, myqp[from + lane], max_probe)) failed = true;
		}
	};

	for (uint64_t base = (wid << 5); base < n_words; base += ACC_THREADS) {
		const uint64_t i = base + lane;
		const bool live = i < n_words;
		const uint64_t w = gw0 + (live ? i : 0);
		uint64_t cur = 0, prev = 0, N64 = ~0ULL, S64 = 0;
		if (live) {
			cur = a.pk[w];
			const uint32_t nm = a.nmask[w], sm = a.smask[w];
			uint32_t pnm = 0xFFFFFFFFu, psm = 0;
			if (w > 0) { prev = a.pk[w - 1]; pnm = a.nmask[w - 1]; psm = a.smask[w - 1]; }
			N64 = ((uint64_t)nm << 32) | pnm;
			S64 = ((uint64_t)sm << 32) | psm;
		}
		uint64_t fw = prev & (kmask >> 2);
		uint64_t rc = revcomp(fw, k);
		const uint64_t p0 = w << 5;
#pragma unroll 4
		for (int j = 0; j < 32; ++j) {
			const uint64_t b = (cur >> (62 - 2 * j)) & 3;
			fw = ((fw << 2) | b) & kmask;
			rc = (rc >> 2) | ((3 - b) << rsh);
			const uint64_t can = fw < rc ? fw : rc;
			const uint64_t p = p0 + j;
			const bool ok = live && p >= first_end && p < end && ((N64 & (win_n << j)) | (S64 & (win_s << j))) == 0;
			const bool pass = ok && divisible(murmur64(can), a.mt);
			const uint32_t bal = __ballot_sync(0xffffffffu, pass);
			if (bal) {
				if (pass) { const uint32_t o = qn + __popc(bal & lt); myqk[o] = can; myqp[o] = (uint32_t)(p - start); }
				qn += __popc(bal);
				if (qn >= 32) { __syncwarp(); qn -= 32; drain(32, qn); __syncwarp(); }
			}
		}
	}
	__syncwarp();
	drain(qn, 0);
	if (failed) s_fail = 1;
	__syncthreads();
	if (s_fail) {
		if (threadIdx.x == 0) { a.acc_n[r] = PENDING; atomicAdd(&a.scal[SC_CURSOR2], 1ULL); }
		return;
	}
	// positions of the kept k-mers -> bitmap
	for (uint32_t i = threadIdx.x; i < cap; i += ACC_THREADS) {
		const uint64_t v = map[i];
		if (v != EMPTY64) { const uint32_t pos = (uint32_t)v; atomicOr(&bits[pos >> 5], 1u << (pos & 31)); }
	}
	__syncthreads();
	// exclusive prefix popcount per bitmap word
	uint32_t running = 0;
	for (uint32_t base = 0; base < bm_words; base += ACC_THREADS) {
		const uint32_t i = base + threadIdx.x;
		const uint32_t c = i < bm_words ? __popc(bits[i]) : 0;
		uint32_t tot;
		const uint32_t ex = block_excl_scan(c, ws, &tot);
		if (i < bm_words) pref[i] = running + ex;
		running += tot;
	}
	const uint32_t n_kept = running;
	if (threadIdx.x == 0) {
		const unsigned long long b = atomicAdd(&a.scal[SC_CURSOR], (unsigned long long)n_kept);
		s_base = b;
		a.acc_start[r] = b; a.acc_n[r] = n_kept;
		if (b + n_kept > a.acc_cap) atomicOr(&a.scal[SC_OVERFLOW], 1ULL);
	}
	__syncthreads();
	const unsigned long long out = s_base;
	if (out + n_kept > a.acc_cap) return;
	for (uint32_t i = threadIdx.x; i < cap; i += ACC_THREADS) {
		const uint64_t v = map[i];
		if (v == EMPTY64) continue;
		const uint32_t pos = (uint32_t)v;
		const uint32_t rank = pref[pos >> 5] + __popc(bits[pos >> 5] & ((1u << (pos & 31)) - 1));
		a.acc_id[out + rank] = (uint32_t)(v >> 32);
	}
}

// ------------------------------------------------------------------------------------------------
// Posting lists (k-mer -> reference reads).
// ------------------------------------------------------------------------------------------------
template <bool FILL>
__global__ void k_post_pass(const uint64_t* __restrict__ acc_start, const uint32_t* __restrict__ acc_n, const uint32_t* __restrict__ acc_id,
	const uint8_t* __restrict__ is_ref, const uint32_t* __restrict__ ref_before, uint32_t n_reads,
	uint32_t* __restrict__ cnt, const uint64_t* __restrict__ off, uint32_t* __restrict__ post)
{
	const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (r >= n_reads || !is_ref[r]) return;
	const uint64_t s = acc_start[r]; const uint32_t n = acc_n[r];
	const uint32_t ref_id = ref_before[r];
	for (uint32_t e = lane; e < n; e += 32) {
		const uint32_t id = acc_id[s + e];
		const uint32_t slot = atomicAdd(&cnt[id], 1u);
		if (FILL) post[off[id] + slot] = ref_id;
	}
}

// Exclusive scan u32 -> u64 in three launches (tile = 256 threads x 8 items).
constexpr int SCAN_THREADS = 256, SCAN_ITEMS = 8, SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(const uint32_t* __restrict__ in, uint64_t n, uint64_t* __restrict__ tile_sum)
{
	__shared__ uint32_t ws[33];
	const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
	uint32_t s = 0;
#pragma unroll
	for (int i = 0; i < SCAN_ITEMS; ++i) if (base + i < n) s += in[base + i];
	uint32_t tot; block_excl_scan(s, ws, &tot);
	if (threadIdx.x == 0) tile_sum[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(1024) k_scan_sums(uint64_t* __restrict__ tile_sum, uint64_t n_tiles, unsigned long long* __restrict__ total)
{
	// single CTA: sequential chunks of 1024 with a shared running offset
	__shared__ unsigned long long sh[1024];
	__shared__ unsigned long long carry;
	if (threadIdx.x == 0) carry = 0;
	__syncthreads();
	for (uint64_t base = 0; base < n_tiles; base += 1024) {
		const uint64_t i = base + threadIdx.x;
		const unsigned long long v = i < n_tiles ? tile_sum[i] : 0;
		sh[threadIdx.x] = v;
		__syncthreads();
		for (int d = 1; d < 1024; d <<= 1) {
			unsigned long long t = threadIdx.x >= (uint32_t)d ? sh[threadIdx.x - d] : 0;
			__syncthreads();
			sh[threadIdx.x] += t;
			__syncthreads();
		}
		if (i < n_tiles) tile_sum[i] = carry + sh[threadIdx.x] - v;
		__syncthreads();
		if (threadIdx.x == 1023) carry += sh[1023];
		__syncthreads();
	}
	if (threadIdx.x == 0) *total = carry;
}
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const uint32_t* __restrict__ in, uint64_t n, const uint64_t* __restrict__ tile_off, uint64_t* __restrict__ out)
{
	__shared__ uint32_t ws[33];
	const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
	uint32_t v[SCAN_ITEMS]; uint32_t s = 0;
#pragma unroll
	for (int i = 0; i < SCAN_ITEMS; ++i) { v[i] = base + i < n ? in[base + i] : 0; s += v[i]; }
	uint32_t tot; const uint32_t ex = block_excl_scan(s, ws, &tot);
	uint64_t run = tile_off[blockIdx.x] + ex;
#pragma unroll
	for (int i = 0; i < SCAN_ITEMS; ++i) { if (base + i < n) out[base + i] = run; run += v[i]; }
}

// Lists longer than the cap: collect their ids ...
__global__ void k_post_oversize(const uint32_t* __restrict__ cnt, uint64_t n, uint32_t cap, uint32_t* __restrict__ list, unsigned long long* __restrict__ cursor)
{
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n && cnt[i] > cap) list[atomicAdd(cursor, 1ULL)] = (uint32_t)i;
}
// ... sort each ascending (one CTA per list, in place in global memory; a bitonic network written in its
// all-ascending form so that the virtual padding to a power of two never moves) and keep the pseudo-read
// entries plus the first normal reference reads up to the cap (reads_sim_graph.cpp:391-394 applied in id order).
__global__ void __launch_bounds__(256) k_post_truncate(const uint32_t* __restrict__ list, const uint64_t* __restrict__ off,
	uint32_t* __restrict__ cnt, uint32_t* __restrict__ post, uint32_t cap, uint32_t n_pseudo)
{
	const uint32_t id = list[blockIdx.x];
	uint32_t* a = post + off[id];
	const uint32_t n = cnt[id];
	uint32_t p2 = 1; while (p2 < n) p2 <<= 1;
	for (uint32_t k = 2; k <= p2; k <<= 1) {
		for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
			const uint32_t l = i ^ (k - 1);
			if (l > i && l < n) { const uint32_t x = a[i], y = a[l]; if (x > y) { a[i] = y; a[l] = x; } }
		}
		__syncthreads();
		for (uint32_t j = k >> 2; j > 0; j >>= 1) {
			for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
				const uint32_t l = i ^ j;
				if (l > i && l < n) { const uint32_t x = a[i], y = a[l]; if (x > y) { a[i] = y; a[l] = x; } }
			}
			__syncthreads();
		}
	}
	if (threadIdx.x == 0) {
		uint32_t lo = 0, hi = n;                       // number of pseudo-read entries (ids < n_pseudo)
		while (lo < hi) { const uint32_t m = (lo + hi) >> 1; if (a[m] < n_pseudo) lo = m + 1; else hi = m; }
		const uint32_t keep = lo > cap ? lo : cap;
		cnt[id] = n < keep ? n : keep;
	}
}

// ------------------------------------------------------------------------------------------------
// Voting + top-c.  One CTA per read; 8 lanes walk one k-mer's list.
// ------------------------------------------------------------------------------------------------
constexpr int VOTE_THREADS = 128;
struct VoteArgs {
	const uint64_t* acc_start; const uint32_t* acc_n; const uint32_t* acc_id;
	const uint32_t* ref_before; const uint32_t* post_cnt; const uint64_t* post_off; const uint32_t* post;
	uint32_t n_pseudo, max_cand;
	const uint32_t* list; uint32_t n_list;
	uint32_t* cand; uint32_t* cand_votes; uint32_t* cand_n;
	unsigned long long* scal;          // SC_CURSOR2 = #pending
	uint32_t* g_keys; const uint64_t* g_off;   // global scratch variant: keys then vals, capacity from g_off
};

template <int SLOTS, bool GLOBAL>
__global__ void __launch_bounds__(VOTE_THREADS) k_vote(VoteArgs a)
{
	__shared__ uint32_t s_keys[GLOBAL ? 1 : SLOTS];
	__shared__ uint32_t s_vals[GLOBAL ? 1 : SLOTS];
	__shared__ uint32_t s_distinct, s_fail;
	__shared__ unsigned long long s_red[VOTE_THREADS / 32];
	__shared__ unsigned long long s_best;

	const uint32_t bi = blockIdx.x;
	const uint32_t r = a.list ? a.list[bi] : bi;
	const uint32_t n = a.acc_n[r];
	const uint32_t limit = a.ref_before[r];
	if (r < a.n_pseudo || n == 0 || limit == 0) {
		if (threadIdx.x == 0) a.cand_n[r] = 0;
		return;
	}
	uint32_t* keys; uint32_t* vals; uint32_t cap;
	if (GLOBAL) {
		const uint64_t o = a.g_off[bi]; cap = (uint32_t)((a.g_off[bi + 1] - o) / 2);
		keys = a.g_keys + o; vals = keys + cap;
	} else { keys = s_keys; vals = s_vals; cap = SLOTS; }
	const uint32_t mask = cap - 1;
	for (uint32_t i = threadIdx.x; i < cap; i += VOTE_THREADS) { keys[i] = EMPTY32; vals[i] = 0; }
	if (threadIdx.x == 0) { s_distinct = 0; s_fail = 0; }
	__syncthreads();

	const uint64_t s0 = a.acc_start[r];
	const uint32_t grp = threadIdx.x >> 3, sub = threadIdx.x & 7;
	const uint32_t max_distinct = GLOBAL ? cap : (SLOTS / 4) * 3;   // the table can never fill up completely
	volatile uint32_t* vfail = &s_fail;
	for (uint32_t e = grp; e < n; e += VOTE_THREADS / 8) {
		if (*vfail) break;
		const uint32_t id = a.acc_id[s0 + e];
		const uint32_t c = a.post_cnt[id];
		const uint32_t* pl = a.post + a.post_off[id];
		for (uint32_t x = sub; x < c; x += 8) {
			const uint32_t ref = pl[x];
			if (ref >= limit) continue;
			uint32_t s = (ref * 0x9E3779B1u) & mask;
			for (;;) {
				uint32_t cur = keys[s];
				if (cur == EMPTY32) {
					cur = atomicCAS(&keys[s], EMPTY32, ref);
					if (cur == EMPTY32) { if (atomicAdd(&s_distinct, 1u) >= max_distinct) *vfail = 1; cur = ref; }
				}
				if (cur == ref) { atomicAdd(&vals[s], 1u); break; }
				s = (s + 1) & mask;
			}
		}
	}
	__syncthreads();
	if (s_fail) {
		if (threadIdx.x == 0) { a.cand_n[r] = PENDING; atomicAdd(&a.scal[SC_CURSOR2], 1ULL); }
		return;
	}
	// top max_cand by (votes desc, ref id asc): repeated CTA-wide arg-max
	uint32_t found = 0;
	for (uint32_t round = 0; round < a.max_cand; ++round) {
		unsigned long long best = 0; uint32_t best_slot = 0;
		for (uint32_t i = threadIdx.x; i < cap; i += VOTE_THREADS) {
			const uint32_t v = vals[i];
			if (v) {
				const unsigned long long key = ((unsigned long long)v << 32) | (0xFFFFFFFFu - keys[i]);
				if (key > best) { best = key; best_slot = i; }
			}
		}
		unsigned long long wbest = best;
#pragma unroll
		for (int d = 16; d; d >>= 1) { const unsigned long long o = __shfl_xor_sync(0xffffffffu, wbest, d); if (o > wbest) wbest = o; }
		if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = wbest;
		__syncthreads();
		if (threadIdx.x == 0) {
			unsigned long long b = 0;
			for (int w = 0; w < VOTE_THREADS / 32; ++w) if (s_red[w] > b) b = s_red[w];
			s_best = b;
		}
		__syncthreads();
		const unsigned long long gb = s_best;
		if (gb == 0) break;
		if (best == gb) {              // unique owner: reference ids are distinct keys
			a.cand[(uint64_t)r * a.max_cand + round] = 0xFFFFFFFFu - (uint32_t)gb;
			a.cand_votes[(uint64_t)r * a.max_cand + round] = (uint32_t)(gb >> 32);
			vals[best_slot] = 0;
		}
		++found;
		__syncthreads();
	}
	if (threadIdx.x == 0) a.cand_n[r] = found;
}

// Upper bound of distinct neighbours of a read (sum of its k-mers' list lengths) for the global-scratch class.
__global__ void k_vote_bound(const uint32_t* __restrict__ list, uint32_t n_list, const uint64_t* __restrict__ acc_start, const uint32_t* __restrict__ acc_n,
	const uint32_t* __restrict__ acc_id, const uint32_t* __restrict__ post_cnt, unsigned long long* __restrict__ bound)
{
	const uint32_t bi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (bi >= n_list) return;
	const uint32_t r = list[bi];
	const uint64_t s = acc_start[r]; const uint32_t n = acc_n[r];
	unsigned long long t = 0;
	for (uint32_t e = lane; e < n; e += 32) t += post_cnt[acc_id[s + e]];
#pragma unroll
	for (int d = 16; d; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
	if (lane == 0) bound[bi] = t;
}

// ------------------------------------------------------------------------------------------------
// HiFi: shared k-mers of the chosen candidates, in the read's k-mer order (reads_sim_graph.cpp:466-484).
// ------------------------------------------------------------------------------------------------
__global__ void k_common_total(const uint32_t* __restrict__ cand_n, const uint32_t* __restrict__ cand_votes, uint32_t n_reads, uint32_t max_cand,
	unsigned long long* __restrict__ total)
{
	uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
	unsigned long long t = 0;
	if (r < n_reads) for (uint32_t j = 0; j < cand_n[r]; ++j) t += cand_votes[(uint64_t)r * max_cand + j];
#pragma unroll
	for (int d = 16; d; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
	if ((threadIdx.x & 31) == 0 && t) atomicAdd(total, t);
}

constexpr int COMMON_THREADS = 128;
__global__ void __launch_bounds__(COMMON_THREADS) k_common(const uint64_t* __restrict__ acc_start, const uint32_t* __restrict__ acc_n, const uint32_t* __restrict__ acc_id,
	const uint32_t* __restrict__ ref_before, const uint32_t* __restrict__ post_cnt, const uint64_t* __restrict__ post_off, const uint32_t* __restrict__ post,
	const uint64_t* __restrict__ sv_kmer, const uint32_t* __restrict__ cand, const uint32_t* __restrict__ cand_votes, const uint32_t* __restrict__ cand_n,
	uint32_t max_cand, uint64_t* __restrict__ common_off, uint64_t* __restrict__ common, unsigned long long* __restrict__ cursor)
{
	__shared__ uint32_t s_cand[32];
	__shared__ unsigned long long s_base[32];
	__shared__ uint32_t ws[33];
	const uint32_t r = blockIdx.x;
	const uint32_t cn = cand_n[r];
	if (cn == 0) return;
	if (threadIdx.x == 0) {
		unsigned long long tot = 0;
		for (uint32_t j = 0; j < cn; ++j) tot += cand_votes[(uint64_t)r * max_cand + j];
		unsigned long long b = atomicAdd(cursor, tot);
		for (uint32_t j = 0; j < cn; ++j) {
			s_cand[j] = cand[(uint64_t)r * max_cand + j];
			s_base[j] = b; common_off[(uint64_t)r * max_cand + j] = b;
			b += cand_votes[(uint64_t)r * max_cand + j];
		}
	}
	__syncthreads();
	const uint32_t n = acc_n[r], limit = ref_before[r];
	const uint64_t s0 = acc_start[r];
	uint32_t run[32];
#pragma unroll
	for (int j = 0; j < 32; ++j) run[j] = 0;
	for (uint32_t base = 0; base < n; base += COMMON_THREADS) {
		const uint32_t e = base + threadIdx.x;
		uint32_t hit = 0, id = 0;
		if (e < n) {
			id = acc_id[s0 + e];
			const uint32_t c = post_cnt[id];
			const uint32_t* pl = post + post_off[id];
			for (uint32_t x = 0; x < c; ++x) {
				const uint32_t ref = pl[x];
				if (ref >= limit) continue;
				for (uint32_t j = 0; j < cn; ++j) if (s_cand[j] == ref) hit |= 1u << j;
			}
		}
#pragma unroll
		for (int j = 0; j < 32; ++j) {
			if ((uint32_t)j >= cn) break;
			const uint32_t f = (hit >> j) & 1;
			uint32_t tot;
			const uint32_t ex = block_excl_scan(f, ws, &tot);
			if (f) common[s_base[j] + run[j] + ex] = sv_kmer[id];
			run[j] += tot;
		}
	}
}

// ------------------------------------------------------------------------------------------------
// Host orchestration
// ------------------------------------------------------------------------------------------------
static clb_status scal_read(clb_ctx* c, unsigned long long* out)
{
	CLB_CUDA(c, cudaMemcpyAsync(out, c->d_scal, sizeof(unsigned long long) * SC_COUNT, cudaMemcpyDeviceToHost, c->stream));
	CLB_CUDA(c, cudaStreamSynchronize(c->stream));
	return CLB_OK;
}
static clb_status scal_zero(clb_ctx* c, int which)
{
	CLB_CUDA(c, cudaMemsetAsync(&c->d_scal[which], 0, sizeof(unsigned long long), c->stream));
	return CLB_OK;
}

template <typename T> static cudaError_t dev_alloc(T** p, uint64_t n, cudaStream_t s = nullptr) { return dev_malloc((void**)p, sizeof(T) * (n ? n : 1), s); }

clb_status exclusive_scan(clb_ctx* c, const uint32_t* in, uint64_t n, uint64_t* out, uint64_t* total)
{
	const uint64_t n_tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
	uint64_t* tiles = nullptr;
	CLB_CUDA(c, dev_malloc((void**)&tiles, sizeof(uint64_t) * (n_tiles + 1), c->stream));
	clb_status st = CLB_OK;
	if (n_tiles) {
		k_scan_tiles<<<(uint32_t)n_tiles, SCAN_THREADS, 0, c->stream>>>(in, n, tiles); ++c->launches;
		k_scan_sums<<<1, 1024, 0, c->stream>>>(tiles, n_tiles, &c->d_scal[SC_SUM_TRUE]); ++c->launches;
		k_scan_apply<<<(uint32_t)n_tiles, SCAN_THREADS, 0, c->stream>>>(in, n, tiles, out); ++c->launches;
		cudaError_t e = cudaGetLastError();
		if (e != cudaSuccess) st = cuda_fail(c, e, "exclusive_scan");
	} else {
		cudaMemsetAsync(&c->d_scal[SC_SUM_TRUE], 0, sizeof(unsigned long long), c->stream);
	}
	unsigned long long sc[SC_COUNT];
	if (st == CLB_OK) st = scal_read(c, sc);
	dev_free_async(tiles, c->stream);
	if (st == CLB_OK) *total = sc[SC_SUM_TRUE];
	return st;
}

// smem bytes of the two shared-memory classes of k_accept
constexpr int ACC_A_MAP = 4096, ACC_A_BM = 2048;       // reads up to 65 536 bases
constexpr int ACC_B_MAP = 16384, ACC_B_BM = 8192;      // reads up to 262 144 bases
static constexpr size_t acc_smem(int map, int bm) { return sizeof(uint64_t) * map + sizeof(uint32_t) * 2 * bm; }

static clb_status collect_pending(clb_ctx* c, const uint32_t* d_field, std::vector<uint32_t>& out)
{
	std::vector<uint32_t> h(c->n_reads);
	CLB_CUDA(c, cudaMemcpyAsync(h.data(), d_field, sizeof(uint32_t) * c->n_reads, cudaMemcpyDeviceToHost, c->stream));
	CLB_CUDA(c, cudaStreamSynchronize(c->stream));
	for (uint64_t i = 0; i < c->n_reads; ++i) if (h[i] == PENDING) out.push_back((uint32_t)i);
	return CLB_OK;
}

static clb_status run_accept(clb_ctx* c)
{
	cudaStream_t s = c->stream;
	const uint64_t n = c->n_reads;
	c->acc_cap = c->sum_true + 1;
	CLB_CUDA(c, dev_alloc(&c->acc_start, n, c->stream));
	CLB_CUDA(c, dev_alloc(&c->acc_n, n, c->stream));
	CLB_CUDA(c, dev_alloc(&c->acc_id, c->acc_cap, c->stream));
	CLB_CUDA(c, cudaMemsetAsync(c->acc_n, 0, sizeof(uint32_t) * (n ? n : 1), s));
	scal_zero(c, SC_CURSOR); scal_zero(c, SC_CURSOR2); scal_zero(c, SC_OVERFLOW);

	AccArgs a{};
	a.pk = c->pk.p; a.nmask = c->nmask.p; a.smask = c->smask.p; a.rd_start = c->rd_start.p; a.rd_len = c->rd_len.p; a.has_n = c->d_has_n;
	a.sv_keys = c->sv_keys; a.sv_ids = c->sv_ids; a.sv_log2 = c->sv_log2; a.k = c->prm.kmer_len; a.mt = c->mt; a.modulo = c->prm.modulo;
	a.acc_start = c->acc_start; a.acc_n = c->acc_n; a.acc_id = c->acc_id; a.acc_cap = c->acc_cap; a.scal = c->d_scal;

	// classes by read length
	std::vector<uint32_t> cls[3];
	for (uint64_t i = 0; i < n; ++i) {
		const uint32_t len = c->h_rd_len[i];
		cls[len <= 32u * ACC_A_BM ? 0 : (len <= 32u * ACC_B_BM ? 1 : 2)].push_back((uint32_t)i);
	}
	CLB_CUDA(c, cudaFuncSetAttribute(k_accept<ACC_A_MAP, ACC_A_BM, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)acc_smem(ACC_A_MAP, ACC_A_BM)));
	CLB_CUDA(c, cudaFuncSetAttribute(k_accept<ACC_B_MAP, ACC_B_BM, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)acc_smem(ACC_B_MAP, ACC_B_BM)));

	uint32_t* d_list = nullptr;
	auto upload = [&](const std::vector<uint32_t>& v) -> clb_status {
		if (d_list) { dev_free(d_list, c->stream); d_list = nullptr; }
		CLB_CUDA(c, dev_alloc(&d_list, v.size(), c->stream));
		CLB_CUDA(c, cudaMemcpyAsync(d_list, v.data(), sizeof(uint32_t) * v.size(), cudaMemcpyHostToDevice, s));
		return CLB_OK;
	};
	clb_status st = CLB_OK;
	for (int pass = 0; pass < 3 && st == CLB_OK; ++pass) {
		std::vector<uint32_t>& v = cls[pass];
		if (v.empty()) continue;
		const bool all = pass == 0 && v.size() == n;
		if (!all) { st = upload(v); if (st != CLB_OK) break; }
		a.list = all ? nullptr : d_list; a.n_list = (uint32_t)v.size();
		scal_zero(c, SC_CURSOR2);
		uint64_t* g_map = nullptr; uint32_t* g_bits = nullptr; uint64_t* g_moff = nullptr; uint64_t* g_boff = nullptr;
		if (pass == 0) {
			CLB_TIMED(c, K_ACCEPT, (k_accept<ACC_A_MAP, ACC_A_BM, false><<<(uint32_t)v.size(), ACC_THREADS, acc_smem(ACC_A_MAP, ACC_A_BM), s>>>(a)));
		} else if (pass == 1) {
			CLB_TIMED(c, K_ACCEPT, (k_accept<ACC_B_MAP, ACC_B_BM, false><<<(uint32_t)v.size(), ACC_THREADS, acc_smem(ACC_B_MAP, ACC_B_BM), s>>>(a)));
		} else {
			// global scratch: a map of >= 2 * (#k-mers of the read) slots and the two bitmap arrays per read
			std::vector<uint64_t> moff(v.size() + 1, 0), boff(v.size() + 1, 0);
			for (size_t i = 0; i < v.size(); ++i) {
				const uint64_t len = c->h_rd_len[v[i]];
				uint64_t cap = 128; while (cap < 2 * len) cap <<= 1;
				moff[i + 1] = moff[i] + cap;
				boff[i + 1] = boff[i] + 2 * ((len + 31) / 32) + 2;
			}
			cudaError_t e = dev_alloc(&g_map, moff.back(), c->stream);
			if (e == cudaSuccess) e = dev_alloc(&g_bits, boff.back(), c->stream);
			if (e == cudaSuccess) e = dev_alloc(&g_moff, moff.size(), c->stream);
			if (e == cudaSuccess) e = dev_alloc(&g_boff, boff.size(), c->stream);
			if (e == cudaSuccess) e = cudaMemcpyAsync(g_moff, moff.data(), sizeof(uint64_t) * moff.size(), cudaMemcpyHostToDevice, s);
			if (e == cudaSuccess) e = cudaMemcpyAsync(g_boff, boff.data(), sizeof(uint64_t) * boff.size(), cudaMemcpyHostToDevice, s);
			if (e != cudaSuccess) { st = cuda_fail(c, e, "accept scratch"); }
			else {
				a.g_map = g_map; a.g_bits = g_bits; a.g_map_off = g_moff; a.g_bits_off = g_boff;
				CLB_TIMED(c, K_ACCEPT, (k_accept<128, 32, true><<<(uint32_t)v.size(), ACC_THREADS, 0, s>>>(a)));
			}
		}
		if (st == CLB_OK) {
			++c->launches;
			cudaError_t e = cudaGetLastError();
			if (e != cudaSuccess) st = cuda_fail(c, e, "k_accept");
		}
		unsigned long long sc[SC_COUNT];
		if (st == CLB_OK) st = scal_read(c, sc);
		dev_free(g_map, c->stream); dev_free(g_bits, c->stream); dev_free(g_moff, c->stream); dev_free(g_boff, c->stream);
		if (st != CLB_OK) break;
		if (sc[SC_OVERFLOW]) { st = fail(c, CLB_ERR_CUDA, "accepted k-mer arena overflow"); break; }
		if (sc[SC_CURSOR2]) {
			if (pass == 2) { st = fail(c, CLB_ERR_CUDA, "k_accept: global scratch class failed"); break; }
			std::vector<uint32_t> pend;
			st = collect_pending(c, c->acc_n, pend);
			cls[pass + 1].insert(cls[pass + 1].end(), pend.begin(), pend.end());
		}
		c->acc_total = sc[SC_CURSOR];
	}
	if (d_list) dev_free(d_list, c->stream);
	return st;
}

static clb_status run_postings(clb_ctx* c, uint32_t n_pseudo)
{
	cudaStream_t s = c->stream;
	const uint64_t ns = c->n_surv, n = c->n_reads;
	CLB_CUDA(c, dev_alloc(&c->post_cnt, ns + 1, c->stream));
	CLB_CUDA(c, dev_alloc(&c->post_off, ns + 1, c->stream));
	CLB_CUDA(c, cudaMemsetAsync(c->post_cnt, 0, sizeof(uint32_t) * (ns + 1), s));
	const uint32_t warps_grid = (uint32_t)((n * 32 + 255) / 256);
	if (n) {
		CLB_TIMED(c, K_POSTINGS, (k_post_pass<false><<<warps_grid, 256, 0, s>>>(c->acc_start, c->acc_n, c->acc_id, c->d_is_ref, c->d_ref_before, (uint32_t)n, c->post_cnt, nullptr, nullptr)));
		CLB_LAUNCH_CHECK(c, "k_post_pass<count>");
	}
	clb_status st = exclusive_scan(c, c->post_cnt, ns, c->post_off, &c->post_total);
	if (st != CLB_OK) return st;
	CLB_CUDA(c, dev_alloc(&c->post, c->post_total, c->stream));
	if (n && c->post_total) {
		CLB_CUDA(c, cudaMemsetAsync(c->post_cnt, 0, sizeof(uint32_t) * (ns + 1), s));
		CLB_TIMED(c, K_POSTINGS, (k_post_pass<true><<<warps_grid, 256, 0, s>>>(c->acc_start, c->acc_n, c->acc_id, c->d_is_ref, c->d_ref_before, (uint32_t)n, c->post_cnt, c->post_off, c->post)));
		CLB_LAUNCH_CHECK(c, "k_post_pass<fill>");
		// lists over the cap
		uint32_t* d_over = nullptr;
		CLB_CUDA(c, dev_alloc(&d_over, ns, c->stream));
		scal_zero(c, SC_CURSOR);
		k_post_oversize<<<(uint32_t)((ns + 255) / 256), 256, 0, s>>>(c->post_cnt, ns, c->prm.max_count, d_over, &c->d_scal[SC_CURSOR]);
		++c->launches;
		unsigned long long sc[SC_COUNT];
		st = scal_read(c, sc);
		if (st == CLB_OK && sc[SC_CURSOR]) {
			k_post_truncate<<<(uint32_t)sc[SC_CURSOR], 256, 0, s>>>(d_over, c->post_off, c->post_cnt, c->post, c->prm.max_count, n_pseudo);
			++c->launches;
			cudaError_t e = cudaStreamSynchronize(s);
			if (e != cudaSuccess) st = cuda_fail(c, e, "k_post_truncate");
		}
		dev_free(d_over, c->stream);
	}
	return st;
}

static clb_status run_votes(clb_ctx* c, uint32_t n_pseudo)
{
	cudaStream_t s = c->stream;
	const uint64_t n = c->n_reads; const uint32_t mc = c->prm.max_candidates;
	CLB_CUDA(c, dev_alloc(&c->cand, n * mc, c->stream));
	CLB_CUDA(c, dev_alloc(&c->cand_votes, n * mc, c->stream));
	CLB_CUDA(c, dev_alloc(&c->cand_n, n, c->stream));
	CLB_CUDA(c, cudaMemsetAsync(c->cand_n, 0, sizeof(uint32_t) * (n ? n : 1), s));
	if (!n) return CLB_OK;
	VoteArgs a{};
	a.acc_start = c->acc_start; a.acc_n = c->acc_n; a.acc_id = c->acc_id; a.ref_before = c->d_ref_before;
	a.post_cnt = c->post_cnt; a.post_off = c->post_off; a.post = c->post; a.n_pseudo = std::max<uint32_t>(n_pseudo, (uint32_t)c->n_context); a.max_cand = mc;   // neither kind is ever queried
	a.cand = c->cand; a.cand_votes = c->cand_votes; a.cand_n = c->cand_n; a.scal = c->d_scal;
	scal_zero(c, SC_CURSOR2);
	CLB_TIMED(c, K_VOTE, (k_vote<4096, false><<<(uint32_t)n, VOTE_THREADS, 0, s>>>(a)));
	CLB_LAUNCH_CHECK(c, "k_vote");
	unsigned long long sc[SC_COUNT];
	clb_status st = scal_read(c, sc);
	if (st != CLB_OK || sc[SC_CURSOR2] == 0) return st;
	// reads with more distinct neighbours than the shared table holds: global scratch sized from a bound
	std::vector<uint32_t> pend;
	st = collect_pending(c, c->cand_n, pend);
	if (st != CLB_OK) return st;
	uint32_t* d_list = nullptr; unsigned long long* d_bound = nullptr; uint32_t* g_keys = nullptr; uint64_t* g_off = nullptr;
	CLB_CUDA(c, dev_alloc(&d_list, pend.size(), c->stream));
	CLB_CUDA(c, dev_alloc(&d_bound, pend.size(), c->stream));
	CLB_CUDA(c, cudaMemcpyAsync(d_list, pend.data(), sizeof(uint32_t) * pend.size(), cudaMemcpyHostToDevice, s));
	k_vote_bound<<<(uint32_t)((pend.size() * 32 + 255) / 256), 256, 0, s>>>(d_list, (uint32_t)pend.size(), c->acc_start, c->acc_n, c->acc_id, c->post_cnt, d_bound);
	++c->launches;
	std::vector<unsigned long long> bound(pend.size());
	cudaError_t e = cudaMemcpyAsync(bound.data(), d_bound, sizeof(unsigned long long) * pend.size(), cudaMemcpyDeviceToHost, s);
	if (e == cudaSuccess) e = cudaStreamSynchronize(s);
	std::vector<uint64_t> off(pend.size() + 1, 0);
	for (size_t i = 0; i < pend.size(); ++i) {
		uint64_t cap = 8192; while (cap < 2 * bound[i]) cap <<= 1;
		if (cap > (1ULL << 31)) cap = 1ULL << 31;
		off[i + 1] = off[i] + 2 * cap;
	}
	if (e == cudaSuccess) e = dev_alloc(&g_keys, off.back(), c->stream);
	if (e == cudaSuccess) e = dev_alloc(&g_off, off.size(), c->stream);
	if (e == cudaSuccess) e = cudaMemcpyAsync(g_off, off.data(), sizeof(uint64_t) * off.size(), cudaMemcpyHostToDevice, s);
	if (e == cudaSuccess) {
		a.list = d_list; a.n_list = (uint32_t)pend.size(); a.g_keys = g_keys; a.g_off = g_off;
		scal_zero(c, SC_CURSOR2);
		CLB_TIMED(c, K_VOTE, (k_vote<1, true><<<(uint32_t)pend.size(), VOTE_THREADS, 0, s>>>(a)));
		++c->launches;
		e = cudaGetLastError();
	}
	if (e == cudaSuccess) { st = scal_read(c, sc); if (st == CLB_OK && sc[SC_CURSOR2]) st = fail(c, CLB_ERR_CUDA, "k_vote: global scratch class failed"); }
	else st = cuda_fail(c, e, "k_vote<global>");
	dev_free(d_list, c->stream); dev_free(d_bound, c->stream); dev_free(g_keys, c->stream); dev_free(g_off, c->stream);
	return st;
}

static clb_status run_common(clb_ctx* c)
{
	cudaStream_t s = c->stream;
	const uint64_t n = c->n_reads; const uint32_t mc = c->prm.max_candidates;
	if (mc > 32) return fail(c, CLB_ERR_BAD_ARG, "HiFi path supports max_candidates <= 32");
	CLB_CUDA(c, dev_alloc(&c->common_off, n * mc, c->stream));
	CLB_CUDA(c, cudaMemsetAsync(c->common_off, 0, sizeof(uint64_t) * (n * mc ? n * mc : 1), s));
	scal_zero(c, SC_CURSOR);
	if (n) { k_common_total<<<(uint32_t)((n + 255) / 256), 256, 0, s>>>(c->cand_n, c->cand_votes, (uint32_t)n, mc, &c->d_scal[SC_CURSOR]); ++c->launches; }
	unsigned long long sc[SC_COUNT];
	clb_status st = scal_read(c, sc);
	if (st != CLB_OK) return st;
	c->common_total = sc[SC_CURSOR];
	CLB_CUDA(c, dev_alloc(&c->common, c->common_total, c->stream));
	scal_zero(c, SC_CURSOR);
	if (n && c->common_total) {
		CLB_TIMED(c, K_COMMON, (k_common<<<(uint32_t)n, COMMON_THREADS, 0, s>>>(c->acc_start, c->acc_n, c->acc_id, c->d_ref_before, c->post_cnt, c->post_off, c->post,
			c->sv_kmer, c->cand, c->cand_votes, c->cand_n, mc, c->common_off, c->common, &c->d_scal[SC_CURSOR])));
		CLB_LAUNCH_CHECK(c, "k_common");
	}
	CLB_CUDA(c, cudaStreamSynchronize(s));
	return CLB_OK;
}

clb_status s1b_build(clb_ctx* c, const uint8_t* is_reference, uint32_t n_pseudo)
{
	if (!c->finalized) return fail(c, CLB_ERR_STATE, "clb_graph_build before clb_count_finalize");
	if (c->graph_done) return fail(c, CLB_ERR_STATE, "clb_graph_build called twice");
	if (c->n_reads >= (1ULL << 30)) return fail(c, CLB_ERR_BAD_ARG, "reference ids must stay below 2^30 (hm_compact.h:545-566)");
	cudaStream_t s = c->stream;
	const uint64_t n = c->n_reads;
	CLB_CUDA(c, dev_alloc(&c->d_has_n, n, c->stream));
	if (n) {
		k_read_flags<<<(uint32_t)((n * 32 + 255) / 256), 256, 0, s>>>(c->nmask.p, c->rd_start.p, c->rd_len.p, (uint32_t)n, c->d_has_n);
		CLB_LAUNCH_CHECK(c, "k_read_flags");
	}
	c->h_has_n.resize(n);
	CLB_CUDA(c, cudaMemcpyAsync(c->h_has_n.data(), c->d_has_n, n, cudaMemcpyDeviceToHost, s));
	CLB_CUDA(c, cudaStreamSynchronize(s));
	// reference ids in input order (id_in_reference, reads_sim_graph.cpp:343-347)
	c->h_is_ref.resize(n); c->h_ref_before.resize(n);
	uint32_t nref = 0;
	for (uint64_t i = 0; i < n; ++i) {
		// context reads (ids below n_context) are reference reads of earlier shards; is_reference covers the reads after them
		const bool ref = i < n_pseudo ? true : ((i < c->n_context || is_reference == nullptr || is_reference[i - c->n_context] != 0) && !c->h_has_n[i]);
		c->h_is_ref[i] = ref; c->h_ref_before[i] = nref; nref += ref;
	}
	c->n_ref = nref;
	CLB_CUDA(c, dev_alloc(&c->d_is_ref, n, c->stream));
	CLB_CUDA(c, dev_alloc(&c->d_ref_before, n, c->stream));
	CLB_CUDA(c, cudaMemcpyAsync(c->d_is_ref, c->h_is_ref.data(), n, cudaMemcpyHostToDevice, s));
	CLB_CUDA(c, cudaMemcpyAsync(c->d_ref_before, c->h_ref_before.data(), sizeof(uint32_t) * n, cudaMemcpyHostToDevice, s));

	clb_status st = run_accept(c);
	if (st == CLB_OK) st = run_postings(c, n_pseudo);
	if (st == CLB_OK) st = run_votes(c, n_pseudo);
	if (st == CLB_OK && c->prm.is_hifi) st = run_common(c);
	if (st == CLB_OK) { CLB_CUDA(c, cudaStreamSynchronize(s)); c->graph_done = true; }
	return st;
}

// has-N flag of every read in the store (context reads included), HOST buffer
clb_status s1b_reads_have_n(clb_ctx* c, uint8_t* flags)
{
	cudaStream_t s = c->stream;
	const uint64_t n = c->n_reads;
	if (!n) return CLB_OK;
	uint8_t* d = nullptr;
	CLB_CUDA(c, dev_malloc((void**)&d, n, s));
	k_read_flags<<<(uint32_t)((n * 32 + 255) / 256), 256, 0, s>>>(c->nmask.p, c->rd_start.p, c->rd_len.p, (uint32_t)n, d);
	CLB_LAUNCH_CHECK(c, "k_read_flags");
	CLB_CUDA(c, cudaMemcpyAsync(flags, d, n, cudaMemcpyDeviceToHost, s));
	CLB_CUDA(c, dev_free_async(d, s));
	CLB_CUDA(c, cudaStreamSynchronize(s));
	return CLB_OK;
}

// ASCII bases of the listed reads back to back (one CTA per read)
__global__ void __launch_bounds__(256) k_reads_export(const uint64_t* __restrict__ pk, const uint32_t* __restrict__ nmask, const uint64_t* __restrict__ rd_start,
	const uint32_t* __restrict__ rd_len, const uint32_t* __restrict__ ids, const uint64_t* __restrict__ out_off, uint8_t* __restrict__ out)
{
	const uint32_t r = ids[blockIdx.x];
	const uint64_t s0 = rd_start[r]; const uint32_t len = rd_len[r];
	uint8_t* o = out + out_off[blockIdx.x];
	for (uint32_t j = threadIdx.x; j < len; j += blockDim.x) {
		const uint64_t p = s0 + j;
		o[j] = ((nmask[p >> 5] >> (p & 31)) & 1u) ? (uint8_t)'N' : (uint8_t)"ACGT"[base_at(pk, p)];
	}
}
clb_status s1b_reads_export(clb_ctx* c, const uint32_t* read_ids, uint32_t n, uint8_t* bases, uint64_t cap, int on_device)
{
	cudaStream_t s = c->stream;
	if (!n) return CLB_OK;
	std::vector<uint64_t> off(n + 1, 0);
	for (uint32_t i = 0; i < n; ++i) {
		if (read_ids[i] >= c->n_reads) return fail(c, CLB_ERR_BAD_ARG, "clb_reads_export: read id out of range");
		off[i + 1] = off[i] + c->h_rd_len[read_ids[i]];
	}
	if (off[n] > cap) return fail(c, CLB_ERR_CAPACITY, "clb_reads_export: buffer too small");
	uint32_t* d_ids = nullptr; uint64_t* d_off = nullptr; uint8_t* d_out = bases;
	CLB_CUDA(c, dev_malloc((void**)&d_ids, sizeof(uint32_t) * n, s)); CLB_CUDA(c, dev_malloc((void**)&d_off, sizeof(uint64_t) * (n + 1), s));
	if (!on_device) CLB_CUDA(c, dev_malloc((void**)&d_out, off[n] + 1, s));
	CLB_CUDA(c, cudaMemcpyAsync(d_ids, read_ids, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, s));
	CLB_CUDA(c, cudaMemcpyAsync(d_off, off.data(), sizeof(uint64_t) * (n + 1), cudaMemcpyHostToDevice, s));
	k_reads_export<<<n, 256, 0, s>>>(c->pk.p, c->nmask.p, c->rd_start.p, c->rd_len.p, d_ids, d_off, d_out);
	CLB_LAUNCH_CHECK(c, "k_reads_export");
	if (!on_device) { CLB_CUDA(c, cudaMemcpyAsync(bases, d_out, off[n], cudaMemcpyDeviceToHost, s)); CLB_CUDA(c, dev_free_async(d_out, s)); }
	CLB_CUDA(c, dev_free_async(d_ids, s)); CLB_CUDA(c, dev_free_async(d_off, s));
	CLB_CUDA(c, cudaStreamSynchronize(s));
	return CLB_OK;
}

void s1_free(clb_ctx* c)
{
	c->pk.release(); c->nmask.release(); c->smask.release(); c->rd_start.release(); c->rd_len.release();
	c->stage_in[0].release(); c->stage_in[1].release(); c->stage_off.release();
	prof_resolve(c);
	if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
	for (int b = 0; b < 2; ++b) { if (c->ev_copied[b]) cudaEventDestroy(c->ev_copied[b]); if (c->ev_consumed[b]) cudaEventDestroy(c->ev_consumed[b]); }
	void* ptrs[] = { c->tab, c->d_scal, c->sv_keys, c->sv_ids, c->sv_kmer, c->sv_count, c->d_has_n, c->d_ref_before, c->d_is_ref,
		c->acc_start, c->acc_n, c->acc_id, c->post_cnt, c->post_off, c->post, c->cand, c->cand_votes, c->cand_n, c->common_off, c->common };
	for (void* p : ptrs) if (p) dev_free(p, c->stream);
}

} // namespace clb
