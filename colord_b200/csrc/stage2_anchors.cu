// stage2_anchors.cu — rows E1-E4 of SURVEY.md §8: m-mer anchors of every read against its candidate reference reads.
//
// Reference (results restated, data structures not): CMmersHashMapDuplicateOptimizedLP + CBloomFilter of the read being
// encoded (encoder.cpp:326-352, encoder.h:72-238), scan of each oriented reference read (encoder.cpp:354-390), intersection
// and the match-count cap (:392-492, :684-695, :1016-1056), LIS over reference positions (:617-661, utils.cpp:157-209),
// MergeAnchors (:731-776), orientation choice (:1149-1192), overlap fix (:1577-1622) and the sort by total anchor length.
//
// Kernels
//   k_anchor_match   one CTA per read: "position table" of the read's m-mers (a cell holds a read position; the key is the
//                    m-mer read back from the packed read, so a cell is 4 bytes and an 8 kb read's table fits in 64 KB of
//                    shared memory) + a 1-hash Bloom filter; both strands of every candidate are scanned in one pass over
//                    its packed words (the reverse-complement read's m-mer at rl-m-p is the complement of the forward
//                    m-mer at p); pairs (read position, reference position) are counted, the CTA reserves exactly that many
//                    slots of the pair arena with one atomic, and a second scan writes them.
//   k_pairs_sort     one CTA per segment: bitonic sort of the pairs by (read position asc, reference position desc) — the
//                    order in which the reference feeds its LIS — in shared memory (<= 2048 pairs) or in place.
//   k_lis            one thread per segment: patience LIS with the reference's tie-breaking, chain read-back, anchor merge.
//   k_select         one thread per read: orientation per candidate, overlap fix, stable sort by total anchor length.
#include "ctx.h"
#include "stage2.h"
#include <algorithm>
#include <vector>

namespace clb {

constexpr int MATCH_THREADS = 512;      // 2 CTAs per SM (shared memory): 32 resident warps behind the probes' shared-memory latency
constexpr uint32_t SMEM_TAB_CELLS = 16384;      // 64 KB: reads up to 8192 m-mers keep the table in shared memory
constexpr uint32_t SMEM_BLOOM_WORDS = 4096;     // 16 KB = 128 Ki bits
constexpr uint32_t MAX_C = 32;
constexpr uint32_t SMEM_HIT_WORDS = 2048;       // 16 KB: which (position, strand) probes of a 32-position item found the m-mer in the read's table

CLB_D uint64_t window(const uint64_t* __restrict__ pk, uint64_t p, uint32_t m)
{
	const uint64_t w = p >> 5; const uint32_t s = 2 * (uint32_t)(p & 31);
	const uint64_t hi = pk[w], lo = pk[w + 1];
	const uint64_t x = s ? ((hi << s) | (lo >> (64 - s))) : hi;
	return x >> (64 - 2 * m);
}
// ---- TMA: the packed words of the read being encoded are staged in shared memory by one bulk copy (cp.async.bulk, completion on an
// mbarrier): every hit of the position table re-reads the read's own m-mer to verify the key (tab_matches), a gather that otherwise
// goes to L1 / L2 for every probe.  1-D bulk copies need 16-byte aligned addresses and sizes.
constexpr uint32_t TILE_WORDS = 1024;           // 8 KB: reads up to 32 k bases
CLB_D uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
CLB_D void tma_stage_words(uint64_t* s_dst, const uint64_t* g_src, uint32_t bytes, uint64_t* bar)
{
	if (threadIdx.x == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(bar)) : "memory");
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
		asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
			:: "r"(smem_u32(s_dst)), "l"(g_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
	}
	uint32_t done = 0;
	while (!done)
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(0u) : "memory");
}

CLB_D uint32_t mm_hash(uint64_t x)
{
	return (uint32_t)((x * 0x9E3779B97F4A7C15ULL) >> 32);      // one multiply: the top half mixes every base of the m-mer (table index = low bits of it, Bloom bit from bit 7 up)
}

struct MatchArgs {
	const uint64_t* pk; const uint64_t* rd_start; const uint32_t* rd_len;
	const uint32_t* enc_list; uint32_t n_list;
	const uint32_t* cand; const uint32_t* cand_n; const uint32_t* ref_to_read;
	S2P P;
	const uint64_t* tab_off; uint32_t* g_tab;        // table cells for reads that do not fit in shared memory
	const uint64_t* bloom_off; uint32_t* g_bloom;
	uint8_t* arena; unsigned long long arena_cap; unsigned long long* cursor;     // in pair slots
	SegInfo* seg; uint32_t* slot_dec;                 // per slot: 1 = too few distinct m-mers (no candidates)
	const uint8_t* skip;                              // HiFi: per (slot, candidate) 1 = anchors already found from the shared k-mers
	int tma;                                          // stage the read's packed words in shared memory by a bulk copy
};

struct MatchShared {
	unsigned long long pairs[2 * MAX_C];
	unsigned long long base[2 * MAX_C];
	uint32_t hits[2 * MAX_C];
	uint32_t fill[2 * MAX_C];
	uint32_t word_base[MAX_C + 1];
	uint32_t ref_read[MAX_C];
	uint32_t uniq;
	int decision;
	uint32_t scan[33];
};

// matches of m-mer x in the position table: calls f(pos) for each read position holding x
template <class F>
CLB_D void tab_matches(const uint32_t* tab, uint32_t mask, const uint64_t* __restrict__ pk, uint64_t estart, uint32_t m, uint64_t x, uint32_t h, F f)
{
	uint32_t i = h & mask;
	for (;;) {
		const uint32_t v = tab[i];
		if (v == 0) break;
		if (window(pk, estart + (v - 1), m) == x) f(v - 1);
		i = (i + 1) & mask;
	}
}

// The candidates' m-mers are scanned twice: the counting pass (match caps, arena slots) and the pass that writes the pairs.  The counting
// pass leaves one bit per probe that hit (hit[item], bit 2 * (position in the item) + strand; a few percent of the probes), and the
// writing pass visits only those — when the items fit the mask (hit != nullptr), else it repeats the whole scan.
template <bool EMIT>
__device__ void scan_refs(const MatchArgs& a, MatchShared& sh, const uint32_t* tab, uint32_t tmask, const uint32_t* bloom, uint32_t bmask,
	const uint64_t* epk, uint64_t estart, uint32_t n_cand, uint64_t* hit)
{
	const uint32_t m = a.P.m;
	const uint64_t mmask = m == 32 ? ~0ULL : ((1ULL << (2 * m)) - 1);
	const uint32_t n_items = sh.word_base[n_cand];
	if (EMIT && hit) {
		for (uint32_t it = threadIdx.x; it < n_items; it += blockDim.x) {
			uint64_t hm = hit[it];
			if (!hm) continue;
			uint32_t j = 0;
			while (sh.word_base[j + 1] <= it) ++j;
			const uint32_t rr = sh.ref_read[j];
			const uint64_t rstart = a.rd_start[rr]; const uint32_t rl = a.rd_len[rr];
			const uint32_t p0 = (it - sh.word_base[j]) * 32;
			for (; hm; hm &= hm - 1) {
				const uint32_t bit = (uint32_t)__ffsll((long long)hm) - 1, p = p0 + (bit >> 1), o = bit & 1;
				const unsigned long long b = sh.base[2 * j + o];
				if (b == ~0ULL) continue;
				const uint64_t f = window(a.pk, rstart + p, m);
				const uint64_t x = o ? revcomp(f, m) : f;
				const uint32_t pos_o = o ? (rl - m - p) : p;
				uint64_t* out = reinterpret_cast<uint64_t*>(a.arena + b * PAIR_SLOT_BYTES);
				tab_matches(tab, tmask, epk, estart, m, x, mm_hash(x), [&](uint32_t e) {
					const uint32_t idx = atomicAdd(&sh.fill[2 * j + o], 1u);
					out[idx] = ((uint64_t)e << 32) | (uint64_t)(0xFFFFFFFFu - pos_o);
				});
			}
		}
		return;
	}
	for (uint32_t it = threadIdx.x; it < n_items; it += blockDim.x) {
		uint32_t j = 0;
		while (sh.word_base[j + 1] <= it) ++j;
		const uint32_t w = it - sh.word_base[j];
		const uint32_t rr = sh.ref_read[j];
		const uint64_t rstart = a.rd_start[rr]; const uint32_t rl = a.rd_len[rr];
		const uint32_t n_pos = rl - m + 1;                  // rl >= m guaranteed by word_base
		const uint32_t p0 = w * 32, p1 = min(p0 + 32, n_pos);
		if (EMIT && sh.base[2 * j] == ~0ULL && sh.base[2 * j + 1] == ~0ULL) continue;
		uint64_t f = window(a.pk, rstart + p0, m);
		uint32_t hits_f = 0, hits_r = 0; unsigned long long pairs_f = 0, pairs_r = 0;
		uint64_t hm = 0;
		for (uint32_t p = p0; p < p1; ++p) {
			if (p > p0) f = ((f << 2) | base_at(a.pk, rstart + p + m - 1)) & mmask;
			const uint64_t r = revcomp(f, m);
#pragma unroll
			for (int o = 0; o < 2; ++o) {
				const uint64_t x = o ? r : f;
				const uint32_t h = mm_hash(x);
				const uint32_t bb = (h >> 7) & bmask;
				if (!((bloom[bb >> 5] >> (bb & 31)) & 1u)) continue;
				const uint32_t pos_o = o ? (rl - m - p) : p;
				if (EMIT) {
					const unsigned long long b = sh.base[2 * j + o];
					if (b == ~0ULL) continue;
					uint64_t* out = reinterpret_cast<uint64_t*>(a.arena + b * PAIR_SLOT_BYTES);
					tab_matches(tab, tmask, epk, estart, m, x, h, [&](uint32_t e) {
						const uint32_t idx = atomicAdd(&sh.fill[2 * j + o], 1u);
						out[idx] = ((uint64_t)e << 32) | (uint64_t)(0xFFFFFFFFu - pos_o);
					});
				} else {
					uint32_t cnt = 0;
					tab_matches(tab, tmask, epk, estart, m, x, h, [&](uint32_t) { ++cnt; });
					if (cnt) { hm |= 1ull << (2 * (p - p0) + o); if (o) { ++hits_r; pairs_r += cnt; } else { ++hits_f; pairs_f += cnt; } }
				}
			}
		}
		if (!EMIT) {
			if (hit) hit[it] = hm;
			if (hits_f) { atomicAdd(&sh.hits[2 * j], hits_f); atomicAdd(&sh.pairs[2 * j], pairs_f); }
			if (hits_r) { atomicAdd(&sh.hits[2 * j + 1], hits_r); atomicAdd(&sh.pairs[2 * j + 1], pairs_r); }
		}
	}
}

__global__ void __launch_bounds__(MATCH_THREADS) k_anchor_match(MatchArgs a)
{
	extern __shared__ __align__(16) uint8_t smem_raw[];
	MatchShared& sh = *reinterpret_cast<MatchShared*>(smem_raw);
	uint32_t* s_tab = reinterpret_cast<uint32_t*>(smem_raw + ((sizeof(MatchShared) + 15) & ~15ull));
	uint32_t* s_bloom = s_tab + SMEM_TAB_CELLS;
	uint64_t* s_pk = reinterpret_cast<uint64_t*>(s_bloom + SMEM_BLOOM_WORDS);      // TILE_WORDS + 2 words, then the mbarrier
	uint64_t* s_bar = s_pk + TILE_WORDS + 2;
	uint64_t* s_hit = s_bar + 2;

	const uint32_t slot = blockIdx.x;
	const uint32_t read = a.enc_list[slot];
	const uint32_t c = a.P.c, m = a.P.m;
	const uint64_t estart_g = a.rd_start[read]; const uint32_t elen = a.rd_len[read];
	const uint32_t n_cand = min(a.cand_n[read], c);
	SegInfo* seg = a.seg + (size_t)slot * c * 2;
	for (uint32_t i = threadIdx.x; i < 2 * c; i += blockDim.x) seg[i] = SegInfo{0, 0, 0, 0, 0};
	if (elen < m) { if (threadIdx.x == 0) a.slot_dec[slot] = 1; return; }          // no m-mers: 0 < frac * len -> refused
	// the read's packed words [w0, w0 + n_w): w0 even (16-byte aligned source), one word behind the last base (window() reads two words)
	const uint64_t w0 = (estart_g >> 5) & ~1ull;
	const uint32_t n_w = (uint32_t)((((estart_g + elen - 1) >> 5) + 2 - w0 + 1) & ~1ull);
	const bool staged = a.tma && n_w <= TILE_WORDS + 2;
	if (staged) tma_stage_words(s_pk, a.pk + w0, n_w * 8, s_bar);
	const uint64_t* epk = staged ? s_pk : a.pk;
	const uint64_t estart = staged ? estart_g - 32 * w0 : estart_g;
	const uint32_t n_mm = elen - m + 1;
	uint32_t cap = 64; while (cap < 2 * n_mm) cap <<= 1;
	uint32_t bbits = 1024; while (bbits < 16 * n_mm && bbits < (1u << 30)) bbits <<= 1;
	uint32_t* tab = cap <= SMEM_TAB_CELLS ? s_tab : a.g_tab + a.tab_off[slot];
	uint32_t* bloom = bbits <= SMEM_BLOOM_WORDS * 32 ? s_bloom : a.g_bloom + a.bloom_off[slot];
	const uint32_t tmask = cap - 1, bmask = bbits - 1;
	for (uint32_t i = threadIdx.x; i < cap; i += blockDim.x) tab[i] = 0;
	for (uint32_t i = threadIdx.x; i < bbits / 32; i += blockDim.x) bloom[i] = 0;
	if (threadIdx.x < 2 * MAX_C) { sh.pairs[threadIdx.x] = 0; sh.hits[threadIdx.x] = 0; sh.fill[threadIdx.x] = 0; sh.base[threadIdx.x] = ~0ULL; }
	if (threadIdx.x == 0) {
		sh.uniq = 0;
		uint32_t wb = 0;
		for (uint32_t j = 0; j < n_cand; ++j) {
			const uint32_t rr = a.ref_to_read[a.cand[(size_t)read * c + j]];
			sh.ref_read[j] = rr; sh.word_base[j] = wb;
			const uint32_t rl = a.rd_len[rr];
			if (rl >= m && !(a.skip && a.skip[(size_t)slot * c + j])) wb += (rl - m + 1 + 31) / 32;
		}
		sh.word_base[n_cand] = wb;
	}
	__syncthreads();
	const uint64_t mmask = m == 32 ? ~0ULL : ((1ULL << (2 * m)) - 1);
	// ---- insert every m-mer of the read (duplicates take separate cells) ----
	const uint32_t n_words = (n_mm + 31) / 32;
	for (uint32_t w = threadIdx.x; w < n_words; w += blockDim.x) {
		const uint32_t p0 = w * 32, p1 = min(p0 + 32, n_mm);
		uint64_t f = window(epk, estart + p0, m);
		for (uint32_t p = p0; p < p1; ++p) {
			if (p > p0) f = ((f << 2) | base_at(epk, estart + p + m - 1)) & mmask;
			const uint32_t h = mm_hash(f);
			const uint32_t bb = (h >> 7) & bmask;
			atomicOr(&bloom[bb >> 5], 1u << (bb & 31));
			uint32_t i = h & tmask;
			while (atomicCAS(&tab[i], 0u, p + 1) != 0u) i = (i + 1) & tmask;
		}
	}
	__syncthreads();
	// ---- distinct m-mers: a position counts iff it is the smallest one holding its m-mer (encoder.cpp:1069-1079) ----
	uint32_t uq = 0;
	for (uint32_t w = threadIdx.x; w < n_words; w += blockDim.x) {
		const uint32_t p0 = w * 32, p1 = min(p0 + 32, n_mm);
		uint64_t f = window(epk, estart + p0, m);
		for (uint32_t p = p0; p < p1; ++p) {
			if (p > p0) f = ((f << 2) | base_at(epk, estart + p + m - 1)) & mmask;
			uint32_t mn = 0xFFFFFFFFu;
			tab_matches(tab, tmask, epk, estart, m, f, mm_hash(f), [&](uint32_t e) { mn = min(mn, e); });
			uq += mn == p;
		}
	}
	for (int d = 16; d; d >>= 1) uq += __shfl_xor_sync(0xffffffffu, uq, d);
	if ((threadIdx.x & 31) == 0 && uq) atomicAdd(&sh.uniq, uq);
	__syncthreads();
	if (threadIdx.x == 0) {
		int dec = -1;
		if ((double)sh.uniq > a.P.min_force * (double)elen) dec = 0;
		else if ((double)sh.uniq < a.P.min_frac * (double)elen) dec = 1;
		sh.decision = dec;
		a.slot_dec[slot] = dec == 1;
	}
	__syncthreads();
	if (sh.decision == 1) return;
	// ---- count ----
	uint64_t* hit = sh.word_base[n_cand] <= SMEM_HIT_WORDS ? s_hit : nullptr;
	scan_refs<false>(a, sh, tab, tmask, bloom, bmask, epk, estart, n_cand, hit);
	__syncthreads();
	if (threadIdx.x == 0) {
		unsigned long long total = 0;
		for (uint32_t s = 0; s < 2 * n_cand; ++s) {
			unsigned long long n = sh.pairs[s];
			// encoder.cpp:1030-1040: refuse when the number of (read, reference) m-mer matches exceeds maxMatchesMultiplier * |read|
			if (sh.hits[s] == 0) n = 0;
			else if (sh.decision != 0 && (double)n > a.P.max_mult * (double)(elen + 1)) { n = 0; seg[s].pad = SEG_TOO_MANY_MATCHES; }
			if (n > 0x7FFFFFF0ull) n = 0x7FFFFFF0ull + 2;      // cannot be stored: reported as an overflow below
			sh.pairs[s] = n;
			total += (n + 1) & ~1ull;
		}
		unsigned long long b = total ? atomicAdd(a.cursor, total) : 0;
		const bool fits = b + total <= a.arena_cap;
		for (uint32_t s = 0; s < 2 * n_cand; ++s) {
			const unsigned long long n = sh.pairs[s];
			if (n && fits && n <= 0x7FFFFFF0ull) { sh.base[s] = b; seg[s].off = b; seg[s].n = (uint32_t)n; }
			else sh.base[s] = ~0ULL;
			if (n > 0x7FFFFFF0ull) atomicAdd(a.cursor, 1ull << 62);
			b += (n + 1) & ~1ull;
		}
	}
	__syncthreads();
	// ---- write the pairs ----
	scan_refs<true>(a, sh, tab, tmask, bloom, bmask, epk, estart, n_cand, hit);
}

// ------------------------------------------------------------------------------------------------ HiFi: anchors from shared k-mers
// encoder.cpp:870-1012 (AnalyseRefReadWithKmers) + :1113-1147 (KmerBasedAnchors): a shared k-mer (or its reverse complement)
// that occurs exactly once as a forward k-mer in the read and exactly once in the oriented reference read is an anchor; the
// anchors must be colinear, overlapping ones are dropped, then they are extended over equal bases and merged.
// One CTA per (read, candidate): a small table of the shared k-mers' two strands collects (count, position) from one scan of
// the read and one scan of the forward reference (the reverse-complement read's k-mer at rl-k-p is the complement of the
// forward one at p); two threads then build the anchors of the two orientations.
struct KEntry { unsigned long long key; uint32_t enc_cnt, enc_pos, ref_cnt, ref_pos; };
struct KmerArgs {
	const uint64_t* pk; const uint64_t* rd_start; const uint32_t* rd_len;
	const uint32_t* enc_list; uint32_t n_list;
	const uint32_t* cand; const uint32_t* cand_n; const uint32_t* common_n; const uint64_t* common_off; const uint64_t* common; const uint32_t* ref_to_read;
	uint32_t k, c;
	const uint64_t* tab_off; KEntry* tab;               // per (slot, candidate): first entry / capacity is 4 * common_n rounded up to a power of two
	const uint64_t* anc_off; uint8_t* anc;              // per (slot, candidate): pair-slot offset of 2 anchor lists (one per orientation)
	SegInfo* kseg; uint8_t* skip;
};
CLB_D KEntry* ktab_find(KEntry* tab, uint32_t mask, uint64_t key)
{
	uint32_t i = mm_hash(key) & mask;
	for (;;) { if (tab[i].key == key) return &tab[i]; if (tab[i].key == ~0ULL) return nullptr; i = (i + 1) & mask; }
}
__global__ void __launch_bounds__(128) k_kmer_anchors(KmerArgs a)
{
	const uint32_t slot = blockIdx.x / a.c, j = blockIdx.x % a.c;
	const uint32_t read = a.enc_list[slot];
	const size_t sc = (size_t)slot * a.c + j;
	SegInfo* ks = a.kseg + sc * 2;
	if (threadIdx.x < 2) ks[threadIdx.x] = SegInfo{0, 0, 0, 0, 1};
	if (threadIdx.x == 0) a.skip[sc] = 0;
	if (j >= min(a.cand_n[read], a.c)) return;
	const uint32_t nco = a.common_n[(size_t)read * a.c + j], k = a.k;
	if (!nco) return;
	const uint64_t* common = a.common + a.common_off[(size_t)read * a.c + j];
	uint32_t cap = 8; while (cap < 4 * nco) cap <<= 1;
	const uint32_t mask = cap - 1;
	KEntry* tab = a.tab + a.tab_off[sc];
	for (uint32_t i = threadIdx.x; i < cap; i += blockDim.x) tab[i] = KEntry{~0ULL, 0, 0, 0, 0};
	__syncthreads();
	for (uint32_t i = threadIdx.x; i < 2 * nco; i += blockDim.x) {
		const uint64_t km = common[i >> 1], key = (i & 1) ? revcomp(km, k) : km;
		uint32_t h = mm_hash(key) & mask;
		for (;;) {
			const unsigned long long old = atomicCAS(&tab[h].key, ~0ULL, (unsigned long long)key);
			if (old == ~0ULL || old == key) break;
			h = (h + 1) & mask;
		}
	}
	__syncthreads();
	const uint64_t kmask = k == 32 ? ~0ULL : ((1ULL << (2 * k)) - 1);
	const uint32_t rr = a.ref_to_read[a.cand[(size_t)read * a.c + j]];
	const uint64_t estart = a.rd_start[read], rstart = a.rd_start[rr];
	const uint32_t el = a.rd_len[read], rl = a.rd_len[rr];
	for (int which = 0; which < 2; ++which) {
		const uint64_t st = which ? rstart : estart; const uint32_t len = which ? rl : el;
		if (len < k) continue;
		const uint32_t n_pos = len - k + 1, n_words = (n_pos + 31) / 32;
		for (uint32_t w = threadIdx.x; w < n_words; w += blockDim.x) {
			const uint32_t p0 = w * 32, p1 = min(p0 + 32, n_pos);
			uint64_t f = window(a.pk, st + p0, k);
			for (uint32_t p = p0; p < p1; ++p) {
				if (p > p0) f = ((f << 2) | base_at(a.pk, st + p + k - 1)) & kmask;
				KEntry* e = ktab_find(tab, mask, f);
				if (!e) continue;
				if (which) { atomicAdd(&e->ref_cnt, 1u); e->ref_pos = p; } else { atomicAdd(&e->enc_cnt, 1u); e->enc_pos = p; }
			}
		}
	}
	__syncthreads();
	if (threadIdx.x >= 2) return;
	const uint32_t o = threadIdx.x;                       // 0: forward reference, 1: reverse complement
	const uint64_t seg_slot = a.anc_off[sc] + (uint64_t)o * ((12ull * nco + PAIR_SLOT_BYTES - 1) / PAIR_SLOT_BYTES + 1);
	Anchor* an = reinterpret_cast<Anchor*>(a.anc + seg_slot * PAIR_SLOT_BYTES);
	uint32_t n = 0;
	for (uint32_t i = 0; i < nco; ++i) {
		for (int form = 0; form < 2; ++form) {
			const uint64_t F = form ? revcomp(common[i], k) : common[i];
			const KEntry* e = ktab_find(tab, mask, F);
			if (!e || e->enc_cnt != 1) continue;
			uint32_t ir;
			if (!o) { if (e->ref_cnt != 1) continue; ir = e->ref_pos; }
			else { const KEntry* e2 = ktab_find(tab, mask, revcomp(F, k)); if (!e2 || e2->ref_cnt != 1) continue; ir = rl - k - e2->ref_pos; }
			// insertion by pos_enc (positions are distinct)
			uint32_t q = n++;
			while (q > 0 && an[q - 1].pos_enc > e->enc_pos) { an[q] = an[q - 1]; --q; }
			an[q] = Anchor{k, e->enc_pos, ir};
			break;
		}
	}
	if (!n) return;
	for (uint32_t i = 1; i < n; ++i) if (an[i].pos_ref < an[i - 1].pos_ref) return;          // not colinear
	{	// drop k-mers overlapping their predecessor (:917-926)
		uint32_t w = 1;
		for (uint32_t i = 1; i < n; ++i) { const Anchor& p = an[w - 1]; if (p.pos_enc + p.len > an[i].pos_enc || p.pos_ref + p.len > an[i].pos_ref) continue; an[w++] = an[i]; }
		n = w;
	}
	auto eb = [&](uint32_t x) { return base_at(a.pk, estart + x); };
	auto rb = [&](uint32_t y) { return o ? 3u - base_at(a.pk, rstart + (rl - 1 - y)) : base_at(a.pk, rstart + y); };
	while (an[0].pos_enc > 0 && an[0].pos_ref > 0 && eb(an[0].pos_enc - 1) == rb(an[0].pos_ref - 1)) { --an[0].pos_enc; --an[0].pos_ref; ++an[0].len; }
	// extend / merge (:944-998), the reference's loop on a vector with erase, restated literally
	for (unsigned long long i = 0; i < n; ++i) {
		if (i > 0) {
			const uint32_t pe = an[i - 1].pos_enc + an[i - 1].len, pr = an[i - 1].pos_ref + an[i - 1].len;
			for (;;) {
				const bool re = an[i].pos_enc == pe, rr2 = an[i].pos_ref == pr;
				if (re && rr2) { an[i].len += an[i - 1].len; for (uint32_t x = (uint32_t)i - 1; x + 1 < n; ++x) an[x] = an[x + 1]; --n; break; }
				if (re || rr2) break;
				if (eb(an[i].pos_enc - 1) != rb(an[i].pos_ref - 1)) break;
				an[i].len++; an[i].pos_enc--; an[i].pos_ref--;
			}
		}
		if (i >= n) break;
		if (i != (unsigned long long)n - 1) {
			const uint32_t ne = an[i + 1].pos_enc, nr = an[i + 1].pos_ref;
			uint32_t pe = an[i].pos_enc + an[i].len, pr = an[i].pos_ref + an[i].len;
			for (;;) {
				const bool re = pe == ne, rr2 = pr == nr;
				if (re && rr2) { an[i].len += an[i + 1].len; for (uint32_t x = (uint32_t)i + 1; x + 1 < n; ++x) an[x] = an[x + 1]; --n; --i; break; }
				else if (re || rr2) break;
				if (eb(pe) != rb(pr)) break;
				++pe; ++pr; ++an[i].len;
			}
		}
	}
	{
		Anchor& l = an[n - 1];
		uint32_t pe = l.pos_enc + l.len, pr = l.pos_ref + l.len;
		while (pe < el && pr < rl && eb(pe) == rb(pr)) { ++pe; ++pr; ++l.len; }
	}
	uint32_t tot = 0; for (uint32_t i = 0; i < n; ++i) tot += an[i].len;
	ks[o] = SegInfo{seg_slot, 1, n, tot, 1};
	a.skip[sc] = 1;
}

// ------------------------------------------------------------------------------------------------ sort
constexpr int SORT_THREADS = 128;
constexpr uint32_t SORT_SMEM = 2048;

__device__ __forceinline__ void cmpx(uint64_t* a, uint32_t i, uint32_t j)
{
	const uint64_t x = a[i], y = a[j];
	if (x > y) { a[i] = y; a[j] = x; }
}
// ascending bitonic network on a[0..n) with virtual +inf padding up to the next power of two
__device__ void bitonic_sort(uint64_t* a, uint32_t n)
{
	uint32_t np = 1; while (np < n) np <<= 1;
	for (uint32_t k = 2; k <= np; k <<= 1) {
		const uint32_t half = k >> 1;
		for (uint32_t t = threadIdx.x; t < np / 2; t += blockDim.x) {
			const uint32_t blk = t / half, off = t % half;
			const uint32_t i = blk * k + off, j = blk * k + (k - 1 - off);
			if (j < n) cmpx(a, i, j);
		}
		__syncthreads();
		for (uint32_t jj = half >> 1; jj > 0; jj >>= 1) {
			for (uint32_t t = threadIdx.x; t < np / 2; t += blockDim.x) {
				const uint32_t i = (t / jj) * 2 * jj + (t % jj), j = i + jj;
				if (j < n) cmpx(a, i, j);
			}
			__syncthreads();
		}
	}
}

__global__ void __launch_bounds__(SORT_THREADS) k_pairs_sort(const SegInfo* __restrict__ seg, uint32_t n_seg, uint8_t* __restrict__ arena)
{
	__shared__ uint64_t s_keys[SORT_SMEM];
	const uint32_t s = blockIdx.x;
	if (s >= n_seg) return;
	const uint32_t n = seg[s].n;
	if (n < 2) return;
	uint64_t* keys = reinterpret_cast<uint64_t*>(arena + seg[s].off * PAIR_SLOT_BYTES);
	if (n <= SORT_SMEM) {
		for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) s_keys[i] = keys[i];
		__syncthreads();
		bitonic_sort(s_keys, n);
		for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) keys[i] = s_keys[i];
	} else {
		__syncthreads();
		bitonic_sort(keys, n);
	}
}

// ------------------------------------------------------------------------------------------------ LIS + merge
// utils.cpp:157-209 (LIS: patience with lower_bound, predecessor links, chain read back from the last pile's top),
// encoder.cpp:646-658 (read positions recovered by forward scans), :731-776 (MergeAnchors).
__global__ void __launch_bounds__(128) k_lis(SegInfo* __restrict__ seg, uint32_t n_seg, uint8_t* __restrict__ arena, uint32_t m)
{
	const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= n_seg) return;
	const uint32_t n = seg[s].n;
	if (n == 0) return;
	uint8_t* base = arena + seg[s].off * PAIR_SLOT_BYTES;
	const uint64_t* keys = reinterpret_cast<const uint64_t*>(base);
	uint32_t* tv = reinterpret_cast<uint32_t*>(base + 8ull * n);
	uint32_t* ti = reinterpret_cast<uint32_t*>(base + 12ull * n);
	uint32_t* pred = reinterpret_cast<uint32_t*>(base + 16ull * n);
	auto refpos = [&](uint32_t i) { return 0xFFFFFFFFu - (uint32_t)keys[i]; };
	uint32_t len = 1; tv[0] = refpos(0); ti[0] = 0; pred[0] = 0xFFFFFFFFu;
	for (uint32_t i = 1; i < n; ++i) {
		const uint32_t x = refpos(i);
		uint32_t pos;
		if (tv[len - 1] < x) pos = len;
		else { uint32_t lo = 0, hi = len; while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (x > tv[mid]) lo = mid + 1; else hi = mid; } pos = lo; }
		if (pos == len) ++len;
		tv[pos] = x; ti[pos] = i;
		pred[i] = pos > 0 ? ti[pos - 1] : 0xFFFFFFFFu;
	}
	// chain (pair indices) in forward order -> tv
	{
		uint32_t cur = ti[len - 1];
		for (uint32_t i = len; i-- > 0;) { tv[i] = cur; cur = pred[cur]; }
	}
	// (read position, reference position) of every chain element: the read position is the first pair at or after the
	// scan pointer whose reference position equals the chain element's (encoder.cpp:646-658)
	{
		uint32_t ptr = 0;
		for (uint32_t i = 0; i < len; ++i) {
			const uint32_t p = refpos(tv[i]);
			while (refpos(ptr) != p) ++ptr;
			ti[i] = (uint32_t)(keys[ptr] >> 32); pred[i] = p;
			++ptr;
		}
	}
	// merge runs of (+1, +1) steps; the pair keys are dead now, anchors take their place
	Anchor* out = reinterpret_cast<Anchor*>(base);
	uint32_t n_res = 0, tot = 0, start = 0;
	for (uint32_t i = 1; i <= len; ++i) {
		if (i == len || ti[i - 1] != ti[i] - 1 || pred[i - 1] != pred[i] - 1) {
			const uint32_t l = (i - start) + m - 1;
			const Anchor a{l, ti[start], pred[start]};
			out[n_res++] = a;             // 12 * n_res <= 12 * len <= 12 * n: inside the segment's first region
			tot += l; start = i;
		}
	}
	seg[s].n_anch = n_res; seg[s].tot = tot;
}
// NB: out[n_res] (12 bytes each) may run over ti/pred entries only at indices >= the ones still to be read:
// anchor k is written after chain element i >= k was consumed, and 12*k bytes from the region start end at or before
// 12*n + ... the start of ti (offset 12n) because k <= len <= n.

// ------------------------------------------------------------------------------------------------ select
// encoder.cpp:1149-1192 (MmerBasedAnchors: both orientations, keep the better), :1577-1622 (fix overlaps), :1103-1108 (sort)
__global__ void __launch_bounds__(128) k_select(const SegInfo* __restrict__ seg, uint32_t n_slots, const uint32_t* __restrict__ enc_list,
	const uint32_t* __restrict__ cand, const uint32_t* __restrict__ cand_n, const uint32_t* __restrict__ rd_len, const uint32_t* __restrict__ slot_dec,
	uint8_t* __restrict__ arena, S2P P, Node* __restrict__ nodes, CandView* __restrict__ cviews, const SegInfo* __restrict__ kseg, uint64_t kbase,
	unsigned long long* __restrict__ stats)
{
	const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
	if (slot >= n_slots) return;
	const uint32_t read = enc_list[slot], c = P.c;
	const uint32_t n_cand = slot_dec[slot] ? 0 : min(cand_n[read], c);
	CandView* out = cviews + (size_t)slot * c;
	uint32_t n_out = 0;
	// the reference's refuse reasons / orientation choices (encoder.cpp:1074-1104, :1118-1190): per read, logged only when no candidate is left
	bool too_many = false, too_low = false;
	uint32_t n_fwd = 0, n_rev = 0;
	for (uint32_t j = 0; j < n_cand; ++j) {
		SegInfo f = seg[((size_t)slot * c + j) * 2], r = seg[((size_t)slot * c + j) * 2 + 1];
		bool kmer = false;
		if (kseg) {      // HiFi: the k-mer based anchors win when either orientation gave some (encoder.cpp:1214-1232)
			const SegInfo kf = kseg[((size_t)slot * c + j) * 2], kr = kseg[((size_t)slot * c + j) * 2 + 1];
			if (kf.n || kr.n) { f = kf; r = kr; f.off += kbase; r.off += kbase; kmer = true; }
		}
		const bool af = f.n > 0 && (kmer || f.n_anch >= P.min_anchors), ar = r.n > 0 && (kmer || r.n_anch >= P.min_anchors);
		if (!af && !ar) {
			too_many |= ((f.pad | r.pad) & SEG_TOO_MANY_MATCHES) != 0;
			too_low |= (f.n > 0 && f.n_anch < P.min_anchors) || (r.n > 0 && r.n_anch < P.min_anchors);
			continue;
		}
		const bool use_f = af && (!ar || f.tot > r.tot);
		if (use_f) ++n_fwd; else ++n_rev;
		const SegInfo& g = use_f ? f : r;
		CandView v{};
		v.anc = g.off * PAIR_SLOT_BYTES; v.ref_id = cand[(size_t)read * c + j]; v.rev = use_f ? 0 : 1;
		v.first = 0; v.n = g.n_anch; v.tot = g.tot;
		// fix overlaps (in place, the total is NOT recomputed: encoder.cpp:1601-1622 leaves tot_anchor_len as it was)
		Anchor* an = reinterpret_cast<Anchor*>(arena + v.anc);
		for (uint32_t i = 0; i + 1 < v.n; ++i) {
			const uint32_t end = an[i].pos_ref + an[i].len;
			if (an[i + 1].pos_ref < end) { const uint32_t d = end - an[i + 1].pos_ref; an[i + 1].pos_ref += d; an[i + 1].len -= d; an[i + 1].pos_enc += d; }
		}
		for (uint32_t i = 0; i + 1 < v.n; ++i) {
			const uint32_t end = an[i].pos_enc + an[i].len;
			if (an[i + 1].pos_enc < end) { const uint32_t d = end - an[i + 1].pos_enc; an[i + 1].pos_enc += d; an[i + 1].len -= d; an[i + 1].pos_ref += d; }
		}
		// stable insertion by tot desc (std::sort on <= 16 elements is an insertion sort)
		uint32_t k = n_out++;
		while (k > 0 && out[k - 1].tot < v.tot) { out[k] = out[k - 1]; --k; }
		out[k] = v;
	}
	Node nd{};
	nd.read = read; nd.level = 0; nd.enc_start = 0; nd.enc_len = rd_len[read]; nd.first_task = 0;
	nd.ncand = n_out; nd.n_anch = n_out ? out[0].n : 0; nd.valid = n_out > 0;
	nodes[slot] = nd;
	if (stats) {
		if (slot_dec[slot]) atomicAdd(&stats[ST_NOT_ENOUGH], 1ull);
		else if (n_out == 0) { if (too_many) atomicAdd(&stats[ST_TOO_MANY], 1ull); if (too_low) atomicAdd(&stats[ST_TOO_LOW], 1ull); }
		if (n_fwd) atomicAdd(&stats[ST_NON_REV], (unsigned long long)n_fwd);
		if (n_rev) atomicAdd(&stats[ST_REV], (unsigned long long)n_rev);
	}
}

// ------------------------------------------------------------------------------------------------ host side
// Runs E1-E4 for the reads of enc_list (device + host copies).  On return nodes[0..n_list) / cviews hold the level-0 state.
clb_status s2_anchors(clb_ctx* c, const S2P& P, const std::vector<uint32_t>& h_list, const uint32_t* d_list,
	const uint32_t* d_ref_to_read, DevBuf<uint8_t>& arena, SegInfo* d_seg, uint32_t* d_slot_dec, Node* d_nodes, CandView* d_cviews,
	unsigned long long* d_cursor)
{
	cudaStream_t s = c->stream;
	const uint32_t nb = (uint32_t)h_list.size();
	if (!nb) return CLB_OK;
	if (P.c > MAX_C) return fail(c, CLB_ERR_BAD_ARG, "max_candidates above 32 is not supported by the anchor kernels");
	// global scratch for reads whose table / Bloom filter do not fit in shared memory
	std::vector<uint64_t> tab_off(nb, 0), bloom_off(nb, 0);
	uint64_t tab_total = 0, bloom_total = 0, est_pairs = 0;
	for (uint32_t i = 0; i < nb; ++i) {
		const uint32_t len = c->h_rd_len[h_list[i]];
		est_pairs += (uint64_t)len;
		if (len < P.m) continue;
		const uint64_t n_mm = len - P.m + 1;
		uint64_t cap = 64; while (cap < 2 * n_mm) cap <<= 1;
		uint64_t bbits = 1024; while (bbits < 16 * n_mm && bbits < (1u << 30)) bbits <<= 1;
		if (cap > SMEM_TAB_CELLS) { tab_off[i] = tab_total; tab_total += cap; }
		if (bbits > SMEM_BLOOM_WORDS * 32) { bloom_off[i] = bloom_total; bloom_total += bbits / 32; }
	}
	uint64_t* d_tab_off = nullptr; uint64_t* d_bloom_off = nullptr;
	CLB_CUDA(c, dev_malloc((void**)&d_tab_off, sizeof(uint64_t) * nb, s));
	CLB_CUDA(c, dev_malloc((void**)&d_bloom_off, sizeof(uint64_t) * nb, s));
	// kept from batch to batch, with headroom for the batches to come (batches hold the same number of bases)
	if (tab_total + 1 > c->s2_gtab.cap) CLB_CUDA(c, c->s2_gtab.reserve(tab_total + tab_total / 4 + 1, s, false));
	if (bloom_total + 1 > c->s2_gbloom.cap) CLB_CUDA(c, c->s2_gbloom.reserve(bloom_total + bloom_total / 4 + 1, s, false));
	uint32_t* const g_tab = c->s2_gtab.p; uint32_t* const g_bloom = c->s2_gbloom.p;
	CLB_CUDA(c, cudaMemcpyAsync(d_tab_off, tab_off.data(), sizeof(uint64_t) * nb, cudaMemcpyHostToDevice, s));
	CLB_CUDA(c, cudaMemcpyAsync(d_bloom_off, bloom_off.data(), sizeof(uint64_t) * nb, cudaMemcpyHostToDevice, s));
	// HiFi: anchors from the shared k-mers first; candidates that get some are skipped by the m-mer search
	SegInfo* d_kseg = nullptr; uint8_t* d_skip = nullptr; uint8_t* d_kanc = nullptr; uint64_t kanc_slots = 0, kbase = 0;
	struct KFree { std::vector<void*> v; cudaStream_t s; ~KFree() { for (void* q : v) dev_free_async(q, s); } } kfree{{}, s};
	if (c->prm.is_hifi) {
		const uint64_t nsc = (uint64_t)nb * P.c;
		std::vector<uint32_t> votes(c->n_reads * P.c);
		CLB_CUDA(c, cudaMemcpyAsync(votes.data(), c->cand_votes, sizeof(uint32_t) * votes.size(), cudaMemcpyDeviceToHost, s));
		std::vector<uint32_t> cn(c->n_reads);
		CLB_CUDA(c, cudaMemcpyAsync(cn.data(), c->cand_n, sizeof(uint32_t) * cn.size(), cudaMemcpyDeviceToHost, s));
		CLB_CUDA(c, cudaStreamSynchronize(s));
		std::vector<uint64_t> ktab_off(nsc, 0), kanc_off(nsc, 0);
		uint64_t ktab_total = 0;
		for (uint32_t i = 0; i < nb; ++i) for (uint32_t j = 0; j < P.c; ++j) {
			const uint32_t nco = j < cn[h_list[i]] ? votes[(size_t)h_list[i] * P.c + j] : 0;
			uint64_t cap = 8; while (cap < 4ull * nco) cap <<= 1;
			ktab_off[(size_t)i * P.c + j] = ktab_total; ktab_total += nco ? cap : 0;
			kanc_off[(size_t)i * P.c + j] = kanc_slots; kanc_slots += nco ? 2 * ((12ull * nco + PAIR_SLOT_BYTES - 1) / PAIR_SLOT_BYTES + 1) : 0;
		}
		uint64_t* d_ktab_off = nullptr; uint64_t* d_kanc_off = nullptr; KEntry* d_ktab = nullptr;
		auto kalloc = [&](void** q, uint64_t bytes) { cudaError_t e = dev_malloc(q, bytes ? bytes : 1, s); if (e == cudaSuccess) kfree.v.push_back(*q); return e; };
		CLB_CUDA(c, kalloc((void**)&d_kseg, sizeof(SegInfo) * nsc * 2)); CLB_CUDA(c, kalloc((void**)&d_skip, nsc));
		CLB_CUDA(c, kalloc((void**)&d_ktab_off, sizeof(uint64_t) * nsc)); CLB_CUDA(c, kalloc((void**)&d_kanc_off, sizeof(uint64_t) * nsc));
		CLB_CUDA(c, kalloc((void**)&d_ktab, sizeof(KEntry) * (ktab_total + 1))); CLB_CUDA(c, kalloc((void**)&d_kanc, kanc_slots * PAIR_SLOT_BYTES + 64));
		CLB_CUDA(c, cudaMemcpyAsync(d_ktab_off, ktab_off.data(), sizeof(uint64_t) * nsc, cudaMemcpyHostToDevice, s));
		CLB_CUDA(c, cudaMemcpyAsync(d_kanc_off, kanc_off.data(), sizeof(uint64_t) * nsc, cudaMemcpyHostToDevice, s));
		KmerArgs ka{c->pk.p, c->rd_start.p, c->rd_len.p, d_list, nb, c->cand, c->cand_n, c->cand_votes, c->common_off, c->common, d_ref_to_read,
			c->prm.kmer_len, P.c, d_ktab_off, d_ktab, d_kanc_off, d_kanc, d_kseg, d_skip};
		CLB_TIMED(c, K_ANCHORS, (k_kmer_anchors<<<(uint32_t)nsc, 128, 0, s>>>(ka)));
		CLB_LAUNCH_CHECK(c, "k_kmer_anchors");
		CLB_CUDA(c, cudaStreamSynchronize(s));          // the host vectors above were read by the copies
	}
	const size_t smem = ((sizeof(MatchShared) + 15) & ~15ull) + sizeof(uint32_t) * (SMEM_TAB_CELLS + SMEM_BLOOM_WORDS) + sizeof(uint64_t) * (TILE_WORDS + 2 + 2 + SMEM_HIT_WORDS);
	static const int tma_on = [] { const char* e = std::getenv("CLB_TMA"); return e ? std::atoi(e) : 1; }();      // on a B200: k_anchors 1 283 ms per 25 Gbases with the bulk-copy stage, 1 425 ms without (profiles/r02h_*)
	CLB_CUDA(c, cudaFuncSetAttribute(k_anchor_match, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	// first guess of the arena: one pair per base of every read and candidate half-used; the kernel reports the exact need
	uint64_t cap_pairs = std::max<uint64_t>(1u << 16, est_pairs * 2);
	clb_status st = CLB_OK;
	for (int attempt = 0; attempt < 3; ++attempt) {
		cudaError_t e = arena.reserve((cap_pairs + kanc_slots + 4) * PAIR_SLOT_BYTES + 64, s, false);
		if (e != cudaSuccess) { st = cuda_fail(c, e, "pair arena"); break; }
		CLB_CUDA(c, cudaMemsetAsync(d_cursor, 0, sizeof(unsigned long long), s));
		MatchArgs a{};
		a.pk = c->pk.p; a.rd_start = c->rd_start.p; a.rd_len = c->rd_len.p; a.enc_list = d_list; a.n_list = nb;
		a.cand = c->cand; a.cand_n = c->cand_n; a.ref_to_read = d_ref_to_read; a.P = P;
		a.tab_off = d_tab_off; a.g_tab = g_tab; a.bloom_off = d_bloom_off; a.g_bloom = g_bloom;
		a.arena = arena.p; a.arena_cap = cap_pairs; a.cursor = d_cursor; a.seg = d_seg; a.slot_dec = d_slot_dec; a.skip = d_skip; a.tma = tma_on;
		CLB_TIMED(c, K_ANCHORS, (k_anchor_match<<<nb, MATCH_THREADS, smem, s>>>(a)));
		CLB_LAUNCH_CHECK(c, "k_anchor_match");
		unsigned long long used = 0;
		CLB_CUDA(c, cudaMemcpyAsync(&used, d_cursor, sizeof(used), cudaMemcpyDeviceToHost, s));
		CLB_CUDA(c, cudaStreamSynchronize(s));
		if (used >= (1ull << 62)) { st = fail(c, CLB_ERR_CAPACITY, "a read shares more than 2^31 m-mer matches with one candidate"); break; }
		if (used <= cap_pairs) break;
		if (attempt == 2) { st = fail(c, CLB_ERR_CAPACITY, "pair arena did not converge"); break; }
		cap_pairs = used + 1024;
	}
	dev_free_async(d_tab_off, s); dev_free_async(d_bloom_off, s);
	if (st != CLB_OK) return st;
	if (kanc_slots) {      // the k-mer anchors join the arena behind the pair region
		kbase = (arena.cap - 64) / PAIR_SLOT_BYTES - kanc_slots - 2;
		CLB_CUDA(c, cudaMemcpyAsync(arena.p + kbase * PAIR_SLOT_BYTES, d_kanc, kanc_slots * PAIR_SLOT_BYTES, cudaMemcpyDeviceToDevice, s));
	}
	const uint32_t n_seg = nb * P.c * 2;
	CLB_TIMED(c, K_ANCHORS, (k_pairs_sort<<<n_seg, SORT_THREADS, 0, s>>>(d_seg, n_seg, arena.p)));
	CLB_LAUNCH_CHECK(c, "k_pairs_sort");
	CLB_TIMED(c, K_ANCHORS, (k_lis<<<(n_seg + 127) / 128, 128, 0, s>>>(d_seg, n_seg, arena.p, P.m)));
	CLB_LAUNCH_CHECK(c, "k_lis");
	CLB_TIMED(c, K_ANCHORS, (k_select<<<(nb + 127) / 128, 128, 0, s>>>(d_seg, nb, d_list, c->cand, c->cand_n, c->rd_len.p, d_slot_dec, arena.p, P, d_nodes, d_cviews, d_kseg, kbase, c->collect_stats ? c->d_stats : nullptr)));
	CLB_LAUNCH_CHECK(c, "k_select");
	return CLB_OK;
}

} // namespace clb
