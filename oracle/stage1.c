/* oracle/stage1.c — TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the reference's stage 1 (k-mer filter + similarity graph).  It exists so
 * the CUDA path can be checked on the GPU box, where /root/reference is absent.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it; the product never does.
 * Pinned against the reference itself: tests/golden/ holds dumps written by oracle/_ref/ref_stage_dump
 * (the unmodified reference classes with taps) and tests/test_oracle_stage1.py replays them.
 *
 * Each function cites the reference lines it restates (paths relative to /root/reference/src).
 * The data structures are deliberately naive (sorting, open addressing): results, not speed.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

/* ---------------------------------------------------------------- S1: hash filter ------------- */
/* colord/filter_kmers.cpp:24-32  ==  filtering-KMC/hash_filter.h:8-16 (MurmurHash3 fmix64) */
uint64_t orc_murmur64(uint64_t x)
{
	x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
	x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
	x ^= x >> 33;
	return x;
}

/* colord/kmer_filter.h:129-133 (Possible) / hash_filter.h:28-78 (checkModuloHash) */
int orc_possible(uint64_t kmer, uint32_t modulo) { return orc_murmur64(kmer) % modulo == 0; }

static int sym_of(uint8_t c)
{
	/* colord/in_reads.cpp:24-42 (to_read_t): A,C,G,T,N -> 0..4 */
	switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; default: return 4; }
}

static int cmp_u64(const void* a, const void* b)
{
	uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
	return x < y ? -1 : x > y;
}

/* ---------------------------------------------------------------- S2: counting ---------------- */
/* Semantics of the filtering-KMC run (colord/count_kmers.cpp:59-68 -> filtering-KMC):
 *  - every window of k symbols without N yields its canonical k-mer (KMC splits reads at N),
 *  - only canonical k-mers with murmur64 % modulo == 0 are counted (kb_collector.cpp:66,98; kb_sorter.h:341,358),
 *  - n_total  = sum of counts of all such k-mers, n_unique = number of distinct ones (kb_sorter.h:1018-1047),
 *  - a k-mer survives iff min_count <= count <= 1e9, its stored count saturates at max_count (:1021-1028),
 *  - n_unique_counted = survivors (kmc.h:1474), total_count_filtered = sum of stored counts
 *    (colord/filter_kmers.cpp:68-83).
 * out_kmers/out_counts (capacity cap) receive the survivors in ascending k-mer order; returns their number
 * (or the required capacity if cap is too small).  stats[0..4] = n_reads, n_total, n_unique,
 * n_unique_counted, total_count_filtered. */
uint64_t orc_count_kmers(const uint8_t* bases, const uint64_t* offsets, uint32_t n_reads,
	uint32_t k, uint32_t modulo, uint32_t min_count, uint32_t max_count,
	uint64_t* out_kmers, uint32_t* out_counts, uint64_t cap, uint64_t* stats)
{
	const uint64_t mask = k == 32 ? ~0ULL : ((1ULL << (2 * k)) - 1);
	const uint32_t rev_offset = 2 * (k - 1);
	uint64_t n_alloc = 1024, n = 0;
	uint64_t* v = (uint64_t*)malloc(n_alloc * sizeof(uint64_t));
	for (uint32_t r = 0; r < n_reads; ++r)
	{
		uint64_t str = 0, rev = 0; uint32_t run = 0;
		for (uint64_t p = offsets[r]; p < offsets[r + 1]; ++p)
		{
			int s = sym_of(bases[p]);
			if (s > 3) { run = 0; str = rev = 0; continue; }
			/* colord/in_reads.h:59-73 (CKmerWalker::NextKmer): same rolling update as KMC's */
			str = ((str << 2) + (uint64_t)s) & mask;
			rev = (rev >> 2) + ((uint64_t)(3 - s) << rev_offset);
			if (++run < k) continue;
			uint64_t can = str < rev ? str : rev;
			if (!orc_possible(can, modulo)) continue;
			if (n == n_alloc) { n_alloc *= 2; v = (uint64_t*)realloc(v, n_alloc * sizeof(uint64_t)); }
			v[n++] = can;
		}
	}
	qsort(v, n, sizeof(uint64_t), cmp_u64);
	uint64_t n_unique = 0, n_surv = 0, tot_filtered = 0;
	for (uint64_t i = 0; i < n; )
	{
		uint64_t j = i; while (j < n && v[j] == v[i]) ++j;
		uint64_t c = j - i;
		++n_unique;
		if (c >= min_count && c <= 1000000000ULL)
		{
			if (c > max_count) c = max_count;
			if (n_surv < cap) { out_kmers[n_surv] = v[i]; out_counts[n_surv] = (uint32_t)c; }
			++n_surv; tot_filtered += c;
		}
		i = j;
	}
	stats[0] = n_reads; stats[1] = n; stats[2] = n_unique; stats[3] = n_surv; stats[4] = tot_filtered;
	free(v);
	return n_surv;
}

/* ---------------------------------------------------------------- S3: membership -------------- */
/* colord/kmer_filter.h:60-80 (CCompactedKmers insert/check): a set; here a sorted array + bsearch. */
static int set_has(const uint64_t* set, uint64_t n, uint64_t x)
{
	uint64_t lo = 0, hi = n;
	while (lo < hi) { uint64_t m = (lo + hi) >> 1; if (set[m] < x) lo = m + 1; else hi = m; }
	return lo < n && set[lo] == x;
}

/* ---------------------------------------------------------------- S4: accepted k-mers --------- */
/* colord/reads_sim_graph.cpp:134-164: reads with N or shorter than k give nothing; otherwise canonical
 * k-mers in read order, kept iff Possible && first occurrence in this read && Check.
 * acc_off[n_reads+1] is filled; acc (capacity cap) receives the lists back to back.  Returns total. */
uint64_t orc_accepted_kmers(const uint8_t* bases, const uint64_t* offsets, uint32_t n_reads,
	uint32_t k, uint32_t modulo, const uint64_t* set, uint64_t n_set,
	uint64_t* acc_off, uint64_t* acc, uint64_t cap)
{
	const uint64_t mask = k == 32 ? ~0ULL : ((1ULL << (2 * k)) - 1);
	const uint32_t rev_offset = 2 * (k - 1);
	uint64_t total = 0;
	for (uint32_t r = 0; r < n_reads; ++r)
	{
		acc_off[r] = total;
		uint64_t len = offsets[r + 1] - offsets[r];
		int has_n = 0;
		for (uint64_t p = offsets[r]; p < offsets[r + 1]; ++p) if (sym_of(bases[p]) > 3) { has_n = 1; break; }
		if (has_n || len < k) continue;
		uint64_t str = 0, rev = 0;
		/* per-read "already processed" set (reads_sim_graph.cpp:150): open addressing, key+1 so that 0 = empty */
		uint64_t scap = 64; while (scap < 2 * (len / modulo + 16)) scap <<= 1;
		uint64_t* seen = (uint64_t*)calloc(scap, sizeof(uint64_t));
		uint64_t n_seen = 0;
		for (uint64_t i = 0; i < len; ++i)
		{
			int s = sym_of(bases[offsets[r] + i]);
			str = ((str << 2) + (uint64_t)s) & mask;
			rev = (rev >> 2) + ((uint64_t)(3 - s) << rev_offset);
			if (i + 1 < k) continue;
			uint64_t can = str < rev ? str : rev;
			if (!orc_possible(can, modulo) || !set_has(set, n_set, can)) continue;
			if (2 * (n_seen + 1) > scap)
			{	/* grow */
				uint64_t ncap = scap * 2; uint64_t* ns = (uint64_t*)calloc(ncap, sizeof(uint64_t));
				for (uint64_t q = 0; q < scap; ++q) if (seen[q]) { uint64_t h = orc_murmur64(seen[q]) & (ncap - 1); while (ns[h]) h = (h + 1) & (ncap - 1); ns[h] = seen[q]; }
				free(seen); seen = ns; scap = ncap;
			}
			uint64_t h = orc_murmur64(can + 1) & (scap - 1); int dup = 0;
			while (seen[h]) { if (seen[h] == can + 1) { dup = 1; break; } h = (h + 1) & (scap - 1); }
			if (dup) continue;
			seen[h] = can + 1; ++n_seen;
			if (total < cap) acc[total] = can;
			++total;
		}
		free(seen);
	}
	acc_off[n_reads] = total;
	return total;
}

/* ---------------------------------------------------------------- S5a: sparse sampler --------- */
/* colord/ref_reads_accepter.h:51-57 with a default-seeded std::mt19937 (seed 5489) and libstdc++'s
 * uniform_real_distribution<double>(0,1) = generate_canonical<double,53>: two 32-bit draws,
 * (lo + hi*2^32) / 2^64, clamped below 1.  One draw per call with idx >= n_pseudo. */
typedef struct { uint32_t mt[624]; int idx; } orc_mt;
static void mt_seed(orc_mt* m, uint32_t s)
{
	m->mt[0] = s;
	for (int i = 1; i < 624; ++i) m->mt[i] = 1812433253U * (m->mt[i - 1] ^ (m->mt[i - 1] >> 30)) + (uint32_t)i;
	m->idx = 624;
}
static uint32_t mt_next(orc_mt* m)
{
	if (m->idx >= 624)
	{
		for (int i = 0; i < 624; ++i)
		{
			uint32_t y = (m->mt[i] & 0x80000000U) | (m->mt[(i + 1) % 624] & 0x7fffffffU);
			m->mt[i] = m->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1) ? 0x9908b0dfU : 0);
		}
		m->idx = 0;
	}
	uint32_t y = m->mt[m->idx++];
	y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680U; y ^= (y << 15) & 0xefc60000U; y ^= y >> 18;
	return y;
}
static double mt_uniform01(orc_mt* m)
{
	double lo = (double)mt_next(m), hi = (double)mt_next(m);
	double r = (lo + hi * 4294967296.0) / 18446744073709551616.0;
	if (r >= 1.0) r = nextafter(1.0, 0.0);
	return r;
}
/* decisions[i] = ShouldAddToReference(i) for i in [0, n): exactly the stream the graph consumes when it
 * calls the accepter once per read in input order (reads_sim_graph.cpp:339-341). */
void orc_sampler(uint32_t range, double exponent, uint32_t n_pseudo, uint32_t n, uint8_t* decisions)
{
	orc_mt m; mt_seed(&m, 5489U);
	for (uint32_t i = 0; i < n; ++i)
	{
		if (i < n_pseudo) { decisions[i] = 1; continue; }
		uint32_t range_no = (i - n_pseudo) / range;
		double p = pow(1.0 / ((double)range_no + 1.0), exponent);
		decisions[i] = mt_uniform01(&m) <= p;
	}
}

/* ---------------------------------------------------------------- S5b: similarity graph ------- */
/* colord/reads_sim_graph.cpp:324-427 (processReadsPack) and :429-528 (HiFi), restated serially.
 * For read i in input order: votes[ref] += 1 for every (k-mer -> ref) entry of each accepted k-mer;
 * then, if the read is a reference (no N and, in sparse mode, sampler said yes: `sampled[i]`; pass all
 * ones for -R all), each k-mer whose list is shorter than max_kmer_count gets (k-mer -> ref id of i).
 * Top max_candidates by (votes desc, id asc).  The HiFi double increment (:475-484) does not change the
 * order, so plain counts are used.  cand[i*max_candidates + j], cand_n[i].
 * If common_off != NULL (HiFi) the shared k-mers of every chosen candidate are emitted in the read's
 * k-mer order: common_off[(i*max_candidates + j)] .. +common_n[..] into common (capacity common_cap). */
typedef struct { uint64_t key; uint32_t* refs; uint32_t n, cap; } post_t;

static uint64_t tab_find(post_t* tab, uint64_t tcap, uint64_t key, int* found)
{
	uint64_t h = orc_murmur64(key ^ 0x9E3779B97F4A7C15ULL) & (tcap - 1);
	while (tab[h].refs != NULL)
	{
		if (tab[h].key == key) { *found = 1; return h; }
		h = (h + 1) & (tcap - 1);
	}
	*found = 0;
	return h;
}

typedef struct { uint32_t ref, votes; } vote_t;
static int cmp_vote(const void* a, const void* b)
{
	const vote_t* x = (const vote_t*)a; const vote_t* y = (const vote_t*)b;
	if (x->votes != y->votes) return x->votes > y->votes ? -1 : 1;
	return x->ref < y->ref ? -1 : x->ref > y->ref;
}

uint64_t orc_sim_graph(const uint64_t* acc_off, const uint64_t* acc, uint32_t n_reads,
	const uint8_t* has_n, const uint8_t* sampled, uint32_t max_candidates, uint32_t max_kmer_count,
	uint32_t* cand, uint32_t* cand_n,
	uint64_t* common_off, uint32_t* common_n, uint64_t* common, uint64_t common_cap)
{
	uint64_t total = acc_off[n_reads];
	uint64_t tcap = 64; while (tcap < 2 * total + 64) tcap <<= 1;
	post_t* tab = (post_t*)calloc(tcap, sizeof(post_t));
	uint32_t id_in_reference = 0;
	uint32_t* votes = (uint32_t*)calloc((size_t)n_reads + 1, sizeof(uint32_t));
	uint32_t* touched = (uint32_t*)malloc(((size_t)n_reads + 1) * sizeof(uint32_t));
	vote_t* vv = (vote_t*)malloc(((size_t)n_reads + 1) * sizeof(vote_t));
	uint64_t n_common = 0;
	for (uint32_t i = 0; i < n_reads; ++i)
	{
		int accept = !has_n[i] && sampled[i];
		if (accept) ++id_in_reference;
		uint32_t n_touched = 0;
		for (uint64_t e = acc_off[i]; e < acc_off[i + 1]; ++e)
		{
			int found; uint64_t h = tab_find(tab, tcap, acc[e], &found);
			uint32_t card = 0;
			if (found)
				for (uint32_t j = 0; j < tab[h].n; ++j)
				{
					uint32_t ref = tab[h].refs[j];
					if (votes[ref]++ == 0) touched[n_touched++] = ref;
					++card;
				}
			if (accept && card < max_kmer_count)
			{
				if (!found) { tab[h].key = acc[e]; tab[h].cap = 4; tab[h].n = 0; tab[h].refs = (uint32_t*)malloc(4 * sizeof(uint32_t)); }
				if (tab[h].n == tab[h].cap) { tab[h].cap *= 2; tab[h].refs = (uint32_t*)realloc(tab[h].refs, tab[h].cap * sizeof(uint32_t)); }
				tab[h].refs[tab[h].n++] = id_in_reference - 1;
			}
		}
		for (uint32_t j = 0; j < n_touched; ++j) { vv[j].ref = touched[j]; vv[j].votes = votes[touched[j]]; votes[touched[j]] = 0; }
		qsort(vv, n_touched, sizeof(vote_t), cmp_vote);
		uint32_t nc = n_touched < max_candidates ? n_touched : max_candidates;
		cand_n[i] = nc;
		for (uint32_t j = 0; j < nc; ++j) cand[(uint64_t)i * max_candidates + j] = vv[j].ref;
		if (common_off)
			for (uint32_t j = 0; j < nc; ++j)
			{
				uint64_t slot = (uint64_t)i * max_candidates + j;
				common_off[slot] = n_common; uint32_t cnt = 0;
				/* the read's own insertions (this iteration) carry ref id id_in_reference-1, never a candidate */
				for (uint64_t e = acc_off[i]; e < acc_off[i + 1]; ++e)
				{
					int found; uint64_t h = tab_find(tab, tcap, acc[e], &found);
					if (!found) continue;
					for (uint32_t q = 0; q < tab[h].n; ++q)
						if (tab[h].refs[q] == vv[j].ref) { if (n_common < common_cap) common[n_common] = acc[e]; ++n_common; ++cnt; break; }
				}
				common_n[slot] = cnt;
			}
	}
	for (uint64_t h = 0; h < tcap; ++h) free(tab[h].refs);
	free(tab); free(votes); free(touched); free(vv);
	return n_common;
}

/* ---------------------------------------------------------------- S6: 2-bit reference reads --- */
/* colord/reference_reads.h:35-72 (compact): 4 bases per byte, first base in bits 7..6, plus one trailer
 * byte = number of symbols in the last partial byte (0 if the length is a multiple of 4).
 * out must hold (len+3)/4 + 1 bytes; returns that size. */
uint64_t orc_pack_ref_read(const uint8_t* bases, uint64_t len, uint8_t* out)
{
	uint64_t n = 0; uint8_t b = 0; uint32_t in_byte = 0;
	for (uint64_t i = 0; i < len; ++i)
	{
		b = (uint8_t)((b << 2) | sym_of(bases[i]));
		if (++in_byte == 4) { out[n++] = b; b = 0; in_byte = 0; }
	}
	if (in_byte) out[n++] = (uint8_t)(b << (2 * (4 - in_byte)));
	out[n++] = (uint8_t)in_byte;
	return n;
}
