/* oracle/stage3_qual.c — TEST INFRASTRUCTURE ONLY.
 *
 * Quality stream, the "*-avg" modes (ONT default 4-avg, HiFi default 5-avg, 2-avg).
 *
 * Part 1 restates the reference's lossy transform and context model (paths relative to /root/reference/src/colord):
 *   quality_coder.cpp:250-270   adjust_quality_map_symbols: phred -> bin by the forward thresholds
 *   quality_coder_impl.cpp:191-249 (encode_quad_average; :130-189 quinary, :252-310 binary): per-read per-bin means,
 *                                coded as (uint32)(mean * 256) (:821-835), then one bin symbol per base under the context
 *                                [3 (6 for 2 bins) previous symbols] + [bases i-2 .. i+1] + [match / anchor flags if level > 1]
 *   quality_coder.cpp:528-537   reset_context / update_context
 *   quality_coder_impl.cpp:25-76 analyze_es: per-base flags from the read's tuples
 *   quality_coder_impl.cpp:559-601 decode_quad_average: the decoder rebuilds integer qualities from the bin means by error
 *                                diffusion — this is what `colord decompress` prints, pinned by the reference's own
 *                                test/<name>.quan fixtures (tests/golden/qual_*.bin.gz)
 * Part 2 is the CPU twin of the device's NATIVE container for these symbols (not the reference's adaptive range coder, whose
 * single serial model chain cannot run in lockstep — SURVEY.md §7): two passes, static per-context frequency tables
 * normalised to 2^12, interleaved rANS (31-bit state, 16-bit renormalisation), 64 lanes per read pack.  The device output
 * must equal this byte for byte; the decoder below gives the round trip.  Parity of the native container with the reference
 * is by (a) identical reconstructed qualities and (b) stream size within the north star's 0.5 % — there is no reference
 * bitstream to compare with, which tests/test_oracle_stage3.py states.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define QB_LANES 64
#define QB_PROB_BITS 12
#define QB_M (1u << QB_PROB_BITS)
#define QB_L (1u << 15)        /* states stay below 2^31: the device divides by multiplying with a 32-bit reciprocal */
#define QB_MIN_CTX 32u         /* contexts seen fewer times share the fallback model of their previous-symbol part */

typedef struct {
	uint32_t n_bins;            /* 2, 4 or 5                                                               */
	uint32_t thr[4];            /* forward thresholds (n_bins - 1 used): bin = #thresholds <= phred        */
	uint32_t level;             /* compression level: > 1 adds the match / anchor flags to the context     */
} orc_qual_params;

static uint32_t bits_per_symbol(uint32_t n_bins) { return n_bins == 2 ? 2 : 3; }
static uint32_t ctx_symbols(uint32_t n_bins) { return n_bins == 2 ? 6 : 3; }
uint32_t orc_qual_ctx_bits(const orc_qual_params* P) { return bits_per_symbol(P->n_bins) * ctx_symbols(P->n_bins) + 8 + (P->level > 1 ? 2 : 0); }

static void make_map(const orc_qual_params* P, uint8_t* map /*96*/)
{
	for (uint32_t q = 0; q < 96; ++q) { uint32_t b = 0; while (b + 1 < P->n_bins && q >= P->thr[b]) ++b; map[q] = (uint8_t)b; }
}

/* quality_coder_impl.cpp:25-76; flags: 0 none, 1 match, 2 anchor (plain reads: none of the two) */
static void base_flags(const uint8_t* es, uint64_t es_n, uint8_t* fl, uint32_t n)
{
	memset(fl, 0, n);
	if (!es_n) return;
	const uint32_t t0 = es[0] >> 4;
	if (t0 == 9 || t0 == 11) return;
	uint64_t p = 5; uint32_t at = 0;                       /* start_es is 5 bytes */
	while (p < es_n)
	{
		const uint32_t t = es[p] >> 4;
		if (t == 4) { uint32_t len = ((uint32_t)(es[p] & 15) << 24) | ((uint32_t)es[p + 1] << 16) | ((uint32_t)es[p + 2] << 8) | es[p + 3]; for (uint32_t k = 0; k < len && at < n; ++k) fl[at++] = 2; p += 4; }
		else if (t == 5) p += 4;
		else if (t == 6) p += 5;
		else { if (t == 2) { if (at < n) fl[at] = 1; ++at; } else if (t == 0 || t == 3) ++at; p += 1; }
	}
}

/* symbols of one read: avg16[b] = (uint32)(mean_b * 256), sym[i] = bin of quality i; ctx[i] = model context of position i */
static void read_symbols(const orc_qual_params* P, const uint8_t* map, const uint8_t* bases, const uint8_t* qual, uint32_t n,
	const uint8_t* fl, uint32_t* avg16, uint8_t* sym, uint32_t* ctx)
{
	double sum[5] = {0, 0, 0, 0, 0}; uint32_t cnt[5] = {0, 0, 0, 0, 0};
	uint32_t h[128]; memset(h, 0, sizeof h);
	for (uint32_t i = 0; i < n; ++i) ++h[qual[i] & 127];
	for (uint32_t q = 33; q < 128; ++q) { sum[map[q - 33]] += (double)(q - 33) * h[q]; cnt[map[q - 33]] += h[q]; }
	for (uint32_t b = 0; b < P->n_bins; ++b) { const double avg = cnt[b] ? sum[b] / cnt[b] : 0.0; avg16[b] = (uint32_t)(avg * 256); }
	const uint32_t bps = bits_per_symbol(P->n_bins), cb = bps * ctx_symbols(P->n_bins);
	const uint32_t cmask = (1u << cb) - 1;
	static const uint8_t code[256] = { ['A'] = 0, ['C'] = 1, ['G'] = 2, ['T'] = 3, ['N'] = 0 };   /* valid_sym: x & 3, N = 4 -> 0 */
	uint32_t c = cmask, dna = n ? code[bases[0]] : 0;
	for (uint32_t i = 0; i < n; ++i)
	{
		dna <<= 2; if (i + 1 < n) dna += code[bases[i + 1]]; dna &= 0xff;
		uint32_t x = c + (dna << cb);
		if (P->level > 1) x += (uint32_t)(fl[i] == 1) << (cb + 8), x += (uint32_t)(fl[i] == 2) << (cb + 9);
		ctx[i] = x;
		sym[i] = map[qual[i] - 33];
		c = ((c << bps) + sym[i]) & cmask;
	}
}

/* quality_coder_impl.cpp:559-601: what the decoder prints for (means, symbols) */
static void reconstruct(const orc_qual_params* P, const uint32_t* avg16, const uint8_t* sym, uint32_t n, uint8_t* out)
{
	double avg[5], avg_sum[5] = {0, 0, 0, 0, 0}, qual_sum[5] = {0, 0, 0, 0, 0};
	for (uint32_t b = 0; b < P->n_bins; ++b) avg[b] = (double)avg16[b] / 256.0;
	for (uint32_t i = 0; i < n; ++i)
	{
		const uint32_t d = sym[i];
		avg_sum[d] += avg[d];
		const uint32_t q = (uint32_t)(avg_sum[d] - qual_sum[d]);
		qual_sum[d] += q;
		out[i] = (uint8_t)(q + 33);
	}
}

/* The reference's lossy transform end to end: quality strings as `colord decompress` would print them. */
void orc_qual_lossy(const orc_qual_params* P, const uint8_t* bases, const uint8_t* quals, const uint64_t* offsets, uint32_t n_reads, uint8_t* out)
{
	uint8_t map[96]; make_map(P, map);
	for (uint32_t r = 0; r < n_reads; ++r)
	{
		const uint32_t n = (uint32_t)(offsets[r + 1] - offsets[r]);
		uint8_t* sym = (uint8_t*)malloc(n + 1); uint32_t* ctx = (uint32_t*)malloc(4 * (n + 1)); uint8_t* fl = (uint8_t*)calloc(n + 1, 1);
		uint32_t avg16[5];
		read_symbols(P, map, bases + offsets[r], quals + offsets[r], n, fl, avg16, sym, ctx);
		reconstruct(P, avg16, sym, n, out + offsets[r]);
		free(sym); free(ctx); free(fl);
	}
}

/* ------------------------------------------------------------------------------------------------ native container */
/* counts -> frequencies summing to 2^12, every seen symbol >= 1; the remainder goes to the (first) most frequent symbol */
static void normalise(const uint32_t* cnt, uint32_t n, uint16_t* f)
{
	uint64_t tot = 0; uint32_t best = 0;
	for (uint32_t i = 0; i < n; ++i) { tot += cnt[i]; if (cnt[i] > cnt[best]) best = i; }
	if (!tot) { for (uint32_t i = 0; i < n; ++i) f[i] = 0; return; }
	uint32_t sum = 0;
	for (uint32_t i = 0; i < n; ++i) { uint32_t v = (uint32_t)(((uint64_t)cnt[i] << QB_PROB_BITS) / tot); if (cnt[i] && !v) v = 1; f[i] = (uint16_t)v; sum += v; }
	f[best] = (uint16_t)(f[best] + QB_M - sum);
}

typedef struct { uint8_t* p; uint64_t n, cap; } obuf;
static void ob_put(obuf* o, const void* s, uint64_t k) { if (o->n + k > o->cap) { o->cap = (o->n + k) * 2 + 4096; o->p = (uint8_t*)realloc(o->p, o->cap); } memcpy(o->p + o->n, s, k); o->n += k; }
static void ob_u32(obuf* o, uint32_t v) { ob_put(o, &v, 4); }
static void ob_u64(obuf* o, uint64_t v) { ob_put(o, &v, 8); }

/* one rANS step, 16-bit words are pushed to a stack that is later read backwards */
static inline uint32_t rans_put(uint32_t x, uint32_t f, uint32_t c, uint16_t* w, uint64_t* nw)
{
	const uint64_t x_max = (uint64_t)f << 19;       /* ((L >> PROB_BITS) << 16) * f */
	while (x >= x_max) { w[(*nw)++] = (uint16_t)x; x >>= 16; }
	return ((x / f) << QB_PROB_BITS) + (x % f) + c;
}

/* Stream layout (little endian):
 *   "QB01", n_bins, level, thr[4], n_reads u64, n_packs u32, ctx_bits u32
 *   mean model: n_bins x 128 u16 frequencies (high byte of mean*256 given the bin)
 *   fallback model: n_bins-1 u16 frequencies for each value of the previous-symbols part of the context (last one implied);
 *                   it codes the symbols of every context seen fewer than QB_MIN_CTX times
 *   n_dense u32, then per dense context (ascending): LEB128 gap to the previous one, n_bins-1 u16 frequencies
 *   per pack: n_reads u32, QB_LANES x lane size in bytes (u32), lane streams
 *   lane l codes reads l, l + QB_LANES, ... of its pack; a lane stream = final state (u32) + 16-bit words in decoding order
 * es / es_off may be NULL when level <= 1. */
/* "QB02" (the device's container since round 2): the same header and tables; per pack: n_reads u32, QB2_STREAMS x stream size in bytes (u32),
 * streams.  Stream s codes reads s, s + QB2_STREAMS, ... of its pack with QB2_STATES interleaved rANS states: the symbols of a read are its
 * 2 x n_bins mean bytes, then one bin symbol per base, and symbol k of a read belongs to state k mod QB2_STATES.  A stream = the final
 * states (u32 each) + 16-bit words in decoding order. */
#define QB2_STREAMS 4
#define QB2_STATES 32
static uint64_t qual_encode_v(int version, const orc_qual_params* P, const uint8_t* bases, const uint8_t* quals, const uint64_t* offsets, uint32_t n_reads,
	const uint8_t* es, const uint64_t* es_off, const uint32_t* pack_sizes, uint32_t n_packs, uint8_t* out, uint64_t cap);
uint64_t orc_qual_encode(const orc_qual_params* P, const uint8_t* bases, const uint8_t* quals, const uint64_t* offsets, uint32_t n_reads,
	const uint8_t* es, const uint64_t* es_off, const uint32_t* pack_sizes, uint32_t n_packs, uint8_t* out, uint64_t cap)
{ return qual_encode_v(2, P, bases, quals, offsets, n_reads, es, es_off, pack_sizes, n_packs, out, cap); }
uint64_t orc_qual_encode_qb01(const orc_qual_params* P, const uint8_t* bases, const uint8_t* quals, const uint64_t* offsets, uint32_t n_reads,
	const uint8_t* es, const uint64_t* es_off, const uint32_t* pack_sizes, uint32_t n_packs, uint8_t* out, uint64_t cap)
{ return qual_encode_v(1, P, bases, quals, offsets, n_reads, es, es_off, pack_sizes, n_packs, out, cap); }
static uint64_t qual_encode_v(int version, const orc_qual_params* P, const uint8_t* bases, const uint8_t* quals, const uint64_t* offsets, uint32_t n_reads,
	const uint8_t* es, const uint64_t* es_off, const uint32_t* pack_sizes, uint32_t n_packs, uint8_t* out, uint64_t cap)
{
	uint8_t map[96]; make_map(P, map);
	const uint32_t cbits = orc_qual_ctx_bits(P), nb = P->n_bins;
	const uint64_t n_ctx = 1ull << cbits;
	uint32_t* hist = (uint32_t*)calloc(n_ctx * nb, 4);
	uint32_t* mh = (uint32_t*)calloc(5 * 128, 4);
	const uint64_t tot = offsets[n_reads];
	uint8_t* sym = (uint8_t*)malloc(tot + 1); uint32_t* ctx = (uint32_t*)malloc(4 * (tot + 1)); uint32_t* avg = (uint32_t*)malloc(4 * 5 * ((size_t)n_reads + 1));
	for (uint32_t r = 0; r < n_reads; ++r)
	{
		const uint32_t n = (uint32_t)(offsets[r + 1] - offsets[r]);
		uint8_t* fl = (uint8_t*)calloc(n + 1, 1);
		if (P->level > 1 && es) base_flags(es + es_off[r], es_off[r + 1] - es_off[r], fl, n);
		read_symbols(P, map, bases + offsets[r], quals + offsets[r], n, fl, avg + 5 * (size_t)r, sym + offsets[r], ctx + offsets[r]);
		free(fl);
		for (uint32_t i = 0; i < n; ++i) ++hist[(uint64_t)ctx[offsets[r] + i] * nb + sym[offsets[r] + i]];
		for (uint32_t b = 0; b < nb; ++b) ++mh[b * 128 + ((avg[5 * (size_t)r + b] >> 8) & 127)];
	}
	/* dense contexts keep their own table; the counts of all the others are pooled per previous-symbols value */
	const uint32_t cb0 = bits_per_symbol(nb) * ctx_symbols(nb); const uint32_t n_fb = 1u << cb0;
	uint32_t* fbh = (uint32_t*)calloc((size_t)n_fb * nb, 4);
	uint8_t* dense = (uint8_t*)calloc(n_ctx, 1);
	for (uint64_t c = 0; c < n_ctx; ++c)
	{
		uint64_t t = 0; for (uint32_t k = 0; k < nb; ++k) t += hist[c * nb + k];
		if (t >= QB_MIN_CTX) dense[c] = 1;
		else for (uint32_t k = 0; k < nb; ++k) fbh[(c & (n_fb - 1)) * nb + k] += hist[c * nb + k];
	}
	uint16_t* freq = (uint16_t*)calloc(n_ctx * nb, 2); uint16_t* cum = (uint16_t*)calloc(n_ctx * nb, 2);
	uint16_t* fbf = (uint16_t*)calloc((size_t)n_fb * nb, 2);
	uint16_t mf[5 * 128], mc[5 * 128];
	for (uint32_t c = 0; c < n_fb; ++c) normalise(fbh + (size_t)c * nb, nb, fbf + (size_t)c * nb);
	for (uint64_t c = 0; c < n_ctx; ++c)
	{
		if (dense[c]) normalise(hist + c * nb, nb, freq + c * nb);
		else memcpy(freq + c * nb, fbf + (c & (n_fb - 1)) * nb, 2 * nb);
		uint32_t a = 0; for (uint32_t k = 0; k < nb; ++k) { cum[c * nb + k] = (uint16_t)a; a += freq[c * nb + k]; }
	}
	for (uint32_t b = 0; b < nb; ++b) { normalise(mh + b * 128, 128, mf + b * 128); uint32_t a = 0; for (uint32_t k = 0; k < 128; ++k) { mc[b * 128 + k] = (uint16_t)a; a += mf[b * 128 + k]; } }

	obuf o = {0, 0, 0};
	ob_put(&o, version == 1 ? "QB01" : "QB02", 4); ob_u32(&o, nb); ob_u32(&o, P->level); ob_put(&o, P->thr, 16); ob_u64(&o, n_reads); ob_u32(&o, n_packs); ob_u32(&o, cbits);
	ob_put(&o, mf, 2ull * nb * 128);
	for (uint32_t c = 0; c < n_fb; ++c) ob_put(&o, fbf + (size_t)c * nb, 2ull * (nb - 1));
	{
		uint32_t nd = 0; for (uint64_t c = 0; c < n_ctx; ++c) nd += dense[c];
		ob_u32(&o, nd);
		uint64_t prev = 0;
		for (uint64_t c = 0; c < n_ctx; ++c) if (dense[c])
		{
			uint64_t gap = c - prev; prev = c;
			do { uint8_t by = (uint8_t)(gap & 127); gap >>= 7; if (gap) by |= 128; ob_put(&o, &by, 1); } while (gap);
			ob_put(&o, freq + c * nb, 2ull * (nb - 1));
		}
	}
	free(fbh); free(fbf); free(dense);
	uint32_t r0 = 0;
	for (uint32_t p = 0; p < n_packs; ++p)
	{
		const uint32_t np = pack_sizes[p];
		ob_u32(&o, np);
		const uint64_t sizes_at = o.n;
		if (version == 2)
		{
			for (int l = 0; l < QB2_STREAMS; ++l) ob_u32(&o, 0);
			for (uint32_t l = 0; l < QB2_STREAMS; ++l)
			{
				uint64_t nsym = 0;
				for (uint32_t r = r0 + l; r < r0 + np; r += QB2_STREAMS) nsym += (offsets[r + 1] - offsets[r]) + 2ull * nb;
				uint16_t* w = (uint16_t*)malloc(2 * (nsym + 4)); uint64_t nw = 0;
				uint32_t x[QB2_STATES]; for (int k = 0; k < QB2_STATES; ++k) x[k] = QB_L;
				uint32_t last = r0 + l; while (last + QB2_STREAMS < r0 + np) last += QB2_STREAMS;
				if (r0 + l < r0 + np) for (int64_t r = last; r >= (int64_t)(r0 + l); r -= QB2_STREAMS)
				{	/* last read first, last symbol first; symbol k of the read -> state k mod QB2_STATES */
					const uint64_t at = offsets[r]; const uint32_t n = (uint32_t)(offsets[r + 1] - at), ns = 2 * nb;
					for (uint32_t k = ns + n; k-- > 0;)
					{
						uint32_t f, cf;
						if (k >= ns) { const uint64_t e = (uint64_t)ctx[at + k - ns] * nb + sym[at + k - ns]; f = freq[e]; cf = cum[e]; }
						else
						{
							const uint32_t b = k >> 1, a = avg[5 * (size_t)r + b], a1 = (a >> 8) & 127, a2 = a & 0xff;
							if (k & 1) { f = QB_M >> 8; cf = a2 * (QB_M >> 8); } else { f = mf[b * 128 + a1]; cf = mc[b * 128 + a1]; }
						}
						x[k % QB2_STATES] = rans_put(x[k % QB2_STATES], f, cf, w, &nw);
					}
				}
				const uint32_t bytes = (uint32_t)(4 * QB2_STATES + 2 * nw);
				memcpy(o.p + sizes_at + 4 * l, &bytes, 4);
				for (int k = 0; k < QB2_STATES; ++k) ob_u32(&o, x[k]);
				for (uint64_t k = nw; k-- > 0;) ob_put(&o, &w[k], 2);
				free(w);
			}
			r0 += np;
			continue;
		}
		for (int l = 0; l < QB_LANES; ++l) ob_u32(&o, 0);
		for (uint32_t l = 0; l < QB_LANES; ++l)
		{
			uint64_t nsym = 0;
			for (uint32_t r = r0 + l; r < r0 + np; r += QB_LANES) nsym += (offsets[r + 1] - offsets[r]) + 2ull * nb;
			uint16_t* w = (uint16_t*)malloc(2 * (nsym + 4)); uint64_t nw = 0;
			uint32_t x = QB_L;
			/* last read first, last symbol first */
			uint32_t last = r0 + l; while (last + QB_LANES < r0 + np) last += QB_LANES;
			if (r0 + l < r0 + np) for (int64_t r = last; r >= (int64_t)(r0 + l); r -= QB_LANES)
			{
				const uint64_t at = offsets[r]; const uint32_t n = (uint32_t)(offsets[r + 1] - at);
				for (uint32_t i = n; i-- > 0;) { const uint64_t k = (uint64_t)ctx[at + i] * nb + sym[at + i]; x = rans_put(x, freq[k], cum[k], w, &nw); }
				for (uint32_t b = nb; b-- > 0;)
				{
					const uint32_t a = avg[5 * (size_t)r + b], a1 = (a >> 8) & 127, a2 = a & 0xff;
					x = rans_put(x, QB_M >> 8, a2 * (QB_M >> 8), w, &nw);            /* low byte: uniform */
					x = rans_put(x, mf[b * 128 + a1], mc[b * 128 + a1], w, &nw);
				}
			}
			const uint32_t bytes = (uint32_t)(4 + 2 * nw);
			memcpy(o.p + sizes_at + 4 * l, &bytes, 4);
			ob_u32(&o, x);
			for (uint64_t k = nw; k-- > 0;) ob_put(&o, &w[k], 2);
			free(w);
		}
		r0 += np;
	}
	const uint64_t total = o.n;
	if (total <= cap) memcpy(out, o.p, total);
	free(o.p); free(hist); free(mh); free(sym); free(ctx); free(avg); free(freq); free(cum);
	return total;
}

/* Decoder of the native container: needs the reads' bases (and tuples when level > 1) exactly as the reference's decoder
 * does (decompression_common.cpp feeds decoded reads to CQualityCoder::Decode).  Writes the reconstructed qualities. */
int orc_qual_decode(const uint8_t* in, uint64_t in_n, const uint8_t* bases, const uint64_t* offsets, uint32_t n_reads,
	const uint8_t* es, const uint64_t* es_off, uint8_t* out)
{
	if (in_n < 48 || (memcmp(in, "QB01", 4) && memcmp(in, "QB02", 4))) return -1;
	const int version = in[3] == '1' ? 1 : 2;
	const uint32_t n_lanes = version == 1 ? QB_LANES : QB2_STREAMS, n_states = version == 1 ? 1 : QB2_STATES;
	orc_qual_params P; uint64_t at = 4;
	memcpy(&P.n_bins, in + at, 4); at += 4; memcpy(&P.level, in + at, 4); at += 4; memcpy(P.thr, in + at, 16); at += 16;
	uint64_t nr; memcpy(&nr, in + at, 8); at += 8; uint32_t n_packs, cbits; memcpy(&n_packs, in + at, 4); at += 4; memcpy(&cbits, in + at, 4); at += 4;
	if (nr != n_reads || cbits != orc_qual_ctx_bits(&P)) return -2;
	const uint32_t nb = P.n_bins; const uint64_t n_ctx = 1ull << cbits;
	uint16_t mf[5 * 128], mc[5 * 128]; (void)mc;
	memcpy(mf, in + at, 2ull * nb * 128); at += 2ull * nb * 128;
	for (uint32_t b = 0; b < nb; ++b) { uint32_t a = 0; for (uint32_t s = 0; s < 128; ++s) { mc[b * 128 + s] = (uint16_t)a; a += mf[b * 128 + s]; } }
	uint16_t* freq = (uint16_t*)calloc(n_ctx * nb, 2);
	{
		const uint32_t cb0 = bits_per_symbol(nb) * ctx_symbols(nb), n_fb = 1u << cb0;
		uint16_t* fbf = (uint16_t*)calloc((size_t)n_fb * nb, 2);
		for (uint32_t c = 0; c < n_fb; ++c)
		{
			uint32_t sm = 0; for (uint32_t k = 0; k + 1 < nb; ++k) { uint16_t v; memcpy(&v, in + at, 2); at += 2; fbf[(size_t)c * nb + k] = v; sm += v; }
			fbf[(size_t)c * nb + nb - 1] = (uint16_t)(sm || 1 ? QB_M - sm : 0);
		}
		for (uint64_t c = 0; c < n_ctx; ++c) memcpy(freq + c * nb, fbf + (c & (n_fb - 1)) * nb, 2 * nb);
		free(fbf);
		uint32_t nd; memcpy(&nd, in + at, 4); at += 4;
		uint64_t c = 0;
		for (uint32_t d = 0; d < nd; ++d)
		{
			uint64_t gap = 0; uint32_t sh = 0; uint8_t by;
			do { by = in[at++]; gap |= (uint64_t)(by & 127) << sh; sh += 7; } while (by & 128);
			c += gap;
			uint32_t sm = 0; for (uint32_t k = 0; k + 1 < nb; ++k) { uint16_t v; memcpy(&v, in + at, 2); at += 2; freq[c * nb + k] = v; sm += v; }
			freq[c * nb + nb - 1] = (uint16_t)(QB_M - sm);
		}
	}
	const uint32_t bps = bits_per_symbol(nb), cb = bps * ctx_symbols(nb), cmask = (1u << cb) - 1;
	static const uint8_t code[256] = { ['A'] = 0, ['C'] = 1, ['G'] = 2, ['T'] = 3, ['N'] = 0 };
	uint32_t r0 = 0;
	for (uint32_t p = 0; p < n_packs; ++p)
	{
		uint32_t np; memcpy(&np, in + at, 4); at += 4;
		uint32_t sizes[QB_LANES]; memcpy(sizes, in + at, 4 * n_lanes); at += 4 * n_lanes;
		for (uint32_t l = 0; l < n_lanes; ++l)
		{
			const uint8_t* s = in + at; at += sizes[l];
			uint32_t xs[QB2_STATES]; memcpy(xs, s, 4 * n_states); uint64_t rp = 4 * n_states;
			for (uint32_t r = r0 + l; r < r0 + np; r += n_lanes)
			{
				uint32_t kk = 0;                          /* symbol index inside the read: picks the state */
#define x xs[(kk) % n_states]
				const uint64_t o = offsets[r]; const uint32_t n = (uint32_t)(offsets[r + 1] - o);
				uint32_t avg16[5];
				for (uint32_t b = 0; b < nb; ++b)
				{
					uint32_t slot = x & (QB_M - 1), a1 = 0, acc = 0;
					while (acc + mf[b * 128 + a1] <= slot) { acc += mf[b * 128 + a1]; ++a1; }
					x = mf[b * 128 + a1] * (x >> QB_PROB_BITS) + slot - acc;
					while (x < QB_L) { uint16_t w; memcpy(&w, s + rp, 2); rp += 2; x = (x << 16) | w; }
					++kk;
					slot = x & (QB_M - 1);
					const uint32_t a2 = slot / (QB_M >> 8);
					x = (QB_M >> 8) * (x >> QB_PROB_BITS) + slot - a2 * (QB_M >> 8);
					while (x < QB_L) { uint16_t w; memcpy(&w, s + rp, 2); rp += 2; x = (x << 16) | w; }
					++kk;
					avg16[b] = (a1 << 8) | a2;
				}
				uint8_t* fl = (uint8_t*)calloc(n + 1, 1);
				if (P.level > 1 && es) base_flags(es + es_off[r], es_off[r + 1] - es_off[r], fl, n);
				uint8_t* sym = (uint8_t*)malloc(n + 1);
				uint32_t c = cmask, dna = n ? code[bases[o]] : 0;
				for (uint32_t i = 0; i < n; ++i)
				{
					dna <<= 2; if (i + 1 < n) dna += code[bases[o + i + 1]]; dna &= 0xff;
					uint32_t cx = c + (dna << cb);
					if (P.level > 1) cx += (uint32_t)(fl[i] == 1) << (cb + 8), cx += (uint32_t)(fl[i] == 2) << (cb + 9);
					const uint16_t* f = freq + (uint64_t)cx * nb;
					const uint32_t slot = x & (QB_M - 1);
					uint32_t d = 0, a = 0;
					while (a + f[d] <= slot) { a += f[d]; ++d; }
					x = f[d] * (x >> QB_PROB_BITS) + slot - a;
					while (x < QB_L) { uint16_t w; memcpy(&w, s + rp, 2); rp += 2; x = (x << 16) | w; }
					++kk;
					sym[i] = (uint8_t)d;
					c = ((c << bps) + d) & cmask;
				}
#undef x
				reconstruct(&P, avg16, sym, n, out + o);
				free(sym); free(fl);
			}
		}
		r0 += np;
	}
	free(freq);
	return 0;
}
