/* stage3_exact.c — TEST INFRASTRUCTURE ONLY (oracle; never linked into or called by the product).
 *
 * Serial CPU restatement of the reference's three entropy coders as they run in `colord compress-*`: adaptive frequency
 * models looked up by context, one 64-bit range coder per stream that is restarted for every pack while the models live on.
 * The byte streams produced here are the reference's own archive parts ("dna", "qual", "header"): pinned against parts
 * written by the unmodified reference (tests/golden/<case>/streams.json, tests/test_oracle_exact.py).
 *
 *   range coder          src/colord/sub_rc.h:72-211        (Start, EncodeFrequency with the unrolled <= 8 byte renormalisation, End)
 *   models               src/colord/rc.h:34-221 (CSimpleModel), :225-480 (fixed size), :487-740 (Fenwick = the same counts),
 *                        Encode / EncodeExcluding rc.h:780-803, :861-893
 *   DNA stream           src/colord/dna_coder.cpp:26-231 (Encode), :440-1239 (events), dna_coder.h:48-60 (model parameters),
 *                        driver entr_read.h:56-80 (Finish / Restart per pack)
 *   quality stream       src/colord/quality_coder.cpp:26-262, :560-604, quality_coder_impl.cpp:25-450, :821-834, driver entr_qual.h:100-126
 *   header stream        src/colord/id_coder.cpp:169-383, id_coder.h:50-59, driver entr_header.cpp:23-46
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------ output + range coder */
typedef struct { uint8_t* p; uint64_t n, cap; } xbuf;
static void xb_put(xbuf* b, uint8_t x) { if (b->n == b->cap) { b->cap = b->cap * 2 + 4096; b->p = (uint8_t*)realloc(b->p, b->cap); } b->p[b->n++] = x; }

typedef struct { uint64_t low, range; xbuf* out; } xrc;
static void rc_start(xrc* r) { r->low = 0; r->range = 0xff00000000000000ULL; }                                   /* sub_rc.h:72-76 */
static void rc_encode(xrc* r, uint32_t freq, uint32_t cum, uint32_t tot)                                           /* sub_rc.h:83-201 */
{
	r->range /= tot;
	r->low += r->range * cum;
	r->range *= freq;
	for (int k = 0; k < 8 && r->range <= 0x00ffffffffffffULL; ++k) {         /* UNROLL_FREQUENCY_CODING + RC_64BIT: at most 8 bytes */
		if ((r->low ^ (r->low + r->range)) & 0xff00000000000000ULL) { const uint64_t x = r->low; r->range = (x | 0x00ffffffffffffULL) - x; }
		xb_put(r->out, (uint8_t)(r->low >> 56));
		r->low <<= 8; r->range <<= 8;
	}
}
static void rc_end(xrc* r) { for (int i = 0; i < 8; ++i) { xb_put(r->out, (uint8_t)(r->low >> 56)); r->low <<= 8; } }   /* sub_rc.h:203-210 */

/* ------------------------------------------------------------------------------------------------ adaptive models by (family, context) */
typedef struct { uint32_t n_sym, max_total, adder; } xfam;
typedef struct { uint64_t ctx; uint32_t fam, used; uint64_t at; } xslot;      /* at: first counter in the arena; [n_sym] = total */
typedef struct { const xfam* fam; xslot* tab; uint64_t cap, n; uint32_t* arena; uint64_t an, acap; } xmodels;

static void xm_init(xmodels* m, const xfam* fam) { memset(m, 0, sizeof *m); m->fam = fam; m->cap = 1u << 12; m->tab = (xslot*)calloc(m->cap, sizeof(xslot)); }
static void xm_free(xmodels* m) { free(m->tab); free(m->arena); }
static uint64_t xm_hash(uint32_t f, uint64_t c) { uint64_t h = c * 0x9E3779B97F4A7C15ULL + f * 0xC2B2AE3D27D4EB4FULL; h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ULL; return h ^ (h >> 32); }
static uint32_t* xm_get(xmodels* m, uint32_t f, uint64_t ctx)
{
	if (m->n * 2 >= m->cap) {
		xslot* old = m->tab; const uint64_t oc = m->cap;
		m->cap *= 2; m->tab = (xslot*)calloc(m->cap, sizeof(xslot));
		for (uint64_t i = 0; i < oc; ++i) if (old[i].used) { uint64_t h = xm_hash(old[i].fam, old[i].ctx) & (m->cap - 1); while (m->tab[h].used) h = (h + 1) & (m->cap - 1); m->tab[h] = old[i]; }
		free(old);
	}
	uint64_t h = xm_hash(f, ctx) & (m->cap - 1);
	while (m->tab[h].used) { if (m->tab[h].fam == f && m->tab[h].ctx == ctx) return m->arena + m->tab[h].at; h = (h + 1) & (m->cap - 1); }
	const uint32_t A = m->fam[f].n_sym;
	if (m->an + A + 1 > m->acap) { m->acap = (m->an + A + 1) * 2 + 4096; m->arena = (uint32_t*)realloc(m->arena, m->acap * sizeof(uint32_t)); }
	m->tab[h].used = 1; m->tab[h].fam = f; m->tab[h].ctx = ctx; m->tab[h].at = m->an; ++m->n;
	uint32_t* c = m->arena + m->an; m->an += A + 1;
	for (uint32_t i = 0; i < A; ++i) c[i] = 1;                                     /* Init(nullptr): all counts 1 (rc.h:122-124, :329-331, :706) */
	c[A] = A;
	return c;
}
/* Encode / EncodeExcluding (rc.h:780-803, :861-893): excl = bit mask of the symbols left out of the alphabet for this one event */
static void xm_encode(xmodels* m, xrc* rc, uint32_t f, uint64_t ctx, uint32_t sym, uint32_t excl)
{
	const xfam* F = &m->fam[f]; const uint32_t A = F->n_sym;
	uint32_t* c = xm_get(m, f, ctx);
	uint32_t cum = 0, tot = c[A];
	for (uint32_t i = 0; i < sym; ++i) if (!(i < 32 && (excl >> i & 1))) cum += c[i];
	for (uint32_t i = 0; i < A && i < 32; ++i) if (excl >> i & 1) tot -= c[i];
	rc_encode(rc, c[sym], cum, tot);
	c[sym] += F->adder; c[A] += F->adder;                                        /* Update (rc.h:178-185) */
	while (c[A] >= F->max_total) { uint32_t t = 0; for (uint32_t i = 0; i < A; ++i) { c[i] = (c[i] + 1) / 2; t += c[i]; } c[A] = t; }
}

static uint32_t ilog2x(uint64_t x) { uint32_t r = 0; for (; x; ++r) x >>= 1; return r; }                          /* basic_coder.h:37-45 */
static uint32_t nbytesx(uint64_t x) { uint32_t r = 1; x >>= 8; for (; x; ++r) x >>= 8; return r; }               /* basic_coder.h:48-58 */

/* every part of a stream: out = parts back to back, part_sizes[p] = bytes of part p */
typedef struct { xbuf b; uint64_t* part_sizes; uint32_t n_parts; uint64_t part_start; xrc rc; } xstream;
static void xs_begin(xstream* s, uint64_t* part_sizes) { memset(s, 0, sizeof *s); s->part_sizes = part_sizes; s->rc.out = &s->b; rc_start(&s->rc); }
static void xs_part(xstream* s) { rc_end(&s->rc); s->part_sizes[s->n_parts++] = s->b.n - s->part_start; s->part_start = s->b.n; rc_start(&s->rc); }
static int64_t xs_finish(xstream* s, uint8_t* out, uint64_t cap) { const uint64_t n = s->b.n; if (n <= cap && n) memcpy(out, s->b.p, n); free(s->b.p); return n <= cap ? (int64_t)n : -(int64_t)n; }

/* ================================================================================================ DNA stream */
enum { D_FLAG, D_LENBITS, D_LENDATA, D_SYM, D_SYMN, D_READID, D_REV, D_TUPLE, D_ANCHOR, D_SKIPL, D_SKIPD, D_SEEN, D_SHORT, D_COUNT };
enum { T_INS = 0, T_DEL, T_MATCH, T_SUBST, T_ANCHOR, T_SKIP, T_ALT, T_MAIN, T_PLAIN, T_START_PLAIN, T_START_ES, T_START_PLAIN_N };

typedef struct { const uint8_t* bases; const uint64_t* off; const uint32_t* ref_to_read; } xreads;
static uint32_t code_of(uint8_t ch) { return ch == 'C' ? 1u : ch == 'G' ? 2u : ch == 'T' ? 3u : 0u; }
typedef struct { const uint8_t* b; uint32_t len; int rev; } xoref;
static xoref xoriented(const xreads* R, uint32_t ref_id, int rev) { const uint32_t r = R->ref_to_read[ref_id]; xoref o = {R->bases + R->off[r], (uint32_t)(R->off[r + 1] - R->off[r]), rev}; return o; }
/* GetRefRead (reference_reads.h:27-72): symbols 0..3 forward or reverse-complemented, guard 255 behind the last one */
static uint32_t xosym(const xoref* o, int pos) { if (pos < 0 || (uint32_t)pos >= o->len) return 255u; return o->rev ? 3u - code_of(o->b[o->len - 1 - pos]) : code_of(o->b[pos]); }
static uint32_t be32x(const uint8_t* t) { return ((uint32_t)t[0] << 24) | ((uint32_t)t[1] << 16) | ((uint32_t)t[2] << 8) | t[3]; }

typedef struct { xmodels M; xrc* rc; uint32_t level, n_t, n_s; uint64_t mask_t, mask_s; uint64_t ctx_read_type; uint32_t cur_read_id; } xdna;

static void xd_read_id(xdna* D, uint32_t id)                                                                       /* dna_coder.cpp:535-551 */
{
	const int n = (int)nbytesx(D->cur_read_id);
	for (int i = n - 1; i >= 0; --i) { const uint64_t add = (i == n - 2) ? ((id >> (8 * (n - 1))) & 0xff) : 0; xm_encode(&D->M, D->rc, D_READID, (uint64_t)i + (add << 3), (id >> (8 * i)) & 0xff, 0); }
}
static void xd_skip(xdna* D, uint32_t len, int local)                                                              /* dna_coder.cpp:1102-1137 */
{
	if (local) { for (uint32_t part = 0; len; ++part) { if (len < 255) { xm_encode(&D->M, D->rc, D_SKIPL, part, len, 0); break; } xm_encode(&D->M, D->rc, D_SKIPL, part, 255, 0); len -= 254; } }
	else { uint32_t enc = 0; for (int i = 3; i >= 0; --i) { const uint32_t x = (len >> (8 * i)) & 0xff; xm_encode(&D->M, D->rc, D_SKIPD, (uint64_t)i * 64 + ilog2x(enc), x, 0); enc = (enc << 8) + x; } }
}
static void xd_read(xdna* D, const xreads* R, const uint8_t* t, uint64_t tn)                                       /* dna_coder.cpp:26-231 */
{
	uint64_t ctx_tuple = D->mask_t, ctx_symbol = D->mask_s, ctx_rev = 0xf;
	uint32_t n_tuples = 0;
	for (uint64_t p = 0; p < tn; ++n_tuples) { const uint32_t ty = t[p] >> 4; p += (ty == T_ANCHOR || ty == T_SKIP) ? 4 : (ty == T_ALT || ty == T_START_ES) ? 5 : 1; }
	const uint32_t t0 = t[0] >> 4, flag = t0 == T_START_PLAIN ? 0u : t0 == T_START_PLAIN_N ? 1u : 2u;
	xm_encode(&D->M, D->rc, D_FLAG, D->ctx_read_type, flag, 0);                                                     /* :440-463 */
	D->ctx_read_type = ((D->ctx_read_type << 2) + flag) & 0xff;
	{	/* encode_read_len(es.size() - 1) :1004-1056 */
		uint32_t len = n_tuples - 1; const uint32_t nbits = ilog2x(len);
		xm_encode(&D->M, D->rc, D_LENBITS, 0, nbits, 0);
		if (nbits >= 2) {
			uint64_t ctx = (uint64_t)nbits << 3;
			len -= 1u << (nbits - 1);
			uint32_t prefix = len, suffix = 0;
			if (nbits > 9) { prefix = len >> (nbits - 9); suffix = len - (prefix << (nbits - 9)); }
			xm_encode(&D->M, D->rc, D_LENDATA, ctx, prefix, 0);
			if (nbits > 9) { ctx += 4; for (int nb = (int)nbits - 9; nb > 0; nb -= 8) { xm_encode(&D->M, D->rc, D_LENDATA, ctx, suffix & 0xff, 0); suffix >>= 8; ++ctx; } }
		}
	}
	if (flag == 0) { for (uint64_t p = 1; p < tn; ++p) { const uint32_t s = t[p] & 15; xm_encode(&D->M, D->rc, D_SYM, ctx_symbol << 2, s, 0); ctx_symbol = ((ctx_symbol << 2) + s) & D->mask_s; } ++D->cur_read_id; return; }
	if (flag == 1) { for (uint64_t p = 1; p < tn; ++p) { const uint32_t s = t[p] & 15; xm_encode(&D->M, D->rc, D_SYMN, ctx_symbol, s, 0); ctx_symbol = ((ctx_symbol << 4) + s) & D->mask_s; } ++D->cur_read_id; return; }

	uint32_t seen_id[64], n_seen = 0;                                           /* uo_rev_comp: at most 1 + max_candidates distinct reference reads per read */
	uint32_t alt_ids[64], alt_rev[64]; int alt_pos_of[64]; uint32_t n_alt = 0; int cur_alt = -1;      /* m_alt_ids (short id = insertion order), m_alt_read, m_alt_pos */
#define PUT_REV(id_, rev_) do { int seen_ = 0; for (uint32_t k_ = 0; k_ < n_seen; ++k_) if (seen_id[k_] == (id_)) seen_ = 1; \
		if (!seen_) { xm_encode(&D->M, D->rc, D_REV, ctx_rev, (rev_) ? 1u : 0u, 0); if (n_seen < 64) seen_id[n_seen++] = (id_); ctx_rev = ((ctx_rev << 2) + ((rev_) ? 1u : 0u)) & 0xf; } } while (0)
	const uint32_t main_id = be32x(t + 1), main_rev = t[0] & 15;
	xd_read_id(D, main_id);
	PUT_REV(main_id, main_rev);
	const xoref main_ref = xoriented(R, main_id, (int)main_rev); xoref alt_ref = main_ref;
	int ref_pos = 0, alt_pos = 0, delta = 0, is_main = 1, first = 1; uint32_t last = 255;
	const uint32_t sh_t = 3 * D->n_t;
	for (uint64_t p = 5; p < tn;) {
		const uint32_t ty = t[p] >> 4, v1 = t[p] & 15; uint32_t v2 = 0;
		if (ty == T_ANCHOR || ty == T_SKIP) { v2 = (v1 << 24) | ((uint32_t)t[p + 1] << 16) | ((uint32_t)t[p + 2] << 8) | t[p + 3]; p += 4; }
		else if (ty == T_ALT) { v2 = be32x(t + p + 1); p += 5; }
		else p += 1;
		const uint32_t rsym = is_main ? xosym(&main_ref, ref_pos) : xosym(&alt_ref, alt_pos);
		{	/* encode_tuple_type :651-717 */
			uint64_t ctx = ctx_tuple + ((ctx_symbol & 0xf) << sh_t) + ((uint64_t)rsym << (sh_t + 4));
			ctx += (uint64_t)(delta < -10 ? 1 : delta < -1 ? 2 : delta > 10 ? 3 : delta > 1 ? 4 : 0) << (sh_t + 6);
			uint32_t excl = 0;
			if (!first) excl = last == T_MATCH ? 1u << T_ANCHOR : last == T_DEL ? 1u << T_SKIP : last == T_ANCHOR ? (1u << T_ANCHOR) | (1u << T_MATCH)
				: last == T_SKIP ? (1u << T_DEL) | (1u << T_SKIP) : (last == T_MAIN || last == T_ALT) ? (1u << T_ALT) | (1u << T_MAIN) : 0;
			xm_encode(&D->M, D->rc, D_TUPLE, ctx, ty, excl);
			ctx_tuple = ((ctx_tuple << 3) + ty) & D->mask_t;
			first = 0;
		}
		if (ty == T_ALT) {
			if (!is_main && cur_alt >= 0) alt_pos_of[cur_alt] = alt_pos;
			int idx = -1;
			for (uint32_t k = 0; k < n_alt; ++k) if (alt_ids[k] == v2) { idx = (int)k; break; }
			if (n_alt == 0) xd_read_id(D, v2);                                                                        /* :572-615 */
			else {
				xm_encode(&D->M, D->rc, D_SEEN, n_alt, idx >= 0, 0);
				if (idx < 0) xd_read_id(D, v2); else xm_encode(&D->M, D->rc, D_SHORT, n_alt, (uint32_t)idx, 0);
			}
			if (idx < 0 && n_alt < 64) { idx = (int)n_alt; alt_ids[n_alt] = v2; alt_rev[n_alt] = v1; alt_pos_of[n_alt] = 0; ++n_alt; }
			cur_alt = idx;
			PUT_REV(v2, v1);
			alt_ref = xoriented(R, v2, (int)alt_rev[idx]);
			alt_pos = 0; is_main = 0; delta = 0;
		} else if (ty == T_ANCHOR) {
			for (uint32_t len = v2, part = 0; len; ++part) { if (len < 23) { xm_encode(&D->M, D->rc, D_ANCHOR, part, len, 0); break; } xm_encode(&D->M, D->rc, D_ANCHOR, part, 23, 0); len -= 22; }   /* :958-978 */
			int* pos = is_main ? &ref_pos : &alt_pos; const xoref* o = is_main ? &main_ref : &alt_ref;
			*pos += (int)v2;
			for (int i = (int)D->n_s; i > 0; --i) ctx_symbol = (ctx_symbol << 2) + xosym(o, *pos - i);
			ctx_symbol &= D->mask_s; delta = 0;
		} else if (ty == T_MATCH) {
			ctx_symbol = ((ctx_symbol << 2) + rsym) & D->mask_s;
			if (is_main) ++ref_pos; else ++alt_pos;
		} else if (ty == T_INS) {                                                                                      /* :772-811 */
			uint64_t ctx = 2; uint32_t sh = 2;
			if (D->level == 1) { ctx += (ctx_symbol & 0xff) << sh; sh += 8; }
			else { ctx += (ctx_symbol & 0x3ff) << sh; sh += 10; if (D->level >= 3) { ctx += (uint64_t)(((ctx_symbol >> 10) & 3) == ((ctx_symbol >> 8) & 3)) << sh; ++sh; } }
			ctx += (uint64_t)rsym << sh; sh += 2;
			ctx += (ctx_tuple & 0777) << sh;
			xm_encode(&D->M, D->rc, D_SYM, ctx, v1, 0);
			ctx_symbol = ((ctx_symbol << 2) + v1) & D->mask_s; ++delta;
		} else if (ty == T_DEL) { if (is_main) ++ref_pos; else ++alt_pos; --delta; }
		else if (ty == T_SUBST) {                                                                                      /* :889-922; subst_to_code dna_coder.h:37 */
			static const uint32_t subst_to_code[4][4] = {{1, 0, 0, 0}, {2, 2, 1, 1}, {3, 3, 3, 2}, {3, 3, 3, 3}};
			const uint32_t symbol = subst_to_code[v1 & 3][rsym & 3];
			uint64_t ctx = 1; uint32_t sh = 2;
			ctx += (ctx_symbol & 0x3f) << sh; sh += 6;
			if (D->level >= 3) { ctx += (uint64_t)(((ctx_symbol >> 6) & 3) == ((ctx_symbol >> 4) & 3)) << sh; ++sh; }
			ctx += (uint64_t)rsym << sh; sh += 2;
			ctx += (ctx_tuple & 07777) << sh;
			xm_encode(&D->M, D->rc, D_SYM, ctx, symbol, 1u << (rsym & 3));
			ctx_symbol = ((ctx_symbol << 2) + symbol) & D->mask_s;
			if (is_main) ++ref_pos; else ++alt_pos;
		} else if (ty == T_SKIP) {                                                                                     /* :166-206 */
			const int skip_len = (int)v2;
			delta -= skip_len;
			if (!is_main && last == T_ALT) {
				const int mod = skip_len - (cur_alt >= 0 ? alt_pos_of[cur_alt] : 0);
				if (mod > 0) xd_skip(D, (uint32_t)mod, 0);
				else { xd_skip(D, 0, 0); xd_skip(D, (uint32_t)(-mod), 0); }
			} else xd_skip(D, (uint32_t)skip_len, last != T_ALT && last != 255);
			if (is_main) ref_pos += skip_len; else alt_pos += skip_len;
		} else if (ty == T_MAIN) { is_main = 1; if (cur_alt >= 0) alt_pos_of[cur_alt] = alt_pos; delta = 0; }
		last = ty;
	}
#undef PUT_REV
	++D->cur_read_id;
}

/* All reads of the input in order; is_ref[r]: read r became a reference read (ids count those in input order); reads
 * [0, n_skip) are not coded (pseudo-reads of a reference genome).  Returns the total bytes (negative: out too small). */
int64_t orc_xdna_encode(uint32_t level, uint32_t max_cand, uint32_t n_skip, const uint8_t* es, const uint64_t* es_off, const uint8_t* bases, const uint64_t* off,
	const uint8_t* is_ref, uint32_t n_reads, const uint32_t* pack_sizes, uint32_t n_packs, uint8_t* out, uint64_t out_cap, uint64_t* part_sizes)
{
	static xfam fam[D_COUNT];
	const xfam f0[D_COUNT] = {{3, 1u << 15, 1}, {32, 1u << 18, 8}, {256, 1u << 18, 8}, {4, 1u << 10, 1}, {5, 1u << 10, 1}, {256, 1u << 13, 1}, {2, 1u << 15, 1},
		{8, 1u << 15, 1}, {24, 1u << 15, 1}, {256, 1u << 15, 1}, {256, 1u << 15, 1}, {2, 1u << 15, 1}, {max_cand, 1u << 13, 1}};      /* dna_coder.h:48-60, dna_coder.cpp:1316-1336 */
	memcpy(fam, f0, sizeof f0);
	uint32_t* ref_to_read = (uint32_t*)malloc(sizeof(uint32_t) * (n_reads + 1)); uint32_t nr = 0;
	for (uint32_t r = 0; r < n_reads; ++r) if (is_ref[r]) ref_to_read[nr++] = r;
	const xreads R = {bases, off, ref_to_read};
	xstream S; xs_begin(&S, part_sizes);
	xdna D; memset(&D, 0, sizeof D);
	xm_init(&D.M, fam); D.rc = &S.rc; D.level = level;
	D.n_t = level >= 3 ? 4 : level == 2 ? 3 : level == 1 ? 2 : 1; D.n_s = level >= 3 ? 8 : level == 2 ? 7 : level == 1 ? 5 : 1;      /* dna_coder.cpp:1253-1280 */
	D.mask_t = (1ull << (3 * D.n_t)) - 1; D.mask_s = (1ull << (2 * D.n_s)) - 1;
	D.cur_read_id = n_skip;
	uint32_t r = n_skip;
	for (uint32_t p = 0; p < n_packs; ++p) {
		for (uint32_t k = 0; k < pack_sizes[p]; ++k, ++r) xd_read(&D, &R, es + es_off[r], es_off[r + 1] - es_off[r]);
		xs_part(&S);
	}
	xm_free(&D.M); free(ref_to_read);
	return xs_finish(&S, out, out_cap);
}

/* ================================================================================================ quality stream */
enum { Q_SYM, Q_BYTE, Q_COUNT };
/* per-base flags of analyze_es (quality_coder_impl.cpp:25-76): 1 = 'M', 2 = 'A', 0 otherwise (plain reads: 'P' -> 0) */
static void xq_flags(const uint8_t* t, uint64_t tn, uint8_t* fl, uint32_t n)
{
	memset(fl, 0, n);
	const uint32_t t0 = t[0] >> 4;
	if (t0 == T_START_PLAIN || t0 == T_START_PLAIN_N) return;
	uint32_t at = 0;
	for (uint64_t p = 5; p < tn;) {
		const uint32_t ty = t[p] >> 4;
		if (ty == T_ANCHOR) { const uint32_t len = ((uint32_t)(t[p] & 15) << 24) | ((uint32_t)t[p + 1] << 16) | ((uint32_t)t[p + 2] << 8) | t[p + 3]; for (uint32_t k = 0; k < len && at < n; ++k) fl[at++] = 2; p += 4; }
		else if (ty == T_SKIP) p += 4;
		else if (ty == T_ALT) p += 5;
		else { if (ty == T_MATCH) { if (at < n) fl[at] = 1; ++at; } else if (ty == T_INS || ty == T_SUBST) ++at; p += 1; }
	}
}
/* lossless quantisers (quality_coder.cpp:276-338 ONT, :356-420 PacBio CLR, :441-505 PacBio HiFi) */
static void xq_quantize(uint32_t source, uint32_t level, uint32_t* q /*96*/)
{
	memset(q, 0, 96 * sizeof(uint32_t));
#define FILL(a, b, v) do { for (int i_ = (a); i_ < (b); ++i_) q[i_] = (uint32_t)(v); } while (0)
	if (source == 0) {
		if (level >= 3) { static const int e[] = {0, 1, 2, 4, 7, 11, 16, 22, 29, 37, 46, 56, 67, 79, 90, 96}; for (int k = 0; k < 15; ++k) FILL(e[k], e[k + 1], k); }
		else { static const int e[] = {0, 1, 2, 5, 10, 15, 20, 25, 35, 50, 70, 96}; for (int k = 0; k < 11; ++k) FILL(e[k], e[k + 1], k); }
	} else {
		const int s = source == 2 ? 1 : 0;
		if (level >= 3) { static const int e[] = {1, 10, 20, 30, 39, 45, 51, 57, 63, 69, 75, 81, 87, 93}; q[0] = (uint32_t)s; for (int k = 0; k < 13; ++k) FILL(e[k], e[k + 1], k + 1 + s); q[93] = s ? 0u : 14u; }
		else { static const int e[] = {1, 15, 29, 41, 53, 63, 72, 80, 87, 93}; q[0] = (uint32_t)s; for (int k = 0; k < 9; ++k) FILL(e[k], e[k + 1], k + 1 + s); q[93] = s ? 0u : 10u; }
	}
#undef FILL
}
/* mode: params.h QualityComprMode — 0 original, 1 quinary average, 2 quad average, 3 binary average, 4 quinary threshold,
 * 5 quad threshold, 6 binary threshold, 7 average, 8 none.  source: 0 ONT, 1 PacBio CLR, 2 PacBio HiFi.  thr: forward thresholds. */
int64_t orc_xqual_encode(uint32_t mode, uint32_t source, uint32_t level, const uint32_t* thr, const uint8_t* bases, const uint8_t* quals, const uint64_t* off,
	const uint8_t* es, const uint64_t* es_off, uint32_t n_reads, const uint32_t* pack_sizes, uint32_t n_packs, uint8_t* out, uint64_t out_cap, uint64_t* part_sizes)
{
	const uint32_t n_bins = (mode == 1 || mode == 4) ? 5 : (mode == 2 || mode == 5) ? 4 : (mode == 3 || mode == 6) ? 2 : 0;
	uint32_t bps, ncs;                                                           /* no_bits_per_symbol, no_ctx_symbols (quality_coder.cpp:59-262) */
	if (mode == 0) { bps = 4; ncs = 2; } else if (mode == 7) { bps = 8; ncs = 2; } else if (n_bins == 2) { bps = 2; ncs = 6; } else { bps = 3; ncs = 3; }
	const uint32_t cbits = bps * ncs; const uint64_t cmask = (1ull << cbits) - 1;
	static xfam fam[Q_COUNT];
	fam[Q_SYM].n_sym = mode == 0 ? 96 : n_bins ? n_bins : 2; fam[Q_SYM].max_total = mode == 0 ? 1u << 20 : 1u << 18; fam[Q_SYM].adder = mode == 0 ? 32 : 8;      /* quality_coder.h:35-39 */
	fam[Q_BYTE].n_sym = 256; fam[Q_BYTE].max_total = 1u << 18; fam[Q_BYTE].adder = 8;
	uint32_t map[96], quant[96];
	if (mode == 0) { for (int i = 0; i < 96; ++i) map[i] = (uint32_t)i; xq_quantize(source, level, quant); }
	else if (n_bins) {                                                           /* adjust_quality_map_symbols quality_coder.cpp:264-283 */
		memset(map, 0, sizeof map);
		for (uint32_t bin = 1; bin + 1 < n_bins; ++bin) for (uint32_t i = thr[bin - 1]; i < thr[bin] && i < 96; ++i) map[i] = bin;
		for (uint32_t i = thr[n_bins - 2]; i < 96; ++i) map[i] = n_bins - 1;
	}
	xstream S; xs_begin(&S, part_sizes);
	xmodels M; xm_init(&M, fam);
	uint8_t* fl = NULL; uint64_t fl_cap = 0;
	uint32_t r = 0;
	for (uint32_t p = 0; p < n_packs; ++p) {
		for (uint32_t k = 0; k < pack_sizes[p]; ++k, ++r) {
			if (mode == 8) continue;
			const uint8_t* b = bases + off[r]; const uint8_t* q = quals + off[r]; const uint32_t n = (uint32_t)(off[r + 1] - off[r]);
			if (level > 1) { if (n > fl_cap) { fl_cap = n * 2ull + 64; fl = (uint8_t*)realloc(fl, fl_cap); } xq_flags(es + es_off[r], es_off[r + 1] - es_off[r], fl, n); }
#define VS(i_) ((uint64_t)(code_of(b[i_])))                                      /* valid_sym: x & 3 on symbols 0..4 (N -> 0) */
#define AVG(ctx_, x_) do { const uint32_t a_ = (uint32_t)((x_) * 256); xm_encode(&M, &S.rc, Q_BYTE, (ctx_), a_ >> 8, 0); xm_encode(&M, &S.rc, Q_BYTE, (uint64_t)(a_ >> 8) + 0x100ull, a_ & 0xff, 0); } while (0)
			if (mode == 7) {                                                         /* encode_average impl:441-450 */
				double avg = 0.0; for (uint32_t i = 0; i < n; ++i) avg += q[i] - 33u; avg /= n;
				AVG(0ull, avg);
				continue;
			}
			uint64_t context = cmask;
			if (mode >= 1 && mode <= 3) {                                            /* encode_*_average impl:130-310 */
				double sum[5] = {0, 0, 0, 0, 0}; uint32_t cnt[5] = {0, 0, 0, 0, 0}, st[128];
				memset(st, 0, sizeof st);
				for (uint32_t i = 0; i < n; ++i) ++st[q[i] & 127];
				for (uint32_t i = 33; i < 128; ++i) { sum[map[i - 33]] += (double)(i - 33u) * st[i]; cnt[map[i - 33]] += st[i]; }
				uint64_t ctx_p = 0;
				for (uint32_t i = 0; i < n_bins; ++i) { const double avg = cnt[i] ? sum[i] / cnt[i] : 0.0; AVG((1ull << 30) + ((uint64_t)i << 24) + (ctx_p << 16), avg); ctx_p = (uint64_t)avg; }
				uint64_t dna = n ? VS(0) : 3;                                            /* read[0] of an empty read is the guard 255 */
				for (uint32_t i = 0; i < n; ++i) {
					uint64_t ctx = context; uint32_t sh = cbits;
					dna <<= 2; if (i + 1 < n) dna += VS(i + 1); dna &= 0xff;
					ctx += dna << sh; sh += 8;
					if (level > 1) { ctx += (uint64_t)(fl[i] == 1) << sh; ++sh; ctx += (uint64_t)(fl[i] == 2) << sh; }
					const uint32_t s = map[q[i] - 33];
					xm_encode(&M, &S.rc, Q_SYM, ctx, s, 0);
					context = ((context << bps) + s) & cmask;
				}
				continue;
			}
			for (uint32_t i = 0; i < n; ++i) {                                       /* encode_original impl:78-128, encode_*_threshold impl:312-438 */
				uint64_t ctx = context; uint32_t sh = cbits;
				ctx += VS(i) << sh; sh += 2;
				if (i > 0) ctx += VS(i - 1) << sh;
				sh += 2;
				if (mode != 0 || level == 3) { if (i > 1) ctx += VS(i - 2) << sh; sh += 2; }
				else { if (i > 1) ctx += (uint64_t)(VS(i - 2) == VS(i - 1)) << sh; sh += 1; }
				if (i + 1 < n) ctx += VS(i + 1) << sh;
				sh += 2;
				if (level > 1) { ctx += (uint64_t)(fl[i] == 1) << sh; ++sh; ctx += (uint64_t)(fl[i] == 2) << sh; }
				const uint32_t s = map[q[i] - 33];
				xm_encode(&M, &S.rc, Q_SYM, ctx, s, 0);
				context = ((context << bps) + (mode == 0 ? quant[s] : s)) & cmask;
			}
#undef VS
#undef AVG
		}
		xs_part(&S);
	}
	free(fl); xm_free(&M);
	return xs_finish(&S, out, out_cap);
}

/* ================================================================================================ header stream */
enum { H_PLUS, H_FLAGS, H_SAME, H_SAMELEN, H_LITERAL, H_PLAIN, H_COUNT };
typedef struct { uint8_t sep; uint32_t b, e; } xtok;
static int is_lit(uint8_t c) { return (c >= '0' && c <= '9') || (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z') || c == '@'; }
/* tokenize (id_coder.cpp:169-207): a_numeric is never set (init_symbol_classes :110-128), so every token is a literal */
static uint32_t xh_tokenize(const uint8_t* s, uint32_t n, xtok** t, uint32_t* cap)
{
	uint32_t nt = 0, start = 0;
	for (uint32_t i = 0; i <= n; ++i) if (i == n || !is_lit(s[i])) {
		if (nt == *cap) { *cap = *cap * 2 + 64; *t = (xtok*)realloc(*t, *cap * sizeof(xtok)); }
		(*t)[nt].sep = i == n ? 0 : s[i]; (*t)[nt].b = start; (*t)[nt].e = i; ++nt; start = i + 1;
	}
	return nt;
}
int64_t orc_xhdr_encode(const uint8_t* bytes, const uint64_t* off, const uint8_t* plus, uint32_t n, const uint32_t* pack_sizes, uint32_t n_packs, uint8_t* out, uint64_t out_cap, uint64_t* part_sizes)
{
	static const xfam fam[H_COUNT] = {{2, 1u << 15, 1}, {2, 1u << 15, 1}, {2, 1u << 15, 1}, {2, 1u << 15, 1}, {256, 1u << 20, 64}, {128, 1u << 19, 32}};      /* id_coder.h:50-59 */
	xstream S; xs_begin(&S, part_sizes);
	xmodels M; xm_init(&M, fam);
	xtok* tc = NULL; xtok* tp = NULL; uint32_t capc = 0, capp = 0, ntp = 0;
	const uint8_t* prv = NULL; uint64_t ctx_flags = 0;
	uint32_t r = 0;
	for (uint32_t p = 0; p < n_packs; ++p) {
		for (uint32_t k = 0; k < pack_sizes[p]; ++k, ++r) {                          /* compress_lossless id_coder.cpp:210-383 */
			const uint8_t* id = bytes + off[r]; const uint32_t len = (uint32_t)(off[r + 1] - off[r]);
			const uint32_t nt = xh_tokenize(id, len, &tc, &capc);
			xm_encode(&M, &S.rc, H_PLUS, 0, plus ? plus[r] : 0, 0);
			int same_types = nt == ntp;
			for (uint32_t i = 0; same_types && i < nt; ++i) if (tc[i].sep != tp[i].sep) same_types = 0;
			if (same_types) {
				xm_encode(&M, &S.rc, H_FLAGS, ctx_flags, 1, 0);
				ctx_flags = ((ctx_flags << 1) + 1) & 0xff;
				for (uint32_t i = 0; i < nt; ++i) {
					const uint32_t lc = tc[i].e - tc[i].b, lp = tp[i].e - tp[i].b;
					const int same_len = lc == lp, same = same_len && memcmp(id + tc[i].b, prv + tp[i].b, lc) == 0;
					xm_encode(&M, &S.rc, H_SAME, i, same ? 1 : 0, 0);
					if (same) continue;
					xm_encode(&M, &S.rc, H_SAMELEN, i, same_len ? 1 : 0, 0);
					if (same_len) for (uint32_t j = 0; j < lc; ++j) { const uint8_t c = id[tc[i].b + j]; xm_encode(&M, &S.rc, H_LITERAL, ctx_flags + (1ull << 32) + j + ((uint64_t)i << 40) + (1ull << 60), c == prv[tp[i].b + j] ? 0 : c, 0); }
					else {
						for (uint32_t j = 0; j < lc; ++j) xm_encode(&M, &S.rc, H_LITERAL, ctx_flags + j + (1ull << 32) + ((uint64_t)i << 40), id[tc[i].b + j], 0);
						xm_encode(&M, &S.rc, H_LITERAL, ctx_flags + lc + (1ull << 32) + ((uint64_t)i << 40), 0, 0);
					}
				}
			} else {
				xm_encode(&M, &S.rc, H_FLAGS, ctx_flags, 0, 0);
				ctx_flags = (ctx_flags << 1) & 0xff;
				for (uint32_t i = 0; i < len; ++i) xm_encode(&M, &S.rc, H_PLAIN, i, id[i], 0);
				xm_encode(&M, &S.rc, H_PLAIN, len, 0, 0);
			}
			{ xtok* t = tp; tp = tc; tc = t; const uint32_t c = capp; capp = capc; capc = c; ntp = nt; prv = id; }
		}
		xs_part(&S);
		ctx_flags = 0;                                                               /* Restart() id_coder.cpp:80-91; the previous header stays */
	}
	free(tc); free(tp); xm_free(&M);
	return xs_finish(&S, out, out_cap);
}
