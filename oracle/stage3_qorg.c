/* oracle/stage3_qorg.c — TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline may use it).
 *
 * CPU twin and decoder of the native lossless-quality container "QO01" written by colord_b200/csrc/stage3_qorg.cu.
 * Context model = the reference's CQualityCoder::encode_original / decode_original (src/colord/quality_coder_impl.cpp:78-128,
 * :603-660): per base one 96-symbol phred value under
 *   [2 previous values quantised to 4 bits (quality_coder.cpp:110-114: no_bits_per_symbol 4, no_ctx_symbols 2; update :534-537,
 *    reset :528-531)] | base i | base i-1 | (level 3: base i-2, else base i-2 == base i-1) | base i+1 | (level > 1: match, anchor)
 * with valid_sym(x) = x & 3 (quality_coder.h:136-139) and the quantiser of the data source and level (quality_coder.cpp:272-338
 * ONT, :356-420 PacBio CLR, :441-505 PacBio HiFi).  The flags come from the read's tuples (quality_coder_impl.cpp:25-76).
 * Differences from the reference's stream ("parity unpinned" for the bytes; parity is twin == device byte for byte, decode == input,
 * size against the reference's own quality stream): static per-context tables instead of adaptive models, 64 coder lanes per pack.
 * Container: "QO01" | source u32 | level u32 | n u64 | n_packs u32 | tables (rc_static.h) | per pack: n_in_pack u32, lane bytes u32 x 64, lanes.
 */
#include "rc_static.h"

#define QO_LANES 64
#define QO_MIN_CTX 256

static void qo_quant(uint32_t source, uint32_t level, uint8_t* q /*96*/)
{
	memset(q, 0, 96);
#define FILL(a, b, v) do { for (int i_ = (a); i_ < (b); ++i_) q[i_] = (uint8_t)(v); } while (0)
	if (source == 0) {
		q[0] = 0; q[1] = 1;
		if (level >= 3) { FILL(2, 4, 2); FILL(4, 7, 3); FILL(7, 11, 4); FILL(11, 16, 5); FILL(16, 22, 6); FILL(22, 29, 7); FILL(29, 37, 8); FILL(37, 46, 9); FILL(46, 56, 10); FILL(56, 67, 11); FILL(67, 79, 12); FILL(79, 90, 13); FILL(90, 96, 14); }
		else { FILL(2, 5, 2); FILL(5, 10, 3); FILL(10, 15, 4); FILL(15, 20, 5); FILL(20, 25, 6); FILL(25, 35, 7); FILL(35, 50, 8); FILL(50, 70, 9); FILL(70, 96, 10); }
	} else {
		const int s = source == 2 ? 1 : 0;      /* HiFi: every code one higher, 93 -> 0 */
		q[0] = (uint8_t)s;
		if (level >= 3) { FILL(1, 10, 1 + s); FILL(10, 20, 2 + s); FILL(20, 30, 3 + s); FILL(30, 39, 4 + s); FILL(39, 45, 5 + s); FILL(45, 51, 6 + s); FILL(51, 57, 7 + s); FILL(57, 63, 8 + s); FILL(63, 69, 9 + s); FILL(69, 75, 10 + s); FILL(75, 81, 11 + s); FILL(81, 87, 12 + s); FILL(87, 93, 13 + s); q[93] = (uint8_t)(s ? 0 : 14); }
		else { FILL(1, 15, 1 + s); FILL(15, 29, 2 + s); FILL(29, 41, 3 + s); FILL(41, 53, 4 + s); FILL(53, 63, 5 + s); FILL(63, 72, 6 + s); FILL(72, 80, 7 + s); FILL(80, 87, 8 + s); FILL(87, 93, 9 + s); q[93] = (uint8_t)(s ? 0 : 10); }
	}
#undef FILL
}
static void qo_model(st_model* m, uint32_t level)
{
	m->n_fam = 1; m->A[0] = 96; m->cbits[0] = 8 + (level >= 3 ? 8 : 7) + (level > 1 ? 2 : 0); m->fbits[0] = 8;
	st_layout(m);
}
static uint32_t bsym(uint8_t ascii) { return ascii == 'C' ? 1 : ascii == 'G' ? 2 : ascii == 'T' ? 3 : 0; }      /* A and N -> 0 (valid_sym) */
/* quality_coder_impl.cpp:25-76; flags: 0 none, 1 match, 2 anchor */
static void qo_flags(const uint8_t* es, uint64_t es_n, uint8_t* fl, uint32_t n)
{
	memset(fl, 0, n);
	if (!es_n) return;
	const uint32_t t0 = es[0] >> 4;
	if (t0 == 9 || t0 == 11) return;
	uint64_t p = 5; uint32_t at = 0;
	while (p < es_n) {
		const uint32_t t = es[p] >> 4;
		if (t == 4) { const uint32_t len = ((uint32_t)(es[p] & 15) << 24) | ((uint32_t)es[p + 1] << 16) | ((uint32_t)es[p + 2] << 8) | es[p + 3]; for (uint32_t k = 0; k < len && at < n; ++k) fl[at++] = 2; p += 4; }
		else if (t == 5) p += 4;
		else if (t == 6) p += 5;
		else { if (t == 2) { if (at < n) fl[at] = 1; ++at; } else if (t == 0 || t == 3) ++at; p += 1; }
	}
}
/* context of position i given the running context of the previous symbols */
static uint32_t qo_ctx(uint32_t prev_ctx, uint32_t level, const uint8_t* b, uint32_t n, const uint8_t* fl, uint32_t i)
{
	uint32_t c = prev_ctx, sh = 8;
	c += bsym(b[i]) << sh; sh += 2;
	if (i > 0) c += bsym(b[i - 1]) << sh;
	sh += 2;
	if (level >= 3) { if (i > 1) c += bsym(b[i - 2]) << sh; sh += 2; }
	else { if (i > 1) c += (uint32_t)(bsym(b[i - 2]) == bsym(b[i - 1])) << sh; sh += 1; }
	if (i + 1 < n) c += bsym(b[i + 1]) << sh;
	sh += 2;
	if (level > 1) { c += (uint32_t)(fl[i] == 1) << sh; ++sh; c += (uint32_t)(fl[i] == 2) << sh; }
	return c;
}

/* CPU twin of clb_qual_encode_original.  es / es_off may be NULL when level <= 1.  Returns the container size (> cap: too small), or < 0. */
int64_t orc_qorg_encode(uint32_t source, uint32_t level, const uint8_t* bases, const uint8_t* quals, const uint64_t* off, uint32_t n,
	const uint8_t* es, const uint64_t* es_off, const uint32_t* pack_sizes, uint32_t n_packs, uint8_t* out, uint64_t cap)
{
	uint32_t* pf = (uint32_t*)calloc((size_t)n_packs + 2, 4); uint32_t np = 0;
	{ uint64_t at = 0; for (uint32_t i = 0; i < n_packs; ++i) { at += pack_sizes[i]; if (pack_sizes[i]) pf[++np] = (uint32_t)at; } if (at != n) { free(pf); return -2; } }
	uint8_t quant[96]; qo_quant(source, level, quant);
	st_model M; qo_model(&M, level);
	uint32_t* hist = (uint32_t*)calloc(M.base[1] + 1, 4);
	uint32_t max_len = 0; for (uint32_t r = 0; r < n; ++r) if (off[r + 1] - off[r] > max_len) max_len = (uint32_t)(off[r + 1] - off[r]);
	uint8_t* fl = (uint8_t*)calloc((size_t)max_len + 1, 1);
	for (uint32_t r = 0; r < n; ++r) {
		const uint8_t* b = bases + off[r]; const uint8_t* q = quals + off[r]; const uint32_t len = (uint32_t)(off[r + 1] - off[r]);
		if (level > 1) qo_flags(es + es_off[r], es_off[r + 1] - es_off[r], fl, len);
		uint32_t pc = 0xff;
		for (uint32_t i = 0; i < len; ++i) {
			const uint32_t s = q[i] - 33u;
			if (s > 95) { free(pf); free(hist); free(fl); return -3; }
			++hist[(qo_ctx(pc, level, b, len, fl, i) & ((1u << M.cbits[0]) - 1)) * 96 + s];
			pc = ((pc << 4) + quant[s]) & 0xff;
		}
	}
	st_buf o = {0, 0, 0};
	const uint64_t n64 = n;
	st_push(&o, "QO01", 4); st_push(&o, &source, 4); st_push(&o, &level, 4); st_push(&o, &n64, 8); st_push(&o, &np, 4);
	st_write_tables(&M, hist, &o, QO_MIN_CTX);
	for (uint32_t p = 0; p < np; ++p) {
		const uint32_t in_pack = pf[p + 1] - pf[p];
		const uint64_t hdr_at = o.n;
		st_push(&o, &in_pack, 4);
		{ uint32_t z = 0; for (int l = 0; l < QO_LANES; ++l) st_push(&o, &z, 4); }
		for (uint32_t l = 0; l < QO_LANES; ++l) {
			const uint64_t lane_at = o.n;
			rcenc e; rce_start(&e, &o);
			for (uint32_t r = pf[p] + l; r < pf[p + 1]; r += QO_LANES) {
				const uint8_t* b = bases + off[r]; const uint8_t* q = quals + off[r]; const uint32_t len = (uint32_t)(off[r + 1] - off[r]);
				if (level > 1) qo_flags(es + es_off[r], es_off[r + 1] - es_off[r], fl, len);
				uint32_t pc = 0xff;
				for (uint32_t i = 0; i < len; ++i) { const uint32_t s = q[i] - 33u; rce_put(&e, &M, 0, qo_ctx(pc, level, b, len, fl, i), s); pc = ((pc << 4) + quant[s]) & 0xff; }
			}
			rce_end(&e);
			const uint32_t nb = (uint32_t)(o.n - lane_at);
			memcpy(o.p + hdr_at + 4 + 4 * l, &nb, 4);
		}
	}
	const int64_t ret = (int64_t)o.n;
	if (o.n <= cap) memcpy(out, o.p, o.n);
	free(o.p); free(pf); free(hist); free(fl); free(M.freq);
	return ret;
}

/* Decoder: phred+33 bytes of all reads into out (layout of `off`).  Returns 0, or < 0 on a malformed container. */
int orc_qorg_decode(const uint8_t* in, uint64_t in_n, const uint8_t* bases, const uint64_t* off, uint32_t n, const uint8_t* es, const uint64_t* es_off, uint8_t* out)
{
	if (in_n < 24 || memcmp(in, "QO01", 4)) return -1;
	uint32_t source, level, np; uint64_t nr, at = 4;
	memcpy(&source, in + at, 4); at += 4; memcpy(&level, in + at, 4); at += 4; memcpy(&nr, in + at, 8); at += 8; memcpy(&np, in + at, 4); at += 4;
	if (nr != n) return -2;
	uint8_t quant[96]; qo_quant(source, level, quant);
	st_model M; qo_model(&M, level);
	at = st_read_tables(&M, in, at);
	uint32_t max_len = 0; for (uint32_t r = 0; r < n; ++r) if (off[r + 1] - off[r] > max_len) max_len = (uint32_t)(off[r + 1] - off[r]);
	uint8_t* fl = (uint8_t*)calloc((size_t)max_len + 1, 1);
	uint32_t r0 = 0;
	for (uint32_t p = 0; p < np; ++p) {
		uint32_t in_pack, lane_bytes[QO_LANES]; rcdec dec[QO_LANES];
		memcpy(&in_pack, in + at, 4); at += 4; memcpy(lane_bytes, in + at, 4 * QO_LANES); at += 4 * QO_LANES;
		for (int l = 0; l < QO_LANES; ++l) { rc_start(&dec[l], in + at, lane_bytes[l]); at += lane_bytes[l]; }
		for (uint32_t r = r0; r < r0 + in_pack; ++r) {
			rcdec* d = &dec[(r - r0) % QO_LANES];
			const uint8_t* b = bases + off[r]; uint8_t* q = out + off[r]; const uint32_t len = (uint32_t)(off[r + 1] - off[r]);
			if (level > 1) qo_flags(es + es_off[r], es_off[r + 1] - es_off[r], fl, len);
			uint32_t pc = 0xff;
			for (uint32_t i = 0; i < len; ++i) { const uint32_t s = rc_get(d, &M, 0, qo_ctx(pc, level, b, len, fl, i)); q[i] = (uint8_t)(s + 33); pc = ((pc << 4) + quant[s]) & 0xff; }
		}
		r0 += in_pack;
	}
	free(fl); free(M.freq);
	return r0 == n ? 0 : -3;
}
