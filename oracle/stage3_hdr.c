/* oracle/stage3_hdr.c — TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline may use it).
 *
 * CPU twin and decoder of the native header container "HB01" written by colord_b200/csrc/stage3_hdr.cu.
 * Event model = the reference's CIDCoder::compress_lossless / decompress_lossless (src/colord/id_coder.cpp:210-383, :407-560)
 * with its dead numeric branch left out (a_numeric is never set, id_coder.cpp:125-127 — every token is a literal):
 *   tokenize            id_coder.cpp:169-208   (cut at every character outside [0-9A-Za-z@], id_coder.cpp:121-138)
 *   token_types_same    id_coder.h:123-133
 *   plus_id, flag       id_coder.cpp:214-221   same / same length / characters  :225-275   plain  :359-373
 * Differences from the reference's stream (why the bytes have no reference counterpart — "parity unpinned" for the bytes; parity
 * is (a) twin == device byte for byte, (b) decode == input, (c) size against the reference's own header stream):
 *   static per-context tables instead of adaptive models, 64 coder lanes per pack, first header of a pack has no predecessor,
 *   contexts folded to table widths (hdr_model.h).
 * Container: "HB01" | n u64 | n_packs u32 | tables (rc_static.h) | per pack: n_in_pack u32, lane bytes u32 x 64, lane streams.
 */
#include "rc_static.h"

enum { H_PLUS = 0, H_FLAG, H_SAME, H_SAMELEN, H_LITEQ, H_LITNEW, H_PLAIN, H_COUNT };
#define HB_LANES 64
#define HB_MIN_CTX 64

static void hdr_model(st_model* m)
{
	const uint32_t A[H_COUNT] = {2, 2, 2, 2, 256, 256, 256};
	const uint32_t cb[H_COUNT] = {0, 8, 6, 6, 14, 10, 8};
	const uint32_t fb[H_COUNT] = {0, 0, 0, 0, 10, 0, 0};
	m->n_fam = H_COUNT;
	for (int f = 0; f < H_COUNT; ++f) { m->A[f] = A[f]; m->cbits[f] = cb[f]; m->fbits[f] = fb[f]; }
	st_layout(m);
}
static int is_lit(uint8_t c) { return (c >= '0' && c <= '9') || (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z') || c == '@'; }
static uint32_t mn(uint32_t a, uint32_t b) { return a < b ? a : b; }

/* tokens of a header: (separator, begin, end); the last one has separator 0 (id_coder.cpp:169-208) */
typedef struct { uint8_t sep; uint32_t b, e; } tok_t;
static uint32_t tokenize(const uint8_t* s, uint32_t n, tok_t* t, uint32_t cap)
{
	uint32_t nt = 0, start = 0;
	for (uint32_t i = 0; i < n; ++i) if (!is_lit(s[i])) { if (nt < cap) { t[nt].sep = s[i]; t[nt].b = start; t[nt].e = i; } ++nt; start = i + 1; }
	if (nt < cap) { t[nt].sep = 0; t[nt].b = start; t[nt].e = n; }
	return nt + 1;
}

/* events of one header -> sink(user, family, context, symbol) */
typedef void (*sink_fn)(void*, uint32_t, uint64_t, uint32_t);
static uint32_t hdr_events(const uint8_t* cur, uint32_t nc, const uint8_t* prv, uint32_t np, int has_prev, int plus, uint32_t fctx, tok_t* tc, tok_t* tp, uint32_t cap, sink_fn put, void* u)
{
	put(u, H_PLUS, 0, plus != 0);
	uint32_t flag = 0, ntc = 0;
	if (has_prev) {
		ntc = tokenize(cur, nc, tc, cap); const uint32_t ntp = tokenize(prv, np, tp, cap);
		flag = ntc == ntp;
		for (uint32_t i = 0; flag && i < ntc; ++i) if (tc[i].sep != tp[i].sep) flag = 0;
	}
	put(u, H_FLAG, fctx, flag);
	if (!flag) {
		for (uint32_t j = 0; j < nc; ++j) put(u, H_PLAIN, mn(j, 255), cur[j]);
		put(u, H_PLAIN, mn(nc, 255), 0);
		return 0;
	}
	for (uint32_t i = 0; i < ntc; ++i) {
		const uint32_t lc = tc[i].e - tc[i].b, lp = tp[i].e - tp[i].b, t5 = mn(i, 31);
		const int same_len = lc == lp, same = same_len && memcmp(cur + tc[i].b, prv + tp[i].b, lc) == 0;
		put(u, H_SAME, mn(i, 63), same);
		if (same) continue;
		put(u, H_SAMELEN, mn(i, 63), same_len);
		if (same_len)
			for (uint32_t k = 0; k < lc; ++k) { const uint32_t c = cur[tc[i].b + k], p = prv[tp[i].b + k]; put(u, H_LITEQ, ((p & 15u) << 10) | (t5 << 5) | mn(k, 31), c == p ? 0 : c); }
		else {
			for (uint32_t k = 0; k < lc; ++k) put(u, H_LITNEW, (t5 << 5) | mn(k, 31), cur[tc[i].b + k]);
			put(u, H_LITNEW, (t5 << 5) | mn(lc, 31), 0);
		}
	}
	return 1;
}

typedef struct { const st_model* m; uint32_t* hist; } count_u;
static void put_count(void* u, uint32_t f, uint64_t ctx, uint32_t sym) { count_u* c = (count_u*)u; ++c->hist[c->m->base[f] + (ctx & ((1ull << c->m->cbits[f]) - 1)) * c->m->A[f] + sym]; }
typedef struct { const st_model* m; rcenc* e; } enc_u;
static void put_enc(void* u, uint32_t f, uint64_t ctx, uint32_t sym) { enc_u* c = (enc_u*)u; rce_put(c->e, c->m, f, ctx, sym); }

/* CPU twin of clb_hdr_encode: out must hold out_cap bytes; returns the container size, < 0 on error (-1: capacity) */
int64_t orc_hdr_encode(const uint8_t* bytes, const uint64_t* off, const uint8_t* plus, uint64_t n, const uint32_t* pack_sizes, uint32_t n_packs, uint8_t* out, uint64_t out_cap)
{
	uint64_t* pf = (uint64_t*)calloc((size_t)n_packs + 2, 8); uint32_t np = 0;
	{ uint64_t at = 0; for (uint32_t i = 0; i < n_packs; ++i) { at += pack_sizes[i]; if (pack_sizes[i]) pf[++np] = at; } if (at != n) { free(pf); return -2; } }
	uint32_t max_len = 0; for (uint64_t r = 0; r < n; ++r) if (off[r + 1] - off[r] > max_len) max_len = (uint32_t)(off[r + 1] - off[r]);
	const uint32_t cap = max_len + 2;
	tok_t* tc = (tok_t*)malloc(sizeof(tok_t) * cap); tok_t* tp = (tok_t*)malloc(sizeof(tok_t) * cap);
	st_model M; hdr_model(&M);
	uint32_t* hist = (uint32_t*)calloc(M.base[H_COUNT] + 1, 4);
	uint8_t* flags = (uint8_t*)calloc((size_t)n + 1, 1);
	/* pass 1: counts (and the flags) */
	count_u cu = {&M, hist};
	for (uint32_t p = 0; p < np; ++p)
		for (uint64_t r = pf[p]; r < pf[p + 1]; ++r) {
			uint32_t fctx = 0; for (uint64_t k = r - pf[p] < 8 ? pf[p] : r - 8; k < r; ++k) fctx = (fctx << 1) + flags[k];
			flags[r] = (uint8_t)hdr_events(bytes + off[r], (uint32_t)(off[r + 1] - off[r]), r > pf[p] ? bytes + off[r - 1] : 0, r > pf[p] ? (uint32_t)(off[r] - off[r - 1]) : 0,
				r > pf[p], plus ? plus[r] : 0, fctx & 0xff, tc, tp, cap, put_count, &cu);
		}
	st_buf o = {0, 0, 0};
	st_push(&o, "HB01", 4); st_push(&o, &n, 8); st_push(&o, &np, 4);
	st_write_tables(&M, hist, &o, HB_MIN_CTX);
	/* pass 2: lanes */
	for (uint32_t p = 0; p < np; ++p) {
		const uint32_t in_pack = (uint32_t)(pf[p + 1] - pf[p]);
		const uint64_t hdr_at = o.n;
		st_push(&o, &in_pack, 4);
		{ uint32_t z = 0; for (int l = 0; l < HB_LANES; ++l) st_push(&o, &z, 4); }
		for (uint32_t l = 0; l < HB_LANES; ++l) {
			const uint64_t lane_at = o.n;
			rcenc e; rce_start(&e, &o); enc_u eu = {&M, &e};
			for (uint64_t r = pf[p] + l; r < pf[p + 1]; r += HB_LANES) {
				uint32_t fctx = 0; for (uint64_t k = r - pf[p] < 8 ? pf[p] : r - 8; k < r; ++k) fctx = (fctx << 1) + flags[k];
				hdr_events(bytes + off[r], (uint32_t)(off[r + 1] - off[r]), r > pf[p] ? bytes + off[r - 1] : 0, r > pf[p] ? (uint32_t)(off[r] - off[r - 1]) : 0,
					r > pf[p], plus ? plus[r] : 0, fctx & 0xff, tc, tp, cap, put_enc, &eu);
			}
			rce_end(&e);
			const uint32_t nb = (uint32_t)(o.n - lane_at);
			memcpy(o.p + hdr_at + 4 + 4 * l, &nb, 4);
		}
	}
	int64_t ret = (int64_t)o.n;
	if (o.n > out_cap) ret = -1; else memcpy(out, o.p, o.n);
	free(o.p); free(pf); free(tc); free(tp); free(hist); free(flags); free(M.freq);
	return ret;
}

/* Decoder: headers back to back into out (capacity out_cap), off[n+1], plus[n].  Returns the number of header bytes,
 * < 0 on error (-1 capacity, -2 bad container, -3 header count mismatch). */
int64_t orc_hdr_decode(const uint8_t* in, uint64_t in_n, uint64_t n_expected, uint8_t* out, uint64_t out_cap, uint64_t* off, uint8_t* plus)
{
	if (in_n < 16 || memcmp(in, "HB01", 4) != 0) return -2;
	uint64_t at = 4, n; uint32_t np;
	memcpy(&n, in + at, 8); at += 8; memcpy(&np, in + at, 4); at += 4;
	if (n != n_expected) return -3;
	st_model M; hdr_model(&M);
	at = st_read_tables(&M, in, at);
	uint64_t w = 0, r = 0; int64_t rc = 0;
	off[0] = 0;
	for (uint32_t p = 0; p < np && rc == 0; ++p) {
		uint32_t in_pack, lane_bytes[HB_LANES]; rcdec dec[HB_LANES];
		memcpy(&in_pack, in + at, 4); at += 4; memcpy(lane_bytes, in + at, 4 * HB_LANES); at += 4 * HB_LANES;
		for (int l = 0; l < HB_LANES; ++l) { rc_start(&dec[l], in + at, lane_bytes[l]); at += lane_bytes[l]; }
		const uint64_t r0 = r;
		uint32_t flag_hist = 0;
		for (uint32_t k = 0; k < in_pack && rc == 0; ++k, ++r) {
			rcdec* d = &dec[k % HB_LANES];
			const uint8_t* prv = out + (r > r0 ? off[r - 1] : 0); const uint32_t lpv = r > r0 ? (uint32_t)(off[r] - off[r - 1]) : 0;
			plus[r] = (uint8_t)rc_get(d, &M, H_PLUS, 0);
			const uint32_t flag = rc_get(d, &M, H_FLAG, flag_hist & 0xff);
			flag_hist = (flag_hist << 1) + flag;
			if (r == r0 && flag) { rc = -2; break; }
#define EMIT(ch) do { if (w >= out_cap) { rc = -1; goto done; } out[w++] = (uint8_t)(ch); } while (0)
			if (!flag) {
				for (uint32_t j = 0;; ++j) { const uint32_t c = rc_get(d, &M, H_PLAIN, mn(j, 255)); if (!c) break; EMIT(c); }
			} else {
				uint32_t j = 0;                     /* cursor in the previous header */
				for (uint32_t t = 0;; ++t) {
					uint32_t je = j; while (je < lpv && is_lit(prv[je])) ++je;
					const uint32_t lp = je - j, t5 = mn(t, 31);
					if (rc_get(d, &M, H_SAME, mn(t, 63))) { for (uint32_t k2 = 0; k2 < lp; ++k2) EMIT(prv[j + k2]); }
					else if (rc_get(d, &M, H_SAMELEN, mn(t, 63))) {
						for (uint32_t k2 = 0; k2 < lp; ++k2) { const uint32_t pc = prv[j + k2]; const uint32_t c = rc_get(d, &M, H_LITEQ, ((pc & 15u) << 10) | (t5 << 5) | mn(k2, 31)); EMIT(c ? c : pc); }
					} else {
						for (uint32_t k2 = 0;; ++k2) { const uint32_t c = rc_get(d, &M, H_LITNEW, (t5 << 5) | mn(k2, 31)); if (!c) break; EMIT(c); }
					}
					if (je == lpv) break;
					EMIT(prv[je]);                  /* the separator is the previous header's (same shape) */
					j = je + 1;
				}
			}
			off[r + 1] = w;
		}
	}
done:
	free(M.freq);
	if (rc) return rc;
	return r == n ? (int64_t)w : -3;
}
