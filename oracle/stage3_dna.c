/* oracle/stage3_dna.c — TEST INFRASTRUCTURE ONLY.
 *
 * Decoder of the device's native DNA / edit-script container "DB01" (colord_b200/csrc/stage3_dna.cu).  It restates, in the
 * decoding direction, the reference's CDNACoder::Decode (src/colord/dna_coder.cpp:237-437: read flag, length, plain symbols,
 * reference read id, reverse-complement flag, tuple types with the symbol / tuple histories and the indel-drift bucket,
 * anchor / skip lengths, insertions, substitutions, alternative reads) and its range decoder (sub_rc.h:262-386: 64-bit
 * low / range / buffer, carry-less renormalisation) over STATIC frequency tables read from the container header, and rebuilds
 * every read from its tuples and the earlier reference reads exactly as the reference's decoder does.  It shares no code with
 * the device encoder: bases that come out equal to the input prove the container and the context model on both sides.
 * The container has no reference bitstream to be compared with ("parity unpinned" for the bytes): parity is this round trip
 * plus the stream size against the reference's own DNA stream on the same input (tests/test_gpu_stage3.py).
 */
#include "rc_static.h"

enum { F_FLAG = 0, F_LENBITS, F_LENDATA, F_SYM, F_SYMN, F_READID, F_REV, F_TUPLE, F_ANCHOR, F_SKIPL, F_SKIPD, F_SEEN, F_SHORT, F_COUNT };
#define DB_LANES 64

typedef struct { st_model t; uint32_t level, n_t, n_s; } model_t;

/* table widths: colord_b200/csrc/dna_model.h make_dna_model (history widths: dna_coder.cpp:1253-1280) */
static void make_model(model_t* m, uint32_t level, uint32_t max_cand)
{
	m->level = level; m->n_t = level >= 3 ? 4 : level == 2 ? 3 : 2; m->n_s = level >= 3 ? 8 : level == 2 ? 7 : 5;
	const uint32_t A[F_COUNT] = {3, 32, 256, 4, 5, 256, 2, 8, 24, 256, 256, 2, max_cand < 2 ? 2 : max_cand};
	const uint32_t sym_bits = level >= 3 ? 24 : level == 2 ? 23 : 22;
	const uint32_t cb[F_COUNT] = {8, 0, 9, sym_bits, 2 * m->n_s, 11, 4, 3 * m->n_t + 9, 6, 6, 8, 6, 6};
	const uint32_t fb[F_COUNT] = {0, 0, 0, 10, 0, 0, 0, 3 * m->n_t + 6, 0, 0, 0, 0, 0};
	m->t.n_fam = F_COUNT;
	for (int f = 0; f < F_COUNT; ++f) { m->t.A[f] = A[f]; m->t.cbits[f] = cb[f]; m->t.fbits[f] = fb[f]; }
	st_layout(&m->t);
}

typedef struct { uint8_t* p; uint64_t n, cap; } bbuf3;
static void b3_push(bbuf3* b, uint8_t x) { if (b->n == b->cap) { b->cap = b->cap * 2 + 4096; b->p = (uint8_t*)realloc(b->p, b->cap); } b->p[b->n++] = x; }

static uint32_t ilog2b(uint64_t x) { uint32_t r = 0; for (; x; ++r) x >>= 1; return r; }
static uint32_t nbytes(uint64_t x) { uint32_t r = 1; x >>= 8; for (; x; ++r) x >>= 8; return r; }

typedef struct { const uint8_t* sym; uint32_t len; int rev; } oref;      /* a stored reference read, oriented on access */
static uint32_t osym(const oref* o, int pos) { if (pos < 0 || (uint32_t)pos >= o->len) return 255; return o->rev ? 3u - o->sym[o->len - 1 - pos] : o->sym[pos]; }

/* Decodes the container.  is_ref[r]: read r joins the reference reads (the sampler decision, false for reads with N).
 * out_bases: ASCII bases of all reads back to back (capacity cap_bases), out_off[n_reads + 1].  Returns 0, or < 0 on a
 * malformed container, or the needed capacity if cap_bases is too small. */
/* ctx_bases / ctx_off / n_ctx: the context reads of a shard's container (multi-GPU: the reference reads of earlier shards, which
 * take the reference ids 0 .. n_ctx-1 and the read ids in front of the container's own reads); n_ctx = 0 for a whole-file container. */
int64_t orc_dna_decode_ctx(const uint8_t* in, uint64_t in_n, uint32_t n_reads, const uint8_t* is_ref, const uint8_t* ctx_bases, const uint64_t* ctx_off, uint32_t n_ctx,
	uint8_t* out_bases, uint64_t cap_bases, uint64_t* out_off)
{
	if (in_n < 28 || memcmp(in, "DB01", 4)) return -1;
	uint32_t level, max_cand, n_packs, hdr_ctx; uint64_t nr, at = 4;
	memcpy(&level, in + at, 4); at += 4; memcpy(&max_cand, in + at, 4); at += 4; memcpy(&nr, in + at, 8); at += 8; memcpy(&n_packs, in + at, 4); at += 4;
	memcpy(&hdr_ctx, in + at, 4); at += 4;
	if (nr != n_reads || hdr_ctx != n_ctx) return -2;
	model_t M; make_model(&M, level, max_cand);
	at = st_read_tables(&M.t, in, at);
	/* reference reads decoded so far (symbols 0..3) */
	uint8_t** ref_sym = (uint8_t**)calloc((size_t)n_reads + n_ctx + 1, sizeof(uint8_t*)); uint32_t* ref_len = (uint32_t*)calloc((size_t)n_reads + n_ctx + 1, 4); uint32_t n_ref = 0;
	for (uint32_t i = 0; i < n_ctx; ++i) {
		const uint64_t b = ctx_off[i], len = ctx_off[i + 1] - b;
		uint8_t* q = (uint8_t*)malloc(len + 1);
		for (uint64_t k = 0; k < len; ++k) { const uint8_t ch = ctx_bases[b + k]; q[k] = ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : 3; }
		ref_sym[n_ref] = q; ref_len[n_ref] = (uint32_t)len; ++n_ref;
	}
	const uint64_t mask_s = (1ull << (2 * M.n_s)) - 1, mask_t = (1ull << (3 * M.n_t)) - 1;
	const uint32_t sh_t = 3 * M.n_t;
	uint64_t w = 0; int64_t rc = 0;
	uint32_t r0 = 0;
	for (uint32_t p = 0; p < n_packs && rc == 0; ++p) {
		uint32_t np; memcpy(&np, in + at, 4); at += 4;
		uint32_t sizes[DB_LANES]; memcpy(sizes, in + at, 4 * DB_LANES); at += 4 * DB_LANES;
		rcdec dec[DB_LANES]; uint32_t fctx[DB_LANES];
		for (int l = 0; l < DB_LANES; ++l) { rc_start(&dec[l], in + at, sizes[l]); at += sizes[l]; fctx[l] = 0; }
		for (uint32_t r = r0; r < r0 + np; ++r) {
			rcdec* d = &dec[(r - r0) % DB_LANES]; uint32_t* fc = &fctx[(r - r0) % DB_LANES];
			bbuf3 rd = {0, 0, 0};                      /* symbols of this read */
			const uint32_t flag = rc_get(d, &M.t, F_FLAG, *fc);
			*fc = ((*fc << 2) + flag) & 0xff;
			uint32_t len;
			{
				const uint32_t nbits = rc_get(d, &M.t, F_LENBITS, 0);
				if (nbits < 2) len = nbits;              /* ilog2: 0 -> 0, 1 -> 1 */
				else {
					uint64_t ctx = (uint64_t)nbits << 3;
					uint32_t v = rc_get(d, &M.t, F_LENDATA, ctx);
					if (nbits > 9) {
						uint32_t suffix = 0, sh = 0; ctx += 4;
						for (int nb = (int)nbits - 9; nb > 0; nb -= 8) { suffix |= rc_get(d, &M.t, F_LENDATA, ctx) << sh; sh += 8; ++ctx; }
						v = (v << (nbits - 9)) + suffix;
					}
					len = v + (1u << (nbits - 1));
				}
			}
			uint64_t ctx_symbol = mask_s, ctx_tuple = mask_t;
			if (flag == 0) { for (uint32_t i = 0; i < len; ++i) { const uint32_t s = rc_get(d, &M.t, F_SYM, ctx_symbol << 2); b3_push(&rd, (uint8_t)s); ctx_symbol = ((ctx_symbol << 2) + s) & mask_s; } }
			else if (flag == 1) { for (uint32_t i = 0; i < len; ++i) { const uint32_t s = rc_get(d, &M.t, F_SYMN, ctx_symbol); b3_push(&rd, (uint8_t)s); ctx_symbol = ((ctx_symbol << 4) + s) & mask_s; } }
			else {
				uint32_t seen_id[34], seen_rev[34], n_seen = 0; uint64_t ctx_rev = 0xf;
				uint32_t alt_ids[32]; int alt_revs[32], alt_saved[32]; uint32_t n_alt = 0; int cur_alt = -1;
#define GET_READ_ID(dst) do { const int nn = (int)nbytes((uint64_t)n_ctx + r); uint32_t id_ = 0; for (int i = nn - 1; i >= 0; --i) { const uint64_t add = (i == nn - 2) ? id_ : 0; id_ = (id_ << 8) + rc_get(d, &M.t, F_READID, (uint64_t)i + (add << 3)); } dst = id_; } while (0)
#define GET_REV(id, dst) do { int fnd = -1; for (uint32_t k = 0; k < n_seen; ++k) if (seen_id[k] == (id)) fnd = (int)k; if (fnd >= 0) dst = (int)seen_rev[fnd]; else { const uint32_t fl_ = rc_get(d, &M.t, F_REV, ctx_rev); if (n_seen < 34) { seen_id[n_seen] = (id); seen_rev[n_seen] = fl_; ++n_seen; } ctx_rev = ((ctx_rev << 2) + fl_) & 0xf; dst = (int)fl_; } } while (0)
				uint32_t main_id; GET_READ_ID(main_id);
				int main_rev; GET_REV(main_id, main_rev);
				if (main_id >= n_ref) { rc = -3; free(rd.p); break; }
				oref mainr = {ref_sym[main_id], ref_len[main_id], main_rev}, altr = mainr;
				int ref_pos = 0, alt_pos = 0, delta = 0, is_main = 1; uint32_t last_tuple = 255;
				for (uint32_t it = 0; it < len; ++it) {
					const oref* o = is_main ? &mainr : &altr; int* pos = is_main ? &ref_pos : &alt_pos;
					const uint32_t rsym = osym(o, *pos);
					uint64_t ctx = ctx_tuple + ((ctx_symbol & 0xf) << sh_t) + ((uint64_t)rsym << (sh_t + 4));
					const uint32_t bucket = delta < -10 ? 1 : delta < -1 ? 2 : delta > 10 ? 3 : delta > 1 ? 4 : 0;
					ctx += (uint64_t)bucket << (sh_t + 6);
					const uint32_t ty = rc_get(d, &M.t, F_TUPLE, ctx);
					ctx_tuple = ((ctx_tuple << 3) + ty) & mask_t;
					if (ty == 6) {
						if (!is_main && cur_alt >= 0) alt_saved[cur_alt] = alt_pos;
						uint32_t id; int idx = -1;
						if (n_alt == 0) GET_READ_ID(id);
						else if (!rc_get(d, &M.t, F_SEEN, n_alt)) GET_READ_ID(id);
						else { idx = (int)rc_get(d, &M.t, F_SHORT, n_alt); id = alt_ids[idx]; }
						int rev; GET_REV(id, rev);
						if (idx < 0) { for (uint32_t k = 0; k < n_alt; ++k) if (alt_ids[k] == id) idx = (int)k; }
						if (idx < 0 && n_alt < 32) { idx = (int)n_alt; alt_ids[n_alt] = id; alt_revs[n_alt] = rev; alt_saved[n_alt] = 0; ++n_alt; }
						cur_alt = idx;
						if (id >= n_ref) { rc = -4; break; }
						altr.sym = ref_sym[id]; altr.len = ref_len[id]; altr.rev = idx >= 0 ? alt_revs[idx] : rev;
						alt_pos = 0; is_main = 0; delta = 0;
					} else if (ty == 4) {
						uint32_t alen = 0;
						for (uint32_t part = 0;; ++part) { const uint32_t v = rc_get(d, &M.t, F_ANCHOR, part < 63 ? part : 63); if (v < 23) { alen += v; break; } alen += 22; }
						for (uint32_t k = 0; k < alen; ++k) b3_push(&rd, (uint8_t)osym(o, *pos + (int)k));
						*pos += (int)alen;
						for (int i = (int)M.n_s; i > 0; --i) ctx_symbol = (ctx_symbol << 2) + osym(o, *pos - i);
						ctx_symbol &= mask_s;
						delta = 0;
					} else if (ty == 2) {
						b3_push(&rd, (uint8_t)rsym);
						ctx_symbol = ((ctx_symbol << 2) + rsym) & mask_s; ++*pos;
					} else if (ty == 0) {
						uint64_t c2 = 2; uint32_t sh = 2;
						if (M.level <= 1) { c2 += (ctx_symbol & 0xff) << sh; sh += 8; }
						else { c2 += (ctx_symbol & 0x3ff) << sh; sh += 10; if (M.level >= 3) { c2 += (uint64_t)(((ctx_symbol >> 10) & 3) == ((ctx_symbol >> 8) & 3)) << sh; ++sh; } }
						c2 += (uint64_t)rsym << sh; sh += 2;
						c2 += (ctx_tuple & 0777) << sh;
						const uint32_t s = rc_get(d, &M.t, F_SYM, c2);
						b3_push(&rd, (uint8_t)s);
						ctx_symbol = ((ctx_symbol << 2) + s) & mask_s; ++delta;
					} else if (ty == 1) { ++*pos; --delta; }
					else if (ty == 3) {
						uint64_t c2 = 1; uint32_t sh = 2;
						c2 += (ctx_symbol & 0x3f) << sh; sh += 6;
						if (M.level >= 3) { c2 += (uint64_t)(((ctx_symbol >> 6) & 3) == ((ctx_symbol >> 4) & 3)) << sh; ++sh; }
						c2 += (uint64_t)rsym << sh; sh += 2;
						c2 += (ctx_tuple & 07777) << sh;
						const uint32_t s = rc_get(d, &M.t, F_SYM, c2);
						b3_push(&rd, (uint8_t)s);
						ctx_symbol = ((ctx_symbol << 2) + s) & mask_s; ++*pos;
					} else if (ty == 5) {
						/* dna_coder.cpp:389-412 mirrors :166-206 */
						uint32_t skip;
						const int distant_after_alt = !is_main && last_tuple == 6;
						const int local = !distant_after_alt && last_tuple != 6 && last_tuple != 255;
#define GET_SKIP(dst, loc) do { uint32_t v_ = 0; if (loc) { for (uint32_t part = 0;; ++part) { const uint32_t x_ = rc_get(d, &M.t, F_SKIPL, part < 63 ? part : 63); if (x_ < 255) { v_ += x_; break; } v_ += 254; } } else { for (int i = 3; i >= 0; --i) { const uint32_t x_ = rc_get(d, &M.t, F_SKIPD, (uint64_t)i * 64 + ilog2b(v_)); v_ = (v_ << 8) + x_; } } dst = v_; } while (0)
						if (distant_after_alt) {
							uint32_t v; GET_SKIP(v, 0);
							const int saved = cur_alt >= 0 ? alt_saved[cur_alt] : 0;
							if (v > 0) skip = v + (uint32_t)saved;
							else { GET_SKIP(v, 0); skip = (uint32_t)(saved - (int)v); }
						} else GET_SKIP(skip, local);
						delta -= (int)skip;
						*pos += (int)skip;
					} else if (ty == 7) { is_main = 1; if (cur_alt >= 0) alt_saved[cur_alt] = alt_pos; delta = 0; }
					last_tuple = ty;
				}
				if (rc) { free(rd.p); break; }
			}
			/* output + reference store */
			out_off[r] = w;
			if (w + rd.n <= cap_bases) for (uint64_t i = 0; i < rd.n; ++i) out_bases[w + i] = "ACGTN"[rd.p[i] > 4 ? 4 : rd.p[i]];
			w += rd.n;
			if (is_ref[r]) { ref_sym[n_ref] = rd.p; ref_len[n_ref] = (uint32_t)rd.n; ++n_ref; } else free(rd.p);
		}
		r0 += np;
	}
	out_off[n_reads] = w;
	for (uint32_t i = 0; i < n_ref; ++i) free(ref_sym[i]);
	free(ref_sym); free(ref_len); free(M.t.freq);
	if (rc) return rc;
	return w > cap_bases ? (int64_t)w : 0;
}

int64_t orc_dna_decode(const uint8_t* in, uint64_t in_n, uint32_t n_reads, const uint8_t* is_ref, uint8_t* out_bases, uint64_t cap_bases, uint64_t* out_off)
{
	return orc_dna_decode_ctx(in, in_n, n_reads, is_ref, 0, 0, 0, out_bases, cap_bases, out_off);
}

/* ------------------------------------------------------------------------------------------------ CPU twin of the encoder
 * (added at the end of round 1; the device-vs-twin byte comparison has not run on a GPU yet — it is tests/test_gpu_stage3.py's
 * next case.)  Restates, in the ENCODING direction, the same events as the decoder above from a read's CompactES tuples
 * (utils.h:69-273: 1 byte per simple tuple = type << 4 | value, 4 bytes for anchor / skip with a 28-bit length, 5 bytes for the
 * id tuples) — CDNACoder::Encode, src/colord/dna_coder.cpp:26-231 — then builds the container exactly as the device does:
 * pass 1 counts (family, context, symbol), st_write_tables turns the counts into the serialised static tables, pass 2 codes
 * every pack with 64 range-coder lanes (lane l: reads l, l + 64, ... of the pack; the read-flag context is the lane's own history).
 * With the tuples the unmodified reference emitted (tests/golden/<case>/es.bin) this closes the loop on the CPU:
 * reference tuples -> container -> the two independent decoders -> the input reads. */
typedef void (*dna_sink_fn)(void*, uint32_t, uint64_t, uint32_t);
typedef struct { const uint8_t* bases; const uint64_t* off; const uint32_t* ref_to_read; uint32_t n_ref; } enc_reads;
static uint32_t code_of(uint8_t ch) { return ch == 'C' ? 1u : ch == 'G' ? 2u : ch == 'T' ? 3u : 0u; }
static uint32_t esym(const enc_reads* R, uint32_t ref_id, int rev, int pos)
{
	if (ref_id >= R->n_ref) return 255;
	const uint32_t rr = R->ref_to_read[ref_id];
	const uint64_t b = R->off[rr]; const uint32_t len = (uint32_t)(R->off[rr + 1] - b);
	if (pos < 0 || (uint32_t)pos >= len) return 255;
	return rev ? 3u - code_of(R->bases[b + (len - 1 - (uint32_t)pos)]) : code_of(R->bases[b + (uint32_t)pos]);
}
static uint32_t be32at(const uint8_t* t, uint64_t p) { return ((uint32_t)t[p] << 24) | ((uint32_t)t[p + 1] << 16) | ((uint32_t)t[p + 2] << 8) | t[p + 3]; }
static void put_id(const model_t* M, uint32_t read_index, uint32_t id, dna_sink_fn put, void* u)
{
	const int n = (int)nbytes(read_index);
	for (int i = n - 1; i >= 0; --i) { const uint64_t add = (i == n - 2) ? ((id >> (8 * (n - 1))) & 0xff) : 0; put(u, F_READID, (uint64_t)i + (add << 3), (id >> (8 * i)) & 0xff); }
	(void)M;
}
static void put_skip_len(uint32_t len, int local, dna_sink_fn put, void* u)
{
	if (local) { for (uint32_t part = 0; len; ++part) { if (len < 255) { put(u, F_SKIPL, part < 63 ? part : 63, len); break; } put(u, F_SKIPL, part < 63 ? part : 63, 255); len -= 254; } }
	else { uint32_t enc = 0; for (int i = 3; i >= 0; --i) { const uint32_t x = (len >> (8 * i)) & 0xff; put(u, F_SKIPD, (uint64_t)i * 64 + ilog2b(enc), x); enc = (enc << 8) + x; } }
}
static uint32_t flag_of_tuples(const uint8_t* t) { const uint32_t t0 = t[0] >> 4; return t0 == 9 ? 0u : t0 == 11 ? 1u : 2u; }

/* events of read r (index among the container's reads) in coding order */
static void dna_events(const model_t* M, const enc_reads* R, uint32_t r, const uint8_t* t, uint64_t tn, uint32_t fctx, dna_sink_fn put, void* u)
{
	uint32_t n_tuples = 0;
	for (uint64_t p = 0; p < tn; ++n_tuples) { const uint32_t ty = t[p] >> 4; p += (ty == 4 || ty == 5) ? 4 : (ty == 6 || ty == 10) ? 5 : 1; }
	const uint32_t flag = flag_of_tuples(t);
	put(u, F_FLAG, fctx, flag);
	{
		uint32_t len = n_tuples - 1;
		const uint32_t nbits = ilog2b(len);
		put(u, F_LENBITS, 0, nbits);
		if (nbits >= 2) {
			uint64_t ctx = (uint64_t)nbits << 3;
			len -= 1u << (nbits - 1);
			uint32_t prefix = len, suffix = 0;
			if (nbits > 9) { prefix = len >> (nbits - 9); suffix = len - (prefix << (nbits - 9)); }
			put(u, F_LENDATA, ctx, prefix);
			if (nbits > 9) { ctx += 4; for (int nb = (int)nbits - 9; nb > 0; nb -= 8) { put(u, F_LENDATA, ctx, suffix & 0xff); suffix >>= 8; ++ctx; } }
		}
	}
	const uint64_t mask_s = (1ull << (2 * M->n_s)) - 1, mask_t = (1ull << (3 * M->n_t)) - 1;
	uint64_t ctx_symbol = mask_s, ctx_tuple = mask_t;
	if (flag == 0) { for (uint64_t p = 1; p < tn; ++p) { const uint32_t s = t[p] & 15; put(u, F_SYM, ctx_symbol << 2, s); ctx_symbol = ((ctx_symbol << 2) + s) & mask_s; } return; }
	if (flag == 1) { for (uint64_t p = 1; p < tn; ++p) { const uint32_t s = t[p] & 15; put(u, F_SYMN, ctx_symbol, s); ctx_symbol = ((ctx_symbol << 4) + s) & mask_s; } return; }
	uint32_t seen[34], n_seen = 0; uint64_t ctx_rev = 0xf;
#define PUT_REV(id_, rev_) do { int f_ = 0; for (uint32_t k_ = 0; k_ < n_seen; ++k_) if (seen[k_] == (id_)) f_ = 1; if (!f_) { put(u, F_REV, ctx_rev, (rev_)); if (n_seen < 34) seen[n_seen++] = (id_); ctx_rev = ((ctx_rev << 2) + (rev_)) & 0xf; } } while (0)
	const uint32_t main_id = be32at(t, 1), main_rev = t[0] & 15;
	put_id(M, r, main_id, put, u);
	PUT_REV(main_id, main_rev);
	uint32_t alt_id = main_id; int alt_rev = (int)main_rev;
	uint32_t alt_ids[32], alt_revs[32]; int alt_saved[32]; uint32_t n_alt = 0; int cur_alt = -1;
	int ref_pos = 0, alt_pos = 0, delta = 0, is_main = 1; uint32_t last_tuple = 255;
	const uint32_t sh_t = 3 * M->n_t;
	for (uint64_t p = 5; p < tn;) {
		const uint32_t ty = t[p] >> 4, v1 = t[p] & 15; uint32_t v2 = 0;
		if (ty == 4 || ty == 5) { v2 = (v1 << 24) | ((uint32_t)t[p + 1] << 16) | ((uint32_t)t[p + 2] << 8) | t[p + 3]; p += 4; }
		else if (ty == 6) { v2 = be32at(t, p + 1); p += 5; }
		else p += 1;
		const uint32_t rsym = is_main ? esym(R, main_id, (int)main_rev, ref_pos) : esym(R, alt_id, alt_rev, alt_pos);
		{
			uint64_t ctx = ctx_tuple + ((ctx_symbol & 0xf) << sh_t) + ((uint64_t)rsym << (sh_t + 4));
			const uint32_t bucket = delta < -10 ? 1 : delta < -1 ? 2 : delta > 10 ? 3 : delta > 1 ? 4 : 0;
			ctx += (uint64_t)bucket << (sh_t + 6);
			put(u, F_TUPLE, ctx, ty);
			ctx_tuple = ((ctx_tuple << 3) + ty) & mask_t;
		}
		if (ty == 6) {
			if (!is_main && cur_alt >= 0) alt_saved[cur_alt] = alt_pos;
			int idx = -1;
			for (uint32_t k = 0; k < n_alt; ++k) if (alt_ids[k] == v2) { idx = (int)k; break; }
			if (n_alt == 0) put_id(M, r, v2, put, u);
			else { put(u, F_SEEN, n_alt, idx >= 0); if (idx < 0) put_id(M, r, v2, put, u); else put(u, F_SHORT, n_alt, (uint32_t)idx); }
			if (idx < 0 && n_alt < 32) { idx = (int)n_alt; alt_ids[n_alt] = v2; alt_revs[n_alt] = v1; alt_saved[n_alt] = 0; ++n_alt; }
			cur_alt = idx;
			PUT_REV(v2, v1);
			alt_id = v2; alt_rev = (int)(idx >= 0 ? alt_revs[idx] : v1);
			alt_pos = 0; is_main = 0; delta = 0;
		} else if (ty == 4) {
			for (uint32_t len = v2, part = 0; len; ++part) { if (len < 23) { put(u, F_ANCHOR, part < 63 ? part : 63, len); break; } put(u, F_ANCHOR, part < 63 ? part : 63, 23); len -= 22; }
			int* pos = is_main ? &ref_pos : &alt_pos;
			*pos += (int)v2;
			for (int i = (int)M->n_s; i > 0; --i) ctx_symbol = (ctx_symbol << 2) + (is_main ? esym(R, main_id, (int)main_rev, *pos - i) : esym(R, alt_id, alt_rev, *pos - i));
			ctx_symbol &= mask_s; delta = 0;
		} else if (ty == 2) { ctx_symbol = ((ctx_symbol << 2) + rsym) & mask_s; if (is_main) ++ref_pos; else ++alt_pos; }
		else if (ty == 0) {
			uint64_t c2 = 2; uint32_t sh = 2;
			if (M->level <= 1) { c2 += (ctx_symbol & 0xff) << sh; sh += 8; }
			else { c2 += (ctx_symbol & 0x3ff) << sh; sh += 10; if (M->level >= 3) { c2 += (uint64_t)(((ctx_symbol >> 10) & 3) == ((ctx_symbol >> 8) & 3)) << sh; ++sh; } }
			c2 += (uint64_t)rsym << sh; sh += 2;
			c2 += (ctx_tuple & 0777) << sh;
			put(u, F_SYM, c2, v1);
			ctx_symbol = ((ctx_symbol << 2) + v1) & mask_s; ++delta;
		} else if (ty == 1) { if (is_main) ++ref_pos; else ++alt_pos; --delta; }
		else if (ty == 3) {
			const uint32_t b = rsym & 3, symbol = v1 + (v1 >= b ? 1u : 0u);      /* the code-th base other than the reference base (dna_coder.h:37) */
			uint64_t c2 = 1; uint32_t sh = 2;
			c2 += (ctx_symbol & 0x3f) << sh; sh += 6;
			if (M->level >= 3) { c2 += (uint64_t)(((ctx_symbol >> 6) & 3) == ((ctx_symbol >> 4) & 3)) << sh; ++sh; }
			c2 += (uint64_t)rsym << sh; sh += 2;
			c2 += (ctx_tuple & 07777) << sh;
			put(u, F_SYM, c2, symbol);
			ctx_symbol = ((ctx_symbol << 2) + symbol) & mask_s; if (is_main) ++ref_pos; else ++alt_pos;
		} else if (ty == 5) {
			const int skip_len = (int)v2;
			delta -= skip_len;
			if (!is_main && last_tuple == 6) {
				const int mod = skip_len - (cur_alt >= 0 ? alt_saved[cur_alt] : 0);
				if (mod > 0) put_skip_len((uint32_t)mod, 0, put, u);
				else { put_skip_len(0, 0, put, u); put_skip_len((uint32_t)(-mod), 0, put, u); }
			} else put_skip_len((uint32_t)skip_len, last_tuple != 6 && last_tuple != 255, put, u);
			if (is_main) ref_pos += skip_len; else alt_pos += skip_len;
		} else if (ty == 7) { is_main = 1; if (cur_alt >= 0) alt_saved[cur_alt] = alt_pos; delta = 0; }
		last_tuple = ty;
	}
#undef PUT_REV
}

typedef struct { const st_model* m; uint32_t* hist; } dcount_u;
static void dput_count(void* u, uint32_t f, uint64_t ctx, uint32_t sym) { dcount_u* c = (dcount_u*)u; ++c->hist[c->m->base[f] + (ctx & ((1ull << c->m->cbits[f]) - 1)) * c->m->A[f] + sym]; }
typedef struct { const st_model* m; rcenc* e; } denc_u;
static void dput_enc(void* u, uint32_t f, uint64_t ctx, uint32_t sym) { denc_u* c = (denc_u*)u; rce_put(c->e, c->m, f, ctx, sym); }

/* CPU twin of clb_dna_encode for a whole-file container (no context reads).  es / es_off: CompactES bytes of the n reads;
 * bases / off: their ASCII bases (reference symbols come from here); is_ref[r]: read r is a reference read.
 * Returns the container size (> cap: too small), < 0 on bad arguments. */
int64_t orc_dna_encode(uint32_t level, uint32_t max_cand, const uint8_t* es, const uint64_t* es_off, const uint8_t* bases, const uint64_t* off,
	const uint8_t* is_ref, uint32_t n, const uint32_t* pack_sizes, uint32_t n_packs, uint8_t* out, uint64_t cap)
{
	uint32_t* pf = (uint32_t*)calloc((size_t)n_packs + 2, 4); uint32_t np = 0;
	{ uint64_t at = 0; for (uint32_t i = 0; i < n_packs; ++i) { at += pack_sizes[i]; if (pack_sizes[i]) pf[++np] = (uint32_t)at; } if (at != n) { free(pf); return -2; } }
	uint32_t* r2r = (uint32_t*)calloc((size_t)n + 1, 4); uint32_t n_ref = 0;
	for (uint32_t r = 0; r < n; ++r) if (is_ref[r]) r2r[n_ref++] = r;
	const enc_reads R = {bases, off, r2r, n_ref};
	model_t M; make_model(&M, level, max_cand);
	uint32_t* hist = (uint32_t*)calloc(M.t.base[F_COUNT] + 1, 4);
	dcount_u cu = {&M.t, hist};
	for (uint32_t p = 0; p < np; ++p)
		for (uint32_t r = pf[p]; r < pf[p + 1]; ++r) {
			uint32_t fctx = 0;
			for (int k = 4; k >= 1; --k) { const long long rr = (long long)r - 64ll * k; if (rr >= (long long)pf[p]) fctx = ((fctx << 2) + flag_of_tuples(es + es_off[rr])) & 0xff; }
			dna_events(&M, &R, r, es + es_off[r], es_off[r + 1] - es_off[r], fctx, dput_count, &cu);
		}
	st_buf o = {0, 0, 0};
	const uint64_t n64 = n; const uint32_t zero = 0;
	st_push(&o, "DB01", 4); st_push(&o, &level, 4); st_push(&o, &max_cand, 4); st_push(&o, &n64, 8); st_push(&o, &np, 4); st_push(&o, &zero, 4);
	st_write_tables(&M.t, hist, &o, 64);
	for (uint32_t p = 0; p < np; ++p) {
		const uint32_t in_pack = pf[p + 1] - pf[p];
		const uint64_t hdr_at = o.n;
		st_push(&o, &in_pack, 4);
		for (int l = 0; l < DB_LANES; ++l) st_push(&o, &zero, 4);
		for (uint32_t l = 0; l < DB_LANES; ++l) {
			const uint64_t lane_at = o.n;
			rcenc e; rce_start(&e, &o); denc_u eu = {&M.t, &e};
			uint32_t fctx = 0;
			for (uint32_t r = pf[p] + l; r < pf[p + 1]; r += DB_LANES) {
				dna_events(&M, &R, r, es + es_off[r], es_off[r + 1] - es_off[r], fctx, dput_enc, &eu);
				fctx = ((fctx << 2) + flag_of_tuples(es + es_off[r])) & 0xff;
			}
			rce_end(&e);
			const uint32_t nb = (uint32_t)(o.n - lane_at);
			memcpy(o.p + hdr_at + 4 + 4 * l, &nb, 4);
		}
	}
	const int64_t ret = (int64_t)o.n;
	if (o.n <= cap) memcpy(out, o.p, o.n);
	free(o.p); free(pf); free(r2r); free(hist); free(M.t.freq);
	return ret;
}
