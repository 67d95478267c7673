/* oracle/rc_static.h — TEST INFRASTRUCTURE ONLY.
 * Shared by the native-container decoders (stage3_dna.c, stage3_hdr.c): the static frequency tables as the device serialises
 * them (colord_b200/csrc/static_tables.h) and the reference's range decoder (src/colord/sub_rc.h:262-386: 64-bit low / range /
 * buffer, carry-less renormalisation byte by byte) with totalFreq = 2^12. */
#ifndef ORC_RC_STATIC_H
#define ORC_RC_STATIC_H
#define ST_FN static __attribute__((unused))
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define ST_M 4096u
#define ST_MAX_FAM 16
typedef struct { uint32_t n_fam, A[ST_MAX_FAM], cbits[ST_MAX_FAM], fbits[ST_MAX_FAM]; uint64_t base[ST_MAX_FAM + 1]; uint16_t* freq; } st_model;

ST_FN void st_layout(st_model* m)
{
	uint64_t at = 0;
	for (uint32_t f = 0; f < m->n_fam; ++f) { m->base[f] = at; at += ((uint64_t)m->A[f]) << m->cbits[f]; }
	m->base[m->n_fam] = at;
}
ST_FN uint64_t st_get_freqs(const uint8_t* in, uint64_t at, uint16_t* f, uint32_t A)
{
	memset(f, 0, 2 * A);
	if (A <= 8) {
		const uint8_t mask = in[at++]; int last = -1; uint32_t sum = 0;
		for (uint32_t k = 0; k < A; ++k) if (mask >> k & 1) last = (int)k;
		for (int k = 0; k < last; ++k) if (mask >> k & 1) { uint16_t v; memcpy(&v, in + at, 2); at += 2; f[k] = v; sum += v; }
		if (last >= 0) f[last] = (uint16_t)(ST_M - sum);
	} else {
		uint16_t nz; memcpy(&nz, in + at, 2); at += 2;
		for (uint32_t i = 0; i < nz; ++i) { const uint8_t k = in[at++]; uint16_t v; memcpy(&v, in + at, 2); at += 2; f[k] = v; }
	}
	return at;
}
/* reads the tables of all families; returns the position after them */
ST_FN uint64_t st_read_tables(st_model* m, const uint8_t* in, uint64_t at)
{
	m->freq = (uint16_t*)calloc(m->base[m->n_fam] + 1, 2);
	uint16_t fr[256];
	for (uint32_t f = 0; f < m->n_fam; ++f) {
		const uint32_t A = m->A[f]; const uint64_t n_ctx = 1ull << m->cbits[f], n_fb = m->fbits[f] ? (1ull << m->fbits[f]) : 0;
		uint16_t* dst = m->freq + m->base[f];
		if (n_fb) {
			uint16_t* fbf = (uint16_t*)calloc(n_fb * A, 2);
			for (uint64_t x = 0; x < n_fb; ++x) at = st_get_freqs(in, at, fbf + x * A, A);
			for (uint64_t x = 0; x < n_ctx; ++x) memcpy(dst + x * A, fbf + (x & (n_fb - 1)) * A, 2 * A);
			free(fbf);
		}
		uint32_t nd; memcpy(&nd, in + at, 4); at += 4;
		uint64_t x = 0;
		for (uint32_t d = 0; d < nd; ++d) {
			uint64_t gap = 0; uint32_t sh = 0; uint8_t by;
			do { by = in[at++]; gap |= (uint64_t)(by & 127) << sh; sh += 7; } while (by & 128);
			x += gap;
			at = st_get_freqs(in, at, fr, A);
			memcpy(dst + x * A, fr, 2 * A);
		}
	}
	return at;
}

typedef struct { const uint8_t* p; uint64_t n, at; uint64_t low, range, buffer; } rcdec;
ST_FN uint8_t rc_byte(rcdec* d) { return d->at < d->n ? d->p[d->at++] : 0; }
ST_FN void rc_start(rcdec* d, const uint8_t* p, uint64_t n) { d->p = p; d->n = n; d->at = 0; d->buffer = 0; for (int i = 0; i < 8; ++i) d->buffer = (d->buffer << 8) + rc_byte(d); d->low = 0; d->range = 0xff00000000000000ULL; }
ST_FN uint32_t rc_get(rcdec* d, const st_model* m, uint32_t f, uint64_t ctx)
{
	const uint16_t* fr = m->freq + m->base[f] + (ctx & ((1ull << m->cbits[f]) - 1)) * m->A[f];
	d->range >>= 12;
	const uint64_t cf = d->buffer / d->range;
	uint32_t s = 0; uint64_t acc = 0;
	while (s + 1 < m->A[f] && acc + fr[s] <= cf) { acc += fr[s]; ++s; }
	const uint64_t r = acc * d->range;
	d->buffer -= r; d->low += r; d->range *= fr[s];
	while (d->range <= 0x0000ffffffffffffULL) {
		if ((d->low ^ (d->low + d->range)) & 0xff00000000000000ULL) { const uint64_t x = d->low; d->range = (x | 0x0000ffffffffffffULL) - x; }
		d->buffer = (d->buffer << 8) + rc_byte(d);
		d->low <<= 8; d->range <<= 8;
	}
	return s;
}

/* ---- encoder side (CPU twins of the device containers): counts -> tables + serialisation, the range encoder ---- */
typedef struct { uint8_t* p; uint64_t n, cap; } st_buf;
ST_FN void st_push(st_buf* b, const void* v, uint64_t k) { if (b->n + k > b->cap) { b->cap = (b->n + k) * 2 + 64; b->p = (uint8_t*)realloc(b->p, b->cap); } memcpy(b->p + b->n, v, k); b->n += k; }
ST_FN void st_push8(st_buf* b, uint8_t v) { st_push(b, &v, 1); }
/* counts of one context -> 12-bit frequencies: floor scaling, present symbols keep at least 1, the remainder goes to the most
 * frequent symbol (first one on ties); an overshoot caused by the +1 floors is taken from the largest frequencies */
ST_FN void st_normalise(const uint32_t* cnt, uint32_t n, uint16_t* f)
{
	uint64_t tot = 0; uint32_t best = 0, sum = 0;
	for (uint32_t i = 0; i < n; ++i) { tot += cnt[i]; if (cnt[i] > cnt[best]) best = i; }
	if (!tot) { memset(f, 0, 2 * n); return; }
	for (uint32_t i = 0; i < n; ++i) { uint32_t v = (uint32_t)(((uint64_t)cnt[i] << 12) / tot); if (cnt[i] && !v) v = 1; f[i] = (uint16_t)v; sum += v; }
	if (sum > ST_M) {
		uint32_t over = sum - ST_M;
		while (over) { uint32_t b = 0; for (uint32_t i = 1; i < n; ++i) if (f[i] > f[b]) b = i; const uint32_t d = over < (uint32_t)f[b] - 1 ? over : (uint32_t)f[b] - 1; f[b] = (uint16_t)(f[b] - d); over -= d; if (!d) break; }
	} else f[best] = (uint16_t)(f[best] + ST_M - sum);
}
ST_FN void st_put_freqs(st_buf* o, const uint16_t* f, uint32_t A)
{
	if (A <= 8) {
		uint8_t mask = 0; int last = -1;
		for (uint32_t k = 0; k < A; ++k) if (f[k]) { mask |= (uint8_t)(1u << k); last = (int)k; }
		st_push8(o, mask);
		for (int k = 0; k < last; ++k) if (f[k]) st_push(o, &f[k], 2);
	} else {
		uint16_t nz = 0; for (uint32_t k = 0; k < A; ++k) nz += f[k] != 0;
		st_push(o, &nz, 2);
		for (uint32_t k = 0; k < A; ++k) if (f[k]) { st_push8(o, (uint8_t)k); st_push(o, &f[k], 2); }
	}
}
ST_FN uint32_t st_bits_q8(uint32_t f)      /* round(-log2(f / 4096) * 256) */
{
	static uint32_t t[ST_M + 1]; static int ready = 0;
	if (!ready) { for (uint32_t i = 1; i <= ST_M; ++i) t[i] = (uint32_t)lround(-log2(i / 4096.0) * 256.0); ready = 1; }
	return t[f];
}
/* hist[base[f] + ctx * A + sym] -> m->freq (allocated here) and the serialised tables appended to hdr */
ST_FN void st_write_tables(st_model* m, const uint32_t* hist, st_buf* hdr, uint32_t min_ctx)
{
	m->freq = (uint16_t*)calloc(m->base[m->n_fam] + 1, 2);
	uint16_t fr[256];
	for (uint32_t f = 0; f < m->n_fam; ++f) {
		const uint32_t A = m->A[f]; const uint64_t n_ctx = 1ull << m->cbits[f], n_fb = m->fbits[f] ? (1ull << m->fbits[f]) : 0;
		const uint32_t* h = hist + m->base[f]; uint16_t* dst = m->freq + m->base[f];
		uint32_t* fbh = (uint32_t*)calloc(n_fb * A + 1, 4); uint16_t* fbf = (uint16_t*)calloc(n_fb * A + 1, 2);
		uint8_t* dense = (uint8_t*)calloc(n_ctx, 1);
		uint32_t nd = 0;
		if (n_fb) {      /* pooled table of every fallback cell over all its contexts: the alternative a context is compared with */
			for (uint64_t x = 0; x < n_ctx; ++x) for (uint32_t k = 0; k < A; ++k) fbh[(x & (n_fb - 1)) * A + k] += h[x * A + k];
			for (uint64_t x = 0; x < n_fb; ++x) st_normalise(fbh + x * A, A, fbf + x * A);
			memset(fbh, 0, (n_fb * A + 1) * 4);
		}
		for (uint64_t x = 0; x < n_ctx; ++x) {
			uint64_t t = 0; for (uint32_t k = 0; k < A; ++k) t += h[x * A + k];
			if (!t) continue;
			int own = !n_fb;
			if (n_fb && t >= min_ctx) {      /* own table iff the bits it saves exceed the bits of its serialisation (1/256-bit units) */
				st_normalise(h + x * A, A, fr);
				const uint16_t* pf = fbf + (x & (n_fb - 1)) * A;
				uint64_t c_own = 0, c_fb = 0; uint32_t nz = 0;
				for (uint32_t k = 0; k < A; ++k) { nz += fr[k] != 0; if (h[x * A + k]) { c_own += (uint64_t)h[x * A + k] * st_bits_q8(fr[k]); c_fb += (uint64_t)h[x * A + k] * st_bits_q8(pf[k]); } }
				const uint64_t bytes = A <= 8 ? 1 + 2ull * (nz ? nz - 1 : 0) : 2 + 3ull * nz;
				own = c_fb > c_own + (bytes + 2) * 8 * 256;
			}
			if (own) { dense[x] = 1; ++nd; }
			else for (uint32_t k = 0; k < A; ++k) fbh[(x & (n_fb - 1)) * A + k] += h[x * A + k];
		}
		for (uint64_t x = 0; x < n_fb; ++x) { st_normalise(fbh + x * A, A, fbf + x * A); st_put_freqs(hdr, fbf + x * A, A); }
		st_push(hdr, &nd, 4);
		uint64_t prev = 0;
		for (uint64_t x = 0; x < n_ctx; ++x) {
			if (dense[x]) {
				st_normalise(h + x * A, A, fr);
				uint64_t gap = x - prev; prev = x;
				do { uint8_t by = (uint8_t)(gap & 127); gap >>= 7; if (gap) by |= 128; st_push8(hdr, by); } while (gap);
				st_put_freqs(hdr, fr, A);
				memcpy(dst + x * A, fr, 2 * A);
			} else if (n_fb) memcpy(dst + x * A, fbf + (x & (n_fb - 1)) * A, 2 * A);
		}
		free(fbh); free(fbf); free(dense);
	}
}
/* CRangeEncoder (src/colord/sub_rc.h:72-201) with totalFreq = 2^12 */
typedef struct { st_buf* o; uint64_t low, range; } rcenc;
ST_FN void rce_start(rcenc* e, st_buf* o) { e->o = o; e->low = 0; e->range = 0xff00000000000000ULL; }
ST_FN void rce_put(rcenc* e, const st_model* m, uint32_t f, uint64_t ctx, uint32_t sym)
{
	const uint16_t* fr = m->freq + m->base[f] + (ctx & ((1ull << m->cbits[f]) - 1)) * m->A[f];
	uint64_t cum = 0; for (uint32_t k = 0; k < sym; ++k) cum += fr[k];
	e->range >>= 12; e->low += e->range * cum; e->range *= fr[sym];
	while (e->range <= 0x0000ffffffffffffULL) {
		if ((e->low ^ (e->low + e->range)) & 0xff00000000000000ULL) { const uint64_t x = e->low; e->range = (x | 0x0000ffffffffffffULL) - x; }
		st_push8(e->o, (uint8_t)(e->low >> 56));
		e->low <<= 8; e->range <<= 8;
	}
}
ST_FN void rce_end(rcenc* e) { for (int i = 0; i < 8; ++i) { st_push8(e->o, (uint8_t)(e->low >> 56)); e->low <<= 8; } }
#endif
