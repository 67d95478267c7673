/* oracle/rc_static.h — TEST INFRASTRUCTURE ONLY.
 * Shared by the native-container decoders (stage3_dna.c, stage3_hdr.c): the static frequency tables as the device serialises
 * them (colord_b200/csrc/static_tables.h) and the reference's range decoder (src/colord/sub_rc.h:262-386: 64-bit low / range /
 * buffer, carry-less renormalisation byte by byte) with totalFreq = 2^12. */
#ifndef ORC_RC_STATIC_H
#define ORC_RC_STATIC_H
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ST_M 4096u
#define ST_MAX_FAM 16
typedef struct { uint32_t n_fam, A[ST_MAX_FAM], cbits[ST_MAX_FAM], fbits[ST_MAX_FAM]; uint64_t base[ST_MAX_FAM + 1]; uint16_t* freq; } st_model;

static void st_layout(st_model* m)
{
	uint64_t at = 0;
	for (uint32_t f = 0; f < m->n_fam; ++f) { m->base[f] = at; at += ((uint64_t)m->A[f]) << m->cbits[f]; }
	m->base[m->n_fam] = at;
}
static uint64_t st_get_freqs(const uint8_t* in, uint64_t at, uint16_t* f, uint32_t A)
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
static uint64_t st_read_tables(st_model* m, const uint8_t* in, uint64_t at)
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
static uint8_t rc_byte(rcdec* d) { return d->at < d->n ? d->p[d->at++] : 0; }
static void rc_start(rcdec* d, const uint8_t* p, uint64_t n) { d->p = p; d->n = n; d->at = 0; d->buffer = 0; for (int i = 0; i < 8; ++i) d->buffer = (d->buffer << 8) + rc_byte(d); d->low = 0; d->range = 0xff00000000000000ULL; }
static uint32_t rc_get(rcdec* d, const st_model* m, uint32_t f, uint64_t ctx)
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
#endif
