// dna_model.h — the context model of the DNA / edit-script stream, shared by the device kernels.
//
// Restates the modelling part of the reference's CDNACoder (src/colord/dna_coder.cpp): which events a read's tuples turn into
// and under which context each one is coded —
//   :26-231   Encode (the walk over the tuples, reference positions, symbol / tuple histories, indel drift)
//   :440-463  read flag            :1004-1056 read length           :1178-1227 plain symbols (with / without N)
//   :535-551  reference read id    :489-509   reverse-complement flag (once per distinct reference read and read)
//   :651-717  tuple type           :958-978   anchor length          :772-811 / :889-922 insertion / substitution
//   :1102-1137 skip length (local / distant)  :572-615 alternative read id (seen flag, short id, full id)
// The reference feeds the events to adaptive models + one range coder; here they go to a Sink (histogram or static coder).
// Contexts are built with the reference's "+=" arithmetic (a guard symbol 255 carries into higher fields) and then folded
// to the family's table width.
#pragma once
#include <cstdint>
#include "util.cuh"

namespace clb {

enum DnaFam : uint32_t { F_FLAG = 0, F_LENBITS, F_LENDATA, F_SYM, F_SYMN, F_READID, F_REV, F_TUPLE, F_ANCHOR, F_SKIPL, F_SKIPD, F_SEEN, F_SHORT, F_COUNT };

struct DnaModel {
	uint32_t level, n_t, n_s;            // compression level, tuple types / symbols kept in the histories (dna_coder.cpp:1253-1280)
	uint32_t A[F_COUNT];                 // alphabet size per family
	uint32_t cbits[F_COUNT];             // context width per family (table = 2^cbits x A)
	uint32_t fbits[F_COUNT];             // width of the fallback context (contexts seen rarely share it); 0 = none
	uint64_t base[F_COUNT + 1];          // first table entry per family
};

inline DnaModel make_dna_model(uint32_t level, uint32_t max_candidates)
{
	DnaModel m{};
	m.level = level; m.n_t = level >= 3 ? 4 : level == 2 ? 3 : 2; m.n_s = level >= 3 ? 8 : level == 2 ? 7 : 5;
	const uint32_t A[F_COUNT] = {3, 32, 256, 4, 5, 256, 2, 8, 24, 256, 256, 2, max_candidates < 2 ? 2 : max_candidates};
	const uint32_t sym_bits = level >= 3 ? 24 : level == 2 ? 23 : 22;
	const uint32_t cb[F_COUNT] = {8, 0, 9, sym_bits, 2 * m.n_s, 11, 4, 3 * m.n_t + 9, 6, 6, 8, 6, 6};
	const uint32_t fb[F_COUNT] = {0, 0, 0, 10, 0, 0, 0, 3 * m.n_t + 6, 0, 0, 0, 0, 0};
	uint64_t at = 0;
	for (uint32_t f = 0; f < F_COUNT; ++f) { m.A[f] = A[f]; m.cbits[f] = cb[f]; m.fbits[f] = fb[f]; m.base[f] = at; at += ((uint64_t)A[f]) << cb[f]; }
	m.base[F_COUNT] = at;
	return m;
}

CLB_HD uint64_t dna_entry(const DnaModel& m, uint32_t f, uint64_t ctx, uint32_t sym) { return m.base[f] + (ctx & ((1ull << m.cbits[f]) - 1)) * m.A[f] + sym; }
CLB_HD uint32_t ilog2_bits(uint64_t x) { uint32_t r = 0; for (; x; ++r) x >>= 1; return r; }
CLB_HD uint32_t no_bytes_of(uint64_t x) { uint32_t r = 1; x >>= 8; for (; x; ++r) x >>= 8; return r; }

struct DnaReads {                       // what the walk needs from the resident read store
	const uint64_t* pk; const uint64_t* rd_start; const uint32_t* rd_len; const uint32_t* ref_to_read;
	const uint8_t* es; const uint64_t* es_off;
	uint32_t first;                     // reads are numbered from the first non-context read: read r of the walk is read first + r of the store
};
struct OrientedRef { uint64_t start; uint32_t len; uint32_t rev; };
CLB_D OrientedRef oriented(const DnaReads& R, uint32_t ref_id, uint32_t rev) { const uint32_t rr = R.ref_to_read[ref_id]; return OrientedRef{R.rd_start[rr], R.rd_len[rr], rev}; }
CLB_D uint32_t ref_sym(const DnaReads& R, const OrientedRef& o, int pos)
{
	if (pos < 0 || (uint32_t)pos >= o.len) return 255u;               // the guard byte of read_t
	return o.rev ? 3u - base_at(R.pk, o.start + (o.len - 1 - (uint32_t)pos)) : base_at(R.pk, o.start + (uint32_t)pos);
}
CLB_D uint32_t read_flag_of(const DnaReads& R, uint32_t r) { const uint32_t t0 = R.es[R.es_off[R.first + r]] >> 4; return t0 == 9 ? 0u : t0 == 11 ? 1u : 2u; }

// The events of read r, in coding order.  ctx_read_type: the last read flags seen by this coder lane (dna_coder.cpp:459-462).
// sink.put(family, context, symbol).  EXACT: the events as the reference's adaptive coder sees them (stage3_exact.cu) — the symbols a
// tuple type / substitution cannot be are handed over as a mask (EncodeExcluding: dna_coder.cpp:651-717, :889-922; sink.putx) and the
// chunk index of an anchor / local skip length is not capped.
template <bool EXACT = false, class Sink>
__device__ void dna_walk(const DnaModel& M, const DnaReads& R, uint32_t r, uint32_t ctx_read_type, Sink& sink)
{
	const uint8_t* t = R.es + R.es_off[R.first + r];
	const uint64_t tn = R.es_off[R.first + r + 1] - R.es_off[R.first + r];
	uint32_t n_tuples = 0;
	for (uint64_t p = 0; p < tn; ++n_tuples) { const uint32_t ty = t[p] >> 4; p += (ty == 4 || ty == 5) ? 4 : (ty == 6 || ty == 10) ? 5 : 1; }
	const uint32_t flag = (t[0] >> 4) == 9 ? 0u : (t[0] >> 4) == 11 ? 1u : 2u;
	sink.put(F_FLAG, ctx_read_type, flag);
	{	// read length = number of tuples after the start tuple
		uint32_t len = n_tuples - 1;
		const uint32_t nbits = ilog2_bits(len);
		sink.put(F_LENBITS, 0, nbits);
		if (nbits >= 2) {
			uint64_t ctx = (uint64_t)nbits << 3;
			len -= 1u << (nbits - 1);
			uint32_t prefix = len, suffix = 0;
			if (nbits > 9) { prefix = len >> (nbits - 9); suffix = len - (prefix << (nbits - 9)); }
			sink.put(F_LENDATA, ctx, prefix);
			if (nbits > 9) { ctx += 4; for (int nb = (int)nbits - 9; nb > 0; nb -= 8) { sink.put(F_LENDATA, ctx, suffix & 0xff); suffix >>= 8; ++ctx; } }
		}
	}
	const uint64_t mask_s = (1ull << (2 * M.n_s)) - 1, mask_t = (1ull << (3 * M.n_t)) - 1;
	uint64_t ctx_symbol = mask_s, ctx_tuple = mask_t;
	if (flag == 0) { for (uint64_t p = 1; p < tn; ++p) { const uint32_t s = t[p] & 15; sink.put(F_SYM, ctx_symbol << 2, s); ctx_symbol = ((ctx_symbol << 2) + s) & mask_s; } return; }
	if (flag == 1) { for (uint64_t p = 1; p < tn; ++p) { const uint32_t s = t[p] & 15; sink.put(F_SYMN, ctx_symbol, s); ctx_symbol = ((ctx_symbol << 4) + s) & mask_s; } return; }

	auto be32 = [&](uint64_t p) { return ((uint32_t)t[p] << 24) | ((uint32_t)t[p + 1] << 16) | ((uint32_t)t[p + 2] << 8) | t[p + 3]; };
	auto put_read_id = [&](uint32_t id) {
		const int n = (int)no_bytes_of(R.first + r);      // reference ids stay below the read's index in the store
		for (int i = n - 1; i >= 0; --i) { const uint64_t add = (i == n - 2) ? ((id >> (8 * (n - 1))) & 0xff) : 0; sink.put(F_READID, (uint64_t)i + (add << 3), (id >> (8 * i)) & 0xff); }
	};
	uint32_t seen_id[34]; uint32_t n_seen = 0; uint64_t ctx_rev = 0xf;         // uo_rev_comp of this read
	auto put_rev = [&](uint32_t id, uint32_t rev) {
		for (uint32_t k = 0; k < n_seen; ++k) if (seen_id[k] == id) return;
		sink.put(F_REV, ctx_rev, rev);
		if (n_seen < 34) seen_id[n_seen++] = id;
		ctx_rev = ((ctx_rev << 2) + rev) & 0xf;
	};
	auto put_skip = [&](uint32_t len, bool local) {
		if (local) { for (uint32_t part = 0; len; ++part) { const uint32_t pc = EXACT ? part : min(part, 63u); if (len < 255) { sink.put(F_SKIPL, pc, len); break; } sink.put(F_SKIPL, pc, 255); len -= 254; } }
		else { uint32_t enc = 0; for (int i = 3; i >= 0; --i) { const uint32_t x = (len >> (8 * i)) & 0xff; sink.put(F_SKIPD, (uint64_t)i * 64 + ilog2_bits(enc), x); enc = (enc << 8) + x; } }
	};
	const uint32_t main_id = be32(1), main_rev = t[0] & 15;
	put_read_id(main_id);
	put_rev(main_id, main_rev);
	const OrientedRef main_ref = oriented(R, main_id, main_rev);
	OrientedRef alt_ref = main_ref;
	uint32_t alt_ids[32], alt_revs[32]; int alt_saved[32]; uint32_t n_alt = 0; int cur_alt = -1;      // m_alt_ids / m_alt_read / m_alt_pos
	int ref_pos = 0, alt_pos = 0, delta = 0;
	bool is_main = true;
	uint32_t last_tuple = 255;
	const uint32_t sh_t = 3 * M.n_t;
	for (uint64_t p = 5; p < tn;) {
		const uint32_t ty = t[p] >> 4, v1 = t[p] & 15;
		uint32_t v2 = 0;
		if (ty == 4 || ty == 5) { v2 = ((uint32_t)v1 << 24) | ((uint32_t)t[p + 1] << 16) | ((uint32_t)t[p + 2] << 8) | t[p + 3]; p += 4; }
		else if (ty == 6) { v2 = be32(p + 1); p += 5; }
		else p += 1;
		const uint32_t rsym = is_main ? ref_sym(R, main_ref, ref_pos) : ref_sym(R, alt_ref, alt_pos);
		{	// tuple type (:651-717)
			uint64_t ctx = ctx_tuple + ((ctx_symbol & 0xf) << sh_t) + ((uint64_t)rsym << (sh_t + 4));
			const uint32_t bucket = delta < -10 ? 1 : delta < -1 ? 2 : delta > 10 ? 3 : delta > 1 ? 4 : 0;
			ctx += (uint64_t)bucket << (sh_t + 6);
			if constexpr (EXACT) {
				const uint32_t excl = last_tuple == 2 ? 1u << 4 : last_tuple == 1 ? 1u << 5 : last_tuple == 4 ? (1u << 4) | (1u << 2)
					: last_tuple == 5 ? (1u << 1) | (1u << 5) : (last_tuple == 7 || last_tuple == 6) ? (1u << 6) | (1u << 7) : 0u;
				sink.putx(F_TUPLE, ctx, ty, excl);
			} else sink.put(F_TUPLE, ctx, ty);
			ctx_tuple = ((ctx_tuple << 3) + ty) & mask_t;
		}
		if (ty == 6) {               // alt_id: v2 = id, v1 = reverse-complement flag
			if (!is_main && cur_alt >= 0) alt_saved[cur_alt] = alt_pos;
			int idx = -1;
			for (uint32_t k = 0; k < n_alt; ++k) if (alt_ids[k] == v2) { idx = (int)k; break; }
			if (n_alt == 0) put_read_id(v2);
			else {
				sink.put(F_SEEN, n_alt, idx >= 0);
				if (idx < 0) put_read_id(v2); else sink.put(F_SHORT, n_alt, (uint32_t)idx);
			}
			if (idx < 0 && n_alt < 32) { idx = (int)n_alt; alt_ids[n_alt] = v2; alt_revs[n_alt] = v1; alt_saved[n_alt] = 0; ++n_alt; }
			cur_alt = idx;
			put_rev(v2, v1);
			alt_ref = oriented(R, v2, idx >= 0 ? alt_revs[idx] : v1);
			alt_pos = 0; is_main = false; delta = 0;
		} else if (ty == 4) {        // anchor
			for (uint32_t len = v2, part = 0; len; ++part) { const uint32_t pc = EXACT ? part : min(part, 63u); if (len < 23) { sink.put(F_ANCHOR, pc, len); break; } sink.put(F_ANCHOR, pc, 23); len -= 22; }
			int& pos = is_main ? ref_pos : alt_pos;
			pos += (int)v2;
			const OrientedRef& o = is_main ? main_ref : alt_ref;
			for (int i = (int)M.n_s; i > 0; --i) ctx_symbol = (ctx_symbol << 2) + ref_sym(R, o, pos - i);
			ctx_symbol &= mask_s;
			delta = 0;
		} else if (ty == 2) {        // match
			ctx_symbol = ((ctx_symbol << 2) + rsym) & mask_s;
			is_main ? ++ref_pos : ++alt_pos;
		} else if (ty == 0) {        // insertion (:772-811)
			uint64_t ctx = 2; uint32_t sh = 2;
			if (M.level <= 1) { ctx += (ctx_symbol & 0xff) << sh; sh += 8; }
			else { ctx += (ctx_symbol & 0x3ff) << sh; sh += 10; if (M.level >= 3) { ctx += (uint64_t)(((ctx_symbol >> 10) & 3) == ((ctx_symbol >> 8) & 3)) << sh; ++sh; } }
			ctx += (uint64_t)rsym << sh; sh += 2;
			ctx += (ctx_tuple & 0777) << sh;
			sink.put(F_SYM, ctx, v1);
			ctx_symbol = ((ctx_symbol << 2) + v1) & mask_s;
			++delta;
		} else if (ty == 1) {        // deletion
			is_main ? ++ref_pos : ++alt_pos;
			--delta;
		} else if (ty == 3) {        // substitution (:889-922): the coded symbol is the base itself
			const uint32_t b = rsym & 3;
			const uint32_t symbol = v1 + (v1 >= b ? 1u : 0u);              // subst_to_code (dna_coder.h:37): code-th base other than the reference base
			uint64_t ctx = 1; uint32_t sh = 2;
			ctx += (ctx_symbol & 0x3f) << sh; sh += 6;
			if (M.level >= 3) { ctx += (uint64_t)(((ctx_symbol >> 6) & 3) == ((ctx_symbol >> 4) & 3)) << sh; ++sh; }
			ctx += (uint64_t)rsym << sh; sh += 2;
			ctx += (ctx_tuple & 07777) << sh;
			if constexpr (EXACT) sink.putx(F_SYM, ctx, symbol, 1u << b); else sink.put(F_SYM, ctx, symbol);
			ctx_symbol = ((ctx_symbol << 2) + symbol) & mask_s;
			is_main ? ++ref_pos : ++alt_pos;
		} else if (ty == 5) {        // skip (:166-206)
			const int skip_len = (int)v2;
			delta -= skip_len;
			if (!is_main && last_tuple == 6) {
				const int mod = skip_len - (cur_alt >= 0 ? alt_saved[cur_alt] : 0);
				if (mod > 0) put_skip((uint32_t)mod, false);
				else { put_skip(0, false); put_skip((uint32_t)(-mod), false); }
			} else put_skip((uint32_t)skip_len, last_tuple != 6 && last_tuple != 255);
			is_main ? (ref_pos += skip_len) : (alt_pos += skip_len);
		} else if (ty == 7) {        // main_ref
			is_main = true;
			if (cur_alt >= 0) alt_saved[cur_alt] = alt_pos;
			delta = 0;
		}
		last_tuple = ty;
	}
}

} // namespace clb
