// hdr_model.h — the event model of the header (read id) stream, shared by the device kernels.
//
// Restates the modelling part of the reference's CIDCoder::compress_lossless (src/colord/id_coder.cpp:210-383):
//   :169-208  tokenize: a header is cut at every character outside [0-9A-Za-z@]; token = (separator after it, begin, end).
//             a_numeric is all false (init_symbol_classes :125-127), so every token is a "literal" and the numeric branch of
//             compress_lossless (:277-355) is dead code: it is not restated.
//   :214-215  the "+ line repeats the header" flag
//   :217-221  flag: do the tokens of this header and of the previous one have the same separators (token_types_same,
//             id_coder.h:123-133), coded under the last 8 flags
//   :225-275  per token: same as the previous header's token? same length? -> the characters that differ (0 = unchanged),
//             or the whole token + terminator
//   :359-373  otherwise the header as plain characters + terminator, context = position
// The reference feeds the events to adaptive models + one range coder; here they go to a Sink (histogram or static coder).
// Contexts: the reference's (token index, position in token) pairs, folded to 5 bits each; a character coded against the
// previous header's character at the same place also sees the low 4 bits of that character (a decimal counter's next digit
// is then almost deterministic).  The first header of a pack is coded without a predecessor, so packs decode independently.
#pragma once
#include <cstdint>
#include "util.cuh"

namespace clb {

enum HdrFam : uint32_t { H_PLUS = 0, H_FLAG, H_SAME, H_SAMELEN, H_LITEQ, H_LITNEW, H_PLAIN, H_COUNT };

struct HdrModel {
	uint32_t A[H_COUNT], cbits[H_COUNT], fbits[H_COUNT];
	uint64_t base[H_COUNT + 1];
};

inline HdrModel make_hdr_model()
{
	HdrModel m{};
	const uint32_t A[H_COUNT] = {2, 2, 2, 2, 256, 256, 256};
	const uint32_t cb[H_COUNT] = {0, 8, 6, 6, 14, 10, 8};
	const uint32_t fb[H_COUNT] = {0, 0, 0, 0, 10, 0, 0};
	uint64_t at = 0;
	for (uint32_t f = 0; f < H_COUNT; ++f) { m.A[f] = A[f]; m.cbits[f] = cb[f]; m.fbits[f] = fb[f]; m.base[f] = at; at += ((uint64_t)A[f]) << cb[f]; }
	m.base[H_COUNT] = at;
	return m;
}

struct HdrInput { const uint8_t* bytes; const uint64_t* off; const uint8_t* plus; };      // header r = bytes[off[r] .. off[r+1])

CLB_HD bool hdr_is_literal(uint8_t c) { return (c >= '0' && c <= '9') || (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z') || c == '@'; }   // id_coder.cpp:121-138

// token_types_same (id_coder.h:123-133) of two headers: same number of tokens and the same separator after each one
CLB_HD bool hdr_same_shape(const uint8_t* a, uint32_t na, const uint8_t* b, uint32_t nb)
{
	uint32_t i = 0, j = 0;
	for (;;) {
		while (i < na && hdr_is_literal(a[i])) ++i;
		while (j < nb && hdr_is_literal(b[j])) ++j;
		if (i == na || j == nb) return i == na && j == nb;
		if (a[i] != b[j]) return false;
		++i; ++j;
	}
}

// The events of header r, in coding order.  has_prev: the header has a predecessor in its pack; flag_ctx: the last 8 flags.
// sink.put(family, context, symbol).  Returns the header's flag.
template <class Sink>
CLB_D uint32_t hdr_walk(const HdrInput& H, uint64_t r, bool has_prev, uint32_t flag_ctx, Sink& sink)
{
	const uint8_t* cur = H.bytes + H.off[r]; const uint32_t nc = (uint32_t)(H.off[r + 1] - H.off[r]);
	const uint8_t* prv = has_prev ? H.bytes + H.off[r - 1] : cur; const uint32_t np = has_prev ? (uint32_t)(H.off[r] - H.off[r - 1]) : 0;
	sink.put(H_PLUS, 0, H.plus ? (H.plus[r] != 0) : 0u);
	const uint32_t flag = has_prev && hdr_same_shape(cur, nc, prv, np);
	sink.put(H_FLAG, flag_ctx, flag);
	if (!flag) {
		for (uint32_t j = 0; j < nc; ++j) sink.put(H_PLAIN, j < 255 ? j : 255, cur[j]);
		sink.put(H_PLAIN, nc < 255 ? nc : 255, 0);
		return 0;
	}
	uint32_t i = 0, j = 0;
	for (uint32_t t = 0;; ++t) {
		uint32_t ie = i, je = j;
		while (ie < nc && hdr_is_literal(cur[ie])) ++ie;
		while (je < np && hdr_is_literal(prv[je])) ++je;
		const uint32_t lc = ie - i, lp = je - j, tc = t < 63 ? t : 63, t5 = t < 31 ? t : 31;
		bool same = lc == lp;
		if (same) for (uint32_t k = 0; k < lc; ++k) if (cur[i + k] != prv[j + k]) { same = false; break; }
		sink.put(H_SAME, tc, same);
		if (!same) {
			sink.put(H_SAMELEN, tc, lc == lp);
			if (lc == lp)
				for (uint32_t k = 0; k < lc; ++k) { const uint32_t c = cur[i + k], p = prv[j + k]; sink.put(H_LITEQ, ((p & 15u) << 10) | (t5 << 5) | (k < 31 ? k : 31), c == p ? 0u : c); }
			else {
				for (uint32_t k = 0; k < lc; ++k) sink.put(H_LITNEW, (t5 << 5) | (k < 31 ? k : 31), cur[i + k]);
				sink.put(H_LITNEW, (t5 << 5) | (lc < 31 ? lc : 31), 0);
			}
		}
		if (ie == nc) break;
		i = ie + 1; j = je + 1;
	}
	return 1;
}

} // namespace clb
