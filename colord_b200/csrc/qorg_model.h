// qorg_model.h — the context model of the lossless quality stream, "-q org" (SURVEY.md §8 row C4), shared by the device kernels.
//
// Restates CQualityCoder::encode_original (src/colord/quality_coder_impl.cpp:78-128): one symbol per base = the phred value
// (alphabet 96, quality_code_map_fwd is the identity: quality_coder.cpp:272-280, :356-365, :441-451) under the context
//   [2 previous symbols, quantised to 4 bits each: no_bits_per_symbol = 4, no_ctx_symbols = 2 (quality_coder.cpp:110-114, :183-187)]
//   | base i | base i-1 | (level 3: base i-2; else: base i-2 == base i-1) | base i+1 | (level > 1: match flag, anchor flag)
// with the quantiser of the data source and level (adjust_quality_map_{ONT,PBRaw,PBHiFi}_lossless, quality_coder.cpp:272-338,
// :356-420, :441-505).  The reference feeds the symbols to adaptive 96-symbol models + one range coder; here they go to a Sink.
#pragma once
#include <cstdint>
#include "util.cuh"

namespace clb {

struct QoModel {
	uint32_t A[1], cbits[1], fbits[1];
	uint64_t base[2];
	uint32_t level, source;             // source: 0 ONT, 1 PacBio CLR, 2 PacBio HiFi (DataSource, params.h)
	uint8_t quant[96];
};

inline QoModel make_qorg_model(uint32_t source, uint32_t level)
{
	QoModel m{};
	m.level = level; m.source = source;
	m.A[0] = 96; m.cbits[0] = 8 + (level >= 3 ? 8 : 7) + (level > 1 ? 2 : 0); m.fbits[0] = 8;      // rare contexts fall back to the two previous symbols
	m.base[0] = 0; m.base[1] = 96ull << m.cbits[0];
	auto fill = [&](int a, int b, int v) { for (int i = a; i < b; ++i) m.quant[i] = (uint8_t)v; };
	if (source == 0) {              // adjust_quality_map_ONT_lossless
		m.quant[0] = 0; m.quant[1] = 1;
		if (level >= 3) { const int e[] = {2, 4, 7, 11, 16, 22, 29, 37, 46, 56, 67, 79, 90, 96}; for (int k = 0; k + 1 < 14; ++k) fill(e[k], e[k + 1], 2 + k); }
		else { const int e[] = {2, 5, 10, 15, 20, 25, 35, 50, 70, 96}; for (int k = 0; k + 1 < 10; ++k) fill(e[k], e[k + 1], 2 + k); }
	} else {                        // adjust_quality_map_PBRaw_lossless / _PBHiFi_lossless: same cuts, HiFi shifts the codes by one and gives 93 the code 0
		const int sh = source == 2 ? 1 : 0;
		m.quant[0] = (uint8_t)(0 + sh);
		if (level >= 3) { const int e[] = {1, 10, 20, 30, 39, 45, 51, 57, 63, 69, 75, 81, 87, 93}; for (int k = 0; k + 1 < 14; ++k) fill(e[k], e[k + 1], 1 + k + sh); m.quant[93] = (uint8_t)(source == 2 ? 0 : 14); }
		else { const int e[] = {1, 15, 29, 41, 53, 63, 72, 80, 87, 93}; for (int k = 0; k + 1 < 10; ++k) fill(e[k], e[k + 1], 1 + k + sh); m.quant[93] = (uint8_t)(source == 2 ? 0 : 10); }
		m.quant[94] = m.quant[95] = 0;      // never set by the reference (array value-initialised)
	}
	return m;
}

struct QoReads {
	const uint64_t* pk; const uint64_t* rd_start; const uint32_t* rd_len;      // views starting at the first non-context read
	const uint8_t* quals; const uint64_t* qoff; const uint8_t* flags;          // flags: level > 1, per base 0 / 1 (match) / 2 (anchor)
};

// context of position i (quality_coder_impl.cpp:88-116); q = phred+33 bytes of the read, n its length, rs its first position
CLB_D uint32_t qorg_context(const QoModel& M, const QoReads& R, uint64_t rs, uint32_t n, const uint8_t* q, const uint8_t* fl, uint32_t i)
{
	uint32_t c = 0xff;                                                     // reset_context: ctx_mask
	if (i >= 2) c = ((uint32_t)M.quant[q[i - 2] - 33u] << 4) | M.quant[q[i - 1] - 33u];
	else if (i == 1) c = 0xf0u | M.quant[q[0] - 33u];
	uint32_t sh = 8;
	const uint32_t b0 = base_at(R.pk, rs + i), b1 = i > 0 ? base_at(R.pk, rs + i - 1) : 0u, b2 = i > 1 ? base_at(R.pk, rs + i - 2) : 0u;
	c += b0 << sh; sh += 2;
	c += b1 << sh; sh += 2;                                                // "if (i > 0)": nothing added otherwise
	if (M.level >= 3) { c += b2 << sh; sh += 2; }
	else { if (i > 1) c += (uint32_t)(b2 == b1) << sh; sh += 1; }
	if (i + 1 < n) c += base_at(R.pk, rs + i + 1) << sh;
	sh += 2;
	if (M.level > 1) { c += (uint32_t)(fl[i] == 1) << sh; ++sh; c += (uint32_t)(fl[i] == 2) << sh; }
	return c;
}

} // namespace clb
