// range_sink.cuh — the two sinks an event walk (dna_model.h, hdr_model.h) feeds on the device: a histogram of
// (family, context, symbol) triples for pass 1, and the reference's range encoder over static 12-bit tables for pass 2.
// Model: anything with A[], cbits[], base[] per family (DnaModel, HdrModel).
#pragma once
#include <cstdint>
#include "util.cuh"

namespace clb {

constexpr uint32_t RS_PROB_BITS = 12;

template <class Model>
CLB_HD uint64_t st_entry(const Model& m, uint32_t f, uint64_t ctx, uint32_t sym) { return m.base[f] + (ctx & ((1ull << m.cbits[f]) - 1)) * m.A[f] + sym; }

template <class Model>
struct HistSinkT {
	uint32_t* hist; const Model* M;
	CLB_D void put(uint32_t f, uint64_t ctx, uint32_t sym) { atomicAdd(&hist[st_entry(*M, f, ctx, sym)], 1u); }
};

// CRangeEncoder (sub_rc.h:72-201) with totalFreqSum = 2^12: tab[entry] = frequency | cumulative frequency << 16
template <class Model>
struct RangeSinkT {
	const uint32_t* tab; const Model* M;
	uint8_t* out; uint64_t n;
	unsigned long long low, range;
	uint64_t cap = ~0ull;                                                    // bytes that may be stored at `out`; n keeps counting beyond it
	CLB_D void start() { low = 0; range = 0xff00000000000000ULL; n = 0; }
	CLB_D void byte(uint8_t b) { if (out && n < cap) out[n] = b; ++n; }     // out == nullptr: sizing pass
	CLB_D void put(uint32_t f, uint64_t ctx, uint32_t sym)
	{
		const uint32_t e = tab[st_entry(*M, f, ctx, sym)];
		range >>= RS_PROB_BITS;
		low += range * (e >> 16);
		range *= (e & 0xffffu);
		while (range <= 0x0000ffffffffffffULL) {
			if ((low ^ (low + range)) & 0xff00000000000000ULL) { const unsigned long long r = low; range = (r | 0x0000ffffffffffffULL) - r; }
			byte((uint8_t)(low >> 56));
			low <<= 8; range <<= 8;
		}
	}
	CLB_D void end() { for (int i = 0; i < 8; ++i) { byte((uint8_t)(low >> 56)); low <<= 8; } }
};

} // namespace clb
