// util.cuh — device/host helpers shared by the stage kernels (sm_100a).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define CLB_HD __host__ __device__ __forceinline__
#define CLB_D __device__ __forceinline__

namespace clb {

constexpr uint64_t EMPTY64 = ~0ULL;
constexpr uint32_t EMPTY32 = 0xFFFFFFFFu;

// MurmurHash3 fmix64 — the filter hash (reference: filter_kmers.cpp:24-32 == hash_filter.h:8-16).
CLB_HD uint64_t murmur64(uint64_t x)
{
	x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
	x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
	x ^= x >> 33;
	return x;
}

// h % f == 0 without a division: f = 2^sh * m (m odd)  =>  low sh bits are zero and
// (h >> sh) * inv(m) mod 2^64 <= floor((2^64-1)/m)   (Granlund–Montgomery exact-division test).
struct ModTest {
	uint64_t inv_m, lim, low_mask;
	uint32_t sh;
};
inline ModTest make_modtest(uint32_t f)
{
	ModTest t{};
	uint32_t sh = 0;
	while ((f & 1u) == 0) { f >>= 1; ++sh; }
	uint64_t m = f, inv = m;              // Newton iteration for the inverse of an odd number mod 2^64
	for (int i = 0; i < 6; ++i) inv *= 2 - m * inv;
	t.inv_m = inv; t.lim = ~0ULL / m; t.sh = sh; t.low_mask = (1ULL << sh) - 1;
	return t;
}
CLB_HD bool divisible(uint64_t h, const ModTest& t)
{
	return ((h & t.low_mask) == 0) & (((h >> t.sh) * t.inv_m) <= t.lim);
}

// Reverse complement of a 2-bit packed k-mer (A,C,G,T = 0..3; complement = 3 - x).
CLB_D uint64_t revcomp(uint64_t x, uint32_t k)
{
	uint64_t y = __brevll(~x);
	y = ((y >> 1) & 0x5555555555555555ULL) | ((y & 0x5555555555555555ULL) << 1);
	return y >> (64 - 2 * k);
}

// Table slot from the (already computed) murmur hash: the filter makes h a multiple of f, so the low
// bits are biased; a Fibonacci multiply of h and the top bits spread it again.
CLB_HD uint64_t slot_of(uint64_t h, uint32_t log2cap)
{
	return (h * 0x9E3779B97F4A7C15ULL) >> (64 - log2cap);
}

// Packed bases: 32 per 64-bit word, base p of the device stream in word p>>5 at bit 62-2*(p&31)
// (MSB first, as CReferenceReads packs 4 per byte: reference_reads.h:35-72).
CLB_D uint32_t base_at(const uint64_t* __restrict__ pk, uint64_t p)
{
	return (uint32_t)(pk[p >> 5] >> (62 - 2 * (p & 31))) & 3u;
}

// Inclusive warp scan (sum).
CLB_D uint32_t warp_incl_scan(uint32_t v)
{
	const uint32_t lane = threadIdx.x & 31;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		uint32_t n = __shfl_up_sync(0xffffffffu, v, d);
		if (lane >= (uint32_t)d) v += n;
	}
	return v;
}

// Block exclusive scan of one value per thread; returns the exclusive prefix, *total = block sum.
// `ws` = shared scratch of 33 uint32.  All threads must call.
CLB_D uint32_t block_excl_scan(uint32_t v, uint32_t* ws, uint32_t* total)
{
	const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
	uint32_t inc = warp_incl_scan(v);
	__syncthreads();                       // protect ws from a previous use
	if (lane == 31) ws[w] = inc;
	__syncthreads();
	if (w == 0) {
		uint32_t x = lane < nw ? ws[lane] : 0;
		uint32_t xi = warp_incl_scan(x);
		ws[lane] = xi - x;
		if (lane == 31) ws[32] = xi;
	}
	__syncthreads();
	*total = ws[32];
	return inc - v + ws[w];
}

} // namespace clb
