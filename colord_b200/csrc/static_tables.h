// static_tables.h — host side of the native entropy containers: (family, context, symbol) counts -> 12-bit frequency tables,
// their serialisation into the container header and the encoder's lookup table.  Metadata-sized work (one pass over the
// count table), shared by the DNA-tuple and header streams.  Layout of the serialised tables (read by oracle/rc_static.h):
//   per family: [fallback table: 2^fbits contexts] n_dense u32, then per dense context (ascending) LEB128 gap + frequencies;
//   frequencies of one context: alphabets <= 8 as a presence mask + every present frequency but the last (implied by the
//   sum 2^12), larger alphabets as count u16 + (symbol u8, frequency u16) pairs.
// A context is "dense" (own table) when its family has no fallback, or when it was seen at least min_ctx times AND its own
// table pays for itself: the bits it saves over the family's pooled table (all contexts sharing the low fbits of the context)
// exceed the bits its serialisation costs.  All other contexts pool their counts in the fallback table indexed by the low
// fbits of the context.  Costs are taken in 1/256 bit from a table of -log2(f / 4096), so that the CPU twins
// (oracle/rc_static.h) reach the same decisions.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

namespace clb {

constexpr uint32_t ST_PROB_BITS = 12, ST_M = 1u << ST_PROB_BITS;

inline void st_normalise(const uint32_t* cnt, uint32_t n, uint16_t* f)
{
	uint64_t tot = 0; uint32_t best = 0;
	for (uint32_t i = 0; i < n; ++i) { tot += cnt[i]; if (cnt[i] > cnt[best]) best = i; }
	if (!tot) { for (uint32_t i = 0; i < n; ++i) f[i] = 0; return; }
	uint32_t sum = 0;
	for (uint32_t i = 0; i < n; ++i) { uint32_t v = (uint32_t)(((uint64_t)cnt[i] << ST_PROB_BITS) / tot); if (cnt[i] && !v) v = 1; f[i] = (uint16_t)v; sum += v; }
	if (sum > ST_M) {                        // the +1 floors of many rare symbols can overshoot: take it from the largest ones
		uint32_t over = sum - ST_M;
		while (over) { uint32_t b = 0; for (uint32_t i = 1; i < n; ++i) if (f[i] > f[b]) b = i; const uint32_t d = std::min<uint32_t>(over, f[b] - 1); f[b] = (uint16_t)(f[b] - d); over -= d; if (!d) break; }
	} else f[best] = (uint16_t)(f[best] + ST_M - sum);
}
template <typename T> inline void st_put(std::vector<uint8_t>& o, const T& v) { const uint8_t* p = reinterpret_cast<const uint8_t*>(&v); o.insert(o.end(), p, p + sizeof(T)); }
inline void st_put_freqs(std::vector<uint8_t>& o, const uint16_t* f, uint32_t A)
{
	if (A <= 8) {
		uint8_t mask = 0; int last = -1;
		for (uint32_t k = 0; k < A; ++k) if (f[k]) { mask |= (uint8_t)(1u << k); last = (int)k; }
		o.push_back(mask);
		for (int k = 0; k < last; ++k) if (f[k]) st_put(o, f[k]);
	} else {
		uint16_t nz = 0; for (uint32_t k = 0; k < A; ++k) nz += f[k] != 0;
		st_put(o, nz);
		for (uint32_t k = 0; k < A; ++k) if (f[k]) { o.push_back((uint8_t)k); st_put(o, f[k]); }
	}
}

inline const uint32_t* st_bits_q8()      // [f] = round(-log2(f / 4096) * 256), f = 1 .. 4096
{
	static const std::vector<uint32_t> t = [] { std::vector<uint32_t> v(ST_M + 1, 0); for (uint32_t f = 1; f <= ST_M; ++f) v[f] = (uint32_t)std::lround(-std::log2(f / 4096.0) * 256.0); return v; }();
	return t.data();
}
inline uint64_t st_freqs_bytes(const uint16_t* f, uint32_t A)
{
	uint32_t nz = 0; for (uint32_t k = 0; k < A; ++k) nz += f[k] != 0;
	return A <= 8 ? 1 + 2ull * (nz ? nz - 1 : 0) : 2 + 3ull * nz;
}

// fn(t, lo, hi) for thread t over its share [lo, hi) of n items; on one thread below `min_items`
template <class Fn>
inline void st_parallel(uint64_t n, uint64_t min_items, Fn&& fn)
{
	unsigned T = std::min<uint64_t>(std::max(1u, std::thread::hardware_concurrency()), 16);
	if (n < min_items) T = 1;
	if (T <= 1) { fn(0u, (uint64_t)0, n); return; }
	std::vector<std::thread> th;
	for (unsigned t = 0; t < T; ++t) th.emplace_back([&fn, t, T, n] { fn(t, n * t / T, n * (t + 1) / T); });
	for (auto& x : th) x.join();
}
inline unsigned st_threads(uint64_t n, uint64_t min_items) { return n < min_items ? 1u : (unsigned)std::min<uint64_t>(std::max(1u, std::thread::hardware_concurrency()), 16); }

// M: anything with A[], cbits[], fbits[], base[] per family.  tab[entry] = frequency | cumulative << 16.
// The big families (millions of contexts) are walked by several host threads: the pooled counts are summed per thread and then
// added up, the per-context decisions and table entries are independent, only the serialisation of the dense contexts is serial.
template <class M>
void st_build_tables(const M& m, uint32_t n_fam, const std::vector<uint32_t>& hist, std::vector<uint32_t>& tab, std::vector<uint8_t>& hdr, uint32_t min_ctx)
{
	constexpr uint64_t PAR = 1u << 16;
	for (uint32_t f = 0; f < n_fam; ++f) {
		const uint32_t A = m.A[f]; const uint64_t n_ctx = 1ull << m.cbits[f], n_fb = m.fbits[f] ? (1ull << m.fbits[f]) : 0;
		const uint32_t* h = hist.data() + m.base[f]; uint32_t* tb = tab.data() + m.base[f];
		std::vector<uint32_t> fbh(n_fb * A, 0); std::vector<uint16_t> fbf(n_fb * A, 0);
		std::vector<uint8_t> dense(n_ctx, 0);
		const unsigned T = st_threads(n_ctx, PAR);
		auto pooled = [&](auto&& keep) {      // fbh = sum over the contexts x with keep(x) of their counts, by fallback cell
			std::vector<std::vector<uint32_t>> part(T > 1 ? T : 0);
			st_parallel(n_ctx, PAR, [&](unsigned t, uint64_t lo, uint64_t hi) {
				std::vector<uint32_t>* acc = &fbh;
				if (T > 1) { part[t].assign(n_fb * A, 0); acc = &part[t]; }
				for (uint64_t x = lo; x < hi; ++x) if (keep(x)) for (uint32_t k = 0; k < A; ++k) (*acc)[(x & (n_fb - 1)) * A + k] += h[x * A + k];
			});
			for (auto& p : part) for (size_t i = 0; i < p.size(); ++i) fbh[i] += p[i];
		};
		if (n_fb) {      // the pooled table of every fallback cell over ALL its contexts: what a context would be coded with otherwise
			pooled([](uint64_t) { return true; });
			for (uint64_t x = 0; x < n_fb; ++x) st_normalise(&fbh[x * A], A, &fbf[x * A]);
			std::fill(fbh.begin(), fbh.end(), 0u);
		}
		const uint32_t* bq = st_bits_q8();
		st_parallel(n_ctx, PAR, [&](unsigned, uint64_t lo, uint64_t hi) {
			uint16_t fr[256];
			for (uint64_t x = lo; x < hi; ++x) {
				uint64_t t = 0; for (uint32_t k = 0; k < A; ++k) t += h[x * A + k];
				if (!t) { dense[x] = 2; continue; }                       // 2: never seen
				bool own = !n_fb;
				if (n_fb && t >= min_ctx) {
					st_normalise(&h[x * A], A, fr);
					const uint16_t* pf = &fbf[(x & (n_fb - 1)) * A];
					uint64_t c_own = 0, c_fb = 0;
					for (uint32_t k = 0; k < A; ++k) if (h[x * A + k]) { c_own += (uint64_t)h[x * A + k] * bq[fr[k]]; c_fb += (uint64_t)h[x * A + k] * bq[pf[k]]; }
					own = c_fb > c_own + (st_freqs_bytes(fr, A) + 2) * 8 * 256;
				}
				dense[x] = own ? 1 : 0;
			}
		});
		if (n_fb) pooled([&](uint64_t x) { return dense[x] == 0; });
		uint32_t nd = 0;
		for (uint64_t x = 0; x < n_ctx; ++x) { if (dense[x] == 2) dense[x] = 0; nd += dense[x]; }
		for (uint64_t x = 0; x < n_fb; ++x) { st_normalise(&fbh[x * A], A, &fbf[x * A]); st_put_freqs(hdr, &fbf[x * A], A); }
		st_put(hdr, nd);
		// the coder's entries of every context (independent), then the dense contexts' tables into the header (in order)
		st_parallel(n_ctx, PAR, [&](unsigned, uint64_t lo, uint64_t hi) {
			uint16_t fr[256];
			for (uint64_t x = lo; x < hi; ++x) {
				const uint16_t* src;
				if (dense[x]) { st_normalise(&h[x * A], A, fr); src = fr; }
				else if (n_fb) src = &fbf[(x & (n_fb - 1)) * A];
				else continue;
				uint32_t acc = 0; for (uint32_t k = 0; k < A; ++k) { tb[x * A + k] = src[k] | (acc << 16); acc += src[k]; }
			}
		});
		uint64_t prev = 0; uint16_t fr[256];
		for (uint64_t x = 0; x < n_ctx; ++x) if (dense[x]) {
			st_normalise(&h[x * A], A, fr);
			uint64_t gap = x - prev; prev = x;
			do { uint8_t by = (uint8_t)(gap & 127); gap >>= 7; if (gap) by |= 128; hdr.push_back(by); } while (gap);
			st_put_freqs(hdr, fr, A);
		}
	}
}

} // namespace clb
