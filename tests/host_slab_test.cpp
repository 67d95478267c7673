// CPU exercise of colord_b200/csrc/slab.h (pure host bookkeeping of the device slab): random allocate / free sequences with
// the invariants checked after every operation.  Compiled and run by tests/test_abi_cpu.py.
#include "../colord_b200/csrc/slab.h"
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "slab test failed: %s (line %d)\n", #c, __LINE__); return 1; } } while (0)

int main()
{
	clb::Slab s;
	CHECK(!s.active() && s.alloc(100) == 0);
	const uint64_t base = 0x7f0000000000ull, size = (1ull << 30) + 300;      // the odd tail is cut to the alignment
	s.init(base, size);
	CHECK(s.active() && s.size() == (size & ~511ull) && s.largest_free() == s.size());
	std::mt19937_64 rng(7);
	std::vector<std::pair<uint64_t, uint64_t>> live;      // (address, requested bytes)
	uint64_t live_bytes = 0;
	for (int step = 0; step < 20000; ++step) {
		const bool do_alloc = live.empty() || (rng() % 100) < 55;
		if (do_alloc) {
			const uint64_t want = 1 + rng() % (rng() % 8 == 0 ? (256ull << 20) : (8ull << 20));
			const uint64_t at = s.alloc(want);
			if (!at) { CHECK(s.largest_free() < ((want + 511) & ~511ull)); continue; }      // refused only when no range fits
			CHECK(at % clb::Slab::ALIGN == 0 && at >= base && at + want <= base + s.size() && s.owns(at));
			for (const auto& b : live) CHECK(at + want <= b.first || b.first + b.second <= at);      // no overlap
			live.emplace_back(at, want); live_bytes += (want + 511) & ~511ull;
		} else {
			const size_t i = rng() % live.size();
			CHECK(s.free(live[i].first));
			CHECK(!s.free(live[i].first));                    // double free is reported, not applied
			live_bytes -= (live[i].second + 511) & ~511ull;
			live[i] = live.back(); live.pop_back();
		}
		CHECK(s.in_use() == live_bytes && s.peak() >= live_bytes);
	}
	CHECK(!s.reset());                                        // blocks in use: the slab stays
	for (const auto& b : live) CHECK(s.free(b.first));
	CHECK(s.in_use() == 0 && s.largest_free() == s.size());     // everything merged back into one range
	CHECK(s.alloc(s.size()) == base && s.alloc(1) == 0);        // the whole slab as one block; then nothing is left
	CHECK(s.free(base) && !s.owns(base + s.size()) && !s.free(base + 512));
	CHECK(s.reset() && !s.active());
	std::puts("slab ok");
	return 0;
}
