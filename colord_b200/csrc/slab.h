// slab.h — bookkeeping of one device slab per GPU, so that a running job asks the driver for nothing.
// Why: cudaMalloc / cudaFree while a job runs cost 0.1-1.6 s of a 9 s step in driver time on a B200, in stalls that move from
// run to run (profiles/r01_summary.md r01e / r01f; profiles/r02_summary.md §8: the same two calls took 29 ms, 488 ms and 1 467 ms
// in three consecutive jobs).  clb_create of the first context on a device takes one block — the free memory minus a reserve
// (CLB_SLAB_RESERVE_GB, default 8) — and every block of DEV_BIG_BYTES or more is cut from it: first fit in address order, freed
// ranges merged with their neighbours; the block goes back to the driver when the last context on the device is destroyed.
// CLB_SLAB_GB=<n>: a slab of that size that stays until clb_release_cached_memory; CLB_SLAB_GB=0: no slab (every large block is
// a cudaMalloc).  Requests the slab cannot serve fall through to cudaMalloc.
// This header is pure host code (no CUDA calls): tests/host_slab_test.cpp exercises it on the CPU.
#pragma once
#include <cstdint>
#include <map>
#include <mutex>

namespace clb {

class Slab {
	uint64_t base_ = 0, size_ = 0, in_use_ = 0, peak_ = 0;
	std::map<uint64_t, uint64_t> free_;          // start -> length, address order, no two ranges adjacent
	std::map<uint64_t, uint64_t> used_;          // start -> length
	mutable std::mutex m_;
public:
	static constexpr uint64_t ALIGN = 512;
	bool active() const { std::lock_guard<std::mutex> g(m_); return size_ != 0; }
	uint64_t base() const { return base_; }
	uint64_t size() const { return size_; }
	uint64_t in_use() const { std::lock_guard<std::mutex> g(m_); return in_use_; }
	uint64_t peak() const { std::lock_guard<std::mutex> g(m_); return peak_; }
	uint64_t largest_free() const { std::lock_guard<std::mutex> g(m_); uint64_t b = 0; for (const auto& r : free_) if (r.second > b) b = r.second; return b; }
	// hands the address range [base, base + size) to the slab; base must be ALIGN-aligned
	void init(uint64_t base, uint64_t size)
	{
		std::lock_guard<std::mutex> g(m_);
		base_ = base; size_ = size & ~(ALIGN - 1); in_use_ = peak_ = 0;
		free_.clear(); used_.clear();
		if (size_) free_[base_] = size_;
	}
	// forgets the range (the caller frees it); false while blocks are still in use
	bool reset()
	{
		std::lock_guard<std::mutex> g(m_);
		if (!used_.empty()) return false;
		base_ = size_ = in_use_ = 0; free_.clear();
		return true;
	}
	// 0: no free range is large enough
	uint64_t alloc(uint64_t bytes)
	{
		std::lock_guard<std::mutex> g(m_);
		if (!size_) return 0;
		const uint64_t need = (bytes ? bytes + ALIGN - 1 : ALIGN) & ~(ALIGN - 1);
		for (auto it = free_.begin(); it != free_.end(); ++it) {
			if (it->second < need) continue;
			const uint64_t at = it->first, len = it->second;
			free_.erase(it);
			if (len > need) free_[at + need] = len - need;
			used_[at] = need;
			in_use_ += need; if (in_use_ > peak_) peak_ = in_use_;
			return at;
		}
		return 0;
	}
	bool owns(uint64_t addr) const { std::lock_guard<std::mutex> g(m_); return size_ && addr >= base_ && addr < base_ + size_; }
	// false: addr is not the start of a block of this slab
	bool free(uint64_t addr)
	{
		std::lock_guard<std::mutex> g(m_);
		auto u = used_.find(addr);
		if (u == used_.end()) return false;
		uint64_t at = u->first, len = u->second;
		used_.erase(u);
		in_use_ -= len;
		auto nxt = free_.lower_bound(at);
		if (nxt != free_.end() && at + len == nxt->first) { len += nxt->second; nxt = free_.erase(nxt); }
		if (nxt != free_.begin()) { auto prv = std::prev(nxt); if (prv->first + prv->second == at) { at = prv->first; len += prv->second; free_.erase(prv); } }
		free_[at] = len;
		return true;
	}
};

constexpr int SLAB_MAX_DEVICES = 16;
inline Slab& job_slab(int device) { static Slab s[SLAB_MAX_DEVICES]; return s[device >= 0 && device < SLAB_MAX_DEVICES ? device : 0]; }

} // namespace clb
