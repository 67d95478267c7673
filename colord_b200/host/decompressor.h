// decompressor.h — archive -> FASTQ / FASTA for archives written by compressor.h (host code).
// The reference decodes with the same serial model chain it encodes with (decompression_common.cpp:27-265 → CDNACoder::Decode
// dna_coder.cpp:237-437, CQualityCoder::Decode quality_coder.cpp:606-670, CIDCoder decompress id_coder.cpp:407-560, output
// through decompression.cpp).  The streams here are the device's native containers — static tables in the container header,
// 64 independent coder lanes per pack — so this is their decoder: the reference's event models in the decoding direction
// over those tables.  Layouts:
//   tables      per family: pooled fallback tables, then the contexts that own a table (LEB128 gaps); 12-bit frequencies
//   "DB01"      level, max candidates, n reads, n packs, n context reads | tables | per pack: n, 64 lane sizes, lane streams
//   "QB01"      bins, level, thresholds, n reads, n packs, context bits | mean model | fallback model | dense contexts | packs (rANS lanes)
//   "QO01"      source, level, n reads, n packs | tables | packs            "HB01"   n headers, n packs | tables | packs
// Lane l of a pack holds reads l, l + 64, ... of the pack.  Range decoder arithmetic = sub_rc.h:262-386 with totalFreq 2^12.
// Every reader checks its bounds: a damaged archive ends in DecodeError, never in an out-of-range access.
#pragma once
#include <memory>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <exception>
#include <iostream>
#include <mutex>
#include <random>
#include <thread>
#include <stdexcept>
#include <string>
#include <vector>
#include "archive_host.h"
#include "ref_genome_host.h"

namespace clbhost {

struct DecodeError : std::runtime_error { using std::runtime_error::runtime_error; };

namespace dec {

struct Bytes {            // bounds-checked little-endian reader over a stream
	const uint8_t* p; uint64_t n, at = 0;
	Bytes(const uint8_t* p_, uint64_t n_) : p(p_), n(n_) {}
	void need(uint64_t k) const { if (k > n - at) throw DecodeError("colord-b200: truncated stream"); }
	uint8_t u8() { need(1); return p[at++]; }
	uint16_t u16() { need(2); uint16_t v; std::memcpy(&v, p + at, 2); at += 2; return v; }
	uint32_t u32() { need(4); uint32_t v; std::memcpy(&v, p + at, 4); at += 4; return v; }
	uint64_t u64() { need(8); uint64_t v; std::memcpy(&v, p + at, 8); at += 8; return v; }
	uint64_t leb() { uint64_t v = 0; uint32_t sh = 0; uint8_t b; do { b = u8(); if (sh > 56) throw DecodeError("colord-b200: bad table index"); v |= static_cast<uint64_t>(b & 127) << sh; sh += 7; } while (b & 128); return v; }
	const uint8_t* take(uint64_t k) { need(k); const uint8_t* q = p + at; at += k; return q; }
	void magic(const char* m) { need(4); if (std::memcmp(p + at, m, 4)) throw DecodeError(std::string("colord-b200: not a ") + m + " stream"); at += 4; }
};

constexpr uint32_t M12 = 4096, LANES = 64;

// Static frequency tables of a container (writer: colord_b200/csrc/static_tables.h)
struct StaticModel {
	struct Family { uint32_t A, cbits, fbits; uint64_t base; };
	std::vector<Family> fam; std::vector<uint16_t> freq;
	void add(uint32_t A, uint32_t cbits, uint32_t fbits) { if (A < 1 || A > 256) throw DecodeError("colord-b200: bad alphabet size"); const uint64_t b = fam.empty() ? 0 : fam.back().base + (static_cast<uint64_t>(fam.back().A) << fam.back().cbits); fam.push_back(Family{A, cbits, fbits, b}); }
	static void get_freqs(Bytes& in, uint16_t* f, uint32_t A)
	{
		if (A > 256) throw DecodeError("colord-b200: bad alphabet size");      // callers hand in uint16_t[256]
		std::memset(f, 0, 2 * A);
		if (A <= 8) {
			const uint8_t mask = in.u8(); int last = -1; uint32_t sum = 0;
			for (uint32_t k = 0; k < A; ++k) if (mask >> k & 1) last = static_cast<int>(k);
			for (int k = 0; k < last; ++k) if (mask >> k & 1) { f[k] = in.u16(); sum += f[k]; }
			if (sum >= M12 && last >= 0) throw DecodeError("colord-b200: bad frequency table");
			if (last >= 0) f[last] = static_cast<uint16_t>(M12 - sum);
		} else {
			const uint16_t nz = in.u16(); uint32_t sum = 0;
			for (uint32_t i = 0; i < nz; ++i) { const uint32_t k = in.u8(); if (k >= A) throw DecodeError("colord-b200: bad frequency table"); f[k] = in.u16(); sum += f[k]; }
			if (nz && sum != M12) throw DecodeError("colord-b200: bad frequency table");
		}
	}
	void read_tables(Bytes& in)
	{
		const uint64_t total = fam.back().base + (static_cast<uint64_t>(fam.back().A) << fam.back().cbits);
		freq.assign(total + 1, 0);
		uint16_t fr[256];
		for (const Family& F : fam) {
			const uint64_t n_ctx = 1ull << F.cbits, n_fb = F.fbits ? (1ull << F.fbits) : 0;
			uint16_t* dst = freq.data() + F.base;
			if (n_fb) {
				std::vector<uint16_t> fb(n_fb * F.A);
				for (uint64_t x = 0; x < n_fb; ++x) get_freqs(in, fb.data() + x * F.A, F.A);
				for (uint64_t x = 0; x < n_ctx; ++x) std::memcpy(dst + x * F.A, fb.data() + (x & (n_fb - 1)) * F.A, 2 * F.A);
			}
			const uint32_t nd = in.u32();
			uint64_t x = 0;
			for (uint32_t d = 0; d < nd; ++d) {
				x += in.leb();
				if (x >= n_ctx) throw DecodeError("colord-b200: table index out of range");
				get_freqs(in, fr, F.A);
				std::memcpy(dst + x * F.A, fr, 2 * F.A);
			}
		}
	}
};

// sub_rc.h:262-386 (CRangeDecoder) over a lane's bytes; reads past the end deliver zeros (the encoder's flush covers real data)
class RangeDecoder {
	const uint8_t* p = nullptr; uint64_t n = 0, at = 0, low = 0, range = 0, buffer = 0;
	uint8_t byte() { return at < n ? p[at++] : 0; }
public:
	void start(const uint8_t* p_, uint64_t n_) { p = p_; n = n_; at = 0; buffer = 0; for (int i = 0; i < 8; ++i) buffer = (buffer << 8) + byte(); low = 0; range = 0xff00000000000000ULL; }
	uint32_t get(const StaticModel& m, uint32_t f, uint64_t ctx)
	{
		const StaticModel::Family& F = m.fam[f];
		const uint16_t* fr = m.freq.data() + F.base + (ctx & ((1ull << F.cbits) - 1)) * F.A;
		range >>= 12;
		const uint64_t cf = buffer / range;
		uint32_t s = 0; uint64_t acc = 0;
		while (s + 1 < F.A && acc + fr[s] <= cf) { acc += fr[s]; ++s; }
		if (!fr[s]) throw DecodeError("colord-b200: damaged stream (symbol with no code space)");
		const uint64_t r = acc * range;
		buffer -= r; low += r; range *= fr[s];
		while (range <= 0x0000ffffffffffffULL) {
			if ((low ^ (low + range)) & 0xff00000000000000ULL) { const uint64_t x = low; range = (x | 0x0000ffffffffffffULL) - x; }
			buffer = (buffer << 8) + byte();
			low <<= 8; range <<= 8;
		}
		return s;
	}
};

struct Pack { uint32_t n_reads; const uint8_t* lane[LANES]; uint32_t lane_bytes[LANES]; };
inline Pack read_pack(Bytes& in, uint32_t n_lanes = LANES)
{
	Pack p{}; p.n_reads = in.u32();
	for (uint32_t l = 0; l < n_lanes; ++l) p.lane_bytes[l] = in.u32();
	for (uint32_t l = 0; l < n_lanes; ++l) p.lane[l] = in.take(p.lane_bytes[l]);
	return p;
}

// all packs of a container with the index of their first read
struct PackAt { Pack pk; uint32_t r0; };
inline std::vector<PackAt> read_packs(Bytes& in, uint32_t n_packs, uint64_t n_units, const char* what, uint32_t n_lanes = LANES)
{
	std::vector<PackAt> v; v.reserve(n_packs);
	uint64_t r0 = 0;
	for (uint32_t p = 0; p < n_packs; ++p) {
		PackAt a{read_pack(in, n_lanes), static_cast<uint32_t>(r0)};
		if (a.pk.n_reads > n_units - r0) throw DecodeError(std::string("colord-b200: pack sizes of the ") + what + " stream exceed the archive's count");
		r0 += a.pk.n_reads;
		v.push_back(a);
	}
	if (r0 != n_units) throw DecodeError(std::string("colord-b200: the ") + what + " stream holds fewer records than the archive says");
	return v;
}
// The lanes of a pack (and the packs) are independent streams: fn(unit) for unit in [0, n) on CLB_DECODE_THREADS threads
// (default: the hardware's count).  The first exception is rethrown on the calling thread.
template <class Fn>
inline void parallel_for(uint64_t n, Fn&& fn)
{
	unsigned T = std::thread::hardware_concurrency();
	if (const char* e = std::getenv("CLB_DECODE_THREADS")) T = static_cast<unsigned>(std::atoi(e));
	T = static_cast<unsigned>(std::min<uint64_t>(std::max(1u, T), n));
	if (T <= 1) { for (uint64_t i = 0; i < n; ++i) fn(i); return; }
	std::atomic<uint64_t> next{0}; std::atomic<bool> failed{false};
	std::exception_ptr err; std::mutex m;
	auto work = [&] {
		for (;;) {
			const uint64_t i = next.fetch_add(1);
			if (i >= n || failed.load()) return;
			try { fn(i); } catch (...) { std::lock_guard<std::mutex> g(m); if (!err) err = std::current_exception(); failed.store(true); return; }
		}
	};
	std::vector<std::thread> th;
	for (unsigned t = 1; t < T; ++t) th.emplace_back(work);
	work();
	for (auto& t : th) t.join();
	if (err) std::rethrow_exception(err);
}

// ref_reads_accepter.h:27-57 — the decisions are part of the format: the decoder replays them to know which reads became references
inline std::vector<uint8_t> sampler_decisions(uint32_t range, double exponent, uint32_t n)
{
	std::vector<uint8_t> d(n);
	std::mt19937 mt; std::uniform_real_distribution<double> dist(0.0, 1.0);
	if (range == 0) range = 1;
	for (uint32_t i = 0; i < n; ++i) d[i] = dist(mt) <= std::pow(1.0 / (i / range + 1), exponent);
	return d;
}

struct Reads {
	std::vector<uint8_t> bases; std::vector<uint64_t> offsets{0};     // ASCII, back to back
	std::vector<uint8_t> flags;                                       // per base: 0 / 1 match / 2 anchor (quality contexts at level > 1)
};

// The tuples of one read (dna_coder.cpp:300-437) pulled from a symbol source:  d.get(family, context) / d.getx(family, context,
// excluded symbols); Src::exact — the reference's own streams (chunk indices of anchor / skip lengths uncapped, tuple types and
// substituted bases coded with exclusions, dna_coder.cpp:651-717, :889-922) against the native container's folded model.
enum DnaFamily { F_FLAG, F_LENBITS, F_LENDATA, F_SYM, F_SYMN, F_READID, F_REV, F_TUPLE, F_ANCHOR, F_SKIPL, F_SKIPD, F_SEEN, F_SHORT };
struct DnaRef { const uint8_t* sym; uint32_t len; bool rev; uint32_t at(int pos) const { if (pos < 0 || static_cast<uint32_t>(pos) >= len) return 255; return rev ? 3u - sym[len - 1 - pos] : sym[pos]; } };
inline uint32_t dna_n_bytes(uint64_t x) { uint32_t r = 1; for (x >>= 8; x; x >>= 8) ++r; return r; }
inline uint32_t dna_n_bits(uint64_t x) { uint32_t r = 0; for (; x; x >>= 1) ++r; return r; }
template <class Src>
inline uint32_t dna_read_id(Src& d, uint32_t r)
{
	const int nn = static_cast<int>(dna_n_bytes(r)); uint32_t id = 0;
	for (int i = nn - 1; i >= 0; --i) { const uint64_t add = i == nn - 2 ? id : 0; id = (id << 8) + d.get(F_READID, static_cast<uint64_t>(i) + (add << 3)); }
	return id;
}
template <class Src>
inline uint32_t dna_skip(Src& d, bool local)
{
	uint32_t v = 0;
	if (local) { for (uint32_t part = 0;; ++part) { const uint32_t x = d.get(F_SKIPL, Src::exact ? part : (part < 63 ? part : 63)); if (x < 255) { v += x; break; } v += 254; if (v > (1u << 30)) throw DecodeError("colord-b200: damaged DNA stream"); } }
	else for (int i = 3; i >= 0; --i) { const uint32_t x = d.get(F_SKIPD, static_cast<uint64_t>(i) * 64 + dna_n_bits(v)); v = (v << 8) + x; }
	return v;
}
// read length = number of tuples after the start tuple (dna_coder.cpp:1004-1056 in the decoding direction)
template <class Src>
inline uint32_t dna_read_len(Src& d)
{
	const uint32_t nbits = d.get(F_LENBITS, 0);
	if (nbits < 2) return nbits;
	uint64_t ctx = static_cast<uint64_t>(nbits) << 3;
	uint32_t v = d.get(F_LENDATA, ctx);
	if (nbits > 9) { uint32_t suffix = 0, sh = 0; ctx += 4; for (int nb = static_cast<int>(nbits) - 9; nb > 0; nb -= 8) { suffix |= d.get(F_LENDATA, ctx) << sh; sh += 8; ++ctx; } v = (v << (nbits - 9)) + suffix; }
	return v + (1u << (nbits - 1));
}
template <class Src>
inline void dna_edit_script(Src& d, uint32_t level, uint32_t n_s, uint32_t r, uint32_t n_tuples, const std::vector<std::vector<uint8_t>>& refs, std::vector<uint8_t>& rd, std::vector<uint8_t>& fl,
		uint64_t ctx_symbol, uint64_t ctx_tuple, uint64_t mask_s, uint64_t mask_t, uint32_t sh_t, uint64_t budget)
	{
		// budget: bases the archive still announces; checked before anything is appended, so a damaged script cannot grow memory past it
		auto room = [&](uint64_t k) { if (k > budget || rd.size() > budget - k) throw DecodeError("colord-b200: damaged DNA stream (more bases than the archive announces)"); };
		uint32_t seen_id[34], seen_rev[34], n_seen = 0; uint64_t ctx_rev = 0xf;
		uint32_t alt_ids[32]; bool alt_revs[32]; int alt_saved[32]; uint32_t n_alt = 0; int cur_alt = -1;
		auto get_rev = [&](uint32_t id) -> bool {
			for (uint32_t k = n_seen; k-- > 0;) if (seen_id[k] == id) return seen_rev[k] != 0;
			const uint32_t f = d.get(F_REV, ctx_rev);
			if (n_seen < 34) { seen_id[n_seen] = id; seen_rev[n_seen] = f; ++n_seen; }
			ctx_rev = ((ctx_rev << 2) + f) & 0xf;
			return f != 0;
		};
		auto ref_of = [&](uint32_t id, bool rev) -> DnaRef { if (id >= refs.size()) throw DecodeError("colord-b200: damaged DNA stream (unknown reference read)"); return DnaRef{refs[id].data(), static_cast<uint32_t>(refs[id].size()), rev}; };
		const uint32_t main_id = dna_read_id(d, r);
		const bool main_rev = get_rev(main_id);
		const DnaRef mainr = ref_of(main_id, main_rev); DnaRef altr = mainr;
		int ref_pos = 0, alt_pos = 0, delta = 0; bool is_main = true; uint32_t last_tuple = 255;
		for (uint32_t it = 0; it < n_tuples; ++it) {
			const DnaRef& o = is_main ? mainr : altr; int& pos = is_main ? ref_pos : alt_pos;
			const uint32_t rsym = o.at(pos);
			uint64_t ctx = ctx_tuple + ((ctx_symbol & 0xf) << sh_t) + (static_cast<uint64_t>(rsym) << (sh_t + 4));
			const uint32_t bucket = delta < -10 ? 1 : delta < -1 ? 2 : delta > 10 ? 3 : delta > 1 ? 4 : 0;
			ctx += static_cast<uint64_t>(bucket) << (sh_t + 6);
			const uint32_t excl = last_tuple == 2 ? 1u << 4 : last_tuple == 1 ? 1u << 5 : last_tuple == 4 ? (1u << 4) | (1u << 2)
				: last_tuple == 5 ? (1u << 1) | (1u << 5) : (last_tuple == 7 || last_tuple == 6) ? (1u << 6) | (1u << 7) : 0u;
			const uint32_t ty = d.getx(F_TUPLE, ctx, excl);
			ctx_tuple = ((ctx_tuple << 3) + ty) & mask_t;
			switch (ty) {
			case 6: {          // alternative reference read
				if (!is_main && cur_alt >= 0) alt_saved[cur_alt] = alt_pos;
				uint32_t id; int idx = -1;
				if (n_alt == 0 || !d.get(F_SEEN, n_alt)) id = dna_read_id(d, r);
				else { idx = static_cast<int>(d.get(F_SHORT, n_alt)); if (static_cast<uint32_t>(idx) >= n_alt) throw DecodeError("colord-b200: damaged DNA stream"); id = alt_ids[idx]; }
				const bool rev = get_rev(id);
				if (idx < 0) for (uint32_t k = 0; k < n_alt; ++k) if (alt_ids[k] == id) idx = static_cast<int>(k);
				if (idx < 0 && n_alt < 32) { idx = static_cast<int>(n_alt); alt_ids[n_alt] = id; alt_revs[n_alt] = rev; alt_saved[n_alt] = 0; ++n_alt; }
				cur_alt = idx;
				altr = ref_of(id, idx >= 0 ? alt_revs[idx] : rev);
				alt_pos = 0; is_main = false; delta = 0;
				break;
			}
			case 4: {          // anchor: a run of matches
				uint32_t alen = 0;
				for (uint32_t part = 0;; ++part) { const uint32_t v = d.get(F_ANCHOR, Src::exact ? part : (part < 63 ? part : 63)); if (v < 23) { alen += v; break; } alen += 22; if (alen > (1u << 30)) throw DecodeError("colord-b200: damaged DNA stream"); }
				if (alen > o.len) throw DecodeError("colord-b200: damaged DNA stream (anchor longer than its reference read)");
				room(alen);
				for (uint32_t k = 0; k < alen; ++k) { rd.push_back(static_cast<uint8_t>(o.at(pos + static_cast<int>(k)))); fl.push_back(2); }
				pos += static_cast<int>(alen);
				for (int i = static_cast<int>(n_s); i > 0; --i) ctx_symbol = (ctx_symbol << 2) + o.at(pos - i);
				ctx_symbol &= mask_s; delta = 0;
				break;
			}
			case 2: room(1); rd.push_back(static_cast<uint8_t>(rsym)); fl.push_back(1); ctx_symbol = ((ctx_symbol << 2) + rsym) & mask_s; ++pos; break;
			case 0: {          // insertion
				uint64_t c2 = 2; uint32_t sh = 2;
				if (level <= 1) { c2 += (ctx_symbol & 0xff) << sh; sh += 8; }
				else { c2 += (ctx_symbol & 0x3ff) << sh; sh += 10; if (level >= 3) { c2 += static_cast<uint64_t>(((ctx_symbol >> 10) & 3) == ((ctx_symbol >> 8) & 3)) << sh; ++sh; } }
				c2 += static_cast<uint64_t>(rsym) << sh; sh += 2;
				c2 += (ctx_tuple & 0777) << sh;
				const uint32_t s = d.get(F_SYM, c2);
				room(1); rd.push_back(static_cast<uint8_t>(s)); fl.push_back(0);
				ctx_symbol = ((ctx_symbol << 2) + s) & mask_s; ++delta;
				break;
			}
			case 1: ++pos; --delta; break;                        // deletion
			case 3: {          // substitution
				uint64_t c2 = 1; uint32_t sh = 2;
				c2 += (ctx_symbol & 0x3f) << sh; sh += 6;
				if (level >= 3) { c2 += static_cast<uint64_t>(((ctx_symbol >> 6) & 3) == ((ctx_symbol >> 4) & 3)) << sh; ++sh; }
				c2 += static_cast<uint64_t>(rsym) << sh; sh += 2;
				c2 += (ctx_tuple & 07777) << sh;
				const uint32_t s = d.getx(F_SYM, c2, 1u << (rsym & 3));
				room(1); rd.push_back(static_cast<uint8_t>(s)); fl.push_back(0);
				ctx_symbol = ((ctx_symbol << 2) + s) & mask_s; ++pos;
				break;
			}
			case 5: {          // skip (dna_coder.cpp:389-412)
				uint32_t skip;
				const bool distant_after_alt = !is_main && last_tuple == 6;
				const bool local = !distant_after_alt && last_tuple != 6 && last_tuple != 255;
				if (distant_after_alt) {
					uint32_t v = dna_skip(d, false);
					const int saved = cur_alt >= 0 ? alt_saved[cur_alt] : 0;
					if (v > 0) skip = v + static_cast<uint32_t>(saved);
					else { v = dna_skip(d, false); skip = static_cast<uint32_t>(saved - static_cast<int>(v)); }
				} else skip = dna_skip(d, local);
				delta -= static_cast<int>(skip); pos += static_cast<int>(skip);
				break;
			}
			default: is_main = true; if (cur_alt >= 0) alt_saved[cur_alt] = alt_pos; delta = 0; break;      // 7: back to the main reference
			}
			last_tuple = ty;
		}
		for (uint8_t s : rd) if (s > 3) throw DecodeError("colord-b200: damaged DNA stream (symbol outside a reference read)");
	}

// ---- DNA / edit-script stream: CDNACoder::Decode (dna_coder.cpp:237-437) over the container's tables ----
class DnaDecoder {
	StaticModel M; uint32_t level = 0, n_t = 0, n_s = 0;
	struct Src {            // a lane's range decoder over the container's static tables
		RangeDecoder& d; const StaticModel& M;
		static constexpr bool exact = false;
		uint32_t get(uint32_t f, uint64_t ctx) { return d.get(M, f, ctx); }
		uint32_t getx(uint32_t f, uint64_t ctx, uint32_t) { return d.get(M, f, ctx); }
	};
public:
	// decisions[r]: the sampler's answer for read r (all ones when every read is a reference); reads holding N never are
	// want_flags false: the per-base flags are not kept (they only feed the quality contexts at level > 1; a third of the memory)
	// max_bases: what the archive's info record announces; a damaged stream that decodes past it is refused instead of growing without bound
	// refs_io: the reference reads that precede the container's reads — the pseudo-reads of a reference genome, the reference reads of
	// the shards before this one in a multi-GPU archive.  The container counts them as its context reads: their ids come first, the
	// read ids start behind them.  The container's own reference reads are appended, so that the next shard finds them.
	// decisions: the sampler's answers for the container's reads (a slice of the archive's).
	Reads decode(const uint8_t* data, uint64_t size, uint32_t n_reads, const uint8_t* decisions, bool want_flags = true, uint64_t max_bases = ~0ull, uint32_t want_max_cand = 0,
		std::vector<std::vector<uint8_t>>* refs_io = nullptr)
	{
		std::vector<std::vector<uint8_t>> own_refs;
		std::vector<std::vector<uint8_t>>& refs = refs_io ? *refs_io : own_refs;                    // reference reads decoded so far (symbols 0..3)
		const uint32_t n_first = static_cast<uint32_t>(refs.size());
		Bytes in(data, size);
		in.magic("DB01");
		level = in.u32(); const uint32_t max_cand = in.u32(); const uint64_t nr = in.u64(); const uint32_t n_packs = in.u32(), n_ctx = in.u32();
		if (nr != n_reads || n_ctx != n_first || level < 1 || level > 3) throw DecodeError("colord-b200: DNA stream does not fit the archive");
		if (max_cand < 1 || max_cand > 32 || (want_max_cand && max_cand != want_max_cand)) throw DecodeError("colord-b200: DNA stream does not fit the archive (candidate limit)");
		n_t = level >= 3 ? 4 : level == 2 ? 3 : 2; n_s = level >= 3 ? 8 : level == 2 ? 7 : 5;
		{	// table widths: colord_b200/csrc/dna_model.h (history widths dna_coder.cpp:1253-1280)
			const uint32_t A[13] = {3, 32, 256, 4, 5, 256, 2, 8, 24, 256, 256, 2, max_cand < 2 ? 2 : max_cand};
			const uint32_t sym_bits = level >= 3 ? 24 : level == 2 ? 23 : 22;
			const uint32_t cb[13] = {8, 0, 9, sym_bits, 2 * n_s, 11, 4, 3 * n_t + 9, 6, 6, 8, 6, 6};
			const uint32_t fb[13] = {0, 0, 0, 10, 0, 0, 0, 3 * n_t + 6, 0, 0, 0, 0, 0};
			for (int f = 0; f < 13; ++f) M.add(A[f], cb[f], fb[f]);
		}
		M.read_tables(in);
		const uint64_t mask_s = (1ull << (2 * n_s)) - 1, mask_t = (1ull << (3 * n_t)) - 1; const uint32_t sh_t = 3 * n_t;
		Reads out; out.offsets.reserve(n_reads + 1);
		std::vector<uint8_t> rd, fl;
		uint32_t r0 = 0;
		for (uint32_t p = 0; p < n_packs; ++p) {
			const Pack pk = read_pack(in);
			if (pk.n_reads > n_reads - r0) throw DecodeError("colord-b200: DNA pack sizes exceed the read count");
			RangeDecoder lanes[LANES]; uint32_t fctx[LANES];
			for (uint32_t l = 0; l < LANES; ++l) { lanes[l].start(pk.lane[l], pk.lane_bytes[l]); fctx[l] = 0; }
			for (uint32_t r = r0; r < r0 + pk.n_reads; ++r) {
				RangeDecoder& d = lanes[(r - r0) % LANES]; uint32_t& fc = fctx[(r - r0) % LANES];
				rd.clear(); fl.clear();
				const uint32_t flag = d.get(M, F_FLAG, fc);
				fc = ((fc << 2) + flag) & 0xff;
				Src src{d, M};
				const uint32_t len = dna_read_len(src);
				if (len > max_bases - std::min<uint64_t>(max_bases, out.bases.size())) throw DecodeError("colord-b200: damaged DNA stream (more bases than the archive announces)");
				uint64_t ctx_symbol = mask_s, ctx_tuple = mask_t;
				if (flag == 0) for (uint32_t i = 0; i < len; ++i) { const uint32_t s = d.get(M, F_SYM, ctx_symbol << 2); rd.push_back(static_cast<uint8_t>(s)); ctx_symbol = ((ctx_symbol << 2) + s) & mask_s; }
				else if (flag == 1) for (uint32_t i = 0; i < len; ++i) { const uint32_t s = d.get(M, F_SYMN, ctx_symbol); rd.push_back(static_cast<uint8_t>(s)); ctx_symbol = ((ctx_symbol << 4) + s) & mask_s; }
				else dna_edit_script(src, level, n_s, n_first + r, len, refs, rd, fl, ctx_symbol, ctx_tuple, mask_s, mask_t, sh_t, max_bases - std::min<uint64_t>(max_bases, out.bases.size()));
				if (rd.size() > max_bases - std::min<uint64_t>(max_bases, out.bases.size())) throw DecodeError("colord-b200: damaged DNA stream (more bases than the archive announces)");
				fl.resize(rd.size(), 0);
				for (uint8_t s : rd) out.bases.push_back("ACGTN"[s > 4 ? 4 : s]);
				if (want_flags) out.flags.insert(out.flags.end(), fl.begin(), fl.end());
				out.offsets.push_back(out.bases.size());
				if (decisions[r] && flag != 1) refs.push_back(rd);
			}
			r0 += pk.n_reads;
		}
		if (r0 != n_reads) throw DecodeError("colord-b200: DNA stream holds fewer reads than the archive says");
		return out;
	}
};

// ---- quality stream, the "*-avg" modes: container "QB01" (interleaved rANS); reconstruction = decode_quad_average & co
// (quality_coder_impl.cpp:559-601): integer qualities from the per-read bin means by error diffusion ----
inline std::vector<uint8_t> decode_qual_avg(const uint8_t* data, uint64_t size, const Reads& reads)
{
	constexpr uint32_t PB = 12, L = 1u << 15;
	Bytes in(data, size);
	// "QB02": 4 streams per pack, 32 interleaved rANS states per stream (symbol k of a read -> state k mod 32); "QB01" (round 1): 64
	// single-state lanes per pack.  Same header, tables and symbol order.
	in.need(4);
	const bool v2 = std::memcmp(in.p + in.at, "QB02", 4) == 0;
	in.magic(v2 ? "QB02" : "QB01");
	const uint32_t n_lanes = v2 ? 4 : LANES, n_states = v2 ? 32 : 1;
	const uint32_t nb = in.u32(), level = in.u32(); uint32_t thr[4]; for (uint32_t& t : thr) t = in.u32();
	const uint64_t nr = in.u64(); const uint32_t n_packs = in.u32(), cbits = in.u32();
	const uint32_t n_reads = static_cast<uint32_t>(reads.offsets.size() - 1);
	if (nb != 2 && nb != 4 && nb != 5) throw DecodeError("colord-b200: bad quality stream");
	const uint32_t bps = nb == 2 ? 2 : 3, cb = bps * (nb == 2 ? 6 : 3), cmask = (1u << cb) - 1;
	if (nr != n_reads || cbits != cb + 8 + (level > 1 ? 2 : 0)) throw DecodeError("colord-b200: quality stream does not fit the archive");
	if (level > 1 && reads.flags.size() != reads.bases.size()) throw DecodeError("colord-b200: the quality stream needs the per-base flags of the DNA stream");
	std::vector<uint16_t> mf(nb * 128);
	for (uint16_t& v : mf) v = in.u16();
	const uint64_t n_ctx = 1ull << cbits; const uint32_t n_fb = 1u << cb;
	std::vector<uint16_t> freq(n_ctx * nb);
	{
		std::vector<uint16_t> fb(static_cast<size_t>(n_fb) * nb);
		for (uint32_t c = 0; c < n_fb; ++c) { uint32_t sm = 0; for (uint32_t k = 0; k + 1 < nb; ++k) { fb[c * nb + k] = in.u16(); sm += fb[c * nb + k]; } if (sm > M12) throw DecodeError("colord-b200: bad quality table"); fb[c * nb + nb - 1] = static_cast<uint16_t>(M12 - sm); }
		for (uint64_t c = 0; c < n_ctx; ++c) std::memcpy(&freq[c * nb], &fb[(c & (n_fb - 1)) * nb], 2 * nb);
		const uint32_t nd = in.u32(); uint64_t c = 0;
		for (uint32_t d = 0; d < nd; ++d) {
			c += in.leb();
			if (c >= n_ctx) throw DecodeError("colord-b200: bad quality table");
			uint32_t sm = 0; for (uint32_t k = 0; k + 1 < nb; ++k) { freq[c * nb + k] = in.u16(); sm += freq[c * nb + k]; }
			if (sm > M12) throw DecodeError("colord-b200: bad quality table");
			freq[c * nb + nb - 1] = static_cast<uint16_t>(M12 - sm);
		}
	}
	std::vector<uint8_t> out(reads.bases.size());
	auto code = [](uint8_t ch) -> uint32_t { return ch == 'C' ? 1 : ch == 'G' ? 2 : ch == 'T' ? 3 : 0; };
	const std::vector<PackAt> packs = read_packs(in, n_packs, n_reads, "quality", n_lanes);
	parallel_for(static_cast<uint64_t>(packs.size()) * n_lanes, [&](uint64_t unit) {
		const Pack& pk = packs[unit / n_lanes].pk; const uint32_t r0 = packs[unit / n_lanes].r0, l = static_cast<uint32_t>(unit % n_lanes);
		{
			if (l >= pk.n_reads) return;
			std::vector<uint8_t> sym;
			Bytes s(pk.lane[l], pk.lane_bytes[l]);
			uint32_t xs[32]; for (uint32_t k = 0; k < n_states; ++k) xs[k] = s.u32();
			uint32_t kk = 0;                                       // symbol index inside the read: picks the state
			auto pull = [&](const uint16_t* f, uint32_t A) -> uint32_t {
				uint32_t& x = xs[kk++ % n_states];
				const uint32_t slot = x & (M12 - 1); uint32_t dsym = 0, acc = 0;
				while (dsym + 1 < A && acc + f[dsym] <= slot) { acc += f[dsym]; ++dsym; }
				if (!f[dsym]) throw DecodeError("colord-b200: damaged quality stream");
				x = f[dsym] * (x >> PB) + slot - acc;
				while (x < L) x = (x << 16) | s.u16();
				return dsym;
			};
			for (uint32_t r = r0 + l; r < r0 + pk.n_reads; r += n_lanes) {
				const uint64_t o = reads.offsets[r]; const uint32_t n = static_cast<uint32_t>(reads.offsets[r + 1] - o);
				uint32_t avg16[5];
				kk = 0;
				for (uint32_t b = 0; b < nb; ++b) {
					const uint32_t a1 = pull(&mf[b * 128], 128);
					uint32_t& x = xs[kk++ % n_states];
					const uint32_t slot = x & (M12 - 1), a2 = slot / (M12 >> 8);
					x = (M12 >> 8) * (x >> PB) + slot - a2 * (M12 >> 8);
					while (x < L) x = (x << 16) | s.u16();
					avg16[b] = (a1 << 8) | a2;
				}
				sym.resize(n);
				uint32_t c = cmask, dna = n ? code(reads.bases[o]) : 0;
				for (uint32_t i = 0; i < n; ++i) {
					dna <<= 2; if (i + 1 < n) dna += code(reads.bases[o + i + 1]); dna &= 0xff;
					uint32_t cx = c + (dna << cb);
					if (level > 1) { cx += static_cast<uint32_t>(reads.flags[o + i] == 1) << (cb + 8); cx += static_cast<uint32_t>(reads.flags[o + i] == 2) << (cb + 9); }
					const uint32_t dsym = pull(&freq[static_cast<uint64_t>(cx) * nb], nb);
					sym[i] = static_cast<uint8_t>(dsym);
					c = ((c << bps) + dsym) & cmask;
				}
				double avg[5], avg_sum[5] = {0, 0, 0, 0, 0}, qual_sum[5] = {0, 0, 0, 0, 0};
				for (uint32_t b = 0; b < nb; ++b) avg[b] = static_cast<double>(avg16[b]) / 256.0;
				for (uint32_t i = 0; i < n; ++i) {
					const uint32_t b = sym[i];
					avg_sum[b] += avg[b];
					const uint32_t q = static_cast<uint32_t>(avg_sum[b] - qual_sum[b]);
					qual_sum[b] += q;
					out[o + i] = static_cast<uint8_t>(q + 33);
				}
			}
		}
	});
	return out;
}

// ---- lossless qualities: container "QO01"; context model of decode_original (quality_coder_impl.cpp:603-660) ----
inline void qorg_quantiser(uint32_t source, uint32_t level, uint8_t* q /*96*/)
{
	std::memset(q, 0, 96);
	auto fill = [&](int a, int b, int v) { for (int i = a; i < b; ++i) q[i] = static_cast<uint8_t>(v); };
	// bin edges: quality_coder.cpp:272-338 (ONT), :356-420 (PacBio CLR), :441-505 (PacBio HiFi: every code one higher, 93 -> 0)
	if (source == 0) {
		static const int e3[] = {0, 1, 2, 4, 7, 11, 16, 22, 29, 37, 46, 56, 67, 79, 90, 96}, e1[] = {0, 1, 2, 5, 10, 15, 20, 25, 35, 50, 70, 96};
		const int* e = level >= 3 ? e3 : e1; const int n = level >= 3 ? 15 : 11;
		for (int k = 0; k < n; ++k) fill(e[k], e[k + 1], k);
	} else {
		const int s = source == 2 ? 1 : 0;
		static const int e3[] = {1, 10, 20, 30, 39, 45, 51, 57, 63, 69, 75, 81, 87, 93}, e1[] = {1, 15, 29, 41, 53, 63, 72, 80, 87, 93};
		const int* e = level >= 3 ? e3 : e1; const int n = level >= 3 ? 13 : 9;
		q[0] = static_cast<uint8_t>(s);
		for (int k = 0; k < n; ++k) fill(e[k], e[k + 1], k + 1 + s);
		q[93] = static_cast<uint8_t>(s ? 0 : n + 1);
	}
}
inline std::vector<uint8_t> decode_qual_org(const uint8_t* data, uint64_t size, const Reads& reads)
{
	Bytes in(data, size);
	in.magic("QO01");
	const uint32_t source = in.u32(), level = in.u32(); const uint64_t nr = in.u64(); const uint32_t n_packs = in.u32();
	const uint32_t n_reads = static_cast<uint32_t>(reads.offsets.size() - 1);
	if (nr != n_reads || source > 2 || level < 1 || level > 3) throw DecodeError("colord-b200: quality stream does not fit the archive");
	if (level > 1 && reads.flags.size() != reads.bases.size()) throw DecodeError("colord-b200: the quality stream needs the per-base flags of the DNA stream");
	uint8_t quant[96]; qorg_quantiser(source, level, quant);
	StaticModel M; M.add(96, 8 + (level >= 3 ? 8 : 7) + (level > 1 ? 2 : 0), 8);
	M.read_tables(in);
	auto bsym = [](uint8_t ch) -> uint32_t { return ch == 'C' ? 1 : ch == 'G' ? 2 : ch == 'T' ? 3 : 0; };
	std::vector<uint8_t> out(reads.bases.size());
	const std::vector<PackAt> packs = read_packs(in, n_packs, n_reads, "quality");
	parallel_for(static_cast<uint64_t>(packs.size()) * LANES, [&](uint64_t unit) {
		const Pack& pk = packs[unit / LANES].pk; const uint32_t r0 = packs[unit / LANES].r0, l = static_cast<uint32_t>(unit % LANES);
		if (l >= pk.n_reads) return;
		RangeDecoder d; d.start(pk.lane[l], pk.lane_bytes[l]);
		for (uint32_t r = r0 + l; r < r0 + pk.n_reads; r += LANES) {
			const uint64_t o = reads.offsets[r]; const uint32_t n = static_cast<uint32_t>(reads.offsets[r + 1] - o);
			const uint8_t* b = reads.bases.data() + o; const uint8_t* fl = level > 1 ? reads.flags.data() + o : nullptr;
			uint32_t pc = 0xff;
			for (uint32_t i = 0; i < n; ++i) {
				uint32_t c = pc, sh = 8;
				c += bsym(b[i]) << sh; sh += 2;
				if (i > 0) c += bsym(b[i - 1]) << sh;
				sh += 2;
				if (level >= 3) { if (i > 1) c += bsym(b[i - 2]) << sh; sh += 2; }
				else { if (i > 1) c += static_cast<uint32_t>(bsym(b[i - 2]) == bsym(b[i - 1])) << sh; sh += 1; }
				if (i + 1 < n) c += bsym(b[i + 1]) << sh;
				sh += 2;
				if (level > 1) { c += static_cast<uint32_t>(fl[i] == 1) << sh; ++sh; c += static_cast<uint32_t>(fl[i] == 2) << sh; }
				const uint32_t s = d.get(M, 0, c);
				out[o + i] = static_cast<uint8_t>(s + 33);
				pc = ((pc << 4) + quant[s]) & 0xff;
			}
		}
	});
	return out;
}

// ---- headers: container "HB01"; event model of CIDCoder::decompress_lossless (id_coder.cpp:407-560) ----
struct Headers { std::vector<uint8_t> bytes; std::vector<uint64_t> offsets{0}; std::vector<uint8_t> plus_id; };
inline Headers decode_headers(const uint8_t* data, uint64_t size, uint64_t n_expected)
{
	enum { H_PLUS, H_FLAG, H_SAME, H_SAMELEN, H_LITEQ, H_LITNEW, H_PLAIN };
	Bytes in(data, size);
	in.magic("HB01");
	const uint64_t n = in.u64(); const uint32_t n_packs = in.u32();
	if (n != n_expected) throw DecodeError("colord-b200: header stream does not fit the archive");
	StaticModel M;
	{ const uint32_t A[7] = {2, 2, 2, 2, 256, 256, 256}, cb[7] = {0, 8, 6, 6, 14, 10, 8}, fb[7] = {0, 0, 0, 0, 10, 0, 0}; for (int f = 0; f < 7; ++f) M.add(A[f], cb[f], fb[f]); }
	M.read_tables(in);
	auto is_lit = [](uint8_t c) { return (c >= '0' && c <= '9') || (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z') || c == '@'; };
	auto mn = [](uint32_t a, uint32_t b) { return a < b ? a : b; };
	Headers H; uint64_t r = 0;
	std::vector<uint8_t> cur, prv;
	for (uint32_t p = 0; p < n_packs; ++p) {
		const Pack pk = read_pack(in);
		if (pk.n_reads > n - r) throw DecodeError("colord-b200: header pack sizes exceed the header count");
		RangeDecoder lanes[LANES];
		for (uint32_t l = 0; l < LANES; ++l) lanes[l].start(pk.lane[l], pk.lane_bytes[l]);
		uint32_t flag_hist = 0;
		prv.clear();
		for (uint32_t k = 0; k < pk.n_reads; ++k, ++r) {
			RangeDecoder& d = lanes[k % LANES];
			cur.clear();
			H.plus_id.push_back(static_cast<uint8_t>(d.get(M, H_PLUS, 0)));
			const uint32_t flag = d.get(M, H_FLAG, flag_hist & 0xff);
			flag_hist = (flag_hist << 1) + flag;
			if (k == 0 && flag) throw DecodeError("colord-b200: damaged header stream");
			auto guard = [&]() { if (cur.size() > (1u << 24)) throw DecodeError("colord-b200: damaged header stream"); };
			if (!flag) { for (uint32_t j = 0;; ++j) { const uint32_t c = d.get(M, H_PLAIN, mn(j, 255)); if (!c) break; cur.push_back(static_cast<uint8_t>(c)); guard(); } }
			else {
				const uint32_t lpv = static_cast<uint32_t>(prv.size()); uint32_t j = 0;
				for (uint32_t t = 0;; ++t) {
					uint32_t je = j; while (je < lpv && is_lit(prv[je])) ++je;
					const uint32_t lp = je - j, t5 = mn(t, 31);
					if (d.get(M, H_SAME, mn(t, 63))) cur.insert(cur.end(), prv.begin() + j, prv.begin() + je);
					else if (d.get(M, H_SAMELEN, mn(t, 63))) for (uint32_t k2 = 0; k2 < lp; ++k2) { const uint32_t pc = prv[j + k2]; const uint32_t c = d.get(M, H_LITEQ, ((pc & 15u) << 10) | (t5 << 5) | mn(k2, 31)); cur.push_back(static_cast<uint8_t>(c ? c : pc)); }
					else for (uint32_t k2 = 0;; ++k2) { const uint32_t c = d.get(M, H_LITNEW, (t5 << 5) | mn(k2, 31)); if (!c) break; cur.push_back(static_cast<uint8_t>(c)); guard(); }
					if (je == lpv) break;
					cur.push_back(prv[je]);                  // same shape: the separator is the previous header's
					j = je + 1;
				}
			}
			H.bytes.insert(H.bytes.end(), cur.begin(), cur.end());
			H.offsets.push_back(H.bytes.size());
			prv.swap(cur);
		}
	}
	if (r != n) throw DecodeError("colord-b200: header stream holds fewer headers than the archive says");
	return H;
}

} // namespace dec
} // namespace clbhost
#include "compat_decoder.h"      // the reference's own streams (version-1 archives)
namespace clbhost {

// decompression_common.cpp:27-265: the archive's records and streams decoded into memory (reads in input order)
struct DecompressedArchive {
	CInfo info; CMeta meta;
	dec::Reads reads; dec::Headers headers; std::vector<uint8_t> quals;      // quals: same layout as reads.bases, empty for FASTA

	// ref_genome_path: the genome of a -G archive that does not carry it (decompression_common.cpp:262-283)
	explicit DecompressedArchive(const std::string& archive_path, bool verbose = false, const std::string& ref_genome_path = std::string())
	{
		CArchive archive(true);
		if (!archive.Open(archive_path)) throw std::runtime_error("Error: cannot open archive: " + archive_path);
		std::vector<uint8_t> raw; size_t md = 0;
		const int s_info = archive.GetStreamId("info"), s_meta = archive.GetStreamId("meta");
		if (s_info < 0 || s_meta < 0 || !archive.ReadPart(s_info, 0, raw, md)) throw DecodeError("Error: not a colord archive (no info / meta record)");
		info.Deserialize(raw);
		// two kinds of archives: the reference's own streams (version 1.x: written by `colord` or by colord-b200's compat mode) and the
		// device's native containers (version B200_VERSION_MAJOR)
		const bool compat = info.version_major == 1;
		const int s_dna = archive.GetStreamId(compat ? "dna" : "dna-b200"), s_qual = archive.GetStreamId(compat ? "qual" : "qual-b200"), s_hdr = archive.GetStreamId(compat ? "header" : "header-b200");
		if ((!compat && info.version_major != B200_VERSION_MAJOR) || s_dna < 0 || s_hdr < 0)
			throw DecodeError("Error: incompatibile archive version");
		if (!archive.ReadPart(s_meta, 0, raw, md)) throw DecodeError("Error: cannot read the meta record");
		meta.Deserialize(raw, s_qual >= 0);
		const uint32_t n_reads = info.total_reads;
		std::vector<std::vector<uint8_t>> pseudo;                   // the genome's pseudo-reads: the first reference reads (symbols 0..3)
		if (meta.ref_genome_available) pseudo = pseudo_reads(archive, compat, ref_genome_path);
		if (verbose) std::cerr << "reads: " << n_reads << "\nbases: " << info.total_bases << "\nquality mode: " << static_cast<int>(meta.qualityComprMode) << "\n";

		if (compat) { decode_reference_streams(archive, s_dna, s_qual, s_hdr, n_reads, pseudo.empty() ? nullptr : &pseudo); return; }
		// native containers: one part per stream and shard (a single-GPU archive is one shard; colord-b200 --gpus N writes N, whose
		// DNA containers name the reference reads of the shards before them as their context reads)
		const std::vector<uint8_t> decisions = meta.referenceReadsMode == ReferenceReadsMode::Sparse ? dec::sampler_decisions(meta.sparseMode_range, meta.sparseMode_exponent, n_reads) : std::vector<uint8_t>(n_reads, 1);
		const bool flags_needed = meta.is_fastq && meta.compressionLevel > 1 && meta.qualityComprMode != QualityComprMode::None;
		const size_t n_shards = archive.Parts(s_dna).size();
		if (!n_shards || archive.Parts(s_hdr).size() != n_shards || (meta.is_fastq && archive.Parts(s_qual).size() != n_shards)) throw DecodeError("Error: the streams of the archive do not have the same shards");
		std::vector<std::vector<uint8_t>> refs = pseudo;
		std::vector<uint8_t> stream;
		uint32_t r0 = 0;
		for (size_t sh = 0; sh < n_shards; ++sh) {
			if (!archive.ReadPart(s_dna, sh, stream, md) || md > n_reads - r0) throw DecodeError("Error: cannot read the DNA stream");
			const uint32_t n_sh = static_cast<uint32_t>(md);
			dec::DnaDecoder dna;
			dec::Reads part = dna.decode(stream.data(), stream.size(), n_sh, decisions.data() + r0, flags_needed, info.total_bases - std::min<uint64_t>(info.total_bases, reads.bases.size()), meta.maxCandidates, &refs);
			if (meta.is_fastq) {
				if (!archive.ReadPart(s_qual, sh, stream, md)) throw DecodeError("Error: cannot read the quality stream");
				std::vector<uint8_t> q;
				switch (meta.qualityComprMode) {
				case QualityComprMode::None: q.assign(part.bases.size(), static_cast<uint8_t>(33 + meta.qualityRevThresholds.at(0))); break;      // quality_coder.cpp:611-617
				case QualityComprMode::Original: q = dec::decode_qual_org(stream.data(), stream.size(), part); break;
				case QualityComprMode::BinaryAverage: case QualityComprMode::QuadAverage: case QualityComprMode::QuinaryAverage: q = dec::decode_qual_avg(stream.data(), stream.size(), part); break;
				default: throw DecodeError("Error: quality mode of the archive is not available in this build");
				}
				quals.insert(quals.end(), q.begin(), q.end());
			}
			if (meta.headerComprMode == HeaderComprMode::Original) {
				if (!archive.ReadPart(s_hdr, sh, stream, md)) throw DecodeError("Error: cannot read the header stream");
				const dec::Headers h = dec::decode_headers(stream.data(), stream.size(), n_sh);
				const uint64_t h0 = headers.bytes.size();
				headers.bytes.insert(headers.bytes.end(), h.bytes.begin(), h.bytes.end());
				for (size_t i = 1; i < h.offsets.size(); ++i) headers.offsets.push_back(h0 + h.offsets[i]);
				headers.plus_id.insert(headers.plus_id.end(), h.plus_id.begin(), h.plus_id.end());
			} else for (uint32_t r = 0; r < n_sh; ++r) {	// no header bytes were stored: the reference prints the id "@" for `none` (id_coder.cpp:393-396) and an empty id for `main` (:588-591)
				if (meta.headerComprMode == HeaderComprMode::None) headers.bytes.push_back('@');
				headers.offsets.push_back(headers.bytes.size());
				headers.plus_id.push_back(0);
			}
			const uint64_t b0 = reads.bases.size();
			reads.bases.insert(reads.bases.end(), part.bases.begin(), part.bases.end());
			for (size_t i = 1; i < part.offsets.size(); ++i) reads.offsets.push_back(b0 + part.offsets[i]);
			r0 += n_sh;
		}
		if (r0 != n_reads) throw DecodeError("Error: the shards of the archive hold fewer reads than it announces");
		if (reads.bases.size() != info.total_bases) throw DecodeError("Error: the decoded reads do not add up to the archive's base count");
	}
	uint32_t n_reads() const { return static_cast<uint32_t>(reads.offsets.size() - 1); }
private:
	std::vector<std::vector<uint8_t>> pseudo_reads(CArchive& archive, bool compat, const std::string& ref_genome_path)
	{
		std::unique_ptr<CReferenceGenome> genome;
		if (meta.storeRefGenome) {
			const int s_gen = archive.GetStreamId(compat ? "ref-genome" : "ref-genome-b200");
			if (s_gen < 0) throw DecodeError("Error: the archive announces a stored reference genome but has none");
			std::vector<std::vector<uint8_t>> seqs; std::vector<uint8_t> part; size_t md = 0;
			if (compat) {      // CReferenceGenome(archive): n_seqs plain reads from a DNA coder of their own, "level 9" (reference_genome.cpp:235-279)
				if (!archive.ReadPart(s_gen, 0, part, md)) throw DecodeError("Error: cannot read the reference genome of the archive");
				size_t served = 0;
				const xdec::PartSource one = [&](std::vector<uint8_t>& data, size_t& m) { if (served++) return false; data = part; m = md; return true; };
				std::vector<uint32_t> part_reads;
				if (md >= (1ull << 32)) throw DecodeError("Error: damaged reference-genome stream");
				const dec::Reads g = xdec::decode_dna(one, static_cast<uint32_t>(md), std::vector<uint8_t>(md, 0), 9, 2, false, ~0ull, part_reads);
				for (size_t q = 0; q + 1 < g.offsets.size(); ++q) seqs.emplace_back(g.bases.begin() + g.offsets[q], g.bases.begin() + g.offsets[q + 1]);
			} else {
				const size_t n_parts = archive.Parts(s_gen).size();
				for (size_t q = 0; q < n_parts; ++q) { if (!archive.ReadPart(s_gen, q, part, md)) throw DecodeError("Error: cannot read the reference genome of the archive"); try { seqs.push_back(CReferenceGenome::Unpack(part)); } catch (const std::runtime_error& e) { throw DecodeError(e.what()); } }
			}
			genome = std::make_unique<CReferenceGenome>(std::move(seqs));
		} else {
			if (ref_genome_path.empty()) throw DecodeError("Error: compressed file was created without -s switch, reference genome is required for decompression");
			genome = std::make_unique<CReferenceGenome>(ref_genome_path);
			if (genome->GetChecksum() != meta.ref_genome_checksum) throw DecodeError("Error: different reference genome was used during compression. Decompression impossible.");
		}
		genome->SetReadLen(meta.ref_genome_read_len, meta.ref_genome_overlap_size);
		if (!genome->valid_read_len() || genome->GetNPseudoReads() != meta.n_ref_genome_pseudo_reads) throw DecodeError("Error: the reference genome does not fit the archive");
		std::vector<uint8_t> pb; std::vector<uint64_t> po;
		genome->PseudoReads(pb, po);
		std::vector<std::vector<uint8_t>> out(po.size() - 1);
		for (size_t q = 0; q + 1 < po.size(); ++q) { out[q].resize(po[q + 1] - po[q]); for (uint64_t i = po[q]; i < po[q + 1]; ++i) { const uint8_t ch = pb[i]; out[q][i - po[q]] = ch == 'C' ? 1 : ch == 'G' ? 2 : ch == 'T' ? 3 : 0; } }
		return out;
	}
	// version-1 archives: one part per pack in every stream (entr_read.h:146-184, entr_qual.h:128-190, entr_header.cpp:49-80)
	void decode_reference_streams(CArchive& archive, int s_dna, int s_qual, int s_hdr, uint32_t n_reads, const std::vector<std::vector<uint8_t>>* first_refs)
	{
		auto parts_of = [&](int stream_id) -> xdec::PartSource {
			auto next = std::make_shared<size_t>(0);
			const size_t n_parts = archive.Parts(stream_id).size();
			return [&archive, stream_id, next, n_parts](std::vector<uint8_t>& data, size_t& md) { if (*next >= n_parts) return false; if (!archive.ReadPart(stream_id, (*next)++, data, md)) throw DecodeError("Error: cannot read a part of the archive"); return true; };
		};
		const std::vector<uint8_t> decisions = meta.referenceReadsMode == ReferenceReadsMode::Sparse ? dec::sampler_decisions(meta.sparseMode_range, meta.sparseMode_exponent, n_reads) : std::vector<uint8_t>(n_reads, 1);
		const bool flags_needed = meta.is_fastq && meta.compressionLevel > 1 && meta.qualityComprMode != QualityComprMode::None;
		std::vector<uint32_t> part_reads;
		reads = xdec::decode_dna(parts_of(s_dna), n_reads, decisions, static_cast<uint32_t>(meta.compressionLevel), meta.maxCandidates, flags_needed, info.total_bases, part_reads, first_refs);
		if (reads.bases.size() != info.total_bases) throw DecodeError("Error: the decoded reads do not add up to the archive's base count");
		if (meta.headerComprMode == HeaderComprMode::Original) headers = xdec::decode_headers(parts_of(s_hdr), n_reads, info.total_bytes);
		else for (uint32_t r = 0; r < n_reads; ++r) {
			if (meta.headerComprMode == HeaderComprMode::None) headers.bytes.push_back('@');
			headers.offsets.push_back(headers.bytes.size());
			headers.plus_id.push_back(0);
		}
		if (meta.is_fastq)
			quals = xdec::decode_qual(parts_of(s_qual), reads, part_reads, static_cast<uint32_t>(meta.qualityComprMode), static_cast<uint32_t>(meta.dataSource), static_cast<uint32_t>(meta.compressionLevel), meta.qualityRevThresholds);
	}
};

// decompression.cpp: archive -> file.  FASTQ records: '@' id, bases, '+' [id], qualities; FASTA records: '>' id, bases on one line.
inline void runDecompression(const std::string& archive_path, const std::string& output_path, bool verbose = false, const std::string& ref_genome_path = std::string())
{
	const DecompressedArchive A(archive_path, verbose, ref_genome_path);
	const CMeta& meta = A.meta; const dec::Reads& reads = A.reads; const dec::Headers& H = A.headers; const std::vector<uint8_t>& quals = A.quals;
	const uint32_t n_reads = A.n_reads();
	FILE* out = std::fopen(output_path.c_str(), "wb");
	if (!out) throw std::runtime_error("Error: cannot open output file: " + output_path);
	std::vector<uint8_t> buf; buf.reserve(1u << 24);
	bool ok = true;
	for (uint32_t r = 0; r < n_reads; ++r) {
		const uint8_t* h = H.bytes.data() + H.offsets[r]; const size_t hn = H.offsets[r + 1] - H.offsets[r];
		const uint8_t* b = reads.bases.data() + reads.offsets[r]; const size_t bn = reads.offsets[r + 1] - reads.offsets[r];
		buf.push_back(meta.is_fastq ? '@' : '>'); buf.insert(buf.end(), h, h + hn); buf.push_back('\n');
		buf.insert(buf.end(), b, b + bn); buf.push_back('\n');
		if (meta.is_fastq) {
			buf.push_back('+'); if (H.plus_id[r]) buf.insert(buf.end(), h, h + hn); buf.push_back('\n');
			buf.insert(buf.end(), quals.begin() + reads.offsets[r], quals.begin() + reads.offsets[r + 1]); buf.push_back('\n');
		}
		if (buf.size() >= (1u << 24) - (1u << 20) || r + 1 == n_reads) { ok = ok && std::fwrite(buf.data(), 1, buf.size(), out) == buf.size(); buf.clear(); }
	}
	ok = (std::fclose(out) == 0) && ok;
	if (!ok) throw std::runtime_error("Error: cannot write output file: " + output_path);
}

} // namespace clbhost
