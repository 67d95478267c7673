// compat_decoder.h — host-side decoders of the REFERENCE's own stage-3 streams ("dna", "qual", "header" of a version-1 archive):
// what `colord decompress` does for its archives, so that colord-b200 reads archives written by the unmodified reference and the
// ones its own compat mode writes (colord_b200/csrc/stage3_exact.cu).  SURVEY.md §8f row 2; pure host code.
//   range decoder     src/colord/sub_rc.h:214-386 (CRangeDecoder, 64-bit, at most 8 bytes pulled per symbol)
//   adaptive models   src/colord/rc.h:34-221 (counts, +ADDER, halving at MAX_TOTAL), Decode / DecodeExcluding rc.h:827-857, :928-965;
//                     parameters dna_coder.h:48-60, quality_coder.h:35-41, id_coder.h:50-59
//   DNA stream        src/colord/dna_coder.cpp:237-437 (the tuple walk is shared with the native container: decompressor.h dna_edit_script)
//   quality stream    src/colord/quality_coder_impl.cpp:452-834, quality_coder.cpp:604-660; drivers entr_qual.h:128-190
//   header stream     src/colord/id_coder.cpp:407-560; driver entr_header.cpp:49-80
// Contexts are only identifiers of a model, so they are formed as the reference's ENCODER forms them (the reference's decoder uses
// other numbers for the same classes in two places: the match / anchor flag bits of the quality contexts, quality_coder_impl.cpp:237-240
// against :599-601).  Damaged streams end in DecodeError, never in an out-of-range access.
#pragma once
#include <cstdint>
#include <cstring>
#include <functional>
#include <vector>
#include "decompressor.h"

namespace clbhost {
namespace xdec {

struct Fam { uint32_t n_sym, max_total, adder; };

// adaptive models by (family, context): open addressing, counters in one arena ([n_sym] = total)
class Models {
	struct Slot { uint64_t ctx; uint32_t fam, used; uint64_t at; };
	std::vector<Fam> fam; std::vector<Slot> tab; uint64_t n = 0; std::vector<uint32_t> arena;
	static uint64_t hash(uint32_t f, uint64_t c) { uint64_t h = c * 0x9E3779B97F4A7C15ULL + f * 0xC2B2AE3D27D4EB4FULL; h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ULL; return h ^ (h >> 32); }
public:
	explicit Models(std::vector<Fam> f) : fam(std::move(f)), tab(1u << 12) { for (const Fam& x : fam) if (x.n_sym < 1 || x.n_sym > 256) throw DecodeError("colord-b200: bad alphabet size"); }
	const Fam& family(uint32_t f) const { return fam[f]; }
	uint32_t* get(uint32_t f, uint64_t ctx)
	{
		if (n * 2 >= tab.size()) {
			std::vector<Slot> old(tab.size() * 2); old.swap(tab);
			for (const Slot& s : old) if (s.used) { uint64_t h = hash(s.fam, s.ctx) & (tab.size() - 1); while (tab[h].used) h = (h + 1) & (tab.size() - 1); tab[h] = s; }
		}
		uint64_t h = hash(f, ctx) & (tab.size() - 1);
		while (tab[h].used) { if (tab[h].fam == f && tab[h].ctx == ctx) return arena.data() + tab[h].at; h = (h + 1) & (tab.size() - 1); }
		const uint32_t A = fam[f].n_sym;
		tab[h] = Slot{ctx, f, 1, arena.size()}; ++n;
		arena.resize(arena.size() + A + 1, 1);                    // Init(nullptr): every count 1
		arena.back() = A;
		return arena.data() + tab[h].at;
	}
};

// one stream's decoder: range decoder over the current part + the models that live on from part to part
class Coder {
	Models& M; const uint8_t* p = nullptr; uint64_t n = 0, at = 0, low = 0, range = 0, buffer = 0;
	uint8_t byte() { return at < n ? p[at++] : 0; }
public:
	static constexpr bool exact = true;
	explicit Coder(Models& m) : M(m) {}
	void start(const uint8_t* p_, uint64_t n_) { p = p_; n = n_; at = 0; buffer = 0; for (int i = 0; i < 8; ++i) buffer = (buffer << 8) + byte(); low = 0; range = 0xff00000000000000ULL; }
	uint32_t getx(uint32_t f, uint64_t ctx, uint32_t excl)
	{
		const Fam& F = M.family(f); const uint32_t A = F.n_sym;
		uint32_t* c = M.get(f, ctx);
		uint32_t tot = c[A];
		for (uint32_t i = 0; i < A && i < 32; ++i) if (excl >> i & 1) tot -= c[i];
		if (!tot) throw DecodeError("colord-b200: damaged stream");
		range /= tot;
		const uint64_t cf = buffer / range;
		uint32_t s = 0; uint64_t acc = 0;
		for (;; ++s) {
			if (s >= A) throw DecodeError("colord-b200: damaged stream (symbol outside its model)");
			if (s < 32 && (excl >> s & 1)) continue;
			if (acc + c[s] > cf) break;
			acc += c[s];
		}
		const uint64_t r = acc * range;
		buffer -= r; low += r; range *= c[s];
		for (int k = 0; k < 8 && range <= 0x00ffffffffffffULL; ++k) {
			if ((low ^ (low + range)) & 0xff00000000000000ULL) { const uint64_t x = low; range = (x | 0x00ffffffffffffULL) - x; }
			buffer = (buffer << 8) + byte();
			low <<= 8; range <<= 8;
		}
		if (!range) throw DecodeError("colord-b200: damaged stream");
		c[s] += F.adder; c[A] += F.adder;
		while (c[A] >= F.max_total) { uint32_t t = 0; for (uint32_t i = 0; i < A; ++i) { c[i] = (c[i] + 1) / 2; t += c[i]; } c[A] = t; }
		return s;
	}
	uint32_t get(uint32_t f, uint64_t ctx) { return getx(f, ctx, 0); }
};

// next part of a stream: bytes + metadata; false at the end of the stream
using PartSource = std::function<bool(std::vector<uint8_t>&, size_t&)>;

// ---- "dna": CDNACoder::Decode.  part_reads receives the reads of every part (the quality parts follow them). ----
inline dec::Reads decode_dna(const PartSource& next, uint32_t n_reads, const std::vector<uint8_t>& decisions, uint32_t level, uint32_t max_cand, bool want_flags, uint64_t max_bases,
	std::vector<uint32_t>& part_reads, const std::vector<std::vector<uint8_t>>* first_refs = nullptr)
{
	using namespace dec;
	if (max_cand < 1 || max_cand > 256) throw DecodeError("colord-b200: the DNA stream's parameters are out of range");
	// dna_coder.cpp:1253-1280: levels 1, 2, 3; anything else (the stored reference genome's "level 9") keeps one tuple / one symbol of history
	const uint32_t n_t = level == 3 ? 4 : level == 2 ? 3 : level == 1 ? 2 : 1, n_s = level == 3 ? 8 : level == 2 ? 7 : level == 1 ? 5 : 1;
	// first_refs: the pseudo-reads of a reference genome — reference reads in front of the archive's reads; the coder's read ids start
	// behind them (start_read_id, compression.cpp:641)
	const uint32_t n_first = first_refs ? static_cast<uint32_t>(first_refs->size()) : 0;
	Models M({{3, 1u << 15, 1}, {32, 1u << 18, 8}, {256, 1u << 18, 8}, {4, 1u << 10, 1}, {5, 1u << 10, 1}, {256, 1u << 13, 1}, {2, 1u << 15, 1},
		{8, 1u << 15, 1}, {24, 1u << 15, 1}, {256, 1u << 15, 1}, {256, 1u << 15, 1}, {2, 1u << 15, 1}, {max_cand, 1u << 13, 1}});      // dna_coder.h:48-60, dna_coder.cpp:1316-1336
	Coder d(M);
	const uint64_t mask_s = (1ull << (2 * n_s)) - 1, mask_t = (1ull << (3 * n_t)) - 1; const uint32_t sh_t = 3 * n_t;
	std::vector<std::vector<uint8_t>> refs;
	if (first_refs) refs = *first_refs;
	Reads out; out.offsets.reserve(static_cast<size_t>(n_reads) + 1);
	std::vector<uint8_t> rd, fl, part;
	uint32_t r = 0; uint64_t ctx_read_type = 0; size_t md = 0;
	part_reads.clear();
	while (next(part, md)) {
		if (md > n_reads - r) throw DecodeError("colord-b200: DNA parts hold more reads than the archive announces");
		part_reads.push_back(static_cast<uint32_t>(md));
		d.start(part.data(), part.size());
		for (size_t k = 0; k < md; ++k, ++r) {
			rd.clear(); fl.clear();
			const uint32_t flag = d.get(F_FLAG, ctx_read_type);
			ctx_read_type = ((ctx_read_type << 2) + flag) & 0xff;
			const uint32_t len = dna_read_len(d);
			const uint64_t budget = max_bases - std::min<uint64_t>(max_bases, out.bases.size());
			if (flag != 2 && len > budget) throw DecodeError("colord-b200: damaged DNA stream (more bases than the archive announces)");
			uint64_t ctx_symbol = mask_s;
			if (flag == 0) for (uint32_t i = 0; i < len; ++i) { const uint32_t s = d.get(F_SYM, ctx_symbol << 2); rd.push_back(static_cast<uint8_t>(s)); ctx_symbol = ((ctx_symbol << 2) + s) & mask_s; }
			else if (flag == 1) for (uint32_t i = 0; i < len; ++i) { const uint32_t s = d.get(F_SYMN, ctx_symbol); rd.push_back(static_cast<uint8_t>(s)); ctx_symbol = ((ctx_symbol << 4) + s) & mask_s; }
			else dna_edit_script(d, level, n_s, n_first + r, len, refs, rd, fl, ctx_symbol, mask_t, mask_s, mask_t, sh_t, budget);
			fl.resize(rd.size(), 0);
			for (uint8_t s : rd) out.bases.push_back("ACGTN"[s > 4 ? 4 : s]);
			if (want_flags) out.flags.insert(out.flags.end(), fl.begin(), fl.end());
			out.offsets.push_back(out.bases.size());
			if (decisions[r] && flag != 1) refs.push_back(rd);
		}
	}
	if (r != n_reads) throw DecodeError("colord-b200: DNA stream holds fewer reads than the archive says");
	return out;
}

// ---- "qual": CQualityCoder::Decode.  mode = QualityComprMode value, source = DataSource value, rev = meta's representatives. ----
inline std::vector<uint8_t> decode_qual(const PartSource& next, const dec::Reads& reads, const std::vector<uint32_t>& part_reads, uint32_t mode, uint32_t source, uint32_t level,
	const std::vector<uint32_t>& rev)
{
	using namespace dec;
	enum { Q_SYM, Q_BYTE };
	if (mode > 8 || source > 2 || level < 1 || level > 3) throw DecodeError("colord-b200: the quality stream's parameters are out of range");
	std::vector<uint8_t> out(reads.bases.size());
	if (mode == 8) { std::fill(out.begin(), out.end(), static_cast<uint8_t>(33 + rev.at(0))); return out; }      // quality_coder.cpp:611-617
	const uint32_t n_bins = (mode == 1 || mode == 4) ? 5 : (mode == 2 || mode == 5) ? 4 : (mode == 3 || mode == 6) ? 2 : 0;
	uint32_t bps, ncs;                                                    // quality_coder.cpp:59-262
	if (mode == 0) { bps = 4; ncs = 2; } else if (mode == 7) { bps = 8; ncs = 2; } else if (n_bins == 2) { bps = 2; ncs = 6; } else { bps = 3; ncs = 3; }
	const uint32_t cbits = bps * ncs; const uint64_t cmask = (1ull << cbits) - 1;
	const bool thresholds = mode >= 4 && mode <= 6, averages = mode >= 1 && mode <= 3;
	if (thresholds && rev.size() < n_bins) throw DecodeError("colord-b200: the archive lacks the quality representatives of its mode");
	if (level > 1 && reads.flags.size() != reads.bases.size()) throw DecodeError("colord-b200: the quality stream needs the per-base flags of the DNA stream");
	Models M({{mode == 0 ? 96u : n_bins ? n_bins : 2u, mode == 0 ? 1u << 20 : 1u << 18, mode == 0 ? 32u : 8u}, {256, 1u << 18, 8}});      // quality_coder.h:35-39
	Coder d(M);
	uint8_t quant[96];
	qorg_quantiser(source, level, quant);
	auto vs = [&](uint64_t at) -> uint64_t { const uint8_t ch = reads.bases[at]; return ch == 'C' ? 1 : ch == 'G' ? 2 : ch == 'T' ? 3 : 0; };
	auto avg_of = [&](uint64_t ctx) { const uint32_t a1 = d.get(Q_BYTE, ctx), a2 = d.get(Q_BYTE, a1 + 0x100ull); return static_cast<double>((a1 << 8) + a2) / 256.0; };      // decode_avg impl:836-850
	std::vector<uint8_t> part; size_t md = 0; uint32_t r = 0;
	for (const uint32_t n_in_part : part_reads) {
		if (!next(part, md)) throw DecodeError("colord-b200: the quality stream has fewer parts than the DNA stream");
		d.start(part.data(), part.size());
		for (uint32_t k = 0; k < n_in_part; ++k, ++r) {
			const uint64_t o = reads.offsets[r]; const uint32_t n = static_cast<uint32_t>(reads.offsets[r + 1] - o);
			auto flags_into = [&](uint64_t& ctx, uint32_t sh, uint32_t i) { if (level > 1) { ctx += static_cast<uint64_t>(reads.flags[o + i] == 1) << sh; ctx += static_cast<uint64_t>(reads.flags[o + i] == 2) << (sh + 1); } };
			if (mode == 7) {                                                  // decode_average impl:802-818
				const double avg = avg_of(0ull); double avg_sum = 0.0, qual_sum = 0.0;
				for (uint32_t i = 0; i < n; ++i) { avg_sum += avg; const uint32_t q = static_cast<uint32_t>(avg_sum - qual_sum); qual_sum += q; out[o + i] = static_cast<uint8_t>(q + 33); }
				continue;
			}
			uint64_t context = cmask;
			if (averages) {                                                   // decode_*_average impl:508-672
				double avg[5], avg_sum[5] = {0, 0, 0, 0, 0}, qual_sum[5] = {0, 0, 0, 0, 0}; uint64_t ctx_p = 0;
				for (uint32_t b = 0; b < n_bins; ++b) { avg[b] = avg_of((1ull << 30) + (static_cast<uint64_t>(b) << 24) + (ctx_p << 16)); ctx_p = static_cast<uint64_t>(avg[b]); }
				uint64_t dna = n ? vs(o) : 3;
				for (uint32_t i = 0; i < n; ++i) {
					uint64_t ctx = context;
					dna <<= 2; if (i + 1 < n) dna += vs(o + i + 1); dna &= 0xff;
					ctx += dna << cbits;
					flags_into(ctx, cbits + 8, i);
					const uint32_t s = d.get(Q_SYM, ctx);
					avg_sum[s] += avg[s];
					const uint32_t q = static_cast<uint32_t>(avg_sum[s] - qual_sum[s]);
					qual_sum[s] += q;
					out[o + i] = static_cast<uint8_t>(q + 33);
					context = ((context << bps) + s) & cmask;
				}
				continue;
			}
			for (uint32_t i = 0; i < n; ++i) {                                // decode_original impl:452-505, decode_*_threshold impl:674-800
				uint64_t ctx = context; uint32_t sh = cbits;
				ctx += vs(o + i) << sh; sh += 2;
				if (i > 0) ctx += vs(o + i - 1) << sh;
				sh += 2;
				if (mode != 0 || level == 3) { if (i > 1) ctx += vs(o + i - 2) << sh; sh += 2; }
				else { if (i > 1) ctx += static_cast<uint64_t>(vs(o + i - 2) == vs(o + i - 1)) << sh; sh += 1; }
				if (i + 1 < n) ctx += vs(o + i + 1) << sh;
				sh += 2;
				flags_into(ctx, sh, i);
				const uint32_t s = d.get(Q_SYM, ctx);
				out[o + i] = static_cast<uint8_t>(33 + (mode == 0 ? s : rev[s]));
				context = ((context << bps) + (mode == 0 ? quant[s] : s)) & cmask;
			}
		}
	}
	if (r + 1 != reads.offsets.size()) throw DecodeError("colord-b200: the quality stream covers fewer reads than the archive holds");
	return out;
}

// ---- "header": CIDCoder::Decode (lossless mode) ----
inline dec::Headers decode_headers(const PartSource& next, uint64_t n_expected, uint64_t max_bytes)
{
	using namespace dec;
	enum { H_PLUS, H_FLAGS, H_SAME, H_SAMELEN, H_LITERAL, H_PLAIN };
	Models M({{2, 1u << 15, 1}, {2, 1u << 15, 1}, {2, 1u << 15, 1}, {2, 1u << 15, 1}, {256, 1u << 20, 64}, {128, 1u << 19, 32}});      // id_coder.h:50-59
	Coder d(M);
	auto is_lit = [](uint8_t c) { return (c >= '0' && c <= '9') || (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z') || c == '@'; };
	Headers H;
	std::vector<uint8_t> prev, cur, part; size_t md = 0; uint64_t r = 0;
	auto room = [&](size_t k) { if (H.bytes.size() + cur.size() + k > max_bytes) throw DecodeError("colord-b200: damaged header stream (more bytes than the archive announces)"); };
	while (next(part, md)) {
		if (md > n_expected - r) throw DecodeError("colord-b200: header parts hold more headers than the archive announces");
		d.start(part.data(), part.size());
		uint64_t ctx_flags = 0;                                               // Restart (id_coder.cpp:80-91): the previous header stays
		for (size_t k = 0; k < md; ++k, ++r) {
			cur.clear();
			const uint32_t plus = d.get(H_PLUS, 0);
			const uint32_t flag = d.get(H_FLAGS, ctx_flags);
			if (flag) {
				if (r == 0) throw DecodeError("colord-b200: damaged header stream");
				ctx_flags = ((ctx_flags << 1) + 1) & 0xff;
				size_t b = 0;
				for (uint64_t t = 0;; ++t) {
					size_t e = b;
					while (e < prev.size() && is_lit(prev[e])) ++e;
					const size_t lp = e - b;
					if (d.get(H_SAME, t)) { room(lp); cur.insert(cur.end(), prev.begin() + b, prev.begin() + e); }
					else if (d.get(H_SAMELEN, t)) {
						room(lp);
						for (size_t j = 0; j < lp; ++j) { const uint32_t c = d.get(H_LITERAL, ctx_flags + (1ull << 32) + j + (t << 40) + (1ull << 60)); cur.push_back(c ? static_cast<uint8_t>(c) : prev[b + j]); }
					} else {
						for (uint64_t j = 0;; ++j) { const uint32_t c = d.get(H_LITERAL, ctx_flags + j + (1ull << 32) + (t << 40)); if (!c) break; room(1); cur.push_back(static_cast<uint8_t>(c)); }
					}
					if (e == prev.size()) break;
					room(1); cur.push_back(prev[e]);                                    // the separator: same shape as the previous header
					b = e + 1;
				}
			} else {
				ctx_flags = (ctx_flags << 1) & 0xff;
				for (uint64_t i = 0;; ++i) { const uint32_t c = d.get(H_PLAIN, i); if (!c) break; room(1); cur.push_back(static_cast<uint8_t>(c)); }
			}
			H.bytes.insert(H.bytes.end(), cur.begin(), cur.end());
			H.offsets.push_back(H.bytes.size());
			H.plus_id.push_back(static_cast<uint8_t>(plus));
			prev.swap(cur);
		}
	}
	if (r != n_expected) throw DecodeError("colord-b200: header stream holds fewer headers than the archive says");
	return H;
}

} // namespace xdec
} // namespace clbhost
