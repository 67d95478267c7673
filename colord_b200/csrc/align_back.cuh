// align_back.cuh — the backward half of an edit-script task, ONE THREAD PER TASK (k_align_back in
// stage2_encode.cu): traceback through the history a forward sweep left (align.cuh: Aligner::sweep, layout in AlignScratch), script
// symbols, and the reference's canonical form — for problems below edlib's traceback limit.
//
// Why a thread per task: the walk is a serial chain of tiny steps whose only cost is the latency of the history entries it reads
// (they left the L2 long ago).  A lane group per task leaves its lanes idle behind that latency; a thread per task puts 32 walks
// into a warp, and each thread keeps the next columns of its block on their way from DRAM with cp.async into a ring in shared
// memory (no registers held, no stall until the walk arrives there).
//
// Reference behaviour restated:
//   * edlib.cpp:945-1159   traceback preference: up (query symbol only) if D(i-1,j)+1 == D(i,j), else left, else diagonal — two bit
//                          tests on the stored Pv / Ph (see align.cuh)
//   * edit_script.h:272-413  what the three wrappers do with the path (SHW end column, the all-insertions case, reversed left flank)
//   * edit_script.h:432-447, :591-671  FixInRange / refactor_edit_script.  Pass 1 gathers maximal ranges of M / D symbols over one
//     repeated reference base (insertions and substitutions end a range) and moves the M's to the front; pass 2 does the same for
//     M / insertion symbols over one repeated read base.  The non-M symbols of a range are all equal ('D', or the letter of the
//     range's base), so a range is fully described by two counts: both passes are STREAMING transducers (symbol in; on a range end
//     M x a, then the other symbol x b out), pass 2 fed by pass 1, and no symbol is ever written twice.
#pragma once
#include "align.cuh"

namespace clb {

constexpr int BACK_RING = 8;                 // history columns of the current block a thread keeps in shared memory

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem)
{
	asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

// symbols of a view one after the other (32 per packed word)
template <class V>
struct SymbolReader {
	V v; int pos, len, avail; uint64_t w;
	__device__ SymbolReader(const V& view, int n) : v(view), pos(0), len(n), avail(0), w(0) {}
	__device__ __forceinline__ uint32_t next()
	{
		if (!avail) { avail = min(32, len - pos); w = v.get32(pos, avail); }
		const uint32_t s = (uint32_t)w & 3u; w >>= 2; --avail; ++pos;
		return s;
	}
};

// script bytes, four per store (the script buffer of a task is 4-byte aligned and has room for whole words)
struct ScriptOut {
	uint32_t* p; uint32_t acc, k;
	__device__ explicit ScriptOut(char* out) : p(reinterpret_cast<uint32_t*>(out)), acc(0), k(0) {}
	__device__ __forceinline__ void put(uint32_t c) { acc |= c << (8 * k); if (++k == 4) { *p++ = acc; acc = 0; k = 0; } }
	__device__ __forceinline__ void finish() { if (k) *p = acc; }
};

constexpr uint32_t BASE_ANY = 0x100;         // "the base of the range is that of its first symbol"

// pass 2 of refactor_edit_script (edit_script.h:634-668) as a transducer: ranges of M / insertion symbols over one read base
template <class V>
struct CanonPass2 {
	ScriptOut& o; SymbolReader<V> enc; uint32_t base, n_m, n_x;
	__device__ CanonPass2(ScriptOut& out, const V& e, int el) : o(out), enc(e, el), base(BASE_ANY), n_m(0), n_x(0) {}
	__device__ __forceinline__ void flush() { for (; n_m; --n_m) o.put('M'); for (; n_x; --n_x) o.put((uint32_t)"ACGT"[base & 3]); }
	__device__ __forceinline__ void put(uint32_t s)
	{
		const bool del = s == 'D', mm = s == 'X' || s == 'Y' || s == 'Z';
		uint32_t b = BASE_ANY;
		if (!del) b = enc.next();                                   // the read base this symbol consumes
		if (del || mm || (base != BASE_ANY && base != b)) {
			flush();
			if (del || mm) { o.put(s); base = mm ? b : BASE_ANY; return; }
			base = b;
		}
		if (base == BASE_ANY) base = b;
		if (s == 'M') ++n_m; else ++n_x;
	}
};
// pass 1 (edit_script.h:597-632): ranges of M / D symbols over one reference base; `b` = the reference base the symbol consumes
template <class V>
struct CanonPass1 {
	CanonPass2<V>& nx; uint32_t base, n_m, n_d;
	__device__ explicit CanonPass1(CanonPass2<V>& next) : nx(next), base(BASE_ANY), n_m(0), n_d(0) {}
	__device__ __forceinline__ void flush() { for (; n_m; --n_m) nx.put('M'); for (; n_d; --n_d) nx.put('D'); }
	__device__ __forceinline__ void put(uint32_t s, uint32_t b)
	{
		const bool ins = s == 'A' || s == 'C' || s == 'G' || s == 'T', mm = s == 'X' || s == 'Y' || s == 'Z';
		if (ins || mm || (base != BASE_ANY && base != b)) {
			flush();
			if (ins || mm) { nx.put(s); base = mm ? b : BASE_ANY; return; }
			base = b;
		}
		if (base == BASE_ANY) base = b;
		if (s == 'M') ++n_m; else ++n_d;
	}
};

// What a task's forward kernel ran on, derived again from the task (edit_script_task, PHASE 1).
//   rows / cols of the sweep: kind 2: ref x enc; flank with cut < 2 or el < 2: ref prefix x enc; else (SHW) enc x ref prefix, both
//   reversed for the left flank.
struct BackPlan { int Q, Ts; bool shw, rows_ref; uint32_t cut; };
__device__ __forceinline__ BackPlan back_plan(uint32_t rl, uint32_t el, uint32_t kind)
{
	BackPlan p;
	if (kind == 2) { p.Q = (int)rl; p.Ts = (int)el; p.shw = false; p.rows_ref = true; p.cut = rl; return p; }
	p.cut = rl < 2 * el ? rl : 2 * el;
	if (p.cut < 2 || el < 2) { p.Q = (int)p.cut; p.Ts = (int)el; p.shw = false; p.rows_ref = true; }
	else { p.Q = (int)el; p.Ts = (int)p.cut; p.shw = true; p.rows_ref = false; }
	return p;
}

// ring: BACK_RING x blockDim.x entries of 16 bytes in shared memory; entry of column c of this thread: ring[(c & (BACK_RING - 1)) * blockDim.x]
// lg: log2 of the lane group of the forward sweep (history layout).  Returns the script length; *lead_out = the 'D' run before a
// left flank's script (not written).
template <class V>
__device__ uint32_t edit_script_back(uint8_t* scratch, const AlignScratch& lay, int lg, ulonglong2* ring, V ref, uint32_t rl, V enc, uint32_t el, uint32_t kind,
	char* out, uint32_t* lead_out)
{
	const BackPlan P = back_plan(rl, el, kind);
	const ulonglong2* __restrict__ hist = reinterpret_cast<const ulonglong2*>(scratch + lay.hist);
	uint64_t* __restrict__ ops = reinterpret_cast<uint64_t*>(scratch + lay.ops);      // 2 bits per op, traceback order, 32 per word
	const int G = 1 << lg, B = (P.Q + 63) >> 6, n_steps = P.Ts + G - 1;
	const uint32_t nthr = blockDim.x;
	auto entry = [&](int b, int c) -> const ulonglong2* {
		const int strip = b >> lg, g = b & (G - 1), Gs = min(G, B - (strip << lg));
		return hist + ((size_t)strip * n_steps * G + (size_t)(c + g) * Gs + g);
	};
	int T = P.Ts;
	uint32_t ref_end = P.cut - 1;
	bool all_ins = false;
	if (P.shw) {      // leftmost minimum of the last row, left by the forward kernel (edlib.cpp:660-694, edit_script.h:352)
		const int* res = reinterpret_cast<const int*>(scratch + lay.res);
		const int best = res[0], end = res[1];
		ref_end = (uint32_t)end; T = end + 1;
		if (best >= (int)el) { all_ins = true; ref_end = 0xFFFFFFFFu; }
	}
	// ---- traceback: ops in walk order (from the last vertex to the first) ----
	uint32_t n = 0; uint64_t acc = 0;
	auto emit = [&](uint32_t op) { acc |= (uint64_t)op << (2 * (n & 31)); if ((++n & 31) == 0) { ops[(n >> 5) - 1] = acc; acc = 0; } };
	if (all_ins) { for (uint32_t x = 0; x < el; ++x) emit(1); }
	else {
		int I = P.Q, J = T;
		auto load_block = [&](int b) {      // columns J-1 .. J-BACK_RING of block b
			for (int k = 0; k < BACK_RING; ++k) { const int c = J - 1 - k; if (c >= 0) cp_async16(&ring[(c & (BACK_RING - 1)) * nthr], entry(b, c)); }
			cp_async_commit();
			cp_async_wait<0>();
		};
		int b = (I - 1) >> 6;
		load_block(b);
		while (I > 0 && J > 0) {
			const int i = I - 1, j = J - 1, bit = i & 63;
			cp_async_wait<BACK_RING - 1>();                      // the column asked for BACK_RING - 1 column moves ago has arrived
			const ulonglong2 e = ring[(j & (BACK_RING - 1)) * nthr];
			const bool up = (e.x >> bit) & 1, left = !up && ((e.y >> bit) & 1);
			emit(up ? 1u : left ? 2u : 0u);
			if (bit == 8 && b > 0 && !left) {                      // eight rows below the block above: its entries around the crossing start from DRAM
				for (int k = 1; k <= 24; ++k) if (j - k >= 0) prefetch_l2(entry(b - 1, j - k));
			}
			if (!left) --I;
			if (!up) --J;
			if (!left && bit == 0 && I > 0 && J > 0) {           // into the block above: its columns replace the ring
				cp_async_wait<0>();
				load_block(--b);
			} else if (!up) {                                     // one column to the left: the column BACK_RING further enters the ring
				const int c = j - BACK_RING;
				if (c >= 0) cp_async16(&ring[(c & (BACK_RING - 1)) * nthr], entry(b, c));
				cp_async_commit();
			}
		}
		cp_async_wait<0>();
		for (; I > 0; --I) emit(1);                               // a border was reached: the rest is all up or all left
		for (; J > 0; --J) emit(2);
	}
	if (n & 31) ops[n >> 5] = acc;
	// ---- script symbols in script order (the walk's order for the reversed left flank, its reverse otherwise) + canonical form ----
	uint32_t lead = 0;
	if (kind == 0) lead = (rl - 1) - ref_end;
	*lead_out = lead;
	const bool walk_order = kind == 0;
	SymbolReader<V> rr(kind == 0 ? ref.sub((int)lead) : ref, (int)(rl - lead)), er(enc, (int)el);
	ScriptOut so(out);
	CanonPass2<V> p2(so, enc, (int)el);
	CanonPass1<V> p1(p2);
	uint64_t w = 0;
	for (uint32_t x = 0; x < n; ++x) {
		const uint32_t k = walk_order ? x : n - 1 - x;
		if (x == 0 || (walk_order ? (k & 31) == 0 : (k & 31) == 31)) w = ops[k >> 5];
		const uint32_t op = (uint32_t)(w >> (2 * (k & 31))) & 3u;
		if (op == 0) { const uint32_t a = rr.next(), c = er.next(); p1.put(a == c ? (uint32_t)'M' : (uint32_t)mismatch_symb((uint8_t)a, (uint8_t)c), a); }
		else if ((op == 1) == P.rows_ref) p1.put('D', rr.next());
		else p1.put((uint32_t)"ACGT"[er.next() & 3], BASE_ANY);
	}
	p1.flush(); p2.flush(); so.finish();
	return n;
}

} // namespace clb
