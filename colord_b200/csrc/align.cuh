// align.cuh — device-side unit-cost alignment with the exact path the reference obtains from edlib, plus the
// reference's edit-script canonicalisation.  SURVEY.md §8 rows E6 / E7.
//
// Reference behaviour restated (see oracle/stage2.c for the scalar form and the pinning tests):
//   * src/colord/edit_script.h:272-413   the three wrappers (NW between anchors, SHW for the right flank, SHW on reversed
//                                         strings for the left flank), tiny inputs via find_edit_dist (:156-245)
//   * src/colord/libs/edlib/edlib.cpp     :395-441 calculateBlock (Myers/Hyyrö bit-vector column step)
//                                         :945-1159 traceback: prefer UP (query symbol), then LEFT (target symbol), then diagonal
//                                         :1178-1215 stored-column traceback below 1 MiB of column data, else
//                                         :1234-1377 Hirschberg: halve the target, split at the TOPMOST row on an optimal path
//   * src/colord/edit_script.h:432-447, :591-671   FixInRange / refactor_edit_script
// edlib's Ukkonen band only removes cells no optimal path visits, so the full (unbanded) bit-vector matrix gives the same moves.
//
// Parallel scheme: a task is owned by a group of GROUP lanes (1, 2, 4, 8 or 32).  Rows are cut into 64-row blocks; lane g of
// the group owns block strip*GROUP + g and the group sweeps the columns as an anti-diagonal wavefront (lane g is g columns
// behind lane g-1, the horizontal delta travels by shuffle).  Tasks with more blocks than lanes are strip-mined: the
// horizontal deltas leaving a strip's last block are kept per column (1 byte) and feed the next strip.  For the traceback
// only two bit-vectors per (block, column) are stored: Pv (vertical +1 deltas after the column) and Ph (horizontal +1 deltas
// of the column).  edlib's preference "up if D(i-1,j)+1 == D(i,j), else left if D(i,j-1)+1 == D(i,j), else diagonal" is then
// two bit tests per step (a diagonal step is a match iff the two symbols are equal), and the walk keeps a window of GROUP
// columns of the current block in registers (one column per lane, fetched by shuffle) so that a step costs no memory round
// trip.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace clb {

// -DCLB_ALIGN_PHASES (profiling builds only, `make phases`): cycles of the group's first lane per phase, summed over all tasks
//   0 sweep, 1 traceback, 2 finish_script, 3 whole task, 4 sweep steps, 5 traceback iterations, 6 traceback window loads, 7 tasks
#ifdef CLB_ALIGN_PHASES
__device__ unsigned long long g_align_phase[6 * 8];
#define CLB_PH_BEGIN long long ph_t0_ = clock64();
#define CLB_PH_END(G, i) { if (gl == 0) atomicAdd(&g_align_phase[(G == 32 ? 5 : G == 16 ? 4 : G == 8 ? 3 : G == 4 ? 2 : G == 2 ? 1 : 0) * 8 + (i)], (unsigned long long)(clock64() - ph_t0_)); }
#define CLB_PH_COUNT(G, i, n) { if (gl == 0) atomicAdd(&g_align_phase[(G == 32 ? 5 : G == 16 ? 4 : G == 8 ? 3 : G == 4 ? 2 : G == 2 ? 1 : 0) * 8 + (i)], (unsigned long long)(n)); }
#else
#define CLB_PH_BEGIN
#define CLB_PH_END(G, i)
#define CLB_PH_COUNT(G, i, n)
#endif

struct SeqView {            // element i = base[i * step]; step = -1 gives a reversed view
	const uint8_t* base; int step;
	__device__ __forceinline__ uint8_t operator[](int i) const { return base[(long long)i * step]; }
	// symbols i .. i+n-1 (n <= 32), symbol k in bits 2k, 2k+1
	__device__ __forceinline__ uint64_t get32(int i, int n) const
	{
		uint64_t x = 0;
		for (int k = 0; k < n; ++k) x |= (uint64_t)(base[(long long)(i + k) * step] & 3) << (2 * k);
		return x;
	}
	__device__ __forceinline__ SeqView sub(int off) const { return SeqView{base + (long long)off * step, step}; }
	__device__ __forceinline__ SeqView reversed(int n) const { return SeqView{base + (long long)(n - 1) * step, -step}; }
};

// The same interface over the resident 2-bit packed read store (ctx.h: pk, 32 bases per word, first base in the top bits).
// element i = base at absolute stream position origin + i*step, complemented if comp == 3 (reverse-complement views use
// step = -1, comp = 3); positions outside [lo, hi) — the read's extent — give the reference's 255 guard byte.
struct PackedView {
	const uint64_t* pk; long long origin; int step; uint32_t comp; long long lo, hi;
	__device__ __forceinline__ uint8_t operator[](int i) const
	{
		const long long a = origin + (long long)i * step;
		if (a < lo || a >= hi) return 255;
		return (uint8_t)(((uint32_t)(pk[a >> 5] >> (62 - 2 * (a & 31))) & 3u) ^ comp);
	}
	// symbols i .. i+n-1 (n <= 32, all inside the read), symbol k in bits 2k, 2k+1; bits above 2n are unspecified
	__device__ __forceinline__ uint64_t get32(int i, int n) const
	{
		const long long a0 = origin + (long long)i * step, a1 = a0 + (long long)(n - 1) * step;
		const long long lo = step > 0 ? a0 : a1, hi = step > 0 ? a1 : a0;
		const uint64_t w0 = pk[lo >> 5], w1 = pk[hi >> 5];
		const uint32_t sh = 2 * (uint32_t)(lo & 31);
		uint64_t x = sh ? ((w0 << sh) | (w1 >> (64 - sh))) : w0;        // address lo + k in bits 63-2k, 62-2k
		if (step > 0) {                                                  // element k = address lo + k: reverse the order of the 2-bit fields
			x = __brevll(x);
			x = ((x >> 1) & 0x5555555555555555ULL) | ((x & 0x5555555555555555ULL) << 1);
		} else x >>= (64 - 2 * n);                                       // element k = address hi - k
		return comp ? ~x : x;
	}
	__device__ __forceinline__ PackedView sub(int off) const { PackedView v = *this; v.origin += (long long)off * step; return v; }
	__device__ __forceinline__ PackedView reversed(int n) const { PackedView v = *this; v.origin += (long long)(n - 1) * step; v.step = -step; return v; }
};

__host__ __device__ inline long long edlib_column_bytes(long long q, long long t) { return (2ll * 8 + 4) * ((q + 63) / 64) * t + 2ll * 4 * t; }
constexpr long long EDLIB_TRACEBACK_LIMIT = 1024 * 1024;

// Scratch layout of one task (host and device agree through this function).
// hist: Pv / Ph of every (block, column) of a stored-column sweep, 16 bytes per entry, in the order the wavefront produces
// them (see Aligner::sweep): entry of block b, column c = strip * n_steps * G + (c + g) * Gs + g with G the lane group (<= 32),
// strip = b / G, g = b % G, n_steps = T + G - 1, Gs = blocks of that strip — at most B * (T + 31) entries.
struct AlignScratch {
	unsigned long long hist, carry, res, ops, tmp, F, R, fin, stack, total;
};
__host__ __device__ inline AlignScratch align_scratch_layout(long long q, long long t)
{
	AlignScratch s{};
	const long long B = (q + 63) / 64;
	const bool big = edlib_column_bytes(q, t) >= EDLIB_TRACEBACK_LIMIT;
	unsigned long long o = 0;
	auto take = [&](unsigned long long bytes) { unsigned long long at = o; o += (bytes + 15) & ~15ULL; return at; };
	// a Hirschberg leaf has 20 B T + 8 T < 1 MiB: 16 B (T + 31) < 1 MiB + 496 B
	s.hist = take(big ? (unsigned long long)EDLIB_TRACEBACK_LIMIT + 4096 + 512ull * B : (unsigned long long)(16 * B * ((t > 0 ? t : 1) + 31)));
	s.carry = take(2ull * 8 * ((unsigned long long)(t + 63) / 32 + 4));      // two arrays of 2-bit horizontal deltas, 32 per word
	s.res = take(16);                                                      // what the forward kernel leaves for the backward kernel (last-row minimum)
	s.ops = take((unsigned long long)(q + t + 2));
	s.tmp = take((unsigned long long)(q + t + 2));
	s.F = take(big ? 4ull * (q + 1) : 0);
	s.R = take(big ? 4ull * (q + 1) : 0);
	s.fin = take(big ? 20ull * B : 0);
	s.stack = take(big ? 64ull * 5 * 4 : 0);
	s.total = o;
	return s;
}

// bits 0, 2, 4, .. 62 of x -> bits 0 .. 31
__device__ __forceinline__ uint64_t compress_even(uint64_t x)
{
	x &= 0x5555555555555555ULL;
	x = (x | (x >> 1)) & 0x3333333333333333ULL;
	x = (x | (x >> 2)) & 0x0f0f0f0f0f0f0f0fULL;
	x = (x | (x >> 4)) & 0x00ff00ff00ff00ffULL;
	x = (x | (x >> 8)) & 0x0000ffff0000ffffULL;
	x = (x | (x >> 16)) & 0x00000000ffffffffULL;
	return x;
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }

constexpr int ALIGN_THREADS = 128;          // CTA size of the kernels that own an Aligner (its shared match-vector table)

template <int GROUP>
struct Aligner {
	uint32_t gl;            // lane within the group
	uint32_t gmask;         // warp mask of the group's lanes
	uint8_t* scratch;       // this task's scratch
	AlignScratch lay;
	uint64_t* peq;          // shared memory: match vector of this lane's 64 rows for base x at peq[x * ALIGN_THREADS]

	// the lanes of the group as a warp mask; for a whole-warp group a literal, so that the compiler drops the convergence checks of
	// a run-time mask (MATCH.ANY / VOTEU / BRA.DIV around every shuffle)
	__device__ __forceinline__ uint32_t mask() const { return GROUP == 32 ? 0xffffffffu : gmask; }
	__device__ __forceinline__ void gsync() const { if (GROUP > 1) __syncwarp(mask()); }

	// what a sweep leaves in the lanes' registers (valid in every lane of the group after the sweep)
	struct SweepResult { int score; int best, end; };      // D(Q-1, T-1); leftmost minimum of the last row D(Q-1, .) and its column

	// state of one lane during a strip
	struct Lane {
		uint64_t Pv, Mv;        // vertical deltas of the block after the last column done
		uint64_t cols;          // column symbols of the coming steps of the chunk, 2 bits each, next one in bits 0-1
		uint64_t cin;           // horizontal deltas entering lane 0 from the previous strip, same order
		uint64_t cacc;          // horizontal deltas leaving this lane, last 32 steps (newest in the top bits)
		uint32_t hp;            // horizontal delta leaving the block in the last column done: bit 0 = +1, bit 1 = -1
		int score, best, end;
	};

	// One column of one 64-row block: edlib.cpp:407-441 (calculateBlock).  CHECK: the lane may be outside its columns (the
	// wavefront's ramps); FIRST: no strip above (a +1 enters the top block); HIST: Pv / Ph go to the history; SCORE: the score of
	// the block's last row is followed (bit sbit of the horizontal deltas) and, in the lane with the last block, its leftmost minimum.
	template <bool CHECK, bool FIRST, bool HIST, bool SCORE>
	__device__ __forceinline__ void step(Lane& L, int s, int T, bool act, bool last_blk, uint32_t sbit, ulonglong2* __restrict__ hp_at) const
	{
		uint32_t hin = GROUP > 1 ? __shfl_up_sync(mask(), L.hp, 1, GROUP) : 0u;
		if (gl == 0) hin = FIRST ? 1u : (uint32_t)L.cin & 3u;
		if (!FIRST) L.cin >>= 2;
		const uint64_t Eq0 = peq[((uint32_t)L.cols & 3u) * ALIGN_THREADS];
		L.cols >>= 2;
		const int c = s - (int)gl;
		if (!CHECK || (act && (unsigned)c < (unsigned)T)) {
			const uint64_t hneg = hin >> 1, hpos = hin & 1u;
			const uint64_t Xv = Eq0 | L.Mv;
			const uint64_t Eq = Eq0 | hneg;
			const uint64_t Xh = (((Eq & L.Pv) + L.Pv) ^ L.Pv) | Eq;
			uint64_t Ph = L.Mv | ~(Xh | L.Pv);
			uint64_t Mh = L.Pv & Xh;
			L.hp = (uint32_t)(Ph >> 63) | ((uint32_t)(Mh >> 63) << 1);
			if (SCORE) {
				L.score += (int)((Ph >> sbit) & 1) - (int)((Mh >> sbit) & 1);
				if (last_blk && L.score < L.best) { L.best = L.score; L.end = c; }
			}
			const uint64_t ph0 = Ph;
			Ph = (Ph << 1) | hpos; Mh = (Mh << 1) | hneg;
			L.Pv = Mh | ~(Xv | Ph);
			L.Mv = Ph & Xv;
			if (HIST && (CHECK || act)) *hp_at = make_ulonglong2(L.Pv, ph0);
		}
		if (GROUP == 32) L.cacc = (L.cacc >> 2) | ((uint64_t)L.hp << 62);
	}

	// Forward sweep of rows[0..Q) x cols[0..T): Myers / Hyyrö columns over 64-row blocks, lane g of the group owning block
	// strip * GROUP + g and running g columns behind lane g - 1 (anti-diagonal wavefront; the horizontal delta at the block
	// border travels by shuffle).  Steps come in chunks of 32: per chunk a lane takes its 32 column symbols from two packed
	// words, and (whole-warp groups only, the others have a single strip) lane 0 its 32 incoming deltas of the strip above while
	// the last lane leaves the deltas of its bottom border as one word.  Chunks in which every lane is inside its columns run
	// without range checks.
	// HIST: hist receives Pv / Ph of every (block, column) in wavefront order (layout: AlignScratch) — the lanes of a step
	// write neighbouring 16-byte entries.  SCORE: scores are followed (see step); fin_* (per block, may be null) receive the
	// final column.
	template <bool HIST, bool SCORE, class V>
	__device__ SweepResult sweep(V rows, int Q, V cols, int T, ulonglong2* __restrict__ hist, uint64_t* fin_pv, uint64_t* fin_mv, int32_t* fin_sc) const
	{
		CLB_PH_BEGIN
		const int B = (Q + 63) >> 6;
		const int n_steps = T + GROUP - 1;
		CLB_PH_COUNT(GROUP, 4, (long long)((B + GROUP - 1) / GROUP) * n_steps)
		uint64_t* const carry0 = reinterpret_cast<uint64_t*>(scratch + lay.carry);
		const int carry_words = (T + 63) / 32 + 4;
		SweepResult res{0, 0x7fffffff, 0};
		for (int strip = 0; strip * GROUP < B; ++strip) {
			const int b = strip * GROUP + (int)gl;
			const bool act = b < B;
			const bool last_blk = b == B - 1;
			const bool more = (strip + 1) * GROUP < B;                   // another strip follows: the last lane leaves its deltas
			const uint32_t sbit = last_blk ? (uint32_t)((Q - 1) & 63) : 63u;
			{      // match vectors of the block's 64 rows from two packed words
				uint64_t peq0 = 0, peq1 = 0, peq2 = 0, peq3 = 0;
				if (act) {
					const int r0 = b << 6, nr = min(Q - r0, 64);
#pragma unroll
					for (int h = 0; h < 2; ++h) {
						const int n = min(nr - 32 * h, 32);
						if (n <= 0) break;
						const uint64_t w = rows.get32(r0 + 32 * h, n);
						const uint64_t valid = n == 32 ? 0xffffffffULL : ((1ULL << n) - 1);
						const uint64_t e = compress_even(w), o = compress_even(w >> 1);
						peq0 |= (~e & ~o & valid) << (32 * h); peq1 |= (e & ~o & valid) << (32 * h);
						peq2 |= (~e & o & valid) << (32 * h); peq3 |= (e & o & valid) << (32 * h);
					}
				}
				peq[0] = peq0; peq[ALIGN_THREADS] = peq1; peq[2 * ALIGN_THREADS] = peq2; peq[3 * ALIGN_THREADS] = peq3;     // this thread's own cells
			}
			Lane L;
			L.Pv = ~0ULL; L.Mv = 0; L.hp = 0; L.cacc = 0; L.cin = 0; L.cols = 0;
			L.score = act ? min(Q, (b << 6) + 64) : 0;        // D(last row of block, column -1) = row index + 1
			L.best = 0x7fffffff; L.end = 0;
			const int Gs = min(GROUP, B - strip * GROUP);
			ulonglong2* hp_at = HIST ? hist + ((size_t)strip * n_steps * GROUP + gl) : nullptr;       // entry of step 0; a step is Gs entries
			const uint64_t* cin_arr = carry0 + (size_t)((strip + 1) & 1) * carry_words;          // written by the strip above
			uint64_t* cout_arr = carry0 + (size_t)(strip & 1) * carry_words;
			// column words: [s0 - 32, s0), [s0, s0 + 32) and the prefetched [s0 + 32, s0 + 64) of the chunk starting at step s0
			uint64_t col_a = 0, col_b = T > 0 ? cols.get32(0, min(32, T)) : 0, col_c = T > 32 ? cols.get32(32, min(32, T - 32)) : 0;
			// incoming deltas: the strip above left the delta of column c in slot (c + 31) & 31 of word (c + 31) >> 5
			uint64_t cin_a = 0, cin_b = 0, cin_c = 0;
			if (GROUP == 32 && strip > 0) { cin_a = cin_arr[0]; cin_b = cin_arr[1]; cin_c = cin_arr[2]; }
			for (int s0 = 0; s0 < n_steps; s0 += 32) {
				if (s0) {
					col_a = col_b; col_b = col_c;
					col_c = s0 + 32 < T ? cols.get32(s0 + 32, min(32, T - s0 - 32)) : 0;
					if (GROUP == 32 && strip > 0) { cin_a = cin_b; cin_b = cin_c; cin_c = cin_arr[(s0 >> 5) + 2]; }
				}
				L.cols = gl == 0 ? col_b : (col_a >> (64 - 2 * gl)) | (col_b << (2 * gl));       // columns s0 - gl .. s0 - gl + 31
				if (GROUP == 32 && strip > 0) L.cin = (cin_a >> 62) | (cin_b << 2);
				const int ns = min(32, n_steps - s0);
				if (s0 >= GROUP - 1 && s0 + 31 < T) {
					if (strip == 0) {
#pragma unroll 4
						for (int i = 0; i < 32; ++i) { step<false, true, HIST, SCORE>(L, s0 + i, T, act, last_blk, sbit, hp_at); if (HIST) hp_at += Gs; }
					} else {
#pragma unroll 4
						for (int i = 0; i < 32; ++i) { step<false, false, HIST, SCORE>(L, s0 + i, T, act, last_blk, sbit, hp_at); if (HIST) hp_at += Gs; }
					}
				} else if (strip == 0) {
					for (int i = 0; i < ns; ++i) { step<true, true, HIST, SCORE>(L, s0 + i, T, act, last_blk, sbit, hp_at); if (HIST) hp_at += Gs; }
				} else {
					for (int i = 0; i < ns; ++i) { step<true, false, HIST, SCORE>(L, s0 + i, T, act, last_blk, sbit, hp_at); if (HIST) hp_at += Gs; }
				}
				if (GROUP == 32 && more && gl == GROUP - 1) cout_arr[s0 >> 5] = ns == 32 ? L.cacc : L.cacc >> (64 - 2 * ns);
			}
			if (act && fin_pv) { fin_pv[b] = L.Pv; fin_mv[b] = L.Mv; fin_sc[b] = L.score; }
			if (act && last_blk) { res.score = L.score; res.best = L.best; res.end = L.end; }
			gsync();           // the deltas written by the last lane are read by every lane in the next strip
		}
		if (GROUP > 1) {       // from the lane that owns the last block to all
			const int owner = (B - 1) % GROUP;
			res.score = __shfl_sync(mask(), res.score, owner, GROUP);
			res.best = __shfl_sync(mask(), res.best, owner, GROUP);
			res.end = __shfl_sync(mask(), res.end, owner, GROUP);
		}
		CLB_PH_END(GROUP, 0)
		return res;
	}

	// history entry of block b, column c of a sweep over T columns and B blocks (layout: AlignScratch)
	__device__ __forceinline__ size_t hist_at(int b, int c, int T, int B) const
	{
		const int strip = b / GROUP, g = b % GROUP, Gs = min(GROUP, B - strip * GROUP);
		return (size_t)strip * (T + GROUP - 1) * GROUP + (size_t)(c + g) * Gs + g;
	}

	// ---- traceback on stored columns: edlib.cpp:945-1159 ----
	// Walks from vertex (Q, T) back to (0, 0) through the history of a sweep over Ts >= T columns; every lane of the group runs
	// the same walk, lane l holds column (window top - l) of the current block.  ops: 1 = up (row symbol only), 2 = left (column
	// symbol only), 0 = diagonal.  Written in forward order to out; returns their number.  tmp: Q + T bytes.
	__device__ int traceback(const ulonglong2* __restrict__ hist, int Ts, int Q, int T, uint8_t* out, uint8_t* tmp) const
	{
		CLB_PH_BEGIN
		const int B = (Q + 63) >> 6;
		int I = Q, J = T, n = 0;
		int wb = -1, wj = -0x40000000;
		uint64_t wpv = 0, wph = 0;
		const uint32_t gbase = (threadIdx.x & 31) & ~(uint32_t)(GROUP - 1);       // first lane of the group inside the warp
		const uint32_t gall = GROUP == 32 ? 0xffffffffu : ((1u << GROUP) - 1u);
		while (I > 0 && J > 0) {
			const int i = I - 1, j = J - 1, b = i >> 6, bit = i & 63;
			if (b != wb || j > wj || j < wj - (GROUP - 1)) {
				wb = b; wj = j;
				CLB_PH_COUNT(GROUP, 6, 1)
				const int col = wj - (int)gl;
				if (col >= 0) { const ulonglong2 e = hist[hist_at(b, col, Ts, B)]; wpv = e.x; wph = e.y; }
				// the walk goes on to the left in this block or to the block above: those entries start their way from DRAM now
				if (col - GROUP >= 0) prefetch_l2(&hist[hist_at(b, col - GROUP, Ts, B)]);
				if (b > 0 && col >= 0) { prefetch_l2(&hist[hist_at(b - 1, col, Ts, B)]); if (col - GROUP >= 0) prefetch_l2(&hist[hist_at(b - 1, col - GROUP, Ts, B)]); }
			}
			const int src = wj - j;                       // lane holding column j
			const int k = (int)gl - src;                  // this lane holds column j - k
			const uint64_t pvj = GROUP > 1 ? __shfl_sync(mask(), wpv, src, GROUP) : wpv;
			int run; uint8_t op;
			if ((pvj >> bit) & 1) {
				// up: as long as the vertical delta stays +1 inside this block and column
				run = min(__clzll((long long)~(pvj << (63 - bit))), bit + 1);
				op = 1; I -= run;
			} else {
				const bool in = k >= 0 && j - k >= 0;
				const bool left_k = in && !((wpv >> bit) & 1) && ((wph >> bit) & 1);
				const uint32_t bl = ((__ballot_sync(mask(), left_k) >> gbase) & gall) >> src;
				if (bl & 1) { run = ~bl ? __ffs((int)~bl) - 1 : 32; op = 2; J -= run; }     // left along row i
				else {
					const bool diag_k = in && bit - k >= 0 && !((wpv >> (bit - k)) & 1) && !((wph >> (bit - k)) & 1);
					const uint32_t bd = ((__ballot_sync(mask(), diag_k) >> gbase) & gall) >> src;
					run = ~bd ? __ffs((int)~bd) - 1 : 32; op = 0; I -= run; J -= run;      // run >= 1: the cell itself is neither up nor left
				}
			}
			for (int x = (int)gl; x < run; x += GROUP) tmp[n + x] = op;
			n += run;
			CLB_PH_COUNT(GROUP, 5, 1)
		}
		// a border was reached: the rest is all up or all left
		const int rest = I + J; const uint8_t rop = I > 0 ? 1 : 2;
		for (int x = (int)gl; x < rest; x += GROUP) tmp[n + x] = rop;
		n += rest;
		gsync();
		for (int x = (int)gl; x < n; x += GROUP) out[x] = tmp[n - 1 - x];
		gsync();
		CLB_PH_END(GROUP, 1)
		return n;
	}

	// The same walk for the groups of a warp IN LOCKSTEP (sub-warp groups, backward kernel): one loop for the whole warp whose
	// iterations are one run of every group that is still walking, shuffles and ballots with the full mask (width GROUP).  Groups
	// that walk on their own take turns on the warp's issue slots and pay the convergence checks of a run-time mask around every
	// shuffle; in lockstep a step costs the same whatever the number of groups.  go == false: this group has nothing to walk.
	__device__ int traceback_lockstep(const ulonglong2* __restrict__ hist, int Ts, int Q, int T, bool go, uint8_t* out, uint8_t* tmp) const
	{
		CLB_PH_BEGIN
		constexpr unsigned FULL = 0xffffffffu;
		const int B = (Q + 63) >> 6;
		int I = go ? Q : 0, J = go ? T : 0, n = 0;
		int wb = -1, wj = -0x40000000;
		uint64_t wpv = 0, wph = 0;
		const uint32_t gbase = (threadIdx.x & 31) & ~(uint32_t)(GROUP - 1);
		const uint32_t gall = GROUP == 32 ? 0xffffffffu : ((1u << GROUP) - 1u);
		for (;;) {
			const bool act = I > 0 && J > 0;
			if (!__any_sync(FULL, act)) break;
			const int i = I - 1, j = J - 1, b = i >> 6, bit = i & 63;
			if (act && (b != wb || j > wj || j < wj - (GROUP - 1))) {
				wb = b; wj = j;
				CLB_PH_COUNT(GROUP, 6, 1)
				const int col = wj - (int)gl;
				if (col >= 0) { const ulonglong2 e = hist[hist_at(b, col, Ts, B)]; wpv = e.x; wph = e.y; }
				if (col - GROUP >= 0) prefetch_l2(&hist[hist_at(b, col - GROUP, Ts, B)]);
				if (b > 0 && col >= 0) { prefetch_l2(&hist[hist_at(b - 1, col, Ts, B)]); if (col - GROUP >= 0) prefetch_l2(&hist[hist_at(b - 1, col - GROUP, Ts, B)]); }
			}
			const int src = act ? wj - j : 0;             // lane holding column j
			const int k = (int)gl - src;                  // this lane holds column j - k
			const uint64_t pvj = __shfl_sync(FULL, wpv, src, GROUP);
			const bool in = act && k >= 0 && j - k >= 0;
			const bool left_k = in && !((wpv >> bit) & 1) && ((wph >> bit) & 1);
			const uint32_t bl = ((__ballot_sync(FULL, left_k) >> gbase) & gall) >> src;
			const int bk = (bit - k) & 63;
			const bool diag_k = in && bit - k >= 0 && !((wpv >> bk) & 1) && !((wph >> bk) & 1);
			const uint32_t bd = ((__ballot_sync(FULL, diag_k) >> gbase) & gall) >> src;
			if (act) {
				int run; uint8_t op;
				if ((pvj >> bit) & 1) { run = min(__clzll((long long)~(pvj << (63 - bit))), bit + 1); op = 1; I -= run; }
				else if (bl & 1) { run = ~bl ? __ffs((int)~bl) - 1 : 32; op = 2; J -= run; }
				else { run = ~bd ? __ffs((int)~bd) - 1 : 32; op = 0; I -= run; J -= run; }
				for (int x = (int)gl; x < run; x += GROUP) tmp[n + x] = op;
				n += run;
				CLB_PH_COUNT(GROUP, 5, 1)
			}
		}
		const int rest = I + J; const uint8_t rop = I > 0 ? 1 : 2;
		for (int x = (int)gl; x < rest; x += GROUP) tmp[n + x] = rop;
		n += rest;
		__syncwarp(FULL);
		for (int x = (int)gl; x < n; x += GROUP) out[x] = tmp[n - 1 - x];
		__syncwarp(FULL);
		CLB_PH_END(GROUP, 1)
		return n;
	}

	// decode the final column of a sweep into D(y-1, T-1) for y = 1..Q, F[0] = T  (vertex values of the last column)
	__device__ void decode_final(const uint64_t* fin_pv, const uint64_t* fin_mv, const int32_t* fin_sc, int Q, int T, uint32_t* F) const
	{
		const int B = (Q + 63) >> 6;
		for (int b = (int)gl; b < B; b += GROUP) {
			const int r0 = b << 6, r1 = min(Q, r0 + 64);
			int v = fin_sc[b];
			const uint64_t pv = fin_pv[b], mv = fin_mv[b];
			for (int r = r1 - 1; r >= r0; --r) {
				F[r + 1] = (uint32_t)v;
				v -= (int)((pv >> (r - r0)) & 1) - (int)((mv >> (r - r0)) & 1);
			}
		}
		if (gl == 0) F[0] = (uint32_t)T;
		gsync();
	}

	// edlib's obtainAlignment for rows x cols with known optimal score `best`: ops appended to ops_out (forward order).
	// Iterative Hirschberg with an explicit stack; all lanes of the group call.  Returns the number of ops.
	template <class V>
	__device__ int path(V rows, int Q, V cols, int T, int best, uint8_t* ops_out) const
	{
		uint8_t* tmp = scratch + lay.tmp;
		int n_ops = 0;
		if (edlib_column_bytes(Q, T) < EDLIB_TRACEBACK_LIMIT || Q == 0 || T == 0) {
			n_ops = leaf(rows, Q, cols, T, ops_out, tmp);
			return n_ops;
		}
		int32_t* stack = reinterpret_cast<int32_t*>(scratch + lay.stack);
		uint32_t* F = reinterpret_cast<uint32_t*>(scratch + lay.F);
		uint32_t* R = reinterpret_cast<uint32_t*>(scratch + lay.R);
		uint64_t* fin_pv = reinterpret_cast<uint64_t*>(scratch + lay.fin);
		int sp = 0;
		if (gl == 0) { stack[0] = 0; stack[1] = Q; stack[2] = 0; stack[3] = T; stack[4] = best; }
		sp = 1;
		gsync();
		while (sp > 0) {
			--sp;
			const int qo = stack[sp * 5], ql = stack[sp * 5 + 1], to = stack[sp * 5 + 2], tl = stack[sp * 5 + 3], bs = stack[sp * 5 + 4];
			gsync();
			const V r = rows.sub(qo), c = cols.sub(to);
			if (ql == 0 || tl == 0 || edlib_column_bytes(ql, tl) < EDLIB_TRACEBACK_LIMIT) {
#ifdef CLB_ALIGN_TIMING
				long long tl0 = clock64();
#endif
				n_ops += leaf(r, ql, c, tl, ops_out + n_ops, tmp);
#ifdef CLB_ALIGN_TIMING
				if (gl == 0 && Q > 30000) printf("[leaf] ql %d tl %d: %lld cycles\n", ql, tl, clock64() - tl0);
#endif
				continue;
			}
#ifdef CLB_ALIGN_TIMING
			long long th0 = clock64();
#endif
			const int Bq = (ql + 63) >> 6;
			uint64_t* fpv = fin_pv; uint64_t* fmv = fin_pv + Bq; int32_t* fsc = reinterpret_cast<int32_t*>(fin_pv + 2 * Bq);
			const int lw = tl / 2, rw = tl - lw;
			sweep<false, true>(r, ql, c, lw, nullptr, fpv, fmv, fsc);
			gsync();
			decode_final(fpv, fmv, fsc, ql, lw, F);
			sweep<false, true>(r.reversed(ql), ql, c.sub(lw).reversed(rw), rw, nullptr, fpv, fmv, fsc);
			gsync();
			decode_final(fpv, fmv, fsc, ql, rw, R);           // R[z] = dist(last z rows, right half)
			// topmost y in 1..ql-1 with F[y] + R[ql-y] == bs, then y = 0, then y = ql   (edlib.cpp:1305-1338)
			int y = 0x7fffffff;
			for (int cand = 1 + (int)gl; cand <= ql - 1; cand += GROUP) if ((int)(F[cand] + R[ql - cand]) == bs) { y = cand; break; }
			if (GROUP > 1) for (int d = GROUP / 2; d; d >>= 1) y = min(y, __shfl_xor_sync(mask(), y, d, GROUP));
			if (y == 0x7fffffff) { if ((int)(lw + R[ql]) == bs) y = 0; else y = ql; }
			const int ls = y == 0 ? lw : (int)F[y], rs = y == ql ? rw : (int)R[ql - y];
#ifdef CLB_ALIGN_TIMING
			if (gl == 0 && Q > 30000) printf("[hirsch] ql %d tl %d y %d: %lld cycles\n", ql, tl, y, clock64() - th0);
#endif
			gsync();
			if (gl == 0) {
				int32_t* e = stack + sp * 5;          // right part first so that the left part is processed first
				e[0] = qo + y; e[1] = ql - y; e[2] = to + lw; e[3] = rw; e[4] = rs;
				e[5] = qo; e[6] = y; e[7] = to; e[8] = lw; e[9] = ls;
			}
			sp += 2;
			gsync();
		}
		return n_ops;
	}

	// the two halves of leaf() for the split kernels (Q, T > 0): sweep with history / traceback on it
	template <class V>
	__device__ void leaf_fwd(V rows, int Q, V cols, int T) const { sweep<true, false>(rows, Q, cols, T, reinterpret_cast<ulonglong2*>(scratch + lay.hist), nullptr, nullptr, nullptr); }
	__device__ int leaf_back(int Q, int T, uint8_t* out, uint8_t* tmp) const { return traceback(reinterpret_cast<const ulonglong2*>(scratch + lay.hist), T, Q, T, out, tmp); }

	// stored-column traceback of a problem below edlib's 1 MiB limit
	template <class V>
	__device__ int leaf(V rows, int Q, V cols, int T, uint8_t* out, uint8_t* tmp) const
	{
		int n = 0;
		if (Q == 0 || T == 0) {
			if (gl == 0) for (int i = 0; i < Q + T; ++i) out[i] = Q == 0 ? 2 : 1;
			gsync();
			return Q + T;
		}
		const int B = (Q + 63) >> 6;
		ulonglong2* hist = reinterpret_cast<ulonglong2*>(scratch + lay.hist);
		sweep<true, false>(rows, Q, cols, T, hist, nullptr, nullptr, nullptr);
		gsync();
		n = traceback(hist, T, Q, T, out, tmp);
		return n;
	}
};

// ---- script symbols ----------------------------------------------------------------------------------
__device__ __forceinline__ char mismatch_symb(uint8_t ref, uint8_t enc)      // utils.h:341-352
{
	// the three other bases in ACGT order are X, Y, Z
	const int idx = enc - (enc > ref ? 1 : 0);
	return (char)('X' + idx);
}
__device__ __forceinline__ bool es_is_mm(char c) { return c == 'X' || c == 'Y' || c == 'Z'; }
__device__ __forceinline__ bool es_is_ins(char c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }

__device__ inline void fix_in_range(char* es, uint32_t start, uint32_t end)   // edit_script.h:432-447
{
	if (end < start + 2) return;
	--end;
	for (;;) {
		while (start < end && es[start] == 'M') ++start;
		while (start < end && es[end] != 'M') --end;
		if (start == end) break;
		const char t = es[start]; es[start] = es[end]; es[end] = t;
	}
}
// edit_script.h:591-671; ref/enc may be indexed one past the part (the byte that follows it in the read)
// The two passes on es[lo..hi) with the reference / read positions of symbol lo given.
template <class V>
__device__ inline void refactor_range(V ref, V enc, char* es, uint32_t lo, uint32_t hi, uint32_t ref_at, uint32_t enc_at)
{
	uint32_t ref_start = ref_at, ref_pos = ref_at, es_start = lo;
	for (uint32_t p = lo; p < hi; ++p) {
		const char s = es[p]; const bool mm = es_is_mm(s), ins = es_is_ins(s);
		if (ins || mm || ref[ref_start] != ref[ref_pos]) {
			fix_in_range(es, es_start, p);
			es_start = p; if (ins || mm) ++es_start;
			ref_start = ref_pos;
		}
		if (!ins) ++ref_pos;
	}
	fix_in_range(es, es_start, hi);
	uint32_t enc_start = enc_at, enc_pos = enc_at; es_start = lo;
	for (uint32_t p = lo; p < hi; ++p) {
		const char s = es[p]; const bool mm = es_is_mm(s), del = s == 'D';
		if (del || mm || enc[enc_start] != enc[enc_pos]) {
			fix_in_range(es, es_start, p);
			es_start = p; if (del || mm) ++es_start;
			enc_start = enc_pos;
		}
		if (!del) ++enc_pos;
	}
	fix_in_range(es, es_start, hi);
}
// edit_script.h:591-671; ref/enc may be indexed one past the part (the byte that follows it in the read)
template <class V>
__device__ inline void refactor_edit_script(V ref, V enc, char* es, uint32_t n) { refactor_range(ref, enc, es, 0, n, 0, 0); }

// ops -> script symbols -> canonical form, by all lanes of the group.
// ops[0..n) in script order: 0 = diagonal, 1 = row symbol only, 2 = column symbol only; rows are the reference iff rows_ref.
// The canonicalisation is two passes that only reorder symbols inside ranges: pass 1 over M / D symbols on one repeated
// reference base (insertions and substitutions end a range), pass 2 over M / insertion symbols on one repeated read base
// (deletions and substitutions end a range).  At some positions both passes start a new range whatever came before — next to
// a substitution; between two reference-consuming symbols on different reference bases (pass 1 leaves either a 'D' before
// the position or matches of different bases around it); between insertions of different bases; between a deletion and an
// insertion.  Lanes cut the script at such positions near the n/GROUP boundaries and each runs the reference's serial
// algorithm on its own piece, which gives the same script as one serial run.
template <int GROUP, class V>
__device__ void finish_script(const Aligner<GROUP>& A, const uint8_t* ops, uint32_t n, V ref, V enc, bool rows_ref, char* out)
{
	const uint32_t gl = A.gl;
	CLB_PH_BEGIN
	const uint32_t L = (n + GROUP - 1) / GROUP;
	const uint32_t s = min(n, gl * L), e = min(n, s + L);
	uint32_t cr = 0, ce = 0;
	for (uint32_t i = s; i < e; ++i) { const uint8_t o = ops[i]; cr += (o == 0) | ((o == 1) == rows_ref); ce += (o == 0) | ((o == 1) != rows_ref); }
	uint32_t pr = cr, pe = ce;
	if (GROUP > 1) {
#pragma unroll
		for (int d = 1; d < GROUP; d <<= 1) {
			const uint32_t a = __shfl_up_sync(A.mask(), pr, d, GROUP), b = __shfl_up_sync(A.mask(), pe, d, GROUP);
			if ((int)gl >= d) { pr += a; pe += b; }
		}
	}
	pr -= cr; pe -= ce;                        // reference / read symbols consumed before position s
	{
		uint32_t r = pr, q = pe;
		for (uint32_t i = s; i < e; ++i) {
			const uint8_t o = ops[i];
			if (o == 0) { const uint8_t a = ref[(int)r], b = enc[(int)q]; out[i] = a == b ? 'M' : mismatch_symb(a, b); ++r; ++q; }
			else if ((o == 1) == rows_ref) { out[i] = 'D'; ++r; }
			else { out[i] = "ACGT"[enc[(int)q] & 3]; ++q; }
		}
	}
	if (GROUP == 1) { refactor_range(ref, enc, out, 0, n, 0, 0); CLB_PH_END(GROUP, 2) return; }
	A.gsync();
	uint32_t cut = 0, cut_r = 0, cut_q = 0;
	if (gl > 0) {
		// first position p >= s where both passes start a new range whatever came before (a = symbol p-1, b = symbol p):
		//   a or b is a substitution; a, b both consume the reference (M / D) and the reference base changes;
		//   a, b are insertions of different bases; a is a deletion and b an insertion.
		cut = n;
		uint32_t r = pr, q = pe;
		for (uint32_t p = s; p < n; ++p) {
			const char b = out[p];
			if (p > 0) {
				const char a = out[p - 1];
				const bool a_rc = a == 'M' || a == 'D', b_rc = b == 'M' || b == 'D';
				bool safe = es_is_mm(a) || es_is_mm(b);
				if (!safe && a_rc && b_rc) safe = ref[(int)r - 1] != ref[(int)r];
				if (!safe && es_is_ins(a) && es_is_ins(b)) safe = a != b;
				if (!safe && a == 'D' && es_is_ins(b)) safe = true;
				if (safe) { cut = p; cut_r = r; cut_q = q; break; }
			}
			if (b == 'D') ++r; else if (es_is_ins(b)) ++q; else { ++r; ++q; }
		}
	}
	uint32_t next = __shfl_down_sync(A.mask(), cut, 1, GROUP);
	if (gl == GROUP - 1) next = n;
	A.gsync();                                 // every cut is known before any lane reorders symbols
	if (cut < next) refactor_range(ref, enc, out, cut, next, cut_r, cut_q);
	A.gsync();
	CLB_PH_END(GROUP, 2)
}

// One edit-script task = CEncoder::GetEditDist (encoder.cpp:1255-1283).
// ref/enc: views of the parts inside their reads (index rl / el = the byte after the part, 255 at a read's end).
// kind: 0 left flank, 1 right flank, 2 between anchors.  Writes the script to out (capacity rl + el + 2); returns its length.
// Scratch must be laid out for rows/cols = (max(rl,el), max(rl,el)) — see align_task_dims().
__host__ __device__ inline void align_task_dims(uint32_t rl, uint32_t el, uint32_t kind, long long* q, long long* t)
{
	if (rl == 0 || el == 0) { *q = 1; *t = 1; return; }
	if (kind == 2) { *q = rl; *t = el; return; }
	const uint32_t cut = rl < 2 * el ? rl : 2 * el;
	if (cut < 2 || el < 2) { *q = cut; *t = el; }          // global fallback, rows = ref
	else { *q = el; *t = cut; }                              // SHW: rows = enc, cols = ref prefix
}

// lead_out != nullptr: the 'D' run that precedes a left flank's script (the reference symbols left of the aligned window)
// is not written; its length is returned through *lead_out instead and the returned length excludes it.
// PHASE 0: the whole task.  PHASE 1 / 2: the task in two kernels (k_align_fwd / k_align_back) for problems below edlib's traceback
// limit — 1 runs the sweep and leaves its history (and the last-row minimum of a flank) in the scratch, 2 walks back through it
// and writes the script; the host sends problems above the limit (Hirschberg, sweeps and walks interleaved) to PHASE 0.
template <int GROUP, int PHASE = 0, class V>
__device__ uint32_t edit_script_task(const Aligner<GROUP>& A, V ref, uint32_t rl, V enc, uint32_t el, uint32_t kind, char* out, uint32_t* lead_out = nullptr)
{
	if (lead_out) *lead_out = 0;
	const uint32_t gl = A.gl;
	int* const res = reinterpret_cast<int*>(A.scratch + A.lay.res);
	if (rl == 0 || el == 0) {      // edit_script.h:247-266
		if (gl == 0) {
			if (rl == 0) for (uint32_t i = 0; i < el; ++i) out[i] = "ACGT"[enc[i]];
			else for (uint32_t i = 0; i < rl; ++i) out[i] = 'D';
		}
		A.gsync();
		return rl == 0 ? el : rl;
	}
	uint8_t* ops = A.scratch + A.lay.ops;
	int n_ops = 0;
	uint32_t n_out = 0;
	uint32_t ref_end = 0;
	bool rows_ref = true;
	constexpr bool LOCKSTEP = PHASE == 2 && GROUP > 1 && GROUP < 32;
	if constexpr (LOCKSTEP) {
		// backward kernel, sub-warp groups: what the forward kernel swept is derived per group, then every group of the warp walks
		// back in one lockstep loop (problems above the traceback limit never come here: the host gives them bins of their own)
		const ulonglong2* hist = reinterpret_cast<const ulonglong2*>(A.scratch + A.lay.hist);
		int Q = (int)rl, Ts = (int)el, T = (int)el;
		bool go = true;
		if (kind != 2) {
			const uint32_t cut = rl < 2 * el ? rl : 2 * el;
			if (cut < 2 || el < 2) { Q = (int)cut; ref_end = cut - 1; }
			else {
				rows_ref = false;
				Q = (int)el; Ts = (int)cut;
				const int best = res[0], end = res[1];
				ref_end = (uint32_t)end; T = end + 1;
				if (best >= (int)el) { go = false; ref_end = 0xFFFFFFFFu; }      // the empty prefix wins: |enc| insertions (see below)
			}
		}
		n_ops = A.traceback_lockstep(hist, Ts, Q, T, go, ops, A.scratch + A.lay.tmp);
		if (!go) { for (int x = (int)gl; x < (int)el; x += GROUP) ops[x] = 1; n_ops = (int)el; A.gsync(); }
	} else
	if (kind == 2) {
		// NW, rows = ref, cols = enc: UP = 'D', LEFT = insertion
		int best;
		const bool small = edlib_column_bytes(rl, el) < EDLIB_TRACEBACK_LIMIT;
		if (PHASE == 1) { if (small) A.leaf_fwd(ref, (int)rl, enc, (int)el); return 0; }
		if (small && PHASE == 2) n_ops = A.leaf_back((int)rl, (int)el, ops, A.scratch + A.lay.tmp);
		else if (small) n_ops = A.leaf(ref, (int)rl, enc, (int)el, ops, A.scratch + A.lay.tmp);
		else {
			best = A.template sweep<false, true>(ref, (int)rl, enc, (int)el, nullptr, nullptr, nullptr, nullptr).score;
			A.gsync();
			n_ops = A.path(ref, (int)rl, enc, (int)el, best, ops);
		}
		finish_script<GROUP>(A, ops, (uint32_t)n_ops, ref, enc, true, out);
		return (uint32_t)n_ops;
	}
	// flanks: SHW of enc against a prefix of ref limited to 2*|enc| symbols; the left flank works on reversed strings
	const uint32_t cut = rl < 2 * el ? rl : 2 * el;
	const V r = kind == 0 ? ref.reversed((int)rl) : ref;      // first `cut` symbols are used
	const V e = kind == 0 ? enc.reversed((int)el) : enc;
	if (LOCKSTEP) {}
	else if (cut < 2 || el < 2) {       // edit_script.h:336-343: global alignment of the (cut) ref against enc, rows = ref
		rows_ref = true; ref_end = cut - 1;
		if (PHASE == 1) { A.leaf_fwd(r, (int)cut, e, (int)el); return 0; }
		if (PHASE == 2) n_ops = A.leaf_back((int)cut, (int)el, ops, A.scratch + A.lay.tmp);
		else n_ops = A.leaf(r, (int)cut, e, (int)el, ops, A.scratch + A.lay.tmp);
	} else {
		rows_ref = false;
#ifdef CLB_ALIGN_TIMING
		long long tc0 = clock64();
#endif
		const bool small = edlib_column_bytes(el, cut) < EDLIB_TRACEBACK_LIMIT;
		ulonglong2* hist = reinterpret_cast<ulonglong2*>(A.scratch + A.lay.hist);
		// leftmost column with the minimal last-row score (edlib.cpp:660-674): followed by the lane of the last block during the sweep
		if (!small && PHASE == 1) return 0;
		typename Aligner<GROUP>::SweepResult sw{0, 0, 0};
		if (small && PHASE == 2) { sw.best = res[0]; sw.end = res[1]; }
		else sw = small ? A.template sweep<true, true>(e, (int)el, r, (int)cut, hist, nullptr, nullptr, nullptr)
			: A.template sweep<false, true>(e, (int)el, r, (int)cut, nullptr, nullptr, nullptr, nullptr);
		if (PHASE == 1) { if (gl == 0) { res[0] = sw.best; res[1] = sw.end; } return 0; }
		A.gsync();
		const int best = sw.best, end = sw.end;
		ref_end = (uint32_t)end;
		const int T = end + 1;
#ifdef CLB_ALIGN_TIMING
		long long tc1 = clock64();
#endif
		// edlib pads the query to whole words with wildcards and reads the score of prefix c - W in column c, so the EMPTY prefix
		// (position -1, score |enc|) is a candidate too and comes first: if no prefix beats inserting the whole part,
		// endLocations[0] = -1, the path is |enc| insertions and ref_end wraps around (edlib.cpp:660-694, edit_script.h:352)
		if (best >= (int)el) {
			ref_end = 0xFFFFFFFFu;
			for (int x = (int)gl; x < (int)el; x += GROUP) ops[x] = 1;
			n_ops = (int)el;
			A.gsync();
		}
		else if (small) n_ops = A.traceback(hist, (int)cut, (int)el, T, ops, A.scratch + A.lay.tmp);
		else n_ops = A.path(e, (int)el, r, T, best, ops);
#ifdef CLB_ALIGN_TIMING
		if (gl == 0 && el > 30000) printf("[task] el %u cut %u small %d: sweep %lld path %lld cycles, n_ops %d\n", el, cut, (int)small, tc1 - tc0, clock64() - tc1, n_ops);
#endif
	}
	// the left flank was aligned on reversed strings: its script (and the order in which symbols are consumed) is the reverse
	uint32_t lead = 0;
	char* w = out;
	if (kind == 0) {
		lead = (rl - 1) - ref_end;
		for (int x = (int)gl; x < n_ops / 2; x += GROUP) { const uint8_t t = ops[x]; ops[x] = ops[n_ops - 1 - x]; ops[n_ops - 1 - x] = t; }
		if (lead_out) { *lead_out = lead; }
		else { for (uint32_t i = gl; i < lead; i += GROUP) out[i] = 'D'; w = out + lead; }
		A.gsync();
	}
	finish_script<GROUP>(A, ops, (uint32_t)n_ops, kind == 0 ? ref.sub((int)lead) : ref, enc, rows_ref, w);
	return (lead_out ? 0 : lead) + (uint32_t)n_ops;
}

} // namespace clb
