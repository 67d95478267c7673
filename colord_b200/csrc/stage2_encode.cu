// stage2_encode.cu — rows E6-E9 of SURVEY.md §8 on device, and the driver of the whole stage 2.
//
// The reference encodes read by read, recursing into "alternative" candidates for parts the main candidate covers badly
// (encoder.cpp:1445-1575).  Here the recursion is turned into level waves over a batch of reads:
//   level L:  k_task_count/k_task_fill  every new node (one AddEncodedReadWithCandidates call) lists its even fragments
//             k_task_classify/scatter   parts are binned by lane-group class and width
//             k_align<GROUP>            edit scripts of all parts of a bin (align.cuh)
//             k_decide                  parts >= minPartLenToConsiderAltRead: stateless entropy test (encoder.cpp:1315-1327),
//                                       losers get a child node on the next candidate (AdjustAnchors as a view, :778-868);
//                                       shorter parts only prepare their CEntropyEstimator statistics
//   then      k_estimate                one thread per read pack replays the adaptive estimator (utils.h:1060-1126) in the
//                                       reference's order — the only serial dependency of stage 2 — and settles short parts
//             k_emit_size/k_emit_write  one thread per read walks its node tree and run-length codes the big edit script into
//                                       CompactES bytes (encoder.cpp:1348-1443, utils.h:69-273)
// Double arithmetic follows the reference's operation order with explicit round-to-nearest adds/multiplies (no FMA
// contraction); log2 is CUDA's (<= 1 ulp) where the reference uses libm's.
#include "ctx.h"
#include "stage2.h"
#include "align.cuh"
#include "align_back.cuh"
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <cstdio>
#include <ctime>
#include <vector>

namespace clb {

clb_status s2_anchors(clb_ctx* c, const S2P& P, const std::vector<uint32_t>& h_list, const uint32_t* d_list,
	const uint32_t* d_ref_to_read, DevBuf<uint8_t>& arena, SegInfo* d_seg, uint32_t* d_slot_dec, Node* d_nodes, CandView* d_cviews,
	unsigned long long* d_cursor);

struct ReadStore { const uint64_t* pk; const uint64_t* rd_start; const uint32_t* rd_len; const uint32_t* nmask; const uint32_t* ref_to_read; };

CLB_D PackedView enc_view(const ReadStore& R, uint32_t read, uint32_t at)
{
	const long long s = (long long)R.rd_start[read];
	return PackedView{R.pk, s + at, 1, 0u, s, s + (long long)R.rd_len[read]};
}
// element i of the view = symbol `at + i` of the oriented reference read (reverse-complement if rev)
CLB_D PackedView ref_view(const ReadStore& R, uint32_t ref_id, uint32_t rev, uint32_t at)
{
	const uint32_t rr = R.ref_to_read[ref_id];
	const long long s = (long long)R.rd_start[rr], l = (long long)R.rd_len[rr];
	if (!rev) return PackedView{R.pk, s + at, 1, 0u, s, s + l};
	return PackedView{R.pk, s + l - 1 - at, -1, 3u, s, s + l};
}

CLB_HD uint32_t es_cap(uint32_t rl, uint32_t el, uint32_t kind)
{
	const uint64_t base = el == 0 ? 0 : rl == 0 ? el : kind == 2 ? (uint64_t)rl + el : (uint64_t)(rl < 2 * (uint64_t)el ? rl : 2 * el) + el;
	return (uint32_t)((base + base / 8 + 8 + 3) & ~3ull);
}

// ------------------------------------------------------------------------------------------------ tasks
// fragments of a node: part i lies between anchor i-1 and anchor i (encoder.cpp:1541-1572)
template <bool FILL>
__global__ void __launch_bounds__(128) k_tasks(const Node* nodes_in, Node* nodes, const CandView* __restrict__ cviews, uint32_t n0, uint32_t n1, uint32_t c,
	ReadStore R, const uint8_t* __restrict__ arena, uint32_t* __restrict__ cnt, uint32_t* __restrict__ capu,
	const uint64_t* __restrict__ task_off, const uint64_t* __restrict__ cap_off, uint64_t task_base, uint64_t es_base, Task* __restrict__ tasks)
{
	const uint32_t id = n0 + blockIdx.x * blockDim.x + threadIdx.x;
	if (id >= n1) return;
	const Node N = nodes_in[id];
	if (!N.valid) { if (!FILL) { cnt[id - n0] = 0; capu[id - n0] = 0; } return; }
	const CandView V = cviews[(size_t)id * c + N.level];
	const uint32_t RL = R.rd_len[R.ref_to_read[V.ref_id]];
	uint32_t cur_ref = 0, cur_enc = 0;
	uint64_t cap_sum = 0;
	uint64_t t = 0, eo = 0;
	if (FILL) { t = task_base + task_off[id - n0]; eo = es_base + cap_off[id - n0] * 4; nodes[id].first_task = (uint32_t)t; }
	for (uint32_t i = 0; i <= N.n_anch; ++i) {
		const bool last = i == N.n_anch;
		Anchor a{0, 0, 0};
		if (!last) a = cv_get(arena, V, i);
		const uint32_t end_enc = last ? N.enc_len : a.pos_enc, end_ref = last ? RL : a.pos_ref;
		const uint32_t el = end_enc - cur_enc, rl = end_ref - cur_ref;
		const uint32_t kind = i == 0 ? 0 : (last ? 1 : 2);
		const uint32_t cap = es_cap(rl, el, kind);
		if (FILL) {
			Task T{};
			T.node = id; T.frag = i; T.enc_start = N.enc_start + cur_enc; T.el = el; T.ref_start = cur_ref; T.rl = rl; T.kind = kind;
			T.decision = D_PENDING; T.es_off = eo; T.child = 0xFFFFFFFFu;
			tasks[t + i] = T;
			eo += cap;
		} else cap_sum += cap;
		if (!last) { cur_ref = a.pos_ref + a.len; cur_enc = a.pos_enc + a.len; }
	}
	if (!FILL) { cnt[id - n0] = N.n_anch + 1; capu[id - n0] = (uint32_t)(cap_sum / 4); }
}

// bins: lane-group class (1/2/4/8/16 lanes: rows within a factor of two) x log2(columns); the 32-lane class, whose row count is
// open-ended, is split by log2(rows) as well so that the parts of one launch cost about the same
// problems above edlib's traceback limit (Hirschberg: sweeps and walks interleaved) have bins of their own after those: they run in
// the one-kernel form on whole warps, everything else as a forward and a backward kernel
// Bins are half octaves of the column count and (above two blocks) of the row count: the scratch slot of a bin's wave is sized by its
// largest task, so narrow bins put more tasks into the same scratch (the backward kernel wants many: it is one thread per task).
constexpr int N_TBIN = 40, N_QBIN = 16, N_SMALL = 5, N_SMALL_BINS = 2 * N_SMALL * N_TBIN, N_SPLIT_BINS = N_SMALL_BINS + N_TBIN * N_QBIN, N_BINS = N_SPLIT_BINS + N_TBIN * N_QBIN;
constexpr int N_ALIGN_STREAMS = 16;
struct BinStats { unsigned int cnt[N_BINS], maxq[N_BINS], maxt[N_BINS], fill[N_BINS], base[N_BINS]; unsigned long long sumq[N_BINS], sumt[N_BINS], sumqt[N_BINS]; };

CLB_HD int gclass_of(long long q) { const long long B = (q + 63) / 64; return B <= 1 ? 0 : B <= 2 ? 1 : B <= 4 ? 2 : B <= 8 ? 3 : B <= 16 ? 4 : 5; }
CLB_HD int ilog2_u32(uint32_t x) { int r = 0; while (x >>= 1) ++r; return r; }
CLB_HD int hlog2_u32(uint32_t x) { const int l = ilog2_u32(x); return 2 * l + (l > 0 ? (int)((x >> (l - 1)) & 1u) : 0); }      // half octaves

// Parts with an empty side need no alignment (edit_script.h:247-266): el == 0 -> 'D' x rl (kept as `lead`), rl == 0 -> the
// part's bases as insertions.  Everything else is binned.
__global__ void __launch_bounds__(256) k_task_classify(Task* __restrict__ tasks, uint64_t t0, uint64_t t1, ReadStore R, const Node* __restrict__ nodes,
	char* __restrict__ esbuf, BinStats* __restrict__ bins, uint32_t* __restrict__ bin_of, bool prof)
{
	const uint64_t t = t0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= t1) return;
	Task& T = tasks[t];
	if (T.el == 0) { T.lead = T.rl; T.es_len = 0; bin_of[t - t0] = 0xFFFFFFFFu; return; }
	if (T.rl == 0) {
		const PackedView e = enc_view(R, nodes[T.node].read, T.enc_start);
		char* o = esbuf + T.es_off;
		for (uint32_t i = 0; i < T.el; ++i) o[i] = "ACGT"[e[(int)i] & 3];
		T.lead = 0; T.es_len = T.el; bin_of[t - t0] = 0xFFFFFFFFu; return;
	}
	long long q, tt;
	align_task_dims(T.rl, T.el, T.kind, &q, &tt);
	const int gc = gclass_of(q), tb = min(N_TBIN - 1, hlog2_u32((uint32_t)tt));
	const long long Bq = (q + 63) / 64;
	const int qs = gc >= 2 && Bq > (3ll << (gc - 2)) ? 1 : 0;              // upper half of the class's block range
	const bool big = edlib_column_bytes(q, tt) >= EDLIB_TRACEBACK_LIMIT;
	const int qb = min(N_QBIN - 1, max(0, hlog2_u32((uint32_t)q) - 20)) * N_TBIN + tb;
	const int b = big ? N_SPLIT_BINS + qb : gc < N_SMALL ? (2 * gc + qs) * N_TBIN + tb : N_SMALL_BINS + qb;
	bin_of[t - t0] = (uint32_t)b;
	atomicAdd(&bins->cnt[b], 1u);
	atomicMax(&bins->maxq[b], (unsigned int)q);
	atomicMax(&bins->maxt[b], (unsigned int)tt);
	if (prof) { atomicAdd(&bins->sumq[b], (unsigned long long)q); atomicAdd(&bins->sumt[b], (unsigned long long)tt); atomicAdd(&bins->sumqt[b], (unsigned long long)(q * tt)); }
}
__global__ void __launch_bounds__(256) k_task_scatter(uint64_t t0, uint64_t t1, const uint32_t* __restrict__ bin_of, BinStats* __restrict__ bins, uint32_t* __restrict__ list)
{
	const uint64_t t = t0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= t1) return;
	const uint32_t b = bin_of[t - t0];
	if (b == 0xFFFFFFFFu) return;
	list[bins->base[b] + atomicAdd(&bins->fill[b], 1u)] = (uint32_t)(t - t0);
}

// PHASE 0: whole tasks; 1: forward sweep with history; 2: traceback + script (see edit_script_task).  The backward kernel is
// bound by the latency of its walk through the history (DRAM), so it is compiled for twice the resident warps.
template <int GROUP, int PHASE>
__global__ void __launch_bounds__(ALIGN_THREADS, PHASE == 2 ? 6 : PHASE == 1 ? 5 : 4) k_align(Task* __restrict__ tasks, uint64_t t0, const uint32_t* __restrict__ list, uint32_t n_list, uint64_t stride,
	uint8_t* __restrict__ scratch, ReadStore R, const Node* __restrict__ nodes, const CandView* __restrict__ cviews, uint32_t c, char* __restrict__ esbuf)
{
	const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t slot = tid / GROUP;
	if (slot >= n_list) return;                 // whole groups leave together
	Task& T = tasks[t0 + list[slot]];
	const Node N = nodes[T.node];
	const CandView& V = cviews[(size_t)T.node * c + N.level];
	__shared__ uint64_t s_peq[4 * ALIGN_THREADS];
	Aligner<GROUP> A;
	A.peq = s_peq + threadIdx.x;
	A.gl = threadIdx.x & (GROUP - 1);
#ifdef CLB_ALIGN_PHASES
	const uint32_t gl = A.gl;
#endif
	CLB_PH_BEGIN
	const uint32_t lane = threadIdx.x & 31;
	A.gmask = GROUP == 32 ? 0xffffffffu : (((1u << GROUP) - 1u) << (lane & ~(uint32_t)(GROUP - 1)));
	A.scratch = scratch + (uint64_t)slot * stride;
	long long q, tt;
	align_task_dims(T.rl, T.el, T.kind, &q, &tt);
	A.lay = align_scratch_layout(q, tt);
	const PackedView ref = ref_view(R, V.ref_id, V.rev, T.ref_start), enc = enc_view(R, N.read, T.enc_start);
	uint32_t lead = 0;
	const uint32_t n = edit_script_task<GROUP, PHASE>(A, ref, T.rl, enc, T.el, T.kind, esbuf + T.es_off, &lead);
	if (PHASE != 1 && A.gl == 0) { T.es_len = n; T.lead = lead; }
	CLB_PH_END(GROUP, 3)
	CLB_PH_COUNT(GROUP, 7, 1)
}

// The backward half of the tasks of a bin, one thread per task (align_back.cuh); lg = log2 of the forward kernel's lane group.
__global__ void __launch_bounds__(ALIGN_THREADS) k_align_back(Task* __restrict__ tasks, uint64_t t0, const uint32_t* __restrict__ list, uint32_t n_list, uint64_t stride,
	uint8_t* __restrict__ scratch, ReadStore R, const Node* __restrict__ nodes, const CandView* __restrict__ cviews, uint32_t c, char* __restrict__ esbuf, int lg)
{
	__shared__ ulonglong2 s_ring[BACK_RING * ALIGN_THREADS];
	const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
	if (slot >= n_list) return;
	Task& T = tasks[t0 + list[slot]];
	const Node N = nodes[T.node];
	const CandView& V = cviews[(size_t)T.node * c + N.level];
	long long q, tt;
	align_task_dims(T.rl, T.el, T.kind, &q, &tt);
	const AlignScratch lay = align_scratch_layout(q, tt);
	const PackedView ref = ref_view(R, V.ref_id, V.rev, T.ref_start), enc = enc_view(R, N.read, T.enc_start);
	uint32_t lead = 0;
	const uint32_t n = edit_script_back(scratch + (uint64_t)slot * stride, lay, lg, s_ring + threadIdx.x, ref, T.rl, enc, T.el, T.kind, esbuf + T.es_off, &lead);
	T.es_len = n; T.lead = lead;
}

// ------------------------------------------------------------------------------------------------ decisions
CLB_D int es_code(char ch) { switch (ch) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; case 'D': return 4; case 'M': return 5; case 'X': return 6; case 'Y': return 7; case 'Z': return 8; default: return 11; } }
CLB_D uint32_t ilog2u_bits(uint64_t x) { return x ? 64 - __clzll((long long)x) : 0; }      // utils.h ilog2 convention: number of bits

// utils.h:700-757 (CEntropy::entropy): -sum p log2 p over the listed symbols in the reference's order
CLB_D double entropy_of(const uint32_t* h, int n)
{
	double sum = 0;
	for (int i = 0; i < n; ++i) sum = __dadd_rn(sum, (double)h[i]);
	const double rec = __ddiv_rn(1.0, sum);
	double e = 0;
	for (int i = 0; i < n; ++i) if (h[i]) { const double p = __dmul_rn((double)h[i], rec); e = __dadd_rn(e, __dmul_rn(log2(p), p)); }
	return -e;
}

// AdjustAnchors (encoder.cpp:778-868) on a view: keep the anchors inside [ns, ne) of the current frame, clip the border
// anchors (drop them if less than anchor_len symbols remain inside), shift to the part's frame.
CLB_D void cv_adjust(const uint8_t* __restrict__ arena, const CandView& v, uint32_t ns, uint32_t ne, uint32_t anchor_len, CandView& o)
{
	o = v; o.n = 0; o.tot = 0;
	const uint32_t n = v.n;
	uint32_t first = 0xFFFFFFFFu, last = 0xFFFFFFFFu;
	for (uint32_t i = 0; i < n; ++i) { const Anchor a = cv_get(arena, v, i); if (a.pos_enc + a.len > ns) { first = i; break; } }
	if (first == 0xFFFFFFFFu) return;
	{ const Anchor a = cv_get(arena, v, first); if (a.pos_enc < ns && (a.pos_enc + a.len) - ns < anchor_len) ++first; }
	for (uint32_t i = n; i-- > 0;) { if (cv_get(arena, v, i).pos_enc < ne) { last = i; break; } }
	if (last == 0xFFFFFFFFu) return;
	if (first < n && last < n) {
		const Anchor a = cv_get(arena, v, last);
		if (a.pos_enc + a.len > ne && ne - a.pos_enc < anchor_len) { if (last == 0) return; --last; }
	}
	if (first > last) return;
	o.first = v.first + first; o.n = last - first + 1;
	o.head = first == 0 ? v.head : 0; o.tail = last == n - 1 ? v.tail : 0;
	{ const Anchor a = cv_get(arena, o, o.n - 1); if (a.pos_enc + a.len > ne) o.tail += a.pos_enc + a.len - ne; }
	{ const Anchor a = cv_get(arena, o, 0); if (a.pos_enc < ns) o.head += ns - a.pos_enc; }
	o.shift = v.shift + ns;
	uint32_t tot = 0;
	for (uint32_t i = 0; i < o.n; ++i) tot += cv_get(arena, o, i).len;
	o.tot = tot;
}

struct DecideArgs {
	Task* tasks; uint64_t t0, t1;
	Node* nodes; CandView* cviews; unsigned int* node_cursor; uint32_t node_cap;
	const uint8_t* arena; ReadStore R; char* esbuf; S2P P;
};

__global__ void __launch_bounds__(128) k_decide(DecideArgs a)
{
	const uint64_t t = a.t0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= a.t1) return;
	Task& T = a.tasks[t];
	const Node N = a.nodes[T.node];
	const char* es = a.esbuf + T.es_off;
	const uint32_t n = T.es_len, lead = T.lead, el = T.el;
	const PackedView enc = enc_view(a.R, N.read, T.enc_start);
	if (el >= a.P.min_alt) {
		// encoder.cpp:1300-1327: >= 10 leading deletions are left out of the comparison
		uint32_t k = 0; while (k < n && es[k] == 'D') ++k;
		uint32_t h[11] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};      // order A C D G M T X Y Z S R (utils.h:721)
		uint32_t n_eff;
		if (lead + k >= 10) n_eff = n - k; else { k = 0; n_eff = lead + n; h[2] = lead; }
		for (uint32_t i = k; i < n; ++i) {
			switch (es[i]) { case 'A': ++h[0]; break; case 'C': ++h[1]; break; case 'D': ++h[2]; break; case 'G': ++h[3]; break; case 'M': ++h[4]; break;
			case 'T': ++h[5]; break; case 'X': ++h[6]; break; case 'Y': ++h[7]; break; default: ++h[8]; break; }
		}
		uint32_t hd[4] = {0, 0, 0, 0};
		for (uint32_t i = 0; i < el; ++i) ++hd[enc[(int)i] & 3];
		const double lhs = __dmul_rn(__dmul_rn(entropy_of(h, 11), (double)n_eff), a.P.cost_mult);
		const double rhs = __dmul_rn(entropy_of(hd, 4), (double)el);
		if (lhs < rhs) { T.decision = D_ES; return; }
		// encoder.cpp:1329-1346 (EncodeWithAlternativeRead)
		if (N.ncand <= N.level + 1 || N.level >= a.P.max_rec) { T.decision = D_PLAIN; return; }
		const uint32_t c = a.P.c;
		const CandView* pv = a.cviews + (size_t)T.node * c;
		const uint32_t ns = T.enc_start - N.enc_start, ne = ns + el;
		CandView loc[32];
		for (uint32_t q = N.level + 1; q < N.ncand; ++q) {
			CandView v; cv_adjust(a.arena, pv[q], ns, ne, a.P.m, v);
			uint32_t j = q;
			while (j > N.level + 1 && loc[j - 1].tot < v.tot) { loc[j] = loc[j - 1]; --j; }
			loc[j] = v;
		}
		if (loc[N.level + 1].tot == 0) { T.decision = D_PLAIN; return; }
		const uint32_t id = atomicAdd(a.node_cursor, 1u);
		if (id >= a.node_cap) { T.decision = D_PLAIN; return; }          // cannot happen: capacity = nodes + tasks of the level
		CandView* cv = a.cviews + (size_t)id * c;
		for (uint32_t q = 0; q <= N.level; ++q) cv[q] = pv[q];
		for (uint32_t q = N.level + 1; q < N.ncand; ++q) cv[q] = loc[q];
		Node ch{};
		ch.read = N.read; ch.level = N.level + 1; ch.enc_start = T.enc_start; ch.enc_len = el; ch.first_task = 0;
		ch.n_anch = loc[N.level + 1].n; ch.ncand = N.ncand; ch.valid = 1;
		a.nodes[id] = ch;
		T.child = id; T.decision = D_ALT;
		return;
	}
	// short part: the decision belongs to the adaptive estimator; prepare its inputs (utils.h:838-899 analyze_es)
	uint16_t rd[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
	uint8_t* runs = reinterpret_cast<uint8_t*>(a.esbuf + T.es_off + n);
	uint32_t n_runs = 0;
	char cur = lead ? 'D' : ' '; uint32_t len = lead;
	for (uint32_t i = 0; i <= n; ++i) {
		const char x = i < n ? es[i] : ' ';
		if (x == cur) { ++len; continue; }
		if (cur == 'D') { if (len >= 10) { ++rd[9]; runs[n_runs++] = (uint8_t)(ilog2u_bits(len) + 1); } else rd[4] += (uint16_t)len; }
		else if (cur == 'M') { if (len >= 15) { ++rd[10]; runs[n_runs++] = (uint8_t)(ilog2u_bits(len) + 1); } else rd[5] += (uint16_t)len; }
		else if (cur != ' ') ++rd[es_code(cur)];
		cur = x; len = 1;
	}
	uint16_t rp[4] = {0, 0, 0, 0};
	for (uint32_t i = 0; i < el; ++i) ++rp[enc[(int)i] & 3];
	for (int i = 0; i < 12; ++i) T.rd[i] = rd[i];
	for (int i = 0; i < 4; ++i) T.rp[i] = rp[i];
	T.n_runs = n_runs;
	T.decision = D_PENDING;
}

// ------------------------------------------------------------------------------------------------ estimator
struct Estimator {              // utils.h:760-1126
	uint32_t dna[4], es[12], dec[2];
	double dna_log[4], es_log[12], dec_log[2];
	uint32_t dna_sum, es_sum, dec_sum;
};
CLB_D void est_rescale(uint32_t* a, int n, uint32_t& sum) { while (sum > (1u << 20)) { sum = 0; for (int i = 0; i < n; ++i) { a[i] = (a[i] + 1) / 2; sum += a[i]; } } }
CLB_D void est_logs(const uint32_t* a, double* l, int n, uint32_t sum)
{
	const double rec = __ddiv_rn(1.0, (double)sum);
	for (int i = 0; i < n; ++i) l[i] = a[i] ? -log2(__dmul_rn((double)a[i], rec)) : 0.0;
}
CLB_D void est_reset(Estimator& e)
{
	for (int i = 0; i < 4; ++i) e.dna[i] = 1; e.dna_sum = 4;
	for (int i = 0; i < 12; ++i) e.es[i] = 1; e.es_sum = 12;
	for (int i = 0; i < 2; ++i) e.dec[i] = 1; e.dec_sum = 2;
	est_logs(e.dna, e.dna_log, 4, e.dna_sum); est_logs(e.es, e.es_log, 12, e.es_sum); est_logs(e.dec, e.dec_log, 2, e.dec_sum);
}
// base counts of a read from the packed stream
CLB_D void read_hist(const uint64_t* __restrict__ pk, uint64_t start, uint32_t len, uint32_t* h)
{
	uint64_t p = start; const uint64_t end = start + len;
	uint32_t c1 = 0, c2 = 0, c3 = 0;
	while (p < end) {
		const uint64_t w = pk[p >> 5];
		const uint32_t o = (uint32_t)(p & 31);
		const uint32_t take = (uint32_t)min((uint64_t)(32 - o), end - p);
		// fields o .. o+take-1 (field 0 = top two bits)
		uint64_t m = take == 32 ? ~0ULL : (((1ULL << (2 * take)) - 1) << (64 - 2 * (o + take)));
		m &= 0x5555555555555555ULL;
		const uint64_t lo = w & m, hi = (w >> 1) & m;
		c3 += __popcll(lo & hi); c2 += __popcll(hi & ~lo); c1 += __popcll(lo & ~hi);
		p += take;
	}
	h[0] = len - c1 - c2 - c3; h[1] = c1; h[2] = c2; h[3] = c3;
}

struct PackArgs {
	const uint32_t* pack_first; uint32_t n_packs;       // pack_first[n_packs + 1]: read index (absolute)
	uint32_t read_lo, n_reads;
	const uint32_t* slot_of_read;                        // per batch read: level-0 node or 0xFFFFFFFF
	const uint8_t* has_n;
	Task* tasks; const Node* nodes; const char* esbuf; ReadStore R;
	uint4* hist;                                         // per batch read: base counts
	uint32_t* pend_cnt; const uint64_t* pend_off; uint32_t* pend;     // short parts of every read in the reference's order
};

// base counts of every read of the batch (CEntropyEstimator::LogRead input)
__global__ void __launch_bounds__(128) k_read_hist(PackArgs a)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= a.n_reads) return;
	const uint32_t r = a.read_lo + i;
	uint32_t h[4] = {0, 0, 0, 0};
	if (!a.has_n[r]) read_hist(a.R.pk, a.R.rd_start[r], a.R.rd_len[r], h);
	a.hist[i] = make_uint4(h[0], h[1], h[2], h[3]);
}

// The estimator sees the short parts of a read in the order EncodePart is called: depth first through the alternative reads.
template <bool FILL>
__global__ void __launch_bounds__(128) k_pending(PackArgs a)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= a.n_reads) return;
	const uint32_t root = a.slot_of_read[i];
	uint32_t n = 0;
	if (!a.has_n[a.read_lo + i] && root != 0xFFFFFFFFu && a.nodes[root].valid) {
		uint32_t* out = FILL ? a.pend + a.pend_off[i] : nullptr;
		uint32_t st_node[10], st_i[10]; int sp = 1;
		st_node[0] = root; st_i[0] = 0;
		while (sp > 0) {
			const Node& N = a.nodes[st_node[sp - 1]];
			if (st_i[sp - 1] > N.n_anch) { --sp; continue; }
			const uint32_t t = N.first_task + st_i[sp - 1]++;
			const uint32_t d = a.tasks[t].decision;
			if (d == D_ALT) { st_node[sp] = a.tasks[t].child; st_i[sp] = 0; ++sp; }
			else if (d == D_PENDING) { if (FILL) out[n] = t; ++n; }
		}
	}
	if (!FILL) a.pend_cnt[i] = n;
}

// One warp per read pack replays CEntropyEstimator (utils.h:760-1126): lanes 0-11 own the edit-script symbol counters,
// lanes 12-13 the decision counters, lanes 14-17 the base counters; every lane computes the log2 of its own counter, the
// cost sums are taken in the reference's order through shuffles.  32 parts are loaded at a time (one per lane).
__global__ void __launch_bounds__(32) k_estimate(PackArgs a)
{
	const uint32_t pk_i = blockIdx.x, lane = threadIdx.x;
	const unsigned FULL = 0xffffffffu;
	const bool is_es = lane < 12, is_dec = lane == 12 || lane == 13, is_dna = lane >= 14 && lane < 18;
	uint32_t cnt = 1;
	uint32_t es_sum = 12, dec_sum = 2, dna_sum = 4;
	double dna_lg = is_dna ? -log2(__dmul_rn(1.0, __ddiv_rn(1.0, 4.0))) : 0.0;
	const uint32_t LIM = 1u << 20;
	for (uint32_t r = a.pack_first[pk_i]; r < a.pack_first[pk_i + 1]; ++r) {
		if (a.has_n[r]) continue;
		const uint32_t bi = r - a.read_lo;
		{	// LogRead (utils.h:946)
			const uint4 h = a.hist[bi];
			if (is_dna) cnt += lane == 14 ? h.x : lane == 15 ? h.y : lane == 16 ? h.z : h.w;
			dna_sum += a.R.rd_len[r];
			while (dna_sum > LIM) {
				if (is_dna) cnt = (cnt + 1) / 2;
				dna_sum = __shfl_sync(FULL, cnt, 14) + __shfl_sync(FULL, cnt, 15) + __shfl_sync(FULL, cnt, 16) + __shfl_sync(FULL, cnt, 17);
			}
			if (is_dna) dna_lg = -log2(__dmul_rn((double)cnt, __ddiv_rn(1.0, (double)dna_sum)));
		}
		const uint64_t p0 = a.pend_off[bi], p1 = p0 + a.pend_cnt[bi];
		for (uint64_t base = p0; base < p1; base += 32) {
			const uint32_t m = (uint32_t)min((uint64_t)32, p1 - base);
			uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0}, nr = 0, rl = 0, tid = 0;
			unsigned long long rptr = 0;
			if (lane < m) {
				tid = a.pend[base + lane];
				const Task& T = a.tasks[tid];
				const uint32_t* rw = reinterpret_cast<const uint32_t*>(T.rd);          // rd[12] + rp[4] = 8 words
#pragma unroll
				for (int x = 0; x < 8; ++x) w[x] = rw[x];
				nr = T.n_runs; rl = T.rl; rptr = T.es_off + T.es_len;
			}
			uint32_t my_dec = 0;
			for (uint32_t k = 0; k < m; ++k) {
				uint32_t W[8];
#pragma unroll
				for (int x = 0; x < 8; ++x) W[x] = __shfl_sync(FULL, w[x], k);
				const uint32_t nr_k = __shfl_sync(FULL, nr, k), rl_k = __shfl_sync(FULL, rl, k);
				const unsigned long long rp_k = __shfl_sync(FULL, rptr, k);
				// my share of the part's statistics
				uint32_t wsel = 0;
#pragma unroll
				for (int x = 0; x < 8; ++x) if ((lane >> 1) == (uint32_t)x) wsel = W[x];
				uint32_t mine = (wsel >> (16 * (lane & 1))) & 0xffffu;                   // lanes 0-11: rd[lane]
				if (is_dna) {                                                            // lanes 14-17: rp[lane - 14] = halves of W[6], W[7]
					const uint32_t ww = lane < 16 ? W[6] : W[7];
					mine = (ww >> (16 * (lane & 1))) & 0xffffu;
				}
				uint32_t sum_rd = 0;
#pragma unroll
				for (int x = 0; x < 6; ++x) sum_rd += (W[x] & 0xffffu) + (W[x] >> 16);
				const uint32_t loc = cnt + mine, loc_sum = es_sum + sum_rd;
				double lg = 0.0;
				if (is_es) lg = -log2(__dmul_rn((double)loc, __ddiv_rn(1.0, (double)loc_sum)));
				else if (is_dec) lg = -log2(__dmul_rn((double)cnt, __ddiv_rn(1.0, (double)dec_sum)));
				const double prod = is_es ? __dmul_rn((double)mine, lg) : (is_dna ? __dmul_rn((double)mine, dna_lg) : 0.0);
				double es_cost = __shfl_sync(FULL, lg, 12), plain_cost = __shfl_sync(FULL, lg, 13);
#pragma unroll
				for (int x = 0; x < 12; ++x) es_cost = __dadd_rn(es_cost, __shfl_sync(FULL, prod, x));
				const uint8_t* runs = reinterpret_cast<const uint8_t*>(a.esbuf) + rp_k;
				for (uint32_t x = 0; x < nr_k; ++x) es_cost = __dadd_rn(es_cost, (double)runs[x]);
#pragma unroll
				for (int x = 14; x < 18; ++x) plain_cost = __dadd_rn(plain_cost, __shfl_sync(FULL, prod, x));
				plain_cost = __dadd_rn(plain_cost, (double)(ilog2u_bits(rl_k) + 1));
				const bool plain = plain_cost < es_cost;
				if (plain) { if (lane == 13) ++cnt; }
				else {
					if (lane == 12) ++cnt;
					if (is_es) cnt = loc;
					es_sum = loc_sum;
					while (es_sum > LIM) {
						if (is_es) cnt = (cnt + 1) / 2;
						uint32_t t = is_es ? cnt : 0;
						for (int d = 16; d; d >>= 1) t += __shfl_xor_sync(FULL, t, d);
						es_sum = t;
					}
				}
				++dec_sum;
				while (dec_sum > LIM) {
					if (is_dec) cnt = (cnt + 1) / 2;
					dec_sum = __shfl_sync(FULL, cnt, 12) + __shfl_sync(FULL, cnt, 13);
				}
				if (lane == k) my_dec = plain ? D_PLAIN : D_ES;
			}
			if (lane < m) a.tasks[tid].decision = my_dec;
		}
	}
}

// ------------------------------------------------------------------------------------------------ tuple emission
enum : uint32_t { T_INS = 0, T_DEL = 1, T_MATCH = 2, T_SUB = 3, T_ANCHOR = 4, T_SKIP = 5, T_ALT = 6, T_MAIN = 7, T_PLAIN = 8, T_START_PLAIN = 9, T_START_ES = 10, T_START_N = 11 };

template <bool W>
struct TupleOut {
	uint8_t* p; uint64_t n;
	CLB_D void byte(uint32_t b) { if (W) p[n] = (uint8_t)b; ++n; }
	CLB_D void len28(uint32_t type, uint32_t v) { byte((type << 4) + (v >> 24)); byte((v >> 16) & 0xff); byte((v >> 8) & 0xff); byte(v & 0xff); }
	CLB_D void id(uint32_t type, uint32_t x, uint32_t rev) { byte((type << 4) + rev); byte(x >> 24); byte((x >> 16) & 0xff); byte((x >> 8) & 0xff); byte(x & 0xff); }
	// encoder.cpp:1348-1392 (singleEditScriptSymbolStore)
	CLB_D void run(char s, uint32_t rep)
	{
		if (s == 'M') { if (rep >= 15) len28(T_ANCHOR, rep); else for (uint32_t i = 0; i < rep; ++i) byte(T_MATCH << 4); }
		else if (s == 'D') { if (rep > 16) len28(T_SKIP, rep); else for (uint32_t i = 0; i < rep; ++i) byte(T_DEL << 4); }
		else if (s == 'X' || s == 'Y' || s == 'Z') for (uint32_t i = 0; i < rep; ++i) byte((T_SUB << 4) + (uint32_t)(s - 'X'));
		else for (uint32_t i = 0; i < rep; ++i) byte((T_INS << 4) + (uint32_t)es_code(s));
	}
};

struct EmitArgs {
	uint32_t read_lo, n_reads;                       // batch
	const uint32_t* slot_of_read; const uint8_t* has_n;
	const Task* tasks; const Node* nodes; const CandView* cviews; uint32_t c;
	const uint8_t* arena; const char* esbuf; ReadStore R;
	uint32_t* size; uint32_t* kind;                  // per batch read; kind 0 = edit script, 1 = plain, 2 = plain with N
	const uint64_t* off; uint64_t base; uint8_t* out; uint64_t* es_off;
};

template <bool W>
__device__ uint64_t emit_read(const EmitArgs& a, uint32_t root, uint8_t* dst)
{
	TupleOut<W> o{dst, 0};
	struct Frame { uint32_t node, i, last_pos, cur_ref, after_d, open; } st[10];
	int sp = 1;
	st[0] = Frame{root, 0, 0, 0, 0, 0};
	const CandView& V0 = a.cviews[(size_t)root * a.c];
	const uint32_t main_ref = V0.ref_id;
	o.id(T_START_ES, main_ref, V0.rev);
	bool first = true;
	char ps = 0; uint32_t pr = 0;            // pending run of the open segment
	while (sp > 0) {
		Frame& f = st[sp - 1];
		const Node& N = a.nodes[f.node];
		const CandView& V = a.cviews[(size_t)f.node * a.c + N.level];
		// big_edit_script += symbols (the segment header is written when the first symbol arrives: encoder.cpp:1414-1443)
		auto push = [&](char s, uint32_t rep) {
			if (!rep) return;
			if (!f.open) {
				f.open = 1;
				if (N.level == 0) { if (V.ref_id != main_ref) o.id(T_ALT, V.ref_id, V.rev); else if (!first) o.byte(T_MAIN << 4); }
				else { if (V.ref_id != main_ref) o.id(T_ALT, V.ref_id, V.rev); else o.byte(T_MAIN << 4); }
				ps = 'D'; pr = N.level > 0 ? f.last_pos : 0;
			}
			if (pr && ps == s) pr += rep;
			else { if (pr) o.run(ps, pr); ps = s; pr = rep; }
		};
		auto flush = [&](uint32_t cur_pos) {
			if (f.open) { if (pr) o.run(ps, pr); pr = 0; f.last_pos = cur_pos; first = false; f.open = 0; }
		};
		if (f.after_d) { const uint32_t d = f.after_d; f.after_d = 0; push('D', d); }
		const uint32_t n_frag = 2 * N.n_anch + 1;
		if (f.i >= n_frag) { flush(f.cur_ref); --sp; continue; }
		const uint32_t i = f.i++;
		if (i & 1) {
			const Anchor an = cv_get(a.arena, V, i >> 1);
			push('M', an.len);
			f.cur_ref = an.pos_ref + an.len;
			continue;
		}
		const Task& T = a.tasks[N.first_task + (i >> 1)];
		const bool last = i == n_frag - 1;
		if (T.decision == D_ES) {
			push('D', T.lead);
			const char* es = a.esbuf + T.es_off;
			for (uint32_t k = 0; k < T.es_len; ++k) push(es[k], 1);
		} else if (T.decision == D_ALT) {
			flush(f.cur_ref);
			if (!last) f.after_d = T.rl;
			st[sp] = Frame{T.child, 0, 0, 0, 0, 0};
			++sp;
		} else {
			const PackedView e = enc_view(a.R, N.read, T.enc_start);
			for (uint32_t k = 0; k < T.el; ++k) push("ACGT"[e[(int)k] & 3], 1);
			if (!last) push('D', T.rl);
		}
	}
	return o.n;
}

template <bool W>
__global__ void __launch_bounds__(128) k_emit(EmitArgs a)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= a.n_reads) return;
	const uint32_t r = a.read_lo + i;
	const uint32_t root = a.slot_of_read[i];
	if (!W) {
		if (a.has_n[r]) { a.kind[i] = 2; a.size[i] = 1 + a.R.rd_len[r]; return; }
		if (root == 0xFFFFFFFFu || !a.nodes[root].valid) { a.kind[i] = 1; a.size[i] = 1 + a.R.rd_len[r]; return; }
		a.kind[i] = 0;
		a.size[i] = (uint32_t)emit_read<false>(a, root, nullptr);
	} else {
		a.es_off[r] = a.base + a.off[i];
		if (a.kind[i] == 0) emit_read<true>(a, root, a.out + a.base + a.off[i]);
	}
}

// ---- the default emission since round 2 (run on a B200: CompactES bytes of the goldens, 219 ms per 25 Gbases against 825 ms; CLB_EMIT_THREAD=1 keeps k_emit) ----
// k_emit gives a read to one thread: ncu shows 1.5 of 32 lanes active and every script symbol costs a byte load, a branchy push
// and a byte store of its own.  Here a read belongs to a warp.  The walk over the node tree is the one of emit_read, executed by
// all 32 lanes with identical (uniform) state; what changes is the per-symbol work: a script string or a run of plain bases is
// taken 32 symbols at a time — a ballot marks where runs start, the first segment joins the pending run, every complete run
// inside the chunk is sized by its lanes (one byte per symbol, or one 4-byte anchor / skip tuple), a warp scan places the bytes
// and the last segment becomes the pending run.  The chunk rule was checked against the serial push on 20 000 random strings
// and pending states with a lane-by-lane emulation before it was written down here (profiles/r01_summary.md).
template <bool W>
struct WarpOut {
	uint8_t* p; uint64_t n; uint32_t lane;
	CLB_D void byte(uint32_t b) { if (W && lane == 0) p[n] = (uint8_t)b; ++n; }
	CLB_D void len28(uint32_t type, uint32_t v) { if (W && lane == 0) { p[n] = (uint8_t)((type << 4) + (v >> 24)); p[n + 1] = (uint8_t)(v >> 16); p[n + 2] = (uint8_t)(v >> 8); p[n + 3] = (uint8_t)v; } n += 4; }
	CLB_D void id(uint32_t type, uint32_t x, uint32_t rev) { if (W && lane == 0) { p[n] = (uint8_t)((type << 4) + rev); p[n + 1] = (uint8_t)(x >> 24); p[n + 2] = (uint8_t)(x >> 16); p[n + 3] = (uint8_t)(x >> 8); p[n + 4] = (uint8_t)x; } n += 5; }
	static CLB_D uint32_t sym_byte(char s)
	{
		if (s == 'M') return T_MATCH << 4;
		if (s == 'D') return T_DEL << 4;
		if (s == 'X' || s == 'Y' || s == 'Z') return (T_SUB << 4) + (uint32_t)(s - 'X');
		return (T_INS << 4) + (uint32_t)es_code(s);
	}
	static CLB_D bool four(char s, uint32_t rep) { return (s == 'M' && rep >= 15) || (s == 'D' && rep > 16); }
	CLB_D void run(char s, uint32_t rep)                       // TupleOut::run, the bytes of a short run written side by side
	{
		if (four(s, rep)) { len28(s == 'M' ? T_ANCHOR : T_SKIP, rep); return; }
		if (W) { const uint8_t b = (uint8_t)sym_byte(s); for (uint32_t i = lane; i < rep; i += 32) p[n + i] = b; }
		n += rep;
	}
};

// symbols [0, len) of a string (get(k) = symbol k, called with a lane's own index) through the pending run (ps, pr)
template <bool W, class Get>
CLB_D void push_string(WarpOut<W>& o, char& ps, uint32_t& pr, uint32_t len, Get get)
{
	const unsigned FULL = 0xffffffffu;
	const uint32_t lane = o.lane;
	for (uint32_t k0 = 0; k0 < len; k0 += 32) {
		const uint32_t n = min(32u, len - k0);
		const int c = lane < n ? (int)get(k0 + lane) : 0;
		int prev = __shfl_up_sync(FULL, c, 1);
		if (lane == 0) prev = pr ? (int)ps : 0;                  // 0 is no script symbol: without a pending run lane 0 starts one
		const uint32_t mask = __ballot_sync(FULL, lane < n && c != prev);
		uint32_t lo = 0;
		if (mask & 1u) { if (pr) o.run(ps, pr); pr = 0; }        // the chunk opens a new run: the pending one is complete
		else {
			if (mask == 0) { pr += n; continue; }                 // the whole chunk continues the pending run
			lo = (uint32_t)__ffs((int)mask) - 1u;
			pr += lo; o.run(ps, pr); pr = 0;                      // the first segment completes it
		}
		const uint32_t m2 = mask & ~((1u << lo) - 1u);           // run starts from lo on (bit lo is set)
		const uint32_t last_b = 31u - (uint32_t)__clz((int)m2);
		uint32_t contrib = 0, L = 0; bool f4 = false;
		if (lane >= lo && lane < n) {
			const uint32_t upto = lane == 31 ? 0xffffffffu : ((2u << lane) - 1u);
			const uint32_t b = 31u - (uint32_t)__clz((int)(m2 & upto));
			const uint32_t above = m2 & ~upto;
			const uint32_t e = above ? (uint32_t)__ffs((int)above) - 1u : n;
			if (b != last_b) {                                    // a complete run inside the chunk
				L = e - b;
				f4 = WarpOut<W>::four((char)c, L);
				contrib = f4 ? (lane == b ? 4u : 0u) : 1u;
			}
		}
		uint32_t incl = contrib;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(FULL, incl, d); if ((int)lane >= d) incl += v; }
		const uint32_t total = __shfl_sync(FULL, incl, 31);
		if (W && contrib) {
			uint8_t* q = o.p + o.n + (incl - contrib);
			if (f4) { const uint32_t ty = c == 'M' ? T_ANCHOR : T_SKIP; q[0] = (uint8_t)((ty << 4) + (L >> 24)); q[1] = (uint8_t)(L >> 16); q[2] = (uint8_t)(L >> 8); q[3] = (uint8_t)L; }
			else q[0] = (uint8_t)WarpOut<W>::sym_byte((char)c);
		}
		o.n += total;
		ps = (char)__shfl_sync(FULL, c, (int)last_b); pr = n - last_b;      // the last segment is the pending run now
	}
}

template <bool W>
__device__ uint64_t emit_read_w(const EmitArgs& a, uint32_t root, uint8_t* dst, uint32_t lane)
{
	WarpOut<W> o{dst, 0, lane};
	struct Frame { uint32_t node, i, last_pos, cur_ref, after_d, open; } st[10];
	int sp = 1;
	st[0] = Frame{root, 0, 0, 0, 0, 0};
	const CandView& V0 = a.cviews[(size_t)root * a.c];
	const uint32_t main_ref = V0.ref_id;
	o.id(T_START_ES, main_ref, V0.rev);
	bool first = true;
	char ps = 0; uint32_t pr = 0;
	while (sp > 0) {
		Frame& f = st[sp - 1];
		const Node& N = a.nodes[f.node];
		const CandView& V = a.cviews[(size_t)f.node * a.c + N.level];
		// the segment header is written when the first symbol arrives (encoder.cpp:1414-1443): same as emit_read
		auto open = [&]() {
			if (f.open) return;
			f.open = 1;
			if (N.level == 0) { if (V.ref_id != main_ref) o.id(T_ALT, V.ref_id, V.rev); else if (!first) o.byte(T_MAIN << 4); }
			else { if (V.ref_id != main_ref) o.id(T_ALT, V.ref_id, V.rev); else o.byte(T_MAIN << 4); }
			ps = 'D'; pr = N.level > 0 ? f.last_pos : 0;
		};
		auto push = [&](char s, uint32_t rep) {
			if (!rep) return;
			open();
			if (pr && ps == s) pr += rep;
			else { if (pr) o.run(ps, pr); ps = s; pr = rep; }
		};
		auto flush = [&](uint32_t cur_pos) {
			if (f.open) { if (pr) o.run(ps, pr); pr = 0; f.last_pos = cur_pos; first = false; f.open = 0; }
		};
		if (f.after_d) { const uint32_t d = f.after_d; f.after_d = 0; push('D', d); }
		const uint32_t n_frag = 2 * N.n_anch + 1;
		if (f.i >= n_frag) { flush(f.cur_ref); --sp; continue; }
		const uint32_t i = f.i++;
		if (i & 1) {
			const Anchor an = cv_get(a.arena, V, i >> 1);
			push('M', an.len);
			f.cur_ref = an.pos_ref + an.len;
			continue;
		}
		const Task& T = a.tasks[N.first_task + (i >> 1)];
		const bool last = i == n_frag - 1;
		if (T.decision == D_ES) {
			push('D', T.lead);
			if (T.es_len) {
				open();
				const char* es = a.esbuf + T.es_off;
				push_string<W>(o, ps, pr, T.es_len, [&](uint32_t k) { return es[k]; });
			}
		} else if (T.decision == D_ALT) {
			flush(f.cur_ref);
			if (!last) f.after_d = T.rl;
			st[sp] = Frame{T.child, 0, 0, 0, 0, 0};
			++sp;
		} else {
			if (T.el) {
				open();
				const PackedView e = enc_view(a.R, N.read, T.enc_start);
				push_string<W>(o, ps, pr, T.el, [&](uint32_t k) { return "ACGT"[e[(int)k] & 3]; });
			}
			if (!last) push('D', T.rl);
		}
	}
	return o.n;
}

template <bool W>
__global__ void __launch_bounds__(128) k_emit_w(EmitArgs a)
{
	const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (i >= a.n_reads) return;                              // whole warps leave together
	const uint32_t r = a.read_lo + i;
	const uint32_t root = a.slot_of_read[i];
	if (!W) {
		if (a.has_n[r]) { if (lane == 0) { a.kind[i] = 2; a.size[i] = 1 + a.R.rd_len[r]; } return; }
		if (root == 0xFFFFFFFFu || !a.nodes[root].valid) { if (lane == 0) { a.kind[i] = 1; a.size[i] = 1 + a.R.rd_len[r]; } return; }
		const uint32_t sz = (uint32_t)emit_read_w<false>(a, root, nullptr, lane);
		if (lane == 0) { a.kind[i] = 0; a.size[i] = sz; }
	} else {
		if (lane == 0) a.es_off[r] = a.base + a.off[i];
		if (a.kind[i] == 0) emit_read_w<true>(a, root, a.out + a.base + a.off[i], lane);
	}
}

// plain reads: start tuple + one plain(base) tuple per symbol (encoder.cpp:663-682)
__global__ void __launch_bounds__(256) k_emit_plain(EmitArgs a)
{
	const uint32_t i = blockIdx.x;
	const uint32_t k = a.kind[i];
	if (k == 0) return;
	const uint32_t r = a.read_lo + i;
	uint8_t* o = a.out + a.base + a.off[i];
	const uint64_t s = a.R.rd_start[r]; const uint32_t len = a.R.rd_len[r];
	if (threadIdx.x == 0) o[0] = (uint8_t)((k == 2 ? T_START_N : T_START_PLAIN) << 4);
	for (uint32_t j = threadIdx.x; j < len; j += blockDim.x) {
		const uint64_t p = s + j;
		const uint32_t isn = (a.R.nmask[p >> 5] >> (p & 31)) & 1u;
		o[1 + j] = (uint8_t)((T_PLAIN << 4) + (isn ? 4u : base_at(a.R.pk, p)));
	}
}

// ------------------------------------------------------------------------------------------------ the reference's -v counters
// stats_collector.h:28-75 as logged by encoder.cpp:1445-1575: per recursion level what every EncodePart call decided (edit script with
// its symbol classes / alternative read by fragment position / plain), per AddEncodedReadWithCandidates call its flanks and anchors;
// per read the plain / plain-with-N starts (encoder.cpp:663-676).  Only run when clb_encode_stats_enable asked for them.
struct StatsArgs {
	const Task* tasks; uint64_t n_tasks; const Node* nodes; uint64_t n_nodes; const CandView* cviews; uint32_t c;
	const uint8_t* arena; const char* esbuf; const uint32_t* kind; const uint32_t* rd_len; uint32_t read_lo, n_reads;
	unsigned long long* st;
};
__global__ void __launch_bounds__(128) k_stats_tasks(StatsArgs a)
{
	const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= a.n_tasks) return;
	const Task& T = a.tasks[t];
	const Node N = a.nodes[T.node];
	unsigned long long* L = a.st + ST_LEVEL0 + min(N.level, ST_LEVELS - 1) * ST_LEVEL_FIELDS;
	if (T.decision == D_ES) {
		const char* es = a.esbuf + T.es_off;
		unsigned long long sub = 0, mat = 0, ins = 0, del = T.lead;
		for (uint32_t i = 0; i < T.es_len; ++i) { const char ch = es[i]; if (ch == 'M') ++mat; else if (ch == 'D') ++del; else if (ch == 'X' || ch == 'Y' || ch == 'Z') ++sub; else ++ins; }
		atomicAdd(&L[SL_CODED_SYMB], (unsigned long long)T.el);
		atomicAdd(&L[SL_ES_SYMB], (unsigned long long)T.lead + T.es_len);
		if (sub) atomicAdd(&L[SL_SUBST], sub);
		if (mat) atomicAdd(&L[SL_MATCH], mat);
		if (ins) atomicAdd(&L[SL_INS], ins);
		if (del) atomicAdd(&L[SL_DEL], del);
	} else if (T.decision == D_ALT) {      // encoder.cpp:1478-1483: the last fragment first, then the first one
		atomicAdd(&L[T.frag == N.n_anch ? SL_ALT_RIGHT : T.frag == 0 ? SL_ALT_LEFT : SL_ALT_BETWEEN], 1ull);
	} else if (T.decision == D_PLAIN) atomicAdd(&L[SL_PLAIN_SYMB], (unsigned long long)T.el);
}
__global__ void __launch_bounds__(128) k_stats_nodes(StatsArgs a)
{
	const uint64_t id = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (id >= a.n_nodes) return;
	const Node N = a.nodes[id];
	if (!N.valid || !N.n_anch) return;
	const CandView V = a.cviews[id * a.c + N.level];
	unsigned long long* L = a.st + ST_LEVEL0 + min(N.level, ST_LEVELS - 1) * ST_LEVEL_FIELDS;
	unsigned long long symb = 0;
	for (uint32_t i = 0; i < N.n_anch; ++i) symb += cv_get(a.arena, V, i).len;
	const Anchor a0 = cv_get(a.arena, V, 0), al = cv_get(a.arena, V, N.n_anch - 1);
	atomicAdd(&L[SL_LEFT_FLANK], (unsigned long long)a0.pos_enc);
	atomicAdd(&L[SL_RIGHT_FLANK], (unsigned long long)(N.enc_len - (al.pos_enc + al.len)));
	atomicAdd(&L[SL_ANCHORS], (unsigned long long)N.n_anch);
	atomicAdd(&L[SL_ANCHOR_SYMB], symb);
	atomicMax(&a.st[ST_MAX_LEVEL], (unsigned long long)N.level + 1);
}
__global__ void __launch_bounds__(128) k_stats_reads(StatsArgs a)
{
	const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= a.n_reads) return;
	const uint32_t kind = a.kind[r];
	if (kind == 1) { atomicAdd(&a.st[ST_PLAIN_READS], 1ull); atomicAdd(&a.st[ST_PLAIN_SYMB], (unsigned long long)a.rd_len[a.read_lo + r]); }
	else if (kind == 2) { atomicAdd(&a.st[ST_PLAIN_N_READS], 1ull); atomicAdd(&a.st[ST_PLAIN_N_SYMB], (unsigned long long)a.rd_len[a.read_lo + r]); }
}

// ------------------------------------------------------------------------------------------------ driver
template <typename T> static cudaError_t dmalloc(T** p, uint64_t n, cudaStream_t s = nullptr) { return dev_malloc((void**)p, sizeof(T) * (n ? n : 1), s); }
struct Scoped {            // stream-ordered scratch of a batch / level: freed (back to the pool) on every exit path
	cudaStream_t s;
	std::vector<void*> v;
	explicit Scoped(cudaStream_t st) : s(st) {}
	template <typename T> cudaError_t get(T** p, uint64_t n) { cudaError_t e = dev_malloc((void**)p, sizeof(T) * (n ? n : 1), s); if (e == cudaSuccess) v.push_back(*p); return e; }
	~Scoped() { for (void* p : v) dev_free_async(p, s); }
};

static clb_status align_level(clb_ctx* c, const S2P& P, Task* d_tasks, uint64_t t0, uint64_t t1, const ReadStore& R, const Node* d_nodes, const CandView* d_cviews,
	char* d_esbuf, BinStats* d_bins)
{
	cudaStream_t s = c->stream;
	const uint64_t nt = t1 - t0;
	if (!nt) return CLB_OK;
	Scoped mem(s);
	uint32_t* d_bin_of = nullptr; uint32_t* d_list = nullptr;
	CLB_CUDA(c, mem.get(&d_bin_of, nt));
	CLB_CUDA(c, mem.get(&d_list, nt));
	CLB_CUDA(c, cudaMemsetAsync(d_bins, 0, sizeof(BinStats), s));
	const uint32_t blocks = (uint32_t)((nt + 255) / 256);
	CLB_TIMED(c, K_ENCODE, (k_task_classify<<<blocks, 256, 0, s>>>(d_tasks, t0, t1, R, d_nodes, d_esbuf, d_bins, d_bin_of, std::getenv("CLB_ALIGN_PROFILE") != nullptr)));
	CLB_LAUNCH_CHECK(c, "k_task_classify");
	BinStats hb;
	CLB_CUDA(c, cudaMemcpyAsync(&hb, d_bins, sizeof(BinStats), cudaMemcpyDeviceToHost, s));
	CLB_CUDA(c, cudaStreamSynchronize(s));
	unsigned int at = 0;
	for (int b = 0; b < N_BINS; ++b) { hb.base[b] = at; at += hb.cnt[b]; hb.fill[b] = 0; }
	if (!at) return CLB_OK;
	CLB_CUDA(c, cudaMemcpyAsync(d_bins, &hb, sizeof(BinStats), cudaMemcpyHostToDevice, s));
	CLB_TIMED(c, K_ENCODE, (k_task_scatter<<<blocks, 256, 0, s>>>(t0, t1, d_bin_of, d_bins, d_list)));
	CLB_LAUNCH_CHECK(c, "k_task_scatter");
	const char* env_budget = std::getenv("CLB_ALIGN_SCRATCH_MB");
	uint64_t budget = env_budget ? (uint64_t)std::atoll(env_budget) << 20 : 16ull << 30;
	if (!env_budget && c->s2_budget) budget = c->s2_budget;
	else if (!env_budget) {      // as much as is free beyond a reserve for the later levels and stage 3, within 16 .. 40 GiB; fixed for the job
		const uint64_t have = dev_mem_available() + c->s2_scratch.cap;
		const uint64_t reserve = 60ull << 30;          // the tuples grow by reallocation (old + new alive), stage 3 wants a temp of their size
		budget = std::min<uint64_t>(40ull << 30, std::max<uint64_t>(budget, have > reserve ? have - reserve : 0));
		c->s2_budget = budget;
	}
	const bool bin_prof = std::getenv("CLB_ALIGN_PROFILE") != nullptr;
	// bins run concurrently on a few streams (each with its own slice of the scratch) so that the tail of one bin
	// overlaps the bulk of another; the per-bin profile serialises them on the main stream instead
	const int n_str = bin_prof ? 1 : N_ALIGN_STREAMS;
	if (!c->s2_streams[0]) {
		for (int i = 0; i < N_ALIGN_STREAMS; ++i) CLB_CUDA(c, cudaStreamCreateWithFlags(&c->s2_streams[i], cudaStreamNonBlocking));
		CLB_CUDA(c, cudaEventCreateWithFlags(&c->s2_fork, cudaEventDisableTiming));
		for (int i = 0; i < N_ALIGN_STREAMS; ++i) CLB_CUDA(c, cudaEventCreateWithFlags(&c->s2_join[i], cudaEventDisableTiming));
	}
	const uint64_t slice = (budget / n_str) & ~255ull;
	uint64_t need_total = 0;
	for (int b = 0; b < N_BINS; ++b) if (hb.cnt[b]) {
		const uint64_t stride = (align_scratch_layout(hb.maxq[b], hb.maxt[b]).total + 63) & ~63ull;
		if (stride > slice) return fail(c, CLB_ERR_CAPACITY, "one alignment needs more scratch than CLB_ALIGN_SCRATCH_MB allows");
		need_total = std::max(need_total, std::min<uint64_t>(slice, stride * hb.cnt[b]));
	}
	CLB_CUDA(c, c->s2_scratch.reserve(slice * n_str, s, false));
	cudaEvent_t pe0 = nullptr, pe1 = nullptr;
	if (bin_prof) { cudaEventCreate(&pe0); cudaEventCreate(&pe1); }
	prof_begin(c, K_ALIGN);
	CLB_CUDA(c, cudaEventRecord(c->s2_fork, s));
	for (int i = 0; i < n_str && !bin_prof; ++i) CLB_CUDA(c, cudaStreamWaitEvent(c->s2_streams[i], c->s2_fork, 0));
	// bins with the longest single parts first: a part is one warp's serial work, so the level ends no earlier than its
	// longest part; the bulk bins fill the machine next to them
	std::vector<int> order;
	for (int b = 0; b < N_BINS; ++b) if (hb.cnt[b]) order.push_back(b);
	std::sort(order.begin(), order.end(), [&](int x, int y) { return (double)hb.maxq[x] * hb.maxt[x] > (double)hb.maxq[y] * hb.maxt[y]; });
	int rr = 0;
	for (int b : order) {
		if (bin_prof) cudaEventRecord(pe0, s);
		const uint64_t stride = (align_scratch_layout(hb.maxq[b], hb.maxt[b]).total + 63) & ~63ull;
		const uint64_t per_wave = std::max<uint64_t>(1, slice / stride);
		const int g = b < N_SMALL_BINS ? 1 << (b / (2 * N_TBIN)) : 32;
		const bool split = b < N_SPLIT_BINS && !std::getenv("CLB_ALIGN_ONE_KERNEL");
		const int si = rr++ % n_str;
		cudaStream_t ls = bin_prof ? s : c->s2_streams[si];
		uint8_t* scratch = c->s2_scratch.p + (uint64_t)si * slice;
		for (uint64_t w0 = 0; w0 < hb.cnt[b]; w0 += per_wave) {
			const uint32_t m = (uint32_t)std::min<uint64_t>(per_wave, hb.cnt[b] - w0);
			const uint32_t* list = d_list + hb.base[b] + w0;
			const uint32_t threads = 128;
			const uint32_t grid = (uint32_t)(((uint64_t)m * g + threads - 1) / threads);
#define CLB_ALIGN_LAUNCH(G, PH) k_align<G, PH><<<grid, threads, 0, ls>>>(d_tasks, t0, list, m, stride, scratch, R, d_nodes, d_cviews, P.c, d_esbuf)
#define CLB_ALIGN_GROUPS(PH) switch (g) { case 1: CLB_ALIGN_LAUNCH(1, PH); break; case 2: CLB_ALIGN_LAUNCH(2, PH); break; case 4: CLB_ALIGN_LAUNCH(4, PH); break; \
	case 8: CLB_ALIGN_LAUNCH(8, PH); break; case 16: CLB_ALIGN_LAUNCH(16, PH); break; default: CLB_ALIGN_LAUNCH(32, PH); break; }
			// the backward half: the lane-group kernel; the thread-per-task kernel (align_back.cuh) needs about 10^5 tasks in flight to hide
			// its DRAM round trips, which the scratch of a 25-Gbase job does not hold (measured: k_align 4 194 ms against 2 071 ms)
			if (split && !std::getenv("CLB_ALIGN_THREAD_BACK")) { CLB_ALIGN_GROUPS(1) CLB_ALIGN_GROUPS(2) ++c->launches; }
			else if (split) {
				CLB_ALIGN_GROUPS(1)
				k_align_back<<<(m + ALIGN_THREADS - 1) / ALIGN_THREADS, ALIGN_THREADS, 0, ls>>>(d_tasks, t0, list, m, stride, scratch, R, d_nodes, d_cviews, P.c, d_esbuf, ilog2_u32((uint32_t)g));
				++c->launches;
			}
			else if (b < N_SPLIT_BINS) { CLB_ALIGN_GROUPS(0) }
			else CLB_ALIGN_LAUNCH(32, 0);
			CLB_LAUNCH_CHECK(c, "k_align");
		}
		if (bin_prof) {
			cudaEventRecord(pe1, s); cudaEventSynchronize(pe1);
			float ms = 0; cudaEventElapsedTime(&ms, pe0, pe1);
			fprintf(stderr, "[align] group %2d bin %3d: %9u tasks, max q %7u, max t %7u, mean q %7.0f, mean t %7.0f, Gcells %8.3f, stride %9llu, %9.3f ms\n", g, b, hb.cnt[b], hb.maxq[b], hb.maxt[b],
				(double)hb.sumq[b] / hb.cnt[b], (double)hb.sumt[b] / hb.cnt[b], (double)hb.sumqt[b] * 1e-9, (unsigned long long)stride, ms);
		}
	}
	if (bin_prof) { cudaEventDestroy(pe0); cudaEventDestroy(pe1); }
	else for (int i = 0; i < n_str; ++i) {
		CLB_CUDA(c, cudaEventRecord(c->s2_join[i], c->s2_streams[i]));
		CLB_CUDA(c, cudaStreamWaitEvent(s, c->s2_join[i], 0));
	}
	prof_end(c);
	CLB_CUDA(c, cudaStreamSynchronize(s));        // d_list / d_bin_of die with `mem`
#ifdef CLB_ALIGN_PHASES
	{
		unsigned long long h[48];
		cudaMemcpyFromSymbol(h, g_align_phase, sizeof h);
		static const int gs[6] = {1, 2, 4, 8, 16, 32};
		for (int g = 0; g < 6; ++g) if (h[g * 8 + 7]) fprintf(stderr, "[align phases] group %2d: tasks %llu, Mcycles sweep %.0f traceback %.0f finish %.0f whole %.0f | sweep steps %llu (%.0f cyc/step), traceback iterations %llu (%.0f cyc/it), window loads %llu\n",
			gs[g], h[g * 8 + 7], h[g * 8] * 1e-6, h[g * 8 + 1] * 1e-6, h[g * 8 + 2] * 1e-6, h[g * 8 + 3] * 1e-6, h[g * 8 + 4], (double)h[g * 8] / (h[g * 8 + 4] ? h[g * 8 + 4] : 1), h[g * 8 + 5], (double)h[g * 8 + 1] / (h[g * 8 + 5] ? h[g * 8 + 5] : 1), h[g * 8 + 6]);
	}
#endif
	return CLB_OK;
}

static clb_status dump_candidates(clb_ctx* c, const S2P& P, const std::vector<uint32_t>& h_list, const Node* d_nodes, const CandView* d_cviews)
{
	const uint32_t nb = (uint32_t)h_list.size();
	std::vector<Node> nodes(nb); std::vector<CandView> cv((size_t)nb * P.c);
	if (!nb) return CLB_OK;
	CLB_CUDA(c, cudaMemcpyAsync(nodes.data(), d_nodes, sizeof(Node) * nb, cudaMemcpyDeviceToHost, c->stream));
	CLB_CUDA(c, cudaMemcpyAsync(cv.data(), d_cviews, sizeof(CandView) * nb * P.c, cudaMemcpyDeviceToHost, c->stream));
	CLB_CUDA(c, cudaStreamSynchronize(c->stream));
	std::vector<Anchor> an;
	for (uint32_t i = 0; i < nb; ++i) {
		const uint32_t r = h_list[i];
		std::vector<uint32_t> rec;
		for (uint32_t k = 0; k < nodes[i].ncand; ++k) {
			const CandView& v = cv[(size_t)i * P.c + k];
			an.resize(v.n);
			if (v.n) CLB_CUDA(c, cudaMemcpy(an.data(), c->s2_arena.p + v.anc, sizeof(Anchor) * v.n, cudaMemcpyDeviceToHost));
			rec.insert(rec.end(), {v.ref_id, v.rev, v.tot, v.n});
			for (const Anchor& x : an) rec.insert(rec.end(), {x.len, x.pos_enc, x.pos_ref});
		}
		c->dbg_cand[r] = std::move(rec);
	}
	return CLB_OK;
}

struct Trace {            // CLB_S2_TRACE=1: wall time of every phase (synchronising); =2: host enqueue time + device time from events, no synchronisation
	int mode; cudaStream_t s; double t0, h0; const char* what;
	struct Rec { const char* w; double host; cudaEvent_t ev; };
	std::vector<Rec> recs; cudaEvent_t ev0 = nullptr;
	static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
	Trace(cudaStream_t st) : mode(std::getenv("CLB_S2_TRACE") ? std::atoi(std::getenv("CLB_S2_TRACE")) : 0), s(st), t0(now()), h0(t0), what("start")
	{ if (mode == 2) { cudaEventCreate(&ev0); cudaEventRecord(ev0, s); } }
	void mark(const char* w)
	{
		if (!mode) return;
		if (mode == 2) { Rec r{w, now(), nullptr}; cudaEventCreate(&r.ev); cudaEventRecord(r.ev, s); recs.push_back(r); return; }
		cudaStreamSynchronize(s); const double t = now(); fprintf(stderr, "[s2] %-28s %9.3f ms\n", w, t - t0); t0 = t;
	}
	~Trace()
	{
		if (mode != 2 || !ev0) return;
		cudaStreamSynchronize(s);
		double hp = h0; cudaEvent_t ep = ev0;
		for (auto& r : recs) { float ms = 0; cudaEventElapsedTime(&ms, ep, r.ev); fprintf(stderr, "[s2e] %-28s host %9.3f ms   device %9.3f ms\n", r.w, r.host - hp, ms); hp = r.host; if (ep != ev0) cudaEventDestroy(ep); ep = r.ev; }
		if (ep != ev0) cudaEventDestroy(ep);
		cudaEventDestroy(ev0);
	}
};

// The anchors of the chosen candidates leave the per-batch pair arena for a compact store that lives until the tuples are out.
__global__ void __launch_bounds__(128) k_anchor_count(const Node* __restrict__ nodes, const CandView* __restrict__ cviews, uint32_t n_slots, uint32_t c, uint32_t* __restrict__ cnt)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_slots) return;
	uint32_t n = 0;
	for (uint32_t k = 0; k < nodes[i].ncand; ++k) n += cviews[(size_t)i * c + k].n;
	cnt[i] = n;
}
__global__ void __launch_bounds__(128) k_anchor_copy(const Node* __restrict__ nodes, CandView* __restrict__ cviews, uint32_t n_slots, uint32_t c,
	const uint8_t* __restrict__ arena, uint8_t* __restrict__ store, const uint64_t* __restrict__ off, uint64_t base)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_slots) return;
	uint64_t at = base + off[i];
	for (uint32_t k = 0; k < nodes[i].ncand; ++k) {
		CandView& v = cviews[(size_t)i * c + k];
		const Anchor* src = reinterpret_cast<const Anchor*>(arena + v.anc);
		Anchor* dst = reinterpret_cast<Anchor*>(store) + at;
		for (uint32_t x = 0; x < v.n; ++x) dst[x] = src[x];
		v.anc = at * sizeof(Anchor);
		at += v.n;
	}
}

// E1-E4 for the reads [lo, hi): level-0 nodes slot_base.. and their anchors in the store
static clb_status anchor_batch(clb_ctx* c, const S2P& P, uint32_t lo, uint32_t hi, const std::vector<uint32_t>& h_cand_n, std::vector<uint32_t>& h_slot,
	uint64_t& n_slots, uint64_t& store_used)
{
	cudaStream_t s = c->stream;
	Scoped mem(s);
	Trace tr(s);
	std::vector<uint32_t> h_list;
	for (uint32_t r = lo; r < hi; ++r) if (!c->h_has_n[r] && h_cand_n[r]) { h_slot[r] = (uint32_t)(n_slots + h_list.size()); h_list.push_back(r); }
	const uint32_t nb = (uint32_t)h_list.size();
	if (!nb) return CLB_OK;
	uint32_t* d_list = nullptr; SegInfo* d_seg = nullptr; uint32_t* d_slot_dec = nullptr; unsigned long long* d_cursor = nullptr; uint32_t* d_acnt = nullptr; uint64_t* d_aoff = nullptr;
	CLB_CUDA(c, mem.get(&d_list, nb)); CLB_CUDA(c, mem.get(&d_slot_dec, nb)); CLB_CUDA(c, mem.get(&d_cursor, 2));
	{	// segment records: kept from batch to batch (ctx.h), with headroom
		const uint64_t seg_bytes = sizeof(SegInfo) * (uint64_t)nb * P.c * 2 + 16;
		if (seg_bytes > c->s2_segs.cap) CLB_CUDA(c, c->s2_segs.reserve(seg_bytes + seg_bytes / 4, s, false));
		d_seg = reinterpret_cast<SegInfo*>(c->s2_segs.p);
	}
	CLB_CUDA(c, mem.get(&d_acnt, nb)); CLB_CUDA(c, mem.get(&d_aoff, nb));
	CLB_CUDA(c, cudaMemcpyAsync(d_list, h_list.data(), sizeof(uint32_t) * nb, cudaMemcpyHostToDevice, s));
	CLB_CUDA(c, cudaMemsetAsync(d_slot_dec, 0, sizeof(uint32_t) * nb, s));
	uint64_t slots_job = n_slots + nb;
	if (n_slots == 0) {      // first batch: nodes and candidate views for every read of the job that will get a slot (no regrowth copies)
		slots_job = 0;
		for (uint64_t r = c->n_context; r < c->n_reads; ++r) slots_job += !c->h_has_n[r] && h_cand_n[r];
	}
	CLB_CUDA(c, c->s2_nodes.reserve(std::max<uint64_t>(slots_job, n_slots + nb), s, true, n_slots));
	CLB_CUDA(c, c->s2_cviews.reserve(std::max<uint64_t>(slots_job, n_slots + nb) * P.c, s, true, n_slots * P.c));
	Node* nodes = c->s2_nodes.p + n_slots; CandView* cviews = c->s2_cviews.p + n_slots * P.c;
	tr.mark("anchors: setup");
	clb_status st = s2_anchors(c, P, h_list, d_list, c->d_ref_to_read, c->s2_arena, d_seg, d_slot_dec, nodes, cviews, d_cursor);
	if (st != CLB_OK) return st;
	tr.mark("anchors: s2_anchors");
	if (c->keep_candidates) { st = dump_candidates(c, P, h_list, nodes, cviews); if (st != CLB_OK) return st; }
	CLB_TIMED(c, K_ANCHORS, (k_anchor_count<<<(nb + 127) / 128, 128, 0, s>>>(nodes, cviews, nb, P.c, d_acnt)));
	CLB_LAUNCH_CHECK(c, "k_anchor_count");
	uint64_t total = 0;
	st = exclusive_scan(c, d_acnt, nb, d_aoff, &total); if (st != CLB_OK) return st;
	uint64_t want = store_used + total;
	if (store_used == 0 && hi < c->n_reads) {      // first batch: size the store for the whole job at this batch's anchor density (no regrowth copies)
		uint64_t b_batch = 0, b_rest = 0;
		for (uint32_t r = lo; r < hi; ++r) b_batch += c->h_rd_len[r];
		for (uint64_t r = hi; r < c->n_reads; ++r) b_rest += c->h_rd_len[r];
		want = total + (uint64_t)(1.15 * (double)total * ((double)b_rest / (double)std::max<uint64_t>(1, b_batch))) + 4096;
	}
	CLB_CUDA(c, c->s2_store.reserve(want * sizeof(Anchor) + 16, s, true, store_used * sizeof(Anchor)));
	CLB_TIMED(c, K_ANCHORS, (k_anchor_copy<<<(nb + 127) / 128, 128, 0, s>>>(nodes, cviews, nb, P.c, c->s2_arena.p, c->s2_store.p, d_aoff, store_used)));
	CLB_LAUNCH_CHECK(c, "k_anchor_copy");
	CLB_CUDA(c, cudaStreamSynchronize(s));
	n_slots += nb; store_used += total;
	tr.mark("anchors");
	return CLB_OK;
}

// E6-E9 over all reads at once (packs [0, np)): level waves, estimator, tuples
static clb_status encode_all(clb_ctx* c, const S2P& P, const std::vector<uint32_t>& pack_first, const std::vector<uint32_t>& h_slot, uint64_t n_slots)
{
	cudaStream_t s = c->stream;
	const uint32_t pack_lo = 0, pack_hi = (uint32_t)pack_first.size() - 1;
	const uint32_t lo = (uint32_t)c->n_context, hi = (uint32_t)c->n_reads, nr = hi - lo;
	if (!nr) return CLB_OK;
	Scoped mem(s);
	Trace tr(s);
	const ReadStore R{c->pk.p, c->rd_start.p, c->rd_len.p, c->nmask.p, c->d_ref_to_read};
	const uint8_t* const arena = c->s2_store.p;
	uint32_t* d_slot = nullptr; uint32_t* d_pack_first = nullptr; unsigned long long* d_cursor = nullptr; BinStats* d_bins = nullptr;
	CLB_CUDA(c, mem.get(&d_slot, nr)); CLB_CUDA(c, mem.get(&d_pack_first, pack_hi - pack_lo + 1));
	CLB_CUDA(c, mem.get(&d_cursor, 2)); CLB_CUDA(c, mem.get(&d_bins, 1));
	CLB_CUDA(c, cudaMemcpyAsync(d_slot, h_slot.data() + lo, sizeof(uint32_t) * nr, cudaMemcpyHostToDevice, s));
	CLB_CUDA(c, cudaMemcpyAsync(d_pack_first, pack_first.data() + pack_lo, sizeof(uint32_t) * (pack_hi - pack_lo + 1), cudaMemcpyHostToDevice, s));

	DevBuf<Node>& nodes = c->s2_nodes; DevBuf<CandView>& cviews = c->s2_cviews; DevBuf<Task>& tasks = c->s2_tasks;
	// scripts: one exactly-sized buffer per level wave, addressed absolutely (Task::es_off holds a device address), so that no
	// wave has to copy the earlier ones into a bigger buffer
	struct { char* p; } esbuf{nullptr};
	uint64_t n_nodes = n_slots, n_tasks = 0;
	CLB_CUDA(c, nodes.reserve(std::max<uint64_t>(n_nodes, 1), s, true, n_nodes));
	CLB_CUDA(c, cviews.reserve(std::max<uint64_t>(n_nodes * P.c, 1), s, true, n_nodes * P.c));
	clb_status st = CLB_OK;

	// ---- level waves ----
	uint64_t n0 = 0;
	for (uint32_t level = 0; n0 < n_nodes; ++level) {
		const uint64_t n1 = n_nodes, nn = n1 - n0;
		uint32_t* d_cnt = nullptr; uint32_t* d_capu = nullptr; uint64_t* d_toff = nullptr; uint64_t* d_coff = nullptr;
		Scoped lvl(s);
		CLB_CUDA(c, lvl.get(&d_cnt, nn)); CLB_CUDA(c, lvl.get(&d_capu, nn)); CLB_CUDA(c, lvl.get(&d_toff, nn)); CLB_CUDA(c, lvl.get(&d_coff, nn));
		const uint32_t nblk = (uint32_t)((nn + 127) / 128);
		CLB_TIMED(c, K_ENCODE, (k_tasks<false><<<nblk, 128, 0, s>>>(nodes.p, nodes.p, cviews.p, (uint32_t)n0, (uint32_t)n1, P.c, R, arena, d_cnt, d_capu, nullptr, nullptr, 0, 0, nullptr)));
		CLB_LAUNCH_CHECK(c, "k_tasks<count>");
		uint64_t nt = 0, ncap = 0;
		st = exclusive_scan(c, d_cnt, nn, d_toff, &nt); if (st != CLB_OK) return st;
		st = exclusive_scan(c, d_capu, nn, d_coff, &ncap); if (st != CLB_OK) return st;
		tr.mark("level: count + scans");
		if (n_tasks + nt >= 0xFFFFFFF0ull) return fail(c, CLB_ERR_CAPACITY, "more than 2^32 parts in one batch: lower CLB_BATCH_MBASES");
		CLB_CUDA(c, tasks.reserve(n_tasks + nt, s, true, n_tasks));
		tr.mark("level: tasks.reserve");
		char* level_buf = nullptr;
		CLB_CUDA(c, mem.get(&level_buf, ncap * 4 + 16));
		tr.mark("level: script buffer");
		const uint64_t es_used = reinterpret_cast<uint64_t>(level_buf);
		CLB_TIMED(c, K_ENCODE, (k_tasks<true><<<nblk, 128, 0, s>>>(nodes.p, nodes.p, cviews.p, (uint32_t)n0, (uint32_t)n1, P.c, R, arena, nullptr, nullptr, d_toff, d_coff, n_tasks, es_used, tasks.p)));
		CLB_LAUNCH_CHECK(c, "k_tasks<fill>");
		tr.mark("level: task lists");
		st = align_level(c, P, tasks.p, n_tasks, n_tasks + nt, R, nodes.p, cviews.p, esbuf.p, d_bins);
		if (st != CLB_OK) return st;
		tr.mark("level: align");
		// decisions; children are appended to the node array
		const uint64_t cap_nodes = n_nodes + nt;
		if (cap_nodes >= 0xFFFFFFF0ull) return fail(c, CLB_ERR_CAPACITY, "too many nodes in one batch");
		CLB_CUDA(c, nodes.reserve(cap_nodes, s, true, n_nodes));
		CLB_CUDA(c, cviews.reserve(cap_nodes * P.c, s, true, n_nodes * P.c));
		unsigned int cur = (unsigned int)n_nodes;
		CLB_CUDA(c, cudaMemcpyAsync(d_cursor, &cur, sizeof(cur), cudaMemcpyHostToDevice, s));
		if (nt) {
			DecideArgs da{tasks.p, n_tasks, n_tasks + nt, nodes.p, cviews.p, reinterpret_cast<unsigned int*>(d_cursor), (uint32_t)cap_nodes, arena, R, esbuf.p, P};
			CLB_TIMED(c, K_DECIDE, (k_decide<<<(uint32_t)((nt + 127) / 128), 128, 0, s>>>(da)));
			CLB_LAUNCH_CHECK(c, "k_decide");
		}
		CLB_CUDA(c, cudaMemcpyAsync(&cur, d_cursor, sizeof(cur), cudaMemcpyDeviceToHost, s));
		CLB_CUDA(c, cudaStreamSynchronize(s));
		tr.mark("level: decide");
		n_tasks += nt;
		n0 = n1; n_nodes = std::min<uint64_t>(cur, cap_nodes);
		if (level > P.max_rec + 1) break;
	}

	// ---- adaptive estimator: one warp per pack over the short parts listed in the reference's order ----
	const uint32_t np = pack_hi - pack_lo;
	uint4* d_hist = nullptr; uint32_t* d_pcnt = nullptr; uint64_t* d_poff = nullptr; uint32_t* d_pend = nullptr;
	CLB_CUDA(c, mem.get(&d_hist, nr)); CLB_CUDA(c, mem.get(&d_pcnt, nr)); CLB_CUDA(c, mem.get(&d_poff, nr));
	PackArgs pa{d_pack_first, np, lo, nr, d_slot, c->d_has_n, tasks.p, nodes.p, esbuf.p, R, d_hist, d_pcnt, d_poff, nullptr};
	CLB_TIMED(c, K_ESTIMATE, (k_read_hist<<<(nr + 127) / 128, 128, 0, s>>>(pa)));
	CLB_LAUNCH_CHECK(c, "k_read_hist");
	CLB_TIMED(c, K_ESTIMATE, (k_pending<false><<<(nr + 127) / 128, 128, 0, s>>>(pa)));
	CLB_LAUNCH_CHECK(c, "k_pending<count>");
	uint64_t n_pend = 0;
	st = exclusive_scan(c, d_pcnt, nr, d_poff, &n_pend); if (st != CLB_OK) return st;
	CLB_CUDA(c, mem.get(&d_pend, n_pend));
	pa.pend = d_pend;
	CLB_TIMED(c, K_ESTIMATE, (k_pending<true><<<(nr + 127) / 128, 128, 0, s>>>(pa)));
	CLB_LAUNCH_CHECK(c, "k_pending<fill>");
	CLB_TIMED(c, K_ESTIMATE, (k_estimate<<<np, 32, 0, s>>>(pa)));
	CLB_LAUNCH_CHECK(c, "k_estimate");
	tr.mark("estimator");

	// ---- tuples ----
	uint32_t* d_size = nullptr; uint32_t* d_kind = nullptr; uint64_t* d_off = nullptr;
	CLB_CUDA(c, mem.get(&d_size, nr)); CLB_CUDA(c, mem.get(&d_kind, nr)); CLB_CUDA(c, mem.get(&d_off, nr));
	EmitArgs ea{lo, nr, d_slot, c->d_has_n, tasks.p, nodes.p, cviews.p, P.c, arena, esbuf.p, R, d_size, d_kind, d_off, c->es_total, nullptr, c->es_off};
	static const bool emit_warp = std::getenv("CLB_EMIT_THREAD") == nullptr;   // warp per read (k_emit_w: 219 ms per 25 Gbases on B200 against 825 ms); CLB_EMIT_THREAD=1 keeps the thread-per-read walk
	const uint32_t emit_warp_blocks = (uint32_t)(((uint64_t)nr * 32 + 127) / 128);
	if (emit_warp) CLB_TIMED(c, K_EMIT, (k_emit_w<false><<<emit_warp_blocks, 128, 0, s>>>(ea)));
	else CLB_TIMED(c, K_EMIT, (k_emit<false><<<(nr + 127) / 128, 128, 0, s>>>(ea)));
	CLB_LAUNCH_CHECK(c, "k_emit<size>");
	uint64_t total = 0;
	st = exclusive_scan(c, d_size, nr, d_off, &total); if (st != CLB_OK) return st;
	CLB_CUDA(c, c->es.reserve(c->es_total + total + 16, s, true, c->es_total));
	ea.out = c->es.p;
	if (emit_warp) CLB_TIMED(c, K_EMIT, (k_emit_w<true><<<emit_warp_blocks, 128, 0, s>>>(ea)));
	else CLB_TIMED(c, K_EMIT, (k_emit<true><<<(nr + 127) / 128, 128, 0, s>>>(ea)));
	CLB_LAUNCH_CHECK(c, "k_emit<write>");
	CLB_TIMED(c, K_EMIT, (k_emit_plain<<<nr, 256, 0, s>>>(ea)));
	CLB_LAUNCH_CHECK(c, "k_emit_plain");
	if (c->collect_stats && c->d_stats) {
		StatsArgs sa{tasks.p, n_tasks, nodes.p, n_nodes, cviews.p, P.c, arena, esbuf.p, d_kind, c->rd_len.p, lo, nr, c->d_stats};
		if (n_tasks) { k_stats_tasks<<<(uint32_t)((n_tasks + 127) / 128), 128, 0, s>>>(sa); CLB_LAUNCH_CHECK(c, "k_stats_tasks"); }
		if (n_nodes) { k_stats_nodes<<<(uint32_t)((n_nodes + 127) / 128), 128, 0, s>>>(sa); CLB_LAUNCH_CHECK(c, "k_stats_nodes"); }
		k_stats_reads<<<(nr + 127) / 128, 128, 0, s>>>(sa); CLB_LAUNCH_CHECK(c, "k_stats_reads");
	}
	CLB_CUDA(c, cudaStreamSynchronize(s));
	tr.mark("emit");
	c->es_total += total;
	return CLB_OK;
}

clb_status s2_encode(clb_ctx* c, const clb_encode_params* prm, const uint32_t* pack_sizes, uint32_t n_packs)
{
	if (!c->graph_done) return fail(c, CLB_ERR_STATE, "clb_encode before clb_graph_build");
	if (c->enc_done) return fail(c, CLB_ERR_STATE, "clb_encode called twice");
	if (prm->anchor_len < 8 || prm->anchor_len > 32) return fail(c, CLB_ERR_BAD_ARG, "anchor_len must be in [8, 32]");
	if (prm->max_recurence > 7) return fail(c, CLB_ERR_BAD_ARG, "max_recurence above 7 is not supported");
	cudaStream_t s = c->stream;
	Trace tr(s);
	const uint64_t n = c->n_reads;
	if (c->collect_stats) {
		if (!c->d_stats) CLB_CUDA(c, dev_malloc((void**)&c->d_stats, sizeof(unsigned long long) * ST_COUNT, s));
		CLB_CUDA(c, cudaMemsetAsync(c->d_stats, 0, sizeof(unsigned long long) * ST_COUNT, s));
	}
	S2P P{prm->anchor_len, prm->min_part_len_alt, prm->max_recurence, prm->min_anchors, c->prm.max_candidates,
		prm->min_mmer_frac, prm->min_mmer_force, prm->max_matches_mult, prm->es_cost_mult};
	// packs
	const uint64_t nc = c->n_context;                 // context reads are not encoded: packs cover the reads after them
	std::vector<uint32_t> pack_first{(uint32_t)nc};
	if (pack_sizes) {
		uint64_t at = nc;
		for (uint32_t i = 0; i < n_packs; ++i) { at += pack_sizes[i]; if (pack_sizes[i]) pack_first.push_back((uint32_t)at); }
		if (at != n) return fail(c, CLB_ERR_BAD_ARG, "pack_sizes do not sum to the number of reads");
	} else {
		uint64_t bytes = 0;
		for (uint64_t i = nc; i < n; ++i) {           // in_reads.cpp:62-76
			bytes += (uint64_t)c->h_rd_len[i] + 1;
			if (bytes >= (2u << 21)) { bytes = 0; pack_first.push_back((uint32_t)(i + 1)); }
		}
		if (pack_first.back() != n) pack_first.push_back((uint32_t)n);
	}
	const uint32_t np = (uint32_t)pack_first.size() - 1;
	// reference id -> read
	std::vector<uint32_t> r2r(c->n_ref ? c->n_ref : 1);
	for (uint64_t i = 0, k = 0; i < n; ++i) if (c->h_is_ref[i]) r2r[k++] = (uint32_t)i;
	CLB_CUDA(c, dmalloc(&c->d_ref_to_read, r2r.size(), c->stream));
	CLB_CUDA(c, cudaMemcpyAsync(c->d_ref_to_read, r2r.data(), sizeof(uint32_t) * r2r.size(), cudaMemcpyHostToDevice, s));
	std::vector<uint32_t> h_cand_n(n ? n : 1);
	if (n) CLB_CUDA(c, cudaMemcpyAsync(h_cand_n.data(), c->cand_n, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, s));
	CLB_CUDA(c, cudaStreamSynchronize(s));
	CLB_CUDA(c, dmalloc(&c->es_off, n + 1, c->stream));
	if (nc) CLB_CUDA(c, cudaMemsetAsync(c->es_off, 0, sizeof(uint64_t) * nc, s));
	// (the tuple buffer is allocated when its exact size is known, after the sizing pass of k_emit)
	c->es_total = 0;
	if (c->keep_candidates) c->dbg_cand.assign(n, std::vector<uint32_t>());
	// anchors batch by batch (the pair arena is the big transient), everything after them over all reads at once
	const char* env_batch = std::getenv("CLB_BATCH_MBASES");
	const uint64_t batch_bases = (env_batch ? (uint64_t)std::atoll(env_batch) : 1024) << 20;
	std::vector<uint32_t> h_slot(n ? n : 1, 0xFFFFFFFFu);
	uint64_t n_slots = 0, store_used = 0;
	tr.mark("encode: setup");
	for (uint32_t lo = (uint32_t)nc; lo < n;) {
		uint32_t hi = lo; uint64_t bases = 0;
		while (hi < n && (hi == lo || bases + c->h_rd_len[hi] <= batch_bases)) bases += c->h_rd_len[hi++];
		clb_status st = anchor_batch(c, P, lo, hi, h_cand_n, h_slot, n_slots, store_used);
		if (st != CLB_OK) return st;
		lo = hi;
	}
	c->s2_arena.release(); c->s2_segs.release(); c->s2_gtab.release(); c->s2_gbloom.release();
	{
		clb_status st = encode_all(c, P, pack_first, h_slot, n_slots);
		if (st != CLB_OK) return st;
	}
	CLB_CUDA(c, cudaMemcpyAsync(c->es_off + n, &c->es_total, sizeof(uint64_t), cudaMemcpyHostToDevice, s));
	CLB_CUDA(c, cudaStreamSynchronize(s));
	c->s2_arena.release(); c->s2_store.release(); c->s2_scratch.release(); c->s2_budget = 0; c->s2_nodes.release(); c->s2_cviews.release(); c->s2_tasks.release(); c->s2_esbuf.release();
	c->s2_segs.release(); c->s2_gtab.release(); c->s2_gbloom.release();
	tr.mark("encode: release");
	if (c->collect_stats) {
		unsigned long long h[ST_COUNT];
		CLB_CUDA(c, cudaMemcpyAsync(h, c->d_stats, sizeof h, cudaMemcpyDeviceToHost, s));
		CLB_CUDA(c, cudaStreamSynchronize(s));
		clb_encode_stats& o = c->h_stats;
		o = clb_encode_stats{};
		o.n_not_enough_unique_mmers_in_enc_read = h[ST_NOT_ENOUGH]; o.n_too_many_matches = h[ST_TOO_MANY]; o.n_too_low_anchors = h[ST_TOO_LOW];
		o.n_non_rev_choosen = h[ST_NON_REV]; o.n_rev_choosen = h[ST_REV];
		o.n_plain_reads_tot = h[ST_PLAIN_READS]; o.n_plain_symb = h[ST_PLAIN_SYMB]; o.n_plain_reads_with_n_tot = h[ST_PLAIN_N_READS]; o.n_plain_with_n_symb = h[ST_PLAIN_N_SYMB];
		o.n_levels = (uint32_t)std::min<unsigned long long>(h[ST_MAX_LEVEL], CLB_MAX_STAT_LEVELS);
		for (uint32_t l = 0; l < ST_LEVELS; ++l) {
			const unsigned long long* L = h + ST_LEVEL0 + l * ST_LEVEL_FIELDS;
			clb_level_stats& d = o.level[l];
			d.n_alternative_left_flank = L[SL_ALT_LEFT]; d.n_alternative_in_between = L[SL_ALT_BETWEEN]; d.n_alternative_right_flank = L[SL_ALT_RIGHT];
			d.n_plain_symbols = L[SL_PLAIN_SYMB]; d.n_symb_coded_with_edit_script = L[SL_CODED_SYMB]; d.n_edit_script_symbols = L[SL_ES_SYMB];
			d.n_substitution = L[SL_SUBST]; d.n_match = L[SL_MATCH]; d.n_insertion = L[SL_INS]; d.n_deletion = L[SL_DEL];
			d.n_symb_anchors = L[SL_ANCHOR_SYMB]; d.n_anchors = L[SL_ANCHORS]; d.n_left_flank_symb = L[SL_LEFT_FLANK]; d.n_right_flank_symb = L[SL_RIGHT_FLANK];
		}
	}
	c->enc_done = true;
	return CLB_OK;
}

void s2_free(clb_ctx* c)
{
	for (int i = 0; i < 16; ++i) { if (c->s2_streams[i]) cudaStreamDestroy(c->s2_streams[i]); if (c->s2_join[i]) cudaEventDestroy(c->s2_join[i]); c->s2_streams[i] = nullptr; c->s2_join[i] = nullptr; }
	if (c->s2_fork) cudaEventDestroy(c->s2_fork);
	c->s2_fork = nullptr;
	c->es.release(); c->s2_arena.release(); c->s2_store.release(); c->s2_scratch.release(); c->s2_budget = 0; c->s2_nodes.release(); c->s2_cviews.release(); c->s2_tasks.release(); c->s2_esbuf.release();
	c->s2_segs.release(); c->s2_gtab.release(); c->s2_gbloom.release();
	if (c->d_stats) dev_free(c->d_stats, c->stream);
	c->d_stats = nullptr;
	if (c->es_off) dev_free(c->es_off, c->stream);
	if (c->d_ref_to_read) dev_free(c->d_ref_to_read, c->stream);
	c->es_off = nullptr; c->d_ref_to_read = nullptr;
}

} // namespace clb
