/* oracle/stage2.c — TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the reference's stage 2 (per-read anchors + edit script -> CompactES tuples).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it; the product never does.
 * Pinned against the reference: tests/golden/<case>/es.bin.gz holds the CompactES bytes of every read as the unmodified
 * CEncoder emitted them (oracle/ref_stage_dump.cpp); tests/test_oracle_stage2.py replays them.
 *
 * Paths are relative to /root/reference/src/colord.  The restatement follows the reference's RESULTS, not its
 * data structures: hash maps become sorted arrays, edlib's banded Myers becomes a plain DP — the places where the
 * reference's result depends on an implementation choice (LIS tie-breaking, edlib's traceback preference and its
 * Hirschberg split rule, std::sort being stable for <= 16 elements) are restated as such and commented.
 */
#include <stdint.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <math.h>

typedef struct {
	uint32_t anchor_len;            /* -a                                                     */
	uint32_t kmer_len, modulo;      /* HiFi k-mer anchors                                     */
	uint32_t is_hifi;
	uint32_t min_part_len_alt;      /* minPartLenToConsiderAltRead                            */
	uint32_t max_recurence;
	uint32_t min_anchors;
	double min_mmer_frac;           /* minFractionOfMmersInEncode                             */
	double min_mmer_force;          /* minFractionOfMmersInEncodeToAlwaysEncode               */
	double max_matches_mult;        /* maxMatchesMultiplier                                   */
	double es_cost_mult;            /* editScriptCostMultiplier                               */
} orc_s2_params;

/* ------------------------------------------------------------------------------------------------ buffers */
typedef struct { char* p; size_t n, cap; } sbuf;
static void sb_reserve(sbuf* s, size_t extra) { if (s->n + extra > s->cap) { s->cap = (s->n + extra) * 2 + 64; s->p = (char*)realloc(s->p, s->cap); } }
static void sb_push(sbuf* s, char c) { sb_reserve(s, 1); s->p[s->n++] = c; }
static void sb_fill(sbuf* s, char c, size_t k) { sb_reserve(s, k); memset(s->p + s->n, c, k); s->n += k; }
static void sb_append(sbuf* s, const char* q, size_t k) { sb_reserve(s, k); memcpy(s->p + s->n, q, k); s->n += k; }
static void sb_free(sbuf* s) { free(s->p); s->p = NULL; s->n = s->cap = 0; }

/* ------------------------------------------------------------------------------------------------ E6: alignment */
/* utils.h:281-364 (CMissmatchCoder::encode_missmatch_symb): the three other bases in ACGT order are X, Y, Z */
static char mismatch_symb(uint8_t ref, uint8_t enc)
{
	static const char mm[4][4] = { {'M','X','Y','Z'}, {'X','M','Y','Z'}, {'X','Y','M','Z'}, {'X','Y','Z','M'} };
	return mm[ref][enc];
}

/* Exact unit-cost DP of rows (query) x cols (target) and the traceback edlib performs
 * (libs/edlib/edlib.cpp:945-1159, obtainAlignmentTraceback): from the last cell prefer UP (consume a query symbol,
 * EDLIB_EDOP_INSERT=1), then LEFT (consume a target symbol, EDLIB_EDOP_DELETE=2), then the diagonal (0 match / 3
 * mismatch); on reaching a border the rest is all LEFT or all UP.  edlib's band (k = best score) only drops cells that
 * no optimal path visits, so the moves equal those on the full matrix.  ops are produced in forward order. */
static void traceback_exact(const uint8_t* q, int qlen, const uint8_t* t, int tlen, uint8_t* ops, int* n_ops)
{
	const size_t W = (size_t)tlen + 1;
	uint32_t* D = (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)qlen + 1) * W);
	for (int j = 0; j <= tlen; ++j) D[j] = (uint32_t)j;
	for (int i = 1; i <= qlen; ++i)
	{
		uint32_t* row = D + (size_t)i * W; const uint32_t* up = row - W;
		row[0] = (uint32_t)i;
		for (int j = 1; j <= tlen; ++j)
		{
			uint32_t a = up[j] + 1, b = row[j - 1] + 1, c = up[j - 1] + (q[i - 1] != t[j - 1]);
			uint32_t m = a < b ? a : b; row[j] = m < c ? m : c;
		}
	}
	int i = qlen, j = tlen, n = 0;
	uint8_t* rev = (uint8_t*)malloc((size_t)qlen + tlen + 2);
	while (i > 0 || j > 0)
	{
		if (i == 0) { rev[n++] = 2; --j; continue; }
		if (j == 0) { rev[n++] = 1; --i; continue; }
		const uint32_t cur = D[(size_t)i * W + j];
		if (D[(size_t)(i - 1) * W + j] + 1 == cur) { rev[n++] = 1; --i; }
		else if (D[(size_t)i * W + j - 1] + 1 == cur) { rev[n++] = 2; --j; }
		else { rev[n++] = (D[(size_t)(i - 1) * W + j - 1] == cur) ? 0 : 3; --i; --j; }
	}
	for (int x = 0; x < n; ++x) ops[x] = rev[n - 1 - x];
	*n_ops = n;
	free(rev); free(D);
}

/* last column of the forward DP: F[y] = dist(q[0..y), t[0..tlen)) for y = 0..qlen */
static void dp_last_column(const uint8_t* q, int qlen, const uint8_t* t, int tlen, int reverse, uint32_t* F)
{
	/* reverse != 0: both sequences are read backwards (q[qlen-1-i], t[tlen-1-j]) */
	uint32_t* col = (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)qlen + 1));
	for (int y = 0; y <= qlen; ++y) col[y] = (uint32_t)y;
	for (int j = 1; j <= tlen; ++j)
	{
		const uint8_t tc = reverse ? t[tlen - j] : t[j - 1];
		uint32_t diag = col[0]; col[0] = (uint32_t)j;
		for (int y = 1; y <= qlen; ++y)
		{
			const uint8_t qc = reverse ? q[qlen - y] : q[y - 1];
			uint32_t a = col[y] + 1, b = col[y - 1] + 1, c = diag + (qc != tc);
			diag = col[y];
			uint32_t m = a < b ? a : b; col[y] = m < c ? m : c;
		}
	}
	memcpy(F, col, sizeof(uint32_t) * ((size_t)qlen + 1));
	free(col);
}

/* edlib.cpp:1178-1215 (obtainAlignment) + :1234-1377 (obtainAlignmentHirschberg).  The traceback is used while edlib's
 * estimate of the stored-column memory is below 1 MiB; above it the target is halved and the split vertex is the
 * TOPMOST row y in 1..qlen-1 with F[y] + R[y] == best, then y = 0, then y = qlen (the order edlib searches). */
static void edlib_path(const uint8_t* q, int qlen, const uint8_t* t, int tlen, int best, uint8_t* ops, int* n_ops)
{
	if (qlen == 0 || tlen == 0)
	{
		for (int i = 0; i < qlen + tlen; ++i) ops[i] = qlen == 0 ? 2 : 1;
		*n_ops = qlen + tlen; return;
	}
	const long long blocks = (qlen + 63) / 64;
	const long long est = (2ll * 8 + 4) * blocks * tlen + 2ll * 4 * tlen;
	if (est < 1024 * 1024) { traceback_exact(q, qlen, t, tlen, ops, n_ops); return; }
	const int lw = tlen / 2, rw = tlen - lw;
	uint32_t* F = (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)qlen + 1));
	uint32_t* Rr = (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)qlen + 1));
	dp_last_column(q, qlen, t, lw, 0, F);
	dp_last_column(q, qlen, t + lw, rw, 1, Rr);     /* Rr[z] = dist(last z of q, t[lw..)) ; R[y] = Rr[qlen - y] */
	int y = -1;
	for (int c = 1; c <= qlen - 1; ++c) if ((int)(F[c] + Rr[qlen - c]) == best) { y = c; break; }
	if (y < 0 && (int)(lw + Rr[qlen]) == best) y = 0;
	if (y < 0 && (int)(F[qlen] + rw) == best) y = qlen;
	if (y < 0) { *n_ops = 0; free(F); free(Rr); return; }   /* cannot happen for a correct best */
	const int ls = y == 0 ? lw : (int)F[y], rs = y == qlen ? rw : (int)Rr[qlen - y];
	free(F); free(Rr);
	int n1 = 0, n2 = 0;
	edlib_path(q, y, t, lw, ls, ops, &n1);
	edlib_path(q + y, qlen - y, t + lw, rw, rs, ops + n1, &n2);
	*n_ops = n1 + n2;
}

/* edit_script.h:156-245 (find_edit_dist): plain DP, rows = ref, cols = enc, traceback prefers 'D', then insertion, then
 * the diagonal — the same moves as traceback_exact with q = ref, t = enc. */
static uint32_t script_rows_ref(const uint8_t* ref, int rl, const uint8_t* enc, int el, int use_edlib_path, sbuf* es)
{
	uint8_t* ops = (uint8_t*)malloc((size_t)rl + el + 2); int n = 0;
	if (use_edlib_path)
	{
		uint32_t* F = (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)rl + 1));
		dp_last_column(ref, rl, enc, el, 0, F);
		const int best = (int)F[rl]; free(F);
		edlib_path(ref, rl, enc, el, best, ops, &n);
	}
	else traceback_exact(ref, rl, enc, el, ops, &n);
	uint32_t dist = 0; int pr = 0, pe = 0;
	for (int i = 0; i < n; ++i)
	{
		switch (ops[i])
		{	/* edit_script.h:296-318: query = ref, target = enc */
		case 0: sb_push(es, 'M'); ++pr; ++pe; break;
		case 1: sb_push(es, 'D'); ++pr; ++dist; break;
		case 2: sb_push(es, "ACGT"[enc[pe++]]); ++dist; break;
		default: sb_push(es, mismatch_symb(ref[pr], enc[pe])); ++pr; ++pe; ++dist; break;
		}
	}
	free(ops);
	return dist;
}

/* edit_script.h:272-331 (find_edit_dist_with_edlib_ex, NW): tiny inputs go through find_edit_dist, the rest through
 * edlibAlign(query = ref, target = enc).  Both give the same moves; only edlib switches to Hirschberg for big inputs. */
static uint32_t script_nw(const uint8_t* ref, int rl, const uint8_t* enc, int el, sbuf* es)
{
	const int tiny = rl < 2 || el < 2 || (rl < 15 && el < 15);
	return script_rows_ref(ref, rl, enc, el, !tiny, es);
}

/* edit_script.h:333-398 (find_edit_dist_with_edlib_ex_odwr, SHW): edlibAlign(query = enc, target = ref prefix): best
 * score over all target prefixes, leftmost end (edlib.cpp:660-674), then the NW path of enc vs ref[0..end].
 * Moves: UP consumes enc (insertion), LEFT consumes ref ('D').  Inputs shorter than 2 fall back to the global
 * find_edit_dist(ref, enc) with ref_end = |ref| - 1. */
static uint32_t script_shw(const uint8_t* ref, int rl, const uint8_t* enc, int el, uint32_t* ref_end, sbuf* es)
{
	if (rl < 2 || el < 2) { *ref_end = (uint32_t)(rl - 1); return script_rows_ref(ref, rl, enc, el, 0, es); }
	/* last row of the DP with rows = enc, cols = ref: score of enc against every prefix of ref */
	uint32_t* row = (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)rl + 1));
	uint32_t* col = (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)el + 1));
	for (int y = 0; y <= el; ++y) col[y] = (uint32_t)y;
	row[0] = (uint32_t)el;
	for (int j = 1; j <= rl; ++j)
	{
		uint32_t diag = col[0]; col[0] = (uint32_t)j;
		for (int y = 1; y <= el; ++y)
		{
			uint32_t a = col[y] + 1, b = col[y - 1] + 1, c = diag + (enc[y - 1] != ref[j - 1]);
			diag = col[y];
			uint32_t m = a < b ? a : b; col[y] = m < c ? m : c;
		}
		row[j] = col[el];
	}
	int end = 1; uint32_t best = row[1];
	for (int j = 2; j <= rl; ++j) if (row[j] < best) { best = row[j]; end = j; }   /* columns 0..rl-1 <-> prefixes of length 1..rl */
	free(row); free(col);
	/* edlib pads the query to a multiple of 64 with wildcards and reads the score of target prefix c - W in column c (edlib.cpp:660-674,
	 * :683-694), so the EMPTY prefix (position -1, score |enc|) takes part, first in the list: when no prefix beats inserting the whole
	 * query, endLocations[0] = -1 and the path is |enc| insertions (ref_end wraps to 0xFFFFFFFF as in edit_script.h:352) */
	if (best >= (uint32_t)el) { *ref_end = 0xFFFFFFFFu; for (int i = 0; i < el; ++i) sb_push(es, "ACGT"[enc[i]]); return (uint32_t)el; }
	*ref_end = (uint32_t)(end - 1);
	uint8_t* ops = (uint8_t*)malloc((size_t)end + el + 2); int n = 0;
	edlib_path(enc, el, ref, end, (int)best, ops, &n);
	int pr = 0, pe = 0;
	for (int i = 0; i < n; ++i)
	{
		switch (ops[i])
		{	/* edit_script.h:360-383: query = enc, target = ref */
		case 0: sb_push(es, 'M'); ++pr; ++pe; break;
		case 1: sb_push(es, "ACGT"[enc[pe++]]); break;
		case 2: sb_push(es, 'D'); ++pr; break;
		default: sb_push(es, mismatch_symb(ref[pr], enc[pe])); ++pr; ++pe; break;
		}
	}
	free(ops);
	return best;
}

/* ------------------------------------------------------------------------------------------------ E7: canonical form */
static int is_mm(char c) { return c == 'X' || c == 'Y' || c == 'Z'; }
static int is_ins(char c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }
/* edit_script.h:432-447 */
static void fix_in_range(char* es, uint64_t start, uint64_t end)
{
	if (end < start + 2) return;
	--end;
	for (;;)
	{
		while (start < end && es[start] == 'M') ++start;
		while (start < end && es[end] != 'M') --end;
		if (start == end) break;
		char t = es[start]; es[start] = es[end]; es[end] = t;
	}
}
/* edit_script.h:591-659.  ref/enc are pointers INTO the reads: the reference indexes one symbol past the part
 * (read_view::operator[] does not check), which lands on the next base of the read or on its 255 guard. */
static void refactor_edit_script(const uint8_t* ref, const uint8_t* enc, char* es, uint32_t n)
{
	uint32_t ref_start = 0, ref_pos = 0, es_start = 0;
	for (uint32_t p = 0; p < n; ++p)
	{
		const char s = es[p]; const int mm = is_mm(s), ins = is_ins(s);
		if (ins || mm || ref[ref_start] != ref[ref_pos])
		{
			fix_in_range(es, es_start, p);
			es_start = p; if (ins || mm) ++es_start;
			ref_start = ref_pos;
		}
		if (!ins) ++ref_pos;
	}
	fix_in_range(es, es_start, n);
	uint32_t enc_start = 0, enc_pos = 0; es_start = 0;
	for (uint32_t p = 0; p < n; ++p)
	{
		const char s = es[p]; const int mm = is_mm(s), del = s == 'D';
		if (del || mm || enc[enc_start] != enc[enc_pos])
		{
			fix_in_range(es, es_start, p);
			es_start = p; if (del || mm) ++es_start;
			enc_start = enc_pos;
		}
		if (!del) ++enc_pos;
	}
	fix_in_range(es, es_start, n);
}

/* encoder.cpp:1255-1283 (GetEditDist).  ref/enc point into the (oriented) reference read and the read being encoded.
 * kind: 0 = left flank (frag 0), 1 = right flank (last frag), 2 = between anchors.  Appends the script to `es`. */
static void get_edit_dist_inner(const uint8_t* ref, uint32_t rl, const uint8_t* enc, uint32_t el, int kind, sbuf* es);
/* ORC_ALIGN_LOG=<file>: one line per alignment problem (kind, reference symbols, read symbols, non-match script symbols) — the
 * problem-size statistics DESIGN.md quotes for the device aligner */
static void get_edit_dist(const uint8_t* ref, uint32_t rl, const uint8_t* enc, uint32_t el, int kind, sbuf* es)
{
	static FILE* logf = NULL; static int tried = 0;
	if (!tried) { tried = 1; const char* p = getenv("ORC_ALIGN_LOG"); if (p) logf = fopen(p, "w"); }
	const size_t at = es->n;
	get_edit_dist_inner(ref, rl, enc, el, kind, es);
	if (logf) { uint32_t d = 0; for (size_t i = at; i < es->n; ++i) d += es->p[i] != 'M'; fprintf(logf, "%d %u %u %u\n", kind, rl, el, d); fflush(logf); }
}
static void get_edit_dist_inner(const uint8_t* ref, uint32_t rl, const uint8_t* enc, uint32_t el, int kind, sbuf* es)
{
	if (rl == 0 || el == 0)
	{	/* edit_script.h:247-266 */
		if (rl == 0) for (uint32_t i = 0; i < el; ++i) sb_push(es, "ACGT"[enc[i]]);
		else sb_fill(es, 'D', rl);
		return;
	}
	const uint32_t max_flank = el * 2;
	if (kind == 0)
	{	/* edit_script.h:400-413: both reversed, ref cut to max_flank symbols, script reversed back */
		uint8_t* rr = (uint8_t*)malloc(rl + 1); uint8_t* re = (uint8_t*)malloc(el + 1);
		for (uint32_t i = 0; i < rl; ++i) rr[i] = ref[rl - 1 - i];
		for (uint32_t i = 0; i < el; ++i) re[i] = enc[el - 1 - i];
		const uint32_t cut = rl < max_flank ? rl : max_flank;
		sbuf t = {0, 0, 0}; uint32_t ref_end = 0;
		script_shw(rr, (int)cut, re, (int)el, &ref_end, &t);
		for (size_t i = 0; i < t.n / 2; ++i) { char c = t.p[i]; t.p[i] = t.p[t.n - 1 - i]; t.p[t.n - 1 - i] = c; }
		const uint32_t ref_offset = (rl - 1) - ref_end;
		refactor_edit_script(ref + ref_offset, enc, t.p, (uint32_t)t.n);
		sb_fill(es, 'D', ref_offset);
		sb_append(es, t.p, t.n);
		free(rr); free(re); sb_free(&t);
	}
	else if (kind == 1)
	{
		const uint32_t cut = rl < max_flank ? rl : max_flank;
		const size_t at = es->n; uint32_t ref_end;
		script_shw(ref, (int)cut, enc, (int)el, &ref_end, es);
		refactor_edit_script(ref, enc, es->p + at, (uint32_t)(es->n - at));
	}
	else
	{
		const size_t at = es->n;
		script_nw(ref, (int)rl, enc, (int)el, es);
		refactor_edit_script(ref, enc, es->p + at, (uint32_t)(es->n - at));
	}
}

/* exported for unit pinning: symbols 0..3 (guard byte after each buffer is the caller's duty) */
uint32_t orc_edit_script(const uint8_t* ref, uint32_t rl, const uint8_t* enc, uint32_t el, int kind, char* out, uint32_t cap)
{
	sbuf s = {0, 0, 0};
	get_edit_dist(ref, rl, enc, el, kind, &s);
	uint32_t n = (uint32_t)s.n;
	if (n <= cap) memcpy(out, s.p, n);
	sb_free(&s);
	return n;
}

/* ------------------------------------------------------------------------------------------------ E8: cost models */
/* utils.h:700-757 (CEntropy) */
static double entropy_dna(const uint8_t* s, size_t n)
{
	uint32_t h[4] = {0, 0, 0, 0};
	for (size_t i = 0; i < n; ++i) ++h[s[i]];
	double sum = 0; for (int i = 0; i < 4; ++i) sum += h[i];
	double rec = 1.0 / sum, e = 0;
	for (int c = 0; c < 4; ++c) if (h[c]) { double p = (double)h[c] * rec; e += log2(p) * p; }
	return -e;
}
static double entropy_es(const char* s, size_t n)
{
	static const uint8_t sym[11] = { 'A', 'C', 'D', 'G', 'M', 'T', 'X', 'Y', 'Z', 'S', 'R' };
	uint32_t h[128]; memset(h, 0, sizeof h);
	for (size_t i = 0; i < n; ++i) ++h[(uint8_t)s[i]];
	double sum = 0; for (int i = 0; i < 11; ++i) sum += h[sym[i]];
	double rec = 1.0 / sum, e = 0;
	for (int c = 0; c < 11; ++c) if (h[sym[c]]) { double p = (double)h[sym[c]] * rec; e += log2(p) * p; }
	return -e;
}
/* encoder.cpp:1315-1327 with :1300-1312 (skip >= 10 leading deletions) */
static int stateless_use_es(const char* es, size_t n, const uint8_t* enc, size_t el, double mult)
{
	size_t nd = 0; while (nd < n && es[nd] == 'D') ++nd;
	if (nd >= 10) { es += nd; n -= nd; }
	return entropy_es(es, n) * (double)n * mult < entropy_dna(enc, el) * (double)el;
}

/* utils.h:760-1126 (CEntropyEstimator): adaptive per pack */
typedef struct {
	uint32_t dna[4], es[12], dec[2];
	double dna_log[4], es_log[12], dec_log[2];
	uint32_t dna_sum, es_sum, dec_sum;
} estimator;
static void est_rescale(uint32_t* a, int n, uint32_t* sum, uint32_t max) { while (*sum > max) { *sum = 0; for (int i = 0; i < n; ++i) { a[i] = (a[i] + 1) / 2; *sum += a[i]; } } }
static void est_logs(const uint32_t* a, double* l, int n, uint32_t sum) { double rec = 1.0 / sum; for (int i = 0; i < n; ++i) l[i] = a[i] ? -log2((double)a[i] * rec) : 0.0; }
static void est_reset(estimator* e)
{
	for (int i = 0; i < 4; ++i) e->dna[i] = 1; e->dna_sum = 4;
	for (int i = 0; i < 12; ++i) e->es[i] = 1; e->es_sum = 12;
	for (int i = 0; i < 2; ++i) e->dec[i] = 1; e->dec_sum = 2;
	est_logs(e->dna, e->dna_log, 4, e->dna_sum); est_logs(e->es, e->es_log, 12, e->es_sum); est_logs(e->dec, e->dec_log, 2, e->dec_sum);
}
static void est_log_read(estimator* e, const uint8_t* r, size_t n)
{
	for (size_t i = 0; i < n; ++i) ++e->dna[r[i]];
	e->dna_sum += (uint32_t)n;
	est_rescale(e->dna, 4, &e->dna_sum, 1u << 20);
	est_logs(e->dna, e->dna_log, 4, e->dna_sum);
}
static uint64_t ilog2u(uint64_t x) { uint64_t r = 0; for (; x; ++r) x >>= 1; return r; }
static int es_code(char c) { switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; case 'D': return 4; case 'M': return 5; case 'X': return 6; case 'Y': return 7; case 'Z': return 8; case 'S': return 9; case 'R': return 10; default: return 11; } }
/* utils.h:1060-1126 */
static int est_use_es(estimator* e, const char* es, size_t n, const uint8_t* plain, size_t pl, size_t ref_len)
{
	uint32_t loc[12], rd[12], rp[4] = {0, 0, 0, 0}; uint32_t loc_sum = e->es_sum;
	memcpy(loc, e->es, sizeof loc); memset(rd, 0, sizeof rd);
	double es_cost = e->dec_log[0], plain_cost = e->dec_log[1];
	/* analyze_es (utils.h:838-899): runs of D >= 10 count as one skip 'S', runs of M >= 15 as one anchor 'R' */
	uint32_t* run_lens = (uint32_t*)malloc(sizeof(uint32_t) * (n + 1)); uint32_t n_runs = 0;
	{
		char c = ' '; uint32_t len = 0;
		for (size_t i = 0; i <= n; ++i)
		{
			char x = i < n ? es[i] : ' ';
			if (x == c) { ++len; continue; }
			if (c == 'D')
			{
				if (len >= 10) { ++loc[9]; ++loc_sum; ++rd[9]; run_lens[n_runs++] = len; }
				else { loc[4] += len; loc_sum += len; rd[4] += len; }
			}
			else if (c == 'M')
			{
				if (len >= 15) { ++loc[10]; ++loc_sum; ++rd[10]; run_lens[n_runs++] = len; }
				else { loc[5] += len; loc_sum += len; rd[5] += len; }
			}
			else if (c != ' ') { ++loc[es_code(c)]; ++loc_sum; ++rd[es_code(c)]; }
			c = x; len = 1;
		}
	}
	for (size_t i = 0; i < pl; ++i) ++rp[plain[i]];
	est_logs(loc, e->es_log, 12, loc_sum);                /* NB: the member logs are overwritten even if plain wins */
	for (int i = 0; i < 12; ++i) es_cost += rd[i] * e->es_log[i];
	for (uint32_t i = 0; i < n_runs; ++i) es_cost += (double)(ilog2u(run_lens[i]) + 1);     /* utils.h:1097-1098, same order */
	free(run_lens);
	for (int i = 0; i < 4; ++i) plain_cost += rp[i] * e->dna_log[i];
	plain_cost += (double)(ilog2u(ref_len) + 1);
	const int choose_plain = plain_cost < es_cost;
	if (choose_plain) { ++e->dec[1]; est_rescale(e->es, 12, &e->es_sum, 1u << 20); }
	else { ++e->dec[0]; memcpy(e->es, loc, sizeof loc); e->es_sum = loc_sum; est_rescale(e->es, 12, &e->es_sum, 1u << 20); }
	++e->dec_sum;
	est_rescale(e->dec, 2, &e->dec_sum, 1u << 20);
	est_logs(e->dec, e->dec_log, 2, e->dec_sum);
	return !choose_plain;
}

/* ------------------------------------------------------------------------------------------------ E1-E4: anchors */
typedef struct { uint32_t len, pos_enc, pos_ref; } anchor_t;
typedef struct { int rev; uint32_t ref_id; anchor_t* a; uint32_t n, tot; } cand_t;
typedef struct { uint64_t mmer; uint32_t pos; } mp_t;
static int cmp_mp(const void* a, const void* b)
{
	const mp_t* x = (const mp_t*)a; const mp_t* y = (const mp_t*)b;
	if (x->mmer != y->mmer) return x->mmer < y->mmer ? -1 : 1;
	return x->pos < y->pos ? -1 : x->pos > y->pos;
}
static int cmp_mp_pos(const void* a, const void* b) { const mp_t* x = (const mp_t*)a; const mp_t* y = (const mp_t*)b; return x->pos < y->pos ? -1 : x->pos > y->pos; }

/* all (m-mer, position) of a read sorted by (m-mer, position); encoder.cpp:326-352 without the hash map */
static uint32_t list_mmers(const uint8_t* r, uint32_t len, uint32_t m, mp_t** out)
{
	if (len < m) { *out = NULL; return 0; }
	const uint64_t mask = m == 32 ? ~0ULL : ((1ULL << (2 * m)) - 1);
	mp_t* v = (mp_t*)malloc(sizeof(mp_t) * (len - m + 1));
	uint64_t x = 0; uint32_t n = 0;
	for (uint32_t p = 0; p < len; ++p) { x = ((x << 2) + r[p]) & mask; if (p + 1 >= m) { v[n].mmer = x; v[n].pos = p + 1 - m; ++n; } }
	qsort(v, n, sizeof(mp_t), cmp_mp);
	*out = v; return n;
}
static int64_t find_first(const mp_t* v, uint32_t n, uint64_t mmer)
{
	uint32_t lo = 0, hi = n;
	while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (v[mid].mmer < mmer) lo = mid + 1; else hi = mid; }
	return lo < n && v[lo].mmer == mmer ? (int64_t)lo : -1;
}

/* utils.cpp:157-209 (LIS): patience with lower_bound (strictly increasing) and predecessor links; the chain is read
 * back from the last pile's top.  Restated literally because the choice among equally long subsequences follows it. */
static uint32_t lis(const int* in, uint32_t n, int* out)
{
	if (!n) return 0;
	int* pred = (int*)malloc(sizeof(int) * n); int* tv = (int*)malloc(sizeof(int) * n); int* ti = (int*)malloc(sizeof(int) * n);
	uint32_t len = 1; tv[0] = in[0]; ti[0] = 0; pred[0] = -1;
	for (uint32_t i = 1; i < n; ++i)
	{
		const int x = in[i]; uint32_t pos;
		if (tv[len - 1] < x) pos = len;
		else { uint32_t lo = 0, hi = len; while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (x > tv[mid]) lo = mid + 1; else hi = mid; } pos = lo; }
		if (pos == len) ++len;
		tv[pos] = x; ti[pos] = (int)i;
		pred[i] = pos > 0 ? ti[pos - 1] : -1;
	}
	int cur = ti[len - 1];
	for (int i = (int)len - 1; i >= 0; --i) { out[i] = in[cur]; cur = pred[cur]; }
	free(pred); free(tv); free(ti);
	return len;
}

enum { RES_EMPTY = 0, RES_TOO_MANY = 1, RES_TOO_LOW = 2, RES_ACCEPT = 3 };

/* encoder.cpp:1016-1056 (AnalyseRefRead) + :392-492 (GetIntersection) + :697-729 (Convert) + :617-661 (LIS alignment) +
 * :731-776 (MergeAnchors).  enc_mm = list_mmers(enc). */
static int analyse_ref_read(const mp_t* enc_mm, uint32_t n_enc_mm, uint32_t enc_len, const uint8_t* ref, uint32_t ref_len,
	const orc_s2_params* P, int decision, cand_t* c)
{
	const uint32_t m = P->anchor_len;
	c->a = NULL; c->n = 0; c->tot = 0;
	/* m-mers of the reference read that also occur in the read being encoded (Bloom + include.check, :368-384) */
	mp_t* all; uint32_t n_all = list_mmers(ref, ref_len, m, &all);
	mp_t* rf = (mp_t*)malloc(sizeof(mp_t) * (n_all + 1)); uint32_t n_rf = 0;
	for (uint32_t i = 0; i < n_all; ++i) if (find_first(enc_mm, n_enc_mm, all[i].mmer) >= 0) rf[n_rf++] = all[i];
	free(all);
	if (n_rf == 0) { free(rf); return RES_EMPTY; }
	/* the read's own occurrences of those m-mers */
	mp_t* ec = (mp_t*)malloc(sizeof(mp_t) * (n_enc_mm + 1)); uint32_t n_ec = 0;
	uint64_t matches = 0;
	for (uint32_t i = 0; i < n_rf; )
	{
		uint32_t j = i; while (j < n_rf && rf[j].mmer == rf[i].mmer) ++j;
		int64_t f = find_first(enc_mm, n_enc_mm, rf[i].mmer); uint32_t e = (uint32_t)f, cnt = 0;
		while (e < n_enc_mm && enc_mm[e].mmer == rf[i].mmer) { ec[n_ec++] = enc_mm[e]; ++e; ++cnt; }
		matches += (uint64_t)cnt * (j - i);
		i = j;
	}
	if (decision != 0) decision = (double)matches > P->max_matches_mult * (double)(enc_len + 1);   /* enc_read.size() counts the guard */
	if (decision == 1) { free(rf); free(ec); return RES_TOO_MANY; }
	/* position-sorted views (Convert) */
	mp_t* enc_pos = (mp_t*)malloc(sizeof(mp_t) * (n_ec + 1)); memcpy(enc_pos, ec, sizeof(mp_t) * n_ec); qsort(enc_pos, n_ec, sizeof(mp_t), cmp_mp_pos);
	mp_t* ref_pos = (mp_t*)malloc(sizeof(mp_t) * (n_rf + 1)); memcpy(ref_pos, rf, sizeof(mp_t) * n_rf); qsort(ref_pos, n_rf, sizeof(mp_t), cmp_mp_pos);
	/* LIS input: for every read position in order, the reference positions of its m-mer, descending if several */
	int* seq = (int*)malloc(sizeof(int) * ((size_t)n_ec * 1 + 16)); size_t seq_cap = (size_t)n_ec + 16, n_seq = 0;
	for (uint32_t i = 0; i < n_ec; ++i)
	{
		int64_t f = find_first(rf, n_rf, enc_pos[i].mmer); uint32_t a = (uint32_t)f, b = a;
		while (b < n_rf && rf[b].mmer == enc_pos[i].mmer) ++b;
		if (n_seq + (b - a) > seq_cap) { seq_cap = (n_seq + (b - a)) * 2; seq = (int*)realloc(seq, sizeof(int) * seq_cap); }
		for (uint32_t x = b; x > a; --x) seq[n_seq++] = (int)rf[x - 1].pos;
	}
	int* chain = (int*)malloc(sizeof(int) * (n_seq + 1));
	const uint32_t n_chain = lis(seq, (uint32_t)n_seq, chain);
	/* back to (read position, reference position): forward scans exactly as :646-658 */
	anchor_t* al = (anchor_t*)malloc(sizeof(anchor_t) * (n_chain + 1));
	uint32_t ep = 0, rp = 0;
	for (uint32_t i = 0; i < n_chain; ++i)
	{
		while (ref_pos[rp++].pos != (uint32_t)chain[i]) ;
		const uint64_t mmer = ref_pos[rp - 1].mmer;
		while (enc_pos[ep++].mmer != mmer) ;
		al[i].pos_enc = enc_pos[ep - 1].pos; al[i].pos_ref = ref_pos[rp - 1].pos; al[i].len = 0;
	}
	/* MergeAnchors */
	anchor_t* res = (anchor_t*)malloc(sizeof(anchor_t) * (n_chain + 1)); uint32_t n_res = 0, tot = 0, start = 0;
	for (uint32_t i = 1; i <= n_chain; ++i)
	{
		if (i == n_chain || al[i - 1].pos_enc != al[i].pos_enc - 1 || al[i - 1].pos_ref != al[i].pos_ref - 1)
		{
			const uint32_t len = (i - start) + m - 1;
			res[n_res].len = len; res[n_res].pos_enc = al[start].pos_enc; res[n_res].pos_ref = al[start].pos_ref; ++n_res;
			tot += len; start = i;
		}
	}
	free(rf); free(ec); free(enc_pos); free(ref_pos); free(seq); free(chain); free(al);
	c->a = res; c->n = n_res; c->tot = tot;
	if (n_res < P->min_anchors) return RES_TOO_LOW;
	return RES_ACCEPT;
}

static void revcomp_read(const uint8_t* r, uint32_t len, uint8_t* out) { for (uint32_t i = 0; i < len; ++i) out[i] = (uint8_t)(3 - r[len - 1 - i]); out[len] = 255; }

/* ---- HiFi: k-mer based anchors (encoder.cpp:870-1012) ------------------------------------------------ */
static uint64_t rc_kmer(uint64_t x, uint32_t k) { uint64_t r = 0; for (uint32_t i = 0; i < k; ++i) { r = (r << 2) + (3 - ((x >> (2 * i)) & 3)); } return r; }
static uint64_t murmur64_s2(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }
static int cmp_u64_s2(const void* a, const void* b) { uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b; return x < y ? -1 : x > y; }
static int has_u64(const uint64_t* v, uint32_t n, uint64_t x) { uint32_t lo = 0, hi = n; while (lo < hi) { uint32_t m = (lo + hi) >> 1; if (v[m] < x) lo = m + 1; else hi = m; } return lo < n && v[lo] == x; }
/* position of x if it occurs exactly once in the (kmer,pos)-sorted list, else -1 (ExistsAndIsUnique, :601-612) */
static int64_t unique_pos(const mp_t* v, uint32_t n, uint64_t x)
{
	int64_t f = find_first(v, n, x);
	if (f < 0) return -1;
	if ((uint32_t)f + 1 < n && v[f + 1].mmer == x) return -1;
	return v[f].pos;
}
static int cmp_anchor_enc(const void* a, const void* b) { const anchor_t* x = (const anchor_t*)a; const anchor_t* y = (const anchor_t*)b; return x->pos_enc < y->pos_enc ? -1 : x->pos_enc > y->pos_enc; }

/* returns 1 on Accept */
static int analyse_ref_read_with_kmers(const mp_t* enc_km, uint32_t n_enc_km, const uint8_t* enc, uint32_t enc_len,
	const uint8_t* ref, uint32_t ref_len, const uint64_t* common_sorted, uint32_t n_common, const orc_s2_params* P, cand_t* c)
{
	const uint32_t k = P->kmer_len;
	c->a = NULL; c->n = 0; c->tot = 0;
	/* forward k-mers of the reference read whose canonical form is a shared k-mer (:549-589) */
	mp_t* rk = NULL; uint32_t n_rk = 0;
	if (ref_len >= k)
	{
		const uint64_t mask = k == 32 ? ~0ULL : ((1ULL << (2 * k)) - 1);
		rk = (mp_t*)malloc(sizeof(mp_t) * (ref_len - k + 1));
		uint64_t f = 0, r = 0;
		for (uint32_t p = 0; p < ref_len; ++p)
		{
			f = ((f << 2) + ref[p]) & mask; r = (r >> 2) + ((uint64_t)(3 - ref[p]) << (2 * (k - 1)));
			if (p + 1 < k) continue;
			const uint64_t can = f < r ? f : r;
			if (murmur64_s2(can) % P->modulo == 0 && has_u64(common_sorted, n_common, can)) { rk[n_rk].mmer = f; rk[n_rk].pos = p + 1 - k; ++n_rk; }
		}
		qsort(rk, n_rk, sizeof(mp_t), cmp_mp);
	}
	anchor_t* an = (anchor_t*)malloc(sizeof(anchor_t) * (n_common + 1)); uint32_t n = 0;
	for (uint32_t i = 0; i < n_common; ++i)
	{
		uint64_t km = common_sorted[i];
		int64_t ie = unique_pos(enc_km, n_enc_km, km), ir = unique_pos(rk, n_rk, km);
		if (ie == -1 || ir == -1) { km = rc_kmer(km, k); ie = unique_pos(enc_km, n_enc_km, km); ir = unique_pos(rk, n_rk, km); }
		if (ie != -1 && ir != -1) { an[n].len = k; an[n].pos_enc = (uint32_t)ie; an[n].pos_ref = (uint32_t)ir; ++n; }
	}
	free(rk);
	if (n == 0) { free(an); return 0; }
	qsort(an, n, sizeof(anchor_t), cmp_anchor_enc);
	for (uint32_t i = 1; i < n; ++i) if (an[i].pos_ref < an[i - 1].pos_ref) { free(an); return 0; }   /* not colinear */
	/* drop k-mers overlapping their predecessor (:917-926) */
	{
		uint32_t w = 1;
		for (uint32_t i = 1; i < n; ++i)
		{
			const anchor_t* p = &an[w - 1];
			if (p->pos_enc + p->len > an[i].pos_enc || p->pos_ref + p->len > an[i].pos_ref) continue;
			an[w++] = an[i];
		}
		n = w;
	}
	/* extend the first anchor to the left */
	while (an[0].pos_enc > 0 && an[0].pos_ref > 0 && enc[an[0].pos_enc - 1] == ref[an[0].pos_ref - 1]) { --an[0].pos_enc; --an[0].pos_ref; ++an[0].len; }
	/* extend / merge (:944-998), literal restatement on a vector with erase */
	for (uint64_t i = 0; i < n; ++i)
	{
		if (i > 0)
		{
			const uint32_t pe = an[i - 1].pos_enc + an[i - 1].len, pr = an[i - 1].pos_ref + an[i - 1].len;
			for (;;)
			{
				const int re = an[i].pos_enc == pe, rr = an[i].pos_ref == pr;
				if (re && rr)
				{
					an[i].len += an[i - 1].len;
					memmove(&an[i - 1], &an[i], sizeof(anchor_t) * (n - i)); --n;     /* erase(i-1): element i moves to i-1 ... */
					break;                                                               /* ... and the loop goes on with index i */
				}
				if (re || rr) break;
				if (enc[an[i].pos_enc - 1] != ref[an[i].pos_ref - 1]) break;
				an[i].len++; an[i].pos_enc--; an[i].pos_ref--;
			}
		}
		if (i >= n) break;
		if (i != (uint64_t)n - 1)
		{
			const uint32_t ne = an[i + 1].pos_enc, nr = an[i + 1].pos_ref;
			uint32_t pe = an[i].pos_enc + an[i].len, pr = an[i].pos_ref + an[i].len;
			for (;;)
			{
				const int re = pe == ne, rr = pr == nr;
				if (re && rr)
				{
					an[i].len += an[i + 1].len;
					memmove(&an[i + 1], &an[i + 2], sizeof(anchor_t) * (n - i - 2)); --n;
					--i;
					break;
				}
				else if (re || rr) break;
				if (enc[pe] != ref[pr]) break;
				++pe; ++pr; ++an[i].len;
			}
		}
	}
	{
		anchor_t* l = &an[n - 1];
		uint32_t pe = l->pos_enc + l->len, pr = l->pos_ref + l->len;
		while (pe < enc_len && pr < ref_len && enc[pe] == ref[pr]) { ++pe; ++pr; ++l->len; }
	}
	uint32_t tot = 0; for (uint32_t i = 0; i < n; ++i) tot += an[i].len;
	c->a = an; c->n = n; c->tot = tot;
	return 1;
}

/* encoder.cpp:1577-1622 */
static void fix_overlaps(anchor_t* a, uint32_t n)
{
	for (uint32_t i = 0; i + 1 < n; ++i)
	{
		const uint32_t end = a[i].pos_ref + a[i].len;
		if (a[i + 1].pos_ref < end) { const uint32_t d = end - a[i + 1].pos_ref; a[i + 1].pos_ref += d; a[i + 1].len -= d; a[i + 1].pos_enc += d; }
	}
	for (uint32_t i = 0; i + 1 < n; ++i)
	{
		const uint32_t end = a[i].pos_enc + a[i].len;
		if (a[i + 1].pos_enc < end) { const uint32_t d = end - a[i + 1].pos_enc; a[i + 1].pos_enc += d; a[i + 1].len -= d; a[i + 1].pos_ref += d; }
	}
}

/* stable insertion sort by tot desc: std::sort on <= 16 elements is an insertion sort that keeps equal elements in order */
static void sort_cands(cand_t* c, uint32_t from, uint32_t n)
{
	for (uint32_t i = from + 1; i < n; ++i) { cand_t x = c[i]; uint32_t j = i; while (j > from && c[j - 1].tot < x.tot) { c[j] = c[j - 1]; --j; } c[j] = x; }
}

/* ------------------------------------------------------------------------------------------------ E9: tuples */
typedef struct { uint8_t* p; size_t n, cap; } bbuf;
static void bb_push(bbuf* b, uint8_t x) { if (b->n == b->cap) { b->cap = b->cap * 2 + 64; b->p = (uint8_t*)realloc(b->p, b->cap); } b->p[b->n++] = x; }
static void t_simple(bbuf* b, int type, int val) { bb_push(b, (uint8_t)((type << 4) + val)); }
static void t_len28(bbuf* b, int type, uint32_t v) { bb_push(b, (uint8_t)((type << 4) + (v >> 24))); bb_push(b, (v >> 16) & 0xff); bb_push(b, (v >> 8) & 0xff); bb_push(b, v & 0xff); }
static void t_id(bbuf* b, int type, uint32_t id, int rev) { bb_push(b, (uint8_t)((type << 4) + rev)); bb_push(b, id >> 24); bb_push(b, (id >> 16) & 0xff); bb_push(b, (id >> 8) & 0xff); bb_push(b, id & 0xff); }
enum { T_INS = 0, T_DEL = 1, T_MATCH = 2, T_SUB = 3, T_ANCHOR = 4, T_SKIP = 5, T_ALT = 6, T_MAIN = 7, T_PLAIN = 8, T_START_PLAIN = 9, T_START_ES = 10, T_START_N = 11 };

/* encoder.cpp:1348-1412 */
static void store_run(bbuf* b, char s, uint32_t rep)
{
	if (s == 'M') { if (rep >= 15) t_len28(b, T_ANCHOR, rep); else for (uint32_t i = 0; i < rep; ++i) t_simple(b, T_MATCH, 0); }
	else if (s == 'D') { if (rep > 16) t_len28(b, T_SKIP, rep); else for (uint32_t i = 0; i < rep; ++i) t_simple(b, T_DEL, 0); }
	else if (is_mm(s)) for (uint32_t i = 0; i < rep; ++i) t_simple(b, T_SUB, s - 'X');
	else for (uint32_t i = 0; i < rep; ++i) t_simple(b, T_INS, es_code(s));
}
static void script_to_tuples(bbuf* b, const char* es, size_t n, uint32_t lead_dels)
{
	/* run-length over ('D' x lead_dels) + es */
	char s = lead_dels ? 'D' : es[0]; uint32_t rep = lead_dels ? lead_dels : 1;
	for (size_t i = lead_dels ? 0 : 1; i < n; ++i) { if (es[i] != s) { store_run(b, s, rep); s = es[i]; rep = 1; } else ++rep; }
	store_run(b, s, rep);
}

typedef struct {
	const orc_s2_params* P;
	estimator est;
	/* oriented reference reads are materialised on demand */
	const uint8_t* sym;            /* all reads, symbols 0..4, each followed by a 255 guard */
	const uint64_t* sym_off;       /* start of read i in sym */
	const uint32_t* len;
	const uint32_t* ref_to_read;   /* reference id -> read index */
} enc_ctx;

static uint8_t* oriented_ref(const enc_ctx* C, uint32_t ref_id, int rev, uint32_t* len)
{
	const uint32_t r = C->ref_to_read[ref_id]; const uint32_t l = C->len[r];
	uint8_t* o = (uint8_t*)malloc(l + 1);
	if (rev) revcomp_read(C->sym + C->sym_off[r], l, o); else { memcpy(o, C->sym + C->sym_off[r], l); o[l] = 255; }
	*len = l; return o;
}

/* encoder.cpp:1414-1443 */
static void store_frag(sbuf* big, uint32_t level, bbuf* out, uint32_t ref_id, uint32_t main_ref, uint32_t* last_pos_in_ref, uint32_t cur_pos_in_ref, int* first, int rev)
{
	if (big->n)
	{
		if (level == 0)
		{
			if (ref_id != main_ref) t_id(out, T_ALT, ref_id, rev);
			else if (!*first) t_simple(out, T_MAIN, 0);
			script_to_tuples(out, big->p, big->n, 0);
		}
		else
		{
			if (ref_id != main_ref) t_id(out, T_ALT, ref_id, rev); else t_simple(out, T_MAIN, 0);
			script_to_tuples(out, big->p, big->n, *last_pos_in_ref);
		}
		*last_pos_in_ref = cur_pos_in_ref;
		*first = 0;
	}
	big->n = 0;
}

/* encoder.cpp:778-868 (AdjustAnchors) */
static uint32_t adjust_anchors(cand_t* c, uint32_t ns, uint32_t ne, uint32_t anchor_len)
{
	anchor_t* a = c->a; uint32_t n = c->n, tot = 0;
	uint32_t first = 0xFFFFFFFFu, last = 0xFFFFFFFFu;
	for (uint32_t i = 0; i < n; ++i) if (a[i].pos_enc + a[i].len > ns) { first = i; break; }
	if (first == 0xFFFFFFFFu) { c->n = 0; return 0; }
	if (a[first].pos_enc < ns && (a[first].pos_enc + a[first].len) - ns < anchor_len) ++first;
	for (int32_t i = (int32_t)n - 1; i >= 0; --i) if (a[i].pos_enc < ne) { last = (uint32_t)i; break; }
	if (last == 0xFFFFFFFFu) { c->n = 0; return 0; }
	if (first < n && last < n && a[last].pos_enc + a[last].len > ne && ne - a[last].pos_enc < anchor_len) { if (last == 0) { c->n = 0; return 0; } --last; }
	if (first > last) { c->n = 0; return 0; }
	const uint32_t m = last - first + 1;
	memmove(a, a + first, sizeof(anchor_t) * m); c->n = n = m;
	if (a[n - 1].pos_enc + a[n - 1].len > ne) a[n - 1].len -= (a[n - 1].pos_enc + a[n - 1].len - ne);
	for (uint32_t i = 0; i < n; ++i)
	{
		if (i == 0 && a[0].pos_enc < ns) { const uint32_t d = ns - a[0].pos_enc; a[0].len -= d; a[0].pos_enc = 0; a[0].pos_ref += d; }
		else a[i].pos_enc -= ns;
		tot += a[i].len;
	}
	return tot;
}

static cand_t* clone_cands(const cand_t* c, uint32_t n)
{
	cand_t* o = (cand_t*)malloc(sizeof(cand_t) * (n + 1));
	for (uint32_t i = 0; i < n; ++i) { o[i] = c[i]; o[i].a = (anchor_t*)malloc(sizeof(anchor_t) * (c[i].n + 1)); memcpy(o[i].a, c[i].a, sizeof(anchor_t) * c[i].n); }
	return o;
}
static void free_cands(cand_t* c, uint32_t n) { for (uint32_t i = 0; i < n; ++i) free(c[i].a); free(c); }

/* encoder.cpp:1513-1575 (AddEncodedReadWithCandidates) with :1445-1511 (EncodePart) inlined */
static void add_encoded(enc_ctx* C, const uint8_t* enc, uint32_t enc_len, cand_t* cands, uint32_t n_cands, uint32_t level, bbuf* out, uint32_t main_ref, int* first)
{
	const orc_s2_params* P = C->P;
	const uint32_t ref_id = cands[level].ref_id; const int rev = cands[level].rev;
	const anchor_t* an = cands[level].a; const uint32_t n_an = cands[level].n;
	uint32_t last_pos_in_ref = 0;
	if (level == 0) t_id(out, T_START_ES, main_ref, cands[0].rev);
	sbuf big = {0, 0, 0};
	uint32_t ref_len; uint8_t* ref = oriented_ref(C, ref_id, rev, &ref_len);
	const uint32_t n_frag = n_an * 2 + 1;
	uint32_t anch = 0, cur_ref = 0, cur_enc = 0;
	for (uint32_t i = 0; i < n_frag; ++i)
	{
		if (i % 2 == 1)
		{
			sb_fill(&big, 'M', an[anch].len);
			cur_ref = an[anch].pos_ref + an[anch].len; cur_enc = an[anch].pos_enc + an[anch].len; ++anch;
			continue;
		}
		const uint32_t end_enc = i == n_frag - 1 ? enc_len : an[anch].pos_enc;
		const uint32_t end_ref = i == n_frag - 1 ? ref_len : an[anch].pos_ref;
		const uint8_t* rp = ref + cur_ref; const uint32_t rl = end_ref - cur_ref;
		const uint8_t* ep = enc + cur_enc; const uint32_t el = end_enc - cur_enc;
		const int kind = i == 0 ? 0 : (i == n_frag - 1 ? 1 : 2);
		sbuf es = {0, 0, 0};
		get_edit_dist(rp, rl, ep, el, kind, &es);
		int use_es;
		if (el < P->min_part_len_alt) use_es = est_use_es(&C->est, es.p, es.n, ep, el, rl);
		else use_es = stateless_use_es(es.p, es.n, ep, el, P->es_cost_mult);
		if (use_es) sb_append(&big, es.p, es.n);
		else
		{
			int alt_ok = 0; cand_t* alt = NULL;
			if (!(n_cands <= level + 1 || el < P->min_part_len_alt || level >= P->max_recurence))
			{	/* encoder.cpp:1329-1346 */
				alt = clone_cands(cands, n_cands);
				for (uint32_t q = level + 1; q < n_cands; ++q) alt[q].tot = adjust_anchors(&alt[q], cur_enc, end_enc, P->anchor_len);
				sort_cands(alt, level + 1, n_cands);
				alt_ok = alt[level + 1].tot != 0;
			}
			if (alt_ok)
			{
				store_frag(&big, level, out, ref_id, main_ref, &last_pos_in_ref, cur_ref, first, rev);
				add_encoded(C, ep, el, alt, n_cands, level + 1, out, main_ref, first);
				if (i != n_frag - 1) sb_fill(&big, 'D', end_ref - cur_ref);
			}
			else
			{
				for (uint32_t q = 0; q < el; ++q) sb_push(&big, "ACGT"[ep[q]]);
				if (i != n_frag - 1) sb_fill(&big, 'D', end_ref - cur_ref);
			}
			if (alt) free_cands(alt, n_cands);
		}
		sb_free(&es);
	}
	store_frag(&big, level, out, ref_id, main_ref, &last_pos_in_ref, cur_ref, first, rev);
	sb_free(&big); free(ref);
}

/* encoder.cpp:1058-1111 / :1194-1253 (prepareEncodeCandidates[HiFi]) */
static uint32_t prepare_candidates(enc_ctx* C, const uint8_t* enc, uint32_t enc_len, const uint32_t* neigh, uint32_t n_neigh,
	const uint64_t* const* common, const uint32_t* common_n, cand_t* out)
{
	const orc_s2_params* P = C->P;
	mp_t* emm; const uint32_t n_emm = list_mmers(enc, enc_len, P->anchor_len, &emm);
	uint32_t uniq = 0; for (uint32_t i = 0; i < n_emm; ++i) if (i == 0 || emm[i].mmer != emm[i - 1].mmer) ++uniq;
	int decision = -1;
	if ((double)uniq > P->min_mmer_force * (double)enc_len) decision = 0;
	else if ((double)uniq < P->min_mmer_frac * (double)enc_len) decision = 1;
	if (decision == 1) { free(emm); return 0; }
	mp_t* ekm = NULL; uint32_t n_ekm = 0;
	if (P->is_hifi) n_ekm = list_mmers(enc, enc_len, P->kmer_len, &ekm);
	uint32_t n_out = 0;
	for (uint32_t q = 0; q < n_neigh; ++q)
	{
		uint32_t rl; uint8_t* fw = oriented_ref(C, neigh[q], 0, &rl); uint8_t* rc = oriented_ref(C, neigh[q], 1, &rl);
		cand_t cf, cr; cf.rev = 0; cr.rev = 1; cf.ref_id = cr.ref_id = neigh[q]; cf.a = cr.a = NULL; cf.n = cr.n = 0; cf.tot = cr.tot = 0;
		int done = 0;
		if (P->is_hifi)
		{	/* KmerBasedAnchors, encoder.cpp:1113-1147 */
			uint64_t* cs = (uint64_t*)malloc(sizeof(uint64_t) * (common_n[q] + 1));
			memcpy(cs, common[q], sizeof(uint64_t) * common_n[q]); qsort(cs, common_n[q], sizeof(uint64_t), cmp_u64_s2);
			const int ar = analyse_ref_read_with_kmers(ekm, n_ekm, enc, enc_len, rc, rl, cs, common_n[q], P, &cr);
			const int af = analyse_ref_read_with_kmers(ekm, n_ekm, enc, enc_len, fw, rl, cs, common_n[q], P, &cf);
			free(cs);
			if (ar && af) { if (cf.tot > cr.tot) { out[n_out++] = cf; free(cr.a); } else { out[n_out++] = cr; free(cf.a); } done = 1; }
			else if (ar) { out[n_out++] = cr; free(cf.a); done = 1; }
			else if (af) { out[n_out++] = cf; free(cr.a); done = 1; }
			else { free(cf.a); free(cr.a); cf.a = cr.a = NULL; }
		}
		if (!done)
		{	/* MmerBasedAnchors, encoder.cpp:1149-1192 */
			const int rr = analyse_ref_read(emm, n_emm, enc_len, rc, rl, P, decision, &cr);
			const int rf = analyse_ref_read(emm, n_emm, enc_len, fw, rl, P, decision, &cf);
			if (rr == RES_ACCEPT && rf == RES_ACCEPT) { if (cf.tot > cr.tot) { out[n_out++] = cf; free(cr.a); } else { out[n_out++] = cr; free(cf.a); } }
			else if (rr == RES_ACCEPT) { out[n_out++] = cr; free(cf.a); }
			else if (rf == RES_ACCEPT) { out[n_out++] = cf; free(cr.a); }
			else { free(cf.a); free(cr.a); }
		}
		free(fw); free(rc);
	}
	free(emm); free(ekm);
	sort_cands(out, 0, n_out);
	return n_out;
}

/* ------------------------------------------------------------------------------------------------ driver */
/* encoder.cpp:1625-1691 (processComprElem / Encode) over all reads; the estimator is reset at every pack boundary.
 * bases: ASCII; is_ref[i]: read i is a reference read (already excludes reads with N); cand[i*max_cand+j] reference ids.
 * common_*: HiFi shared k-mers per (read, candidate) as clb/oracle stage 1 returns them (may be NULL when !is_hifi).
 * Output: CompactES bytes of all reads back to back, es_off[n_reads+1].  Returns the total size (may exceed out_cap). */
uint64_t orc_encode_reads(const uint8_t* bases, const uint64_t* offsets, uint32_t n_reads, const uint8_t* is_ref,
	const uint32_t* cand, const uint32_t* cand_n, uint32_t max_cand,
	const uint64_t* common_off, const uint32_t* common_n, const uint64_t* common,
	const uint32_t* pack_sizes, uint32_t n_packs, const orc_s2_params* P,
	uint8_t* out, uint64_t out_cap, uint64_t* es_off)
{
	enc_ctx C; C.P = P;
	uint64_t* sym_off = (uint64_t*)malloc(sizeof(uint64_t) * ((size_t)n_reads + 1));
	uint32_t* len = (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)n_reads + 1));
	uint8_t* has_n = (uint8_t*)calloc((size_t)n_reads + 1, 1);
	uint64_t tot = 0;
	for (uint32_t i = 0; i < n_reads; ++i) { sym_off[i] = tot; len[i] = (uint32_t)(offsets[i + 1] - offsets[i]); tot += len[i] + 1; }
	uint8_t* sym = (uint8_t*)malloc(tot + 1);
	for (uint32_t i = 0; i < n_reads; ++i)
	{
		for (uint32_t j = 0; j < len[i]; ++j)
		{
			uint8_t c = bases[offsets[i] + j], s;
			switch (c) { case 'A': s = 0; break; case 'C': s = 1; break; case 'G': s = 2; break; case 'T': s = 3; break; default: s = 4; has_n[i] = 1; }
			sym[sym_off[i] + j] = s;
		}
		sym[sym_off[i] + len[i]] = 255;
	}
	uint32_t* ref_to_read = (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)n_reads + 1)); uint32_t n_ref = 0;
	for (uint32_t i = 0; i < n_reads; ++i) if (is_ref[i]) ref_to_read[n_ref++] = i;
	C.sym = sym; C.sym_off = sym_off; C.len = len; C.ref_to_read = ref_to_read;

	bbuf ob = {0, 0, 0};
	uint32_t pack = 0, in_pack = 0;
	est_reset(&C.est);
	cand_t* cs = (cand_t*)malloc(sizeof(cand_t) * (max_cand + 1));
	const uint64_t** cm = (const uint64_t**)malloc(sizeof(uint64_t*) * (max_cand + 1));
	for (uint32_t i = 0; i < n_reads; ++i)
	{
		while (pack < n_packs && in_pack == pack_sizes[pack]) { ++pack; in_pack = 0; est_reset(&C.est); }
		++in_pack;
		es_off[i] = ob.n;
		const uint8_t* r = sym + sym_off[i]; const uint32_t l = len[i];
		if (has_n[i]) { t_simple(&ob, T_START_N, 0); for (uint32_t j = 0; j < l; ++j) t_simple(&ob, T_PLAIN, r[j]); continue; }
		est_log_read(&C.est, r, l);
		uint32_t nc = 0;
		if (cand_n[i])
		{
			for (uint32_t j = 0; j < cand_n[i]; ++j) cm[j] = common ? common + common_off[(uint64_t)i * max_cand + j] : NULL;
			nc = prepare_candidates(&C, r, l, cand + (uint64_t)i * max_cand, cand_n[i], cm, common_n ? common_n + (uint64_t)i * max_cand : NULL, cs);
		}
		if (nc == 0) { t_simple(&ob, T_START_PLAIN, 0); for (uint32_t j = 0; j < l; ++j) t_simple(&ob, T_PLAIN, r[j]); continue; }
		for (uint32_t j = 0; j < nc; ++j) fix_overlaps(cs[j].a, cs[j].n);
		int first = 1;
		add_encoded(&C, r, l, cs, nc, 0, &ob, cs[0].ref_id, &first);
		for (uint32_t j = 0; j < nc; ++j) free(cs[j].a);
	}
	es_off[n_reads] = ob.n;
	if (ob.n <= out_cap) memcpy(out, ob.p, ob.n);
	const uint64_t total = ob.n;
	free(ob.p); free(cs); free(cm); free(sym); free(sym_off); free(len); free(has_n); free(ref_to_read);
	return total;
}

/* Test tap: the candidates (encoder.h:46) of every read after prepareEncodeCandidates + the overlap fixes, best first.
 * Per read cand_off[i] .. cand_off[i+1] index `out` (uint32 words): ref_id, shouldReverse, tot_anchor_len, n_anchors, then
 * n_anchors * (len, pos_enc, pos_ref).  m-mer path only.  Returns the number of words (may exceed cap). */
uint64_t orc_candidates(const uint8_t* bases, const uint64_t* offsets, uint32_t n_reads, const uint8_t* is_ref,
	const uint32_t* cand, const uint32_t* cand_n, uint32_t max_cand, const orc_s2_params* P,
	uint64_t* cand_off, uint32_t* out, uint64_t cap)
{
	enc_ctx C; C.P = P;
	uint64_t* sym_off = (uint64_t*)malloc(sizeof(uint64_t) * ((size_t)n_reads + 1));
	uint32_t* len = (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)n_reads + 1));
	uint8_t* has_n = (uint8_t*)calloc((size_t)n_reads + 1, 1);
	uint64_t tot = 0;
	for (uint32_t i = 0; i < n_reads; ++i) { sym_off[i] = tot; len[i] = (uint32_t)(offsets[i + 1] - offsets[i]); tot += len[i] + 1; }
	uint8_t* sym = (uint8_t*)malloc(tot + 1);
	for (uint32_t i = 0; i < n_reads; ++i)
	{
		for (uint32_t j = 0; j < len[i]; ++j)
		{
			uint8_t c = bases[offsets[i] + j], s;
			switch (c) { case 'A': s = 0; break; case 'C': s = 1; break; case 'G': s = 2; break; case 'T': s = 3; break; default: s = 4; has_n[i] = 1; }
			sym[sym_off[i] + j] = s;
		}
		sym[sym_off[i] + len[i]] = 255;
	}
	uint32_t* ref_to_read = (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)n_reads + 1)); uint32_t n_ref = 0;
	for (uint32_t i = 0; i < n_reads; ++i) if (is_ref[i]) ref_to_read[n_ref++] = i;
	C.sym = sym; C.sym_off = sym_off; C.len = len; C.ref_to_read = ref_to_read;
	cand_t* cs = (cand_t*)malloc(sizeof(cand_t) * (max_cand + 1));
	uint64_t w = 0;
	for (uint32_t i = 0; i < n_reads; ++i)
	{
		cand_off[i] = w;
		if (has_n[i] || !cand_n[i]) continue;
		const uint32_t nc = prepare_candidates(&C, sym + sym_off[i], len[i], cand + (uint64_t)i * max_cand, cand_n[i], NULL, NULL, cs);
		for (uint32_t j = 0; j < nc; ++j)
		{
			fix_overlaps(cs[j].a, cs[j].n);
			if (w + 4 + 3ull * cs[j].n <= cap)
			{
				out[w] = cs[j].ref_id; out[w + 1] = (uint32_t)cs[j].rev; out[w + 2] = cs[j].tot; out[w + 3] = cs[j].n;
				for (uint32_t a = 0; a < cs[j].n; ++a) { out[w + 4 + 3 * a] = cs[j].a[a].len; out[w + 5 + 3 * a] = cs[j].a[a].pos_enc; out[w + 6 + 3 * a] = cs[j].a[a].pos_ref; }
			}
			w += 4 + 3ull * cs[j].n;
			free(cs[j].a);
		}
	}
	cand_off[n_reads] = w;
	free(cs); free(sym); free(sym_off); free(len); free(has_n); free(ref_to_read);
	return w;
}
