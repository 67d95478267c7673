// stage2.h — device-side records of stage 2 shared by stage2_anchors.cu and stage2_encode.cu.
//
// Reference behaviour restated (oracle/stage2.c is the scalar form pinned against the reference's CompactES dumps):
//   src/colord/encoder.cpp:326-492, :617-776, :1016-1192   m-mer anchors of a read against its candidate reference reads
//   src/colord/encoder.cpp:778-868, :1255-1575             fragments, edit scripts, alternative reads, tuple emission
//   src/colord/utils.h:69-273, :700-1126                   CompactES bytes, CEntropy, CEntropyEstimator
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "util.cuh"

namespace clb {

struct S2P {                      // CCompressorParams fields used by CEncoder (encoder.h:371-410)
	uint32_t m;                   // anchor_len
	uint32_t min_alt;             // minPartLenToConsiderAltRead
	uint32_t max_rec;             // maxRecurence
	uint32_t min_anchors;
	uint32_t c;                   // max_candidates
	double min_frac, min_force, max_mult, cost_mult;
};

struct Anchor { uint32_t len, pos_enc, pos_ref; };             // encoder.h:40-44

// A candidate's anchor list as seen by one recursion node: a window [first, first+n) of the anchors computed once for
// the whole read, with the first/last anchor clipped and all read positions shifted to the node's sub-range.
// This is what AdjustAnchors (encoder.cpp:778-868) produces, without copying the vectors.
struct CandView {
	uint64_t anc;                 // byte offset of the candidate's Anchor array in the pair arena
	uint32_t ref_id, rev;
	uint32_t first, n;
	uint32_t head, tail;          // symbols clipped from the first anchor's start / the last anchor's end
	uint32_t shift;               // subtracted from pos_enc
	uint32_t tot;                 // tot_anchor_len
};

CLB_D Anchor cv_get(const uint8_t* __restrict__ arena, const CandView& v, uint32_t i)
{
	Anchor a = reinterpret_cast<const Anchor*>(arena + v.anc)[v.first + i];
	if (i == 0) { a.len -= v.head; a.pos_enc += v.head; a.pos_ref += v.head; }
	if (i == v.n - 1) a.len -= v.tail;
	a.pos_enc -= v.shift;
	return a;
}

// One call of AddEncodedReadWithCandidates (encoder.cpp:1513): a read (level 0) or a part of it handed to the next
// candidate (level > 0).  Candidate views of node x: cviews[x * c + k]; the node encodes against k = level.
struct Node {
	uint32_t read;                // read index
	uint32_t level;
	uint32_t enc_start, enc_len;  // the part of the read this node encodes
	uint32_t first_task;          // tasks of the even fragments 0, 2, .. 2*n_anch are consecutive
	uint32_t n_anch;
	uint32_t ncand;
	uint32_t valid;
};

enum Decision : uint32_t { D_PENDING = 0, D_ES = 1, D_PLAIN = 2, D_ALT = 3 };

// One even fragment = one EncodePart call (encoder.cpp:1445)
struct Task {
	uint32_t node, frag;          // frag = index of the anchor that follows the part (n_anch for the last part)
	uint32_t enc_start, el;       // part of the read (absolute position in the read)
	uint32_t ref_start, rl;       // part of the oriented reference read
	uint32_t kind;                // 0 left flank, 1 right flank, 2 between anchors
	uint32_t decision;
	uint64_t es_off;              // script bytes in the script buffer; long-run cost bytes follow the script (small parts)
	uint32_t es_len;
	uint32_t lead;                // 'D' symbols that precede the stored script
	uint32_t child;               // D_ALT: node that encodes the part
	uint32_t n_runs;              // small parts: number of long D / M runs (utils.h:838-899)
	uint16_t rd[12];              // small parts: CEntropyEstimator symbol counts of this script
	uint16_t rp[4];               // small parts: base counts of the part
};

// segment = (read slot, candidate j, orientation o); index (slot * c + j) * 2 + o
struct SegInfo {
	uint64_t off;                 // first pair slot in the arena (20 bytes per slot: 8 key + 4 spare + 8 scratch)
	uint32_t n;                   // pairs (0: empty / refused)
	uint32_t n_anch, tot;         // after k_lis
	uint32_t pad;
};
constexpr uint32_t PAIR_SLOT_BYTES = 20;

// device counters of the reference's -v report (include/colord_b200.h: clb_encode_stats); per level ST_LEVEL_FIELDS sums from ST_LEVEL0
enum StatIdx : uint32_t { ST_NOT_ENOUGH = 0, ST_TOO_MANY, ST_TOO_LOW, ST_NON_REV, ST_REV, ST_PLAIN_READS, ST_PLAIN_SYMB, ST_PLAIN_N_READS, ST_PLAIN_N_SYMB, ST_MAX_LEVEL,
	ST_LEVEL0 = 16 };
enum StatLevelIdx : uint32_t { SL_ALT_LEFT = 0, SL_ALT_BETWEEN, SL_ALT_RIGHT, SL_PLAIN_SYMB, SL_CODED_SYMB, SL_ES_SYMB, SL_SUBST, SL_MATCH, SL_INS, SL_DEL,
	SL_ANCHOR_SYMB, SL_ANCHORS, SL_LEFT_FLANK, SL_RIGHT_FLANK, ST_LEVEL_FIELDS };
constexpr uint32_t ST_LEVELS = 8, ST_COUNT = ST_LEVEL0 + ST_LEVELS * ST_LEVEL_FIELDS;
constexpr uint32_t SEG_TOO_MANY_MATCHES = 1;      // SegInfo::pad: the match cap refused this (candidate, orientation)

} // namespace clb
