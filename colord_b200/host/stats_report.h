// stats_report.h — the statistics block the reference prints at the end of a verbose compression (src/colord/stats_collector.cpp:100-146):
// the reader's read statistics (in_reads.cpp:71 -> CStatsCollector::LogRead, :152-160) and the encoder's counters, which the device
// collects (clb_encode_stats_enable / clb_encode_stats_get: encoder.cpp:663-676, :1069-1190, :1445-1575).
#pragma once
#include <cstdint>
#include <limits>
#include <initializer_list>
#include <ostream>
#include <string>
#include "../../include/colord_b200.h"

namespace clbhost {

// CStatsCollector::LogRead (stats_collector.cpp:152-160) read by read, in input order — including its rule that a read is only tried as
// the minimum when it is not a new maximum (so the first read never is)
struct ReadStats {
	uint64_t n_reads = 0, min_read_len = std::numeric_limits<uint32_t>::max(), max_read_len = 0, tot_read_len = 0;
	void log(uint64_t len)
	{
		tot_read_len += len;
		if (len > max_read_len) max_read_len = len;
		else if (len < min_read_len) min_read_len = len;
		++n_reads;
	}
	void log_all(const uint64_t* offsets, uint64_t n) { for (uint64_t i = 0; i < n; ++i) log(offsets[i + 1] - offsets[i]); }
};

// The block's lines are "<label padded to 31 columns>: <value>" under three banners and one banner per level; the labels are the
// reference's (a user diffs the two programs' outputs), the fields come in its order.
inline void print_stats_report(std::ostream& os, const ReadStats& r, const clb_encode_stats& s)
{
	struct Line { const char* label; uint64_t value; };
	auto banner = [&os](const char* text) { os << " * * * * * * * * " << text << " * * * * * * * * \n"; };
	auto lines = [&os](std::initializer_list<Line> ls) {
		for (const Line& l : ls) { std::string lab(l.label); lab.resize(31, ' '); os << lab << ": " << l.value << "\n"; }
	};
	banner("READS STATS");
	lines({{"# reads", r.n_reads}, {"min read len", r.min_read_len}, {"max read len", r.max_read_len}, {"# symbols", r.tot_read_len}});
	banner("REFUSE REASONS STATS");
	lines({{"# not enough uniq mmers in enc", s.n_not_enough_unique_mmers_in_enc_read}, {"# too many matches", s.n_too_many_matches}, {"# too low anchors", s.n_too_low_anchors}});
	banner("COMPRESSION STATS");
	lines({{"# plain reads", s.n_plain_reads_tot}, {"# symb plain reads", s.n_plain_symb}, {"# plain reads (reason: N)", s.n_plain_reads_with_n_tot},
		{"# symb plain reads (reason: N)", s.n_plain_with_n_symb}, {"# non rev choosen", s.n_non_rev_choosen}, {"# rev choosen", s.n_rev_choosen}});
	for (uint32_t i = 0; i < s.n_levels && i < CLB_MAX_STAT_LEVELS; ++i) {
		const clb_level_stats& l = s.level[i];
		os << " --------------- level " << i << " --------------- \n";
		lines({{"# alt for left flank", l.n_alternative_left_flank}, {"# alt in between anchors", l.n_alternative_in_between}, {"# alt for right flank", l.n_alternative_right_flank},
			{"# symb plain", l.n_plain_symbols}, {"# symb edit script encoded", l.n_symb_coded_with_edit_script}, {"# symb in edit script", l.n_edit_script_symbols},
			{"# mismatches", l.n_substitution}, {"# matches", l.n_match}, {"# insertions", l.n_insertion}, {"# deletions", l.n_deletion},
			{"# symb anchors", l.n_symb_anchors}, {"# anchors", l.n_anchors}, {"# symb left flank", l.n_left_flank_symb}, {"# symb right flank", l.n_right_flank_symb}});
	}
}

} // namespace clbhost
