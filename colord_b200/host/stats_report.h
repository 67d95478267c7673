// stats_report.h — the statistics block the reference prints at the end of a verbose compression (src/colord/stats_collector.cpp:100-146):
// the reader's read statistics (in_reads.cpp:71 -> CStatsCollector::LogRead, :152-160) and the encoder's counters, which the device
// collects (clb_encode_stats_enable / clb_encode_stats_get: encoder.cpp:663-676, :1069-1190, :1445-1575).
#pragma once
#include <cstdint>
#include <limits>
#include <ostream>
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

inline void print_stats_report(std::ostream& summary, const ReadStats& r, const clb_encode_stats& s)
{
	summary << " * * * * * * * * READS STATS * * * * * * * * \n";
	summary << "# reads                        : " << r.n_reads << "\n";
	summary << "min read len                   : " << r.min_read_len << "\n";
	summary << "max read len                   : " << r.max_read_len << "\n";
	summary << "# symbols                      : " << r.tot_read_len << "\n";
	summary << " * * * * * * * * REFUSE REASONS STATS * * * * * * * * \n";
	summary << "# not enough uniq mmers in enc : " << s.n_not_enough_unique_mmers_in_enc_read << "\n";
	summary << "# too many matches             : " << s.n_too_many_matches << "\n";
	summary << "# too low anchors              : " << s.n_too_low_anchors << "\n";
	summary << " * * * * * * * * COMPRESSION STATS * * * * * * * * \n";
	summary << "# plain reads                  : " << s.n_plain_reads_tot << "\n";
	summary << "# symb plain reads             : " << s.n_plain_symb << "\n";
	summary << "# plain reads (reason: N)      : " << s.n_plain_reads_with_n_tot << "\n";
	summary << "# symb plain reads (reason: N) : " << s.n_plain_with_n_symb << "\n";
	summary << "# non rev choosen              : " << s.n_non_rev_choosen << "\n";
	summary << "# rev choosen                  : " << s.n_rev_choosen << "\n";
	for (uint32_t i = 0; i < s.n_levels && i < CLB_MAX_STAT_LEVELS; ++i) {
		const clb_level_stats& l = s.level[i];
		summary << " --------------- level " << i << " --------------- \n";
		summary << "# alt for left flank           : " << l.n_alternative_left_flank << "\n";
		summary << "# alt in between anchors       : " << l.n_alternative_in_between << "\n";
		summary << "# alt for right flank          : " << l.n_alternative_right_flank << "\n";
		summary << "# symb plain                   : " << l.n_plain_symbols << "\n";
		summary << "# symb edit script encoded     : " << l.n_symb_coded_with_edit_script << "\n";
		summary << "# symb in edit script          : " << l.n_edit_script_symbols << "\n";
		summary << "# mismatches                   : " << l.n_substitution << "\n";
		summary << "# matches                      : " << l.n_match << "\n";
		summary << "# insertions                   : " << l.n_insertion << "\n";
		summary << "# deletions                    : " << l.n_deletion << "\n";
		summary << "# symb anchors                 : " << l.n_symb_anchors << "\n";
		summary << "# anchors                      : " << l.n_anchors << "\n";
		summary << "# symb left flank              : " << l.n_left_flank_symb << "\n";
		summary << "# symb right flank             : " << l.n_right_flank_symb << "\n";
	}
}

} // namespace clbhost
