// Test driver for colord_b200/host/decompressor.h and fastq_reader.h (compiled by the tests with g++; host only, no device).
//   hdr <stream> <n> <out>                          decoded headers: per header "<plus>\t<bytes>\n"
//   qavg|qorg <stream> <bases> <offsets> [flags] <out>   bases: ASCII back to back, offsets: u64[n+1], flags: one byte per base
//   dna <stream> <n_reads> <decisions> <out_bases> <out_offsets> <out_flags>
//   stats-format <numbers>                          the block printed from the numbers of a block (format check against the stock binary's text)
//   stats <input>                                   the -v statistics block (stats_report.h) with the reader's read statistics; the encoder's counters: all reads plain
//   parse <input> <out_prefix> [threads min_piece_bytes]   reader: writes <prefix>.bases .offsets .quals .headers .hoff .plus .packs and prints the statistics as JSON
#include "../colord_b200/host/decompressor.h"
#include "../colord_b200/host/fastq_reader.h"
#include "../colord_b200/host/stats_report.h"
#include <chrono>
#include <cinttypes>
#include <fstream>
#include <iostream>

using namespace clbhost;

static std::vector<uint8_t> slurp(const char* p)
{
	std::ifstream f(p, std::ios::binary);
	if (!f) throw std::runtime_error(std::string("cannot open ") + p);
	return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}
static void spit(const std::string& p, const void* d, size_t n) { std::ofstream f(p, std::ios::binary); f.write(static_cast<const char*>(d), static_cast<std::streamsize>(n)); }

int main(int argc, char** argv)
{
	try {
		const std::string cmd = argc > 1 ? argv[1] : "";
		if (cmd == "hdr" && argc == 5) {
			const auto s = slurp(argv[2]);
			const dec::Headers H = dec::decode_headers(s.data(), s.size(), std::strtoull(argv[3], nullptr, 10));
			std::ofstream o(argv[4], std::ios::binary);
			for (size_t r = 0; r + 1 < H.offsets.size(); ++r) { o << int(H.plus_id[r]) << '\t'; o.write(reinterpret_cast<const char*>(H.bytes.data() + H.offsets[r]), static_cast<std::streamsize>(H.offsets[r + 1] - H.offsets[r])); o << '\n'; }
			return 0;
		}
		if ((cmd == "qavg" || cmd == "qorg") && (argc == 6 || argc == 7)) {
			const auto s = slurp(argv[2]);
			dec::Reads R; R.bases = slurp(argv[3]);
			const auto off = slurp(argv[4]);
			R.offsets.assign(reinterpret_cast<const uint64_t*>(off.data()), reinterpret_cast<const uint64_t*>(off.data() + off.size()));
			if (argc == 7) R.flags = slurp(argv[5]); else R.flags.assign(R.bases.size(), 0);
			const auto q = cmd == "qavg" ? dec::decode_qual_avg(s.data(), s.size(), R) : dec::decode_qual_org(s.data(), s.size(), R);
			spit(argv[argc - 1], q.data(), q.size());
			return 0;
		}
		if (cmd == "dna" && argc == 8) {
			const auto s = slurp(argv[2]);
			const uint32_t n = static_cast<uint32_t>(std::strtoul(argv[3], nullptr, 10));
			const auto decs = slurp(argv[4]);
			dec::DnaDecoder D;
			const dec::Reads R = D.decode(s.data(), s.size(), n, decs.data());
			spit(argv[5], R.bases.data(), R.bases.size()); spit(argv[6], R.offsets.data(), 8 * R.offsets.size()); spit(argv[7], R.flags.data(), R.flags.size());
			return 0;
		}
		if (cmd == "parse-time" && argc == 3) {       // reader throughput: parse only, no dumps
			const auto t0 = std::chrono::steady_clock::now();
			const CInputReads in(argv[2]);
			const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
			std::printf("{\"n_reads\": %u, \"total_bytes\": %" PRIu64 ", \"seconds\": %.4f, \"GBps\": %.3f}\n", in.n_reads(), in.total_bytes, s, in.total_bytes / s * 1e-9);
			return 0;
		}
		if (cmd == "stream" && argc == 6) {            // the streaming form of the reader: pieces collected back into whole arrays; exit code 5 = fallback asked for
			std::vector<uint8_t> bases, quals; std::vector<uint64_t> offsets{0};
			const bool dump = std::string(argv[3]) != "-";
			uint64_t n_sunk = 0;
			try {
				const auto t0 = std::chrono::steady_clock::now();
				const CInputReads in(argv[2], [&](const uint8_t* b, const uint8_t* q, const uint64_t* off, uint32_t n) {
					n_sunk += off[n];
					if (!dump) return;
					const uint64_t base = bases.size();
					bases.insert(bases.end(), b, b + off[n]); quals.insert(quals.end(), q, q + off[n]);
					for (uint32_t i = 0; i < n; ++i) offsets.push_back(base + off[i + 1]);
				}, static_cast<unsigned>(std::atoi(argv[4])), std::strtoull(argv[5], nullptr, 10));
				const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
				const std::string pre = argv[3];
				if (dump) {
					spit(pre + ".bases", bases.data(), bases.size()); spit(pre + ".offsets", offsets.data(), 8 * offsets.size()); spit(pre + ".quals", quals.data(), quals.size());
					spit(pre + ".headers", in.headers.data(), in.headers.size()); spit(pre + ".hoff", in.header_offsets.data(), 8 * in.header_offsets.size()); spit(pre + ".plus", in.plus_id.data(), in.plus_id.size());
				}
				std::printf("{\"threads_used\": %u, \"n_reads\": %u, \"total_bytes\": %" PRIu64 ", \"total_bases\": %" PRIu64 ", \"total_symb_header\": %" PRIu64 ", \"sunk\": %" PRIu64 ", \"seconds\": %.4f, \"GBps\": %.3f, \"read_packs\": [",
					in.threads_used, in.n_reads(), in.total_bytes, in.total_bases, in.total_symb_header, n_sunk, sec, in.total_bytes / sec * 1e-9);
				for (size_t i = 0; i < in.read_pack_sizes.size(); ++i) std::printf("%s%u", i ? ", " : "", in.read_pack_sizes[i]);
				std::printf("], \"header_packs\": [");
				for (size_t i = 0; i < in.header_pack_sizes.size(); ++i) std::printf("%s%u", i ? ", " : "", in.header_pack_sizes[i]);
				std::printf("]}\n");
				return 0;
			} catch (const StreamingFallback&) { return 5; }
		}
		if (cmd == "stats-format" && argc == 3) {      // the block printed from numbers given in its own order (4 + 3 + 6, then 14 per level)
			std::ifstream f(argv[2]);
			std::vector<uint64_t> v; for (uint64_t x; f >> x;) v.push_back(x);
			if (v.size() < 13 || (v.size() - 13) % 14) throw std::runtime_error("stats-format: 13 + 14 per level numbers");
			ReadStats r; r.n_reads = v[0]; r.min_read_len = v[1]; r.max_read_len = v[2]; r.tot_read_len = v[3];
			clb_encode_stats e{};
			e.n_not_enough_unique_mmers_in_enc_read = v[4]; e.n_too_many_matches = v[5]; e.n_too_low_anchors = v[6];
			e.n_plain_reads_tot = v[7]; e.n_plain_symb = v[8]; e.n_plain_reads_with_n_tot = v[9]; e.n_plain_with_n_symb = v[10]; e.n_non_rev_choosen = v[11]; e.n_rev_choosen = v[12];
			e.n_levels = static_cast<uint32_t>((v.size() - 13) / 14);
			for (uint32_t l = 0; l < e.n_levels && l < CLB_MAX_STAT_LEVELS; ++l) {
				const uint64_t* x = v.data() + 13 + 14 * l; clb_level_stats& d = e.level[l];
				d.n_alternative_left_flank = x[0]; d.n_alternative_in_between = x[1]; d.n_alternative_right_flank = x[2]; d.n_plain_symbols = x[3];
				d.n_symb_coded_with_edit_script = x[4]; d.n_edit_script_symbols = x[5]; d.n_substitution = x[6]; d.n_match = x[7]; d.n_insertion = x[8]; d.n_deletion = x[9];
				d.n_symb_anchors = x[10]; d.n_anchors = x[11]; d.n_left_flank_symb = x[12]; d.n_right_flank_symb = x[13];
			}
			print_stats_report(std::cout, r, e);
			return 0;
		}
		if (cmd == "stats" && argc == 3) {
			const CInputReads in(argv[2]);
			ReadStats r; r.log_all(in.offsets.data(), in.n_reads());
			clb_encode_stats e{};      // what a run without any overlap between the reads reports: every read plain (with N where it has one)
			for (uint32_t i = 0; i < in.n_reads(); ++i) {
				const uint64_t len = in.offsets[i + 1] - in.offsets[i];
				if (in.has_n[i]) { ++e.n_plain_reads_with_n_tot; e.n_plain_with_n_symb += len; } else { ++e.n_plain_reads_tot; e.n_plain_symb += len; }
			}
			print_stats_report(std::cout, r, e);
			return 0;
		}
		if (cmd == "parse" && (argc == 4 || argc == 6)) {
			const CInputReads in(argv[2], argc == 6 ? static_cast<unsigned>(std::atoi(argv[4])) : 0u, argc == 6 ? std::strtoull(argv[5], nullptr, 10) : (16u << 20));
			const std::string pre = argv[3];
			spit(pre + ".bases", in.bases.data(), in.bases.size()); spit(pre + ".offsets", in.offsets.data(), 8 * in.offsets.size());
			spit(pre + ".quals", in.quals.data(), in.quals.size()); spit(pre + ".headers", in.headers.data(), in.headers.size());
			spit(pre + ".hoff", in.header_offsets.data(), 8 * in.header_offsets.size()); spit(pre + ".plus", in.plus_id.data(), in.plus_id.size());
			spit(pre + ".hasn", in.has_n.data(), in.has_n.size());
			std::printf("{\"threads_used\": %u, \"is_fastq\": %d, \"is_gzip\": %d, \"n_reads\": %u, \"total_bytes\": %" PRIu64 ", \"total_bases\": %" PRIu64 ", \"total_symb_header\": %" PRIu64 ", \"file_bytes\": %" PRIu64 ", \"read_packs\": [",
				in.threads_used, int(in.is_fastq), int(in.is_gzip), in.n_reads(), in.total_bytes, in.total_bases, in.total_symb_header, in.file_bytes);
			for (size_t i = 0; i < in.read_pack_sizes.size(); ++i) std::printf("%s%u", i ? ", " : "", in.read_pack_sizes[i]);
			std::printf("], \"header_packs\": [");
			for (size_t i = 0; i < in.header_pack_sizes.size(); ++i) std::printf("%s%u", i ? ", " : "", in.header_pack_sizes[i]);
			std::printf("]}\n");
			return 0;
		}
		std::cerr << "host_decode_tool: bad arguments\n";
		return 2;
	} catch (const InputError& e) { std::cerr << e.what() << "\n"; return 1; }
	catch (const DecodeError& e) { std::cerr << e.what() << "\n"; return 3; }
	catch (const std::exception& e) { std::cerr << e.what() << "\n"; return 4; }
}
