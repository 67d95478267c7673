// Test driver for colord_b200/host/archive_host.h (compiled by tests/test_host_archive.py with g++; host only).
//   rewrite <in> <out>   read every part of every stream of <in>, write them to <out> in file order through the writer
//   dump <in>            streams, part tables, the parsed `meta` and `info` records as JSON on stdout
//   meta-roundtrip <in>  parse `meta` and `info`, serialise them again, exit 0 iff the bytes are the same
//   make <out>           write a small archive from scratch (streams in the reference's order, AddPartPrepare/Complete out of order)
#include "../colord_b200/host/archive_host.h"
#include <algorithm>
#include <cinttypes>
#include <cstdlib>
#include <tuple>

using namespace clbhost;

static int fail(const char* what) { std::fprintf(stderr, "host_archive_tool: %s\n", what); return 2; }

static bool load_records(CArchive& in, CMeta& meta, CInfo& info, std::vector<uint8_t>& meta_raw, std::vector<uint8_t>& info_raw)
{
	size_t md;
	const int s_meta = in.GetStreamId("meta"), s_info = in.GetStreamId("info");
	if (s_meta < 0 || s_info < 0) return false;
	if (!in.ReadPart(s_meta, 0, meta_raw, md) || !in.ReadPart(s_info, 0, info_raw, md)) return false;
	meta.Deserialize(meta_raw, in.GetStreamId("qual") >= 0 || in.GetStreamId("qual-b200") >= 0);
	info.Deserialize(info_raw);
	return true;
}

int main(int argc, char** argv)
{
	if (argc < 3) return fail("usage: rewrite <in> <out> | dump <in> | meta-roundtrip <in> | make <out>");
	const std::string cmd = argv[1];
	if (cmd == "make") {
		CArchive out(false);
		if (!out.Open(argv[2])) return fail("cannot create the output");
		const int s_meta = out.RegisterStream("meta"), s_dna = out.RegisterStream("dna"), s_qual = out.RegisterStream("qual");
		const int p0 = out.AddPartPrepare(s_dna), p1 = out.AddPartPrepare(s_dna);
		std::vector<uint8_t> a(300, 7), b(5, 9), e;
		out.AddPartComplete(s_dna, p1, b, 1234567);          // second slot first: part order stays slot order
		out.AddPart(s_qual, e, 0);                           // empty part
		out.AddPartComplete(s_dna, p0, a, 0);
		out.SetRawSize(s_dna, 99);
		CMeta m; m.tot_ref_reads = 3; m.maxCandidates = 5; m.compressionLevel = 1; m.approx_stream_size = 1ull << 33;
		m.qualityComprMode = QualityComprMode::QuadThreshold; m.qualityRevThresholds = {3, 10, 18, 35};
		m.sparseMode_range = 1000; m.sparseMode_exponent = 1.25;
		m.ref_genome_available = true; m.storeRefGenome = false; m.ref_genome_read_len = 160000; m.ref_genome_overlap_size = 230; m.n_ref_genome_pseudo_reads = 7;
		m.ref_genome_checksum.assign(16, 0xab);
		out.AddPart(s_meta, m.Serialize(), 0);
		const int s_info = out.RegisterStream("info");
		CInfo i; i.version_major = 1; i.version_minor = 2; i.version_patch = 1; i.total_bytes = 10; i.total_bases = 4; i.total_reads = 1; i.time = 1700000000; i.full_command_line = "made by host_archive_tool";
		out.AddPart(s_info, i.Serialize(), 0);
		return out.Close() ? 0 : fail("close failed");
	}
	CArchive in(true);
	if (!in.Open(argv[2])) return fail("cannot open the input archive");
	if (cmd == "rewrite") {
		if (argc < 4) return fail("rewrite needs an output path");
		CArchive out(false);
		if (!out.Open(argv[3])) return fail("cannot create the output");
		std::vector<std::tuple<uint64_t, int, size_t>> order;           // (offset, stream, part)
		for (size_t s = 0; s < in.GetNoStreams(); ++s) {
			out.RegisterStream(in.GetStreamName((int)s));
			out.SetRawSize((int)s, in.GetRawSize((int)s));
			const auto parts = in.Parts((int)s);
			for (size_t p = 0; p < parts.size(); ++p) { order.emplace_back(parts[p].offset, (int)s, p); out.AddPartPrepare((int)s); }
		}
		std::sort(order.begin(), order.end());
		std::vector<uint8_t> data; size_t md;
		for (const auto& [off, s, p] : order) {
			(void)off;
			if (!in.ReadPart(s, p, data, md)) return fail("cannot read a part");
			if (!out.AddPartComplete(s, (int)p, data, md)) return fail("cannot write a part");
		}
		return out.Close() ? 0 : fail("close failed");
	}
	if (cmd == "set-header-mode") {      // <in> <out> <0|1|2>: the same archive with the header stream emptied and meta.headerComprMode changed
		if (argc < 5) return fail("set-header-mode needs <in> <out> <mode>");
		CArchive out(false);
		if (!out.Open(argv[3])) return fail("cannot create the output");
		CMeta m; CInfo i; std::vector<uint8_t> mr, ir;
		if (!load_records(in, m, i, mr, ir)) return fail("no meta / info stream");
		m.headerComprMode = static_cast<HeaderComprMode>(std::atoi(argv[4]));
		for (size_t s = 0; s < in.GetNoStreams(); ++s) {
			const std::string name = in.GetStreamName((int)s);
			const int id = out.RegisterStream(name);
			const auto parts = in.Parts((int)s);
			for (size_t p = 0; p < parts.size(); ++p) {
				std::vector<uint8_t> d; size_t md = 0;
				if (!in.ReadPart((int)s, p, d, md)) return fail("cannot read a part");
				if (name == "header-b200") d.clear();
				if (name == "meta") d = m.Serialize();
				out.AddPart(id, d, md);
			}
		}
		return out.Close() ? 0 : fail("close failed");
	}
	CMeta meta; CInfo info; std::vector<uint8_t> meta_raw, info_raw;
	try {
		if (!load_records(in, meta, info, meta_raw, info_raw)) return fail("no meta / info stream");
	} catch (const std::exception& e) { return fail(e.what()); }
	if (cmd == "meta-roundtrip") return (meta.Serialize() == meta_raw && info.Serialize() == info_raw) ? 0 : 1;
	if (cmd != "dump") return fail("unknown command");
	std::printf("{\"streams\": [");
	for (size_t s = 0; s < in.GetNoStreams(); ++s) {
		const auto parts = in.Parts((int)s);
		std::printf("%s{\"name\": \"%s\", \"raw_size\": %zu, \"packed_data_size\": %zu, \"parts\": [", s ? ", " : "", in.GetStreamName((int)s).c_str(), in.GetRawSize((int)s), in.GetStreamPackedDataSize((int)s));
		for (size_t p = 0; p < parts.size(); ++p) {
			std::vector<uint8_t> d; size_t md = 0;
			in.ReadPart((int)s, p, d, md);
			std::printf("%s[%" PRIu64 ", %" PRIu64 ", %zu]", p ? ", " : "", parts[p].offset, parts[p].size, md);
		}
		std::printf("]}");
	}
	std::printf("], \"meta\": {\"tot_ref_reads\": %u, \"maxCandidates\": %u, \"compressionLevel\": %d, \"dataSource\": %u, \"approx_stream_size\": %" PRIu64 ", \"is_fastq\": %d, \"qualityComprMode\": %u, \"qualityRevThresholds\": [",
		meta.tot_ref_reads, meta.maxCandidates, meta.compressionLevel, (unsigned)meta.dataSource, meta.approx_stream_size, (int)meta.is_fastq, (unsigned)meta.qualityComprMode);
	for (size_t i = 0; i < meta.qualityRevThresholds.size(); ++i) std::printf("%s%u", i ? ", " : "", meta.qualityRevThresholds[i]);
	std::printf("], \"headerComprMode\": %u, \"referenceReadsMode\": %u, \"sparseMode_range\": %u, \"sparseMode_exponent\": %.17g, \"ref_genome_available\": %d, \"storeRefGenome\": %d, "
		"\"ref_genome_read_len\": %u, \"ref_genome_overlap_size\": %u, \"n_ref_genome_pseudo_reads\": %u, \"ref_genome_checksum\": \"",
		(unsigned)meta.headerComprMode, (unsigned)meta.referenceReadsMode, meta.sparseMode_range, meta.sparseMode_exponent, (int)meta.ref_genome_available, (int)meta.storeRefGenome,
		meta.ref_genome_read_len, meta.ref_genome_overlap_size, meta.n_ref_genome_pseudo_reads);
	for (uint8_t c : meta.ref_genome_checksum) std::printf("%02x", c);
	std::string cl;
	for (char ch : info.full_command_line) { if (ch == '"' || ch == '\\') cl.push_back('\\'); cl.push_back(ch); }
	std::printf("\"}, \"info\": {\"version\": [%u, %u, %u], \"total_bytes\": %" PRIu64 ", \"total_bases\": %" PRIu64 ", \"total_reads\": %u, \"time\": %" PRIu64 ", \"command\": \"%s\"}}\n",
		info.version_major, info.version_minor, info.version_patch, info.total_bytes, info.total_bases, info.total_reads, info.time, cl.c_str());
	return 0;
}
