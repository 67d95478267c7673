// Test driver for colord_b200/host/presets.h (compiled by tests/test_host_presets.py with g++; host only).
//   preset <compress-ont|compress-pbraw|compress-pbhifi> <ratio|balanced|memory> [quality mode]   the default set as JSON
//   derive <file_bytes> <is_gzip> <is_fastq> <tot_kmers> <modulo> <n_reads> <n_uniq_counted> <range_symbols>   k, anchor, mean read length, sparse range
#include "../colord_b200/host/presets.h"
#include <cinttypes>
#include <cstdio>
#include <cstdlib>

using namespace clbhost;

static void list(const char* name, const std::vector<uint32_t>& v, bool last = false)
{
	std::printf("\"%s\": \"", name);
	for (size_t i = 0; i < v.size(); ++i) std::printf("%s%u", i ? " " : "", v[i]);
	std::printf("\"%s", last ? "" : ", ");
}

int main(int argc, char** argv)
{
	try {
		const std::string cmd = argc > 1 ? argv[1] : "";
		if (cmd == "preset" && argc >= 4) {
			CCompressorParams p = defaultParams(dataSourceFromCommand(argv[2]), compressionPriorityFromString(argv[3]));
			if (argc > 4) { p.qualityComprMode = qualityComprModeFromString(argv[4]); defaultQualityThresholds(p.qualityComprMode, p.qualityFwdThresholds, p.qualityRevThresholds); }
			std::printf("{\"dataSource\": \"%s\", \"priority\": \"%s\", \"compressionLevel\": %d, \"filterHashModulo\": \"%u\", \"maxCandidates\": \"%u\", \"minKmerCount\": \"%u\", \"maxKmerCount\": \"%u\", "
				"\"maxMatchesMultiplier\": \"%g\", \"maxRecurence\": \"%u\", \"minAnchors\": \"%u\", \"minFractionOfMmersInEncode\": \"%g\", \"minFractionOfMmersInEncodeToAlwaysEncode\": \"%g\", "
				"\"minPartLenToConsiderAltRead\": \"%u\", \"qualityComprMode\": \"%s\", \"referenceReadsMode\": \"%s\", \"sparseMode_exponent\": \"%g\", \"sparseMode_range_symbols\": \"%g\", "
				"\"editScriptCostMultiplier\": \"%g\", \"headerComprMode\": \"org\", ",
				dataSourceToString(p.dataSource), compressionPriorityToString(p.priority), p.compressionLevel, p.filterHashModulo, p.maxCandidates, p.minKmerCount, p.maxKmerCount,
				p.maxMatchesMultiplier, p.maxRecurence, p.minAnchors, p.minFractionOfMmersInEncode, p.minFractionOfMmersInEncodeToAlwaysEncode, p.minPartLenToConsiderAltRead,
				qualityComprModeToString(p.qualityComprMode), p.referenceReadsMode == ReferenceReadsMode::All ? "all" : "sparse", p.sparseMode_exponent, p.sparseMode_range_symbols, p.editScriptCostMultiplier);
			list("qualityFwdThresholds", p.qualityFwdThresholds); list("qualityRevThresholds", p.qualityRevThresholds, true);
			std::printf("}\n");
			return 0;
		}
		if (cmd == "derive" && argc == 10) {
			uint32_t k = 0, a = 0;
			adjustKmerAndAnchorLen(k, a, std::atoi(argv[3]) != 0, std::atoi(argv[4]) != 0, std::strtoull(argv[2], nullptr, 10));
			const uint64_t tot_kmers = std::strtoull(argv[5], nullptr, 10), n_reads = std::strtoull(argv[7], nullptr, 10), n_uniq = std::strtoull(argv[8], nullptr, 10);
			const uint32_t modulo = static_cast<uint32_t>(std::atoi(argv[6]));
			const uint64_t mean = meanReadLen(tot_kmers, modulo, n_reads, k);
			std::printf("{\"kmerLen\": %u, \"anchorLen\": %u, \"mean_read_len\": %" PRIu64 ", \"sparse_range_reads\": %u}\n", k, a, mean, sparseModeRange(std::atof(argv[9]), n_uniq, modulo, mean));
			return 0;
		}
		std::fprintf(stderr, "host_presets_tool: bad arguments\n");
		return 2;
	} catch (const std::exception& e) { std::fprintf(stderr, "%s\n", e.what()); return 1; }
}
