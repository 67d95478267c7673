// presets.h — compression parameters: the reference's CCompressorParams with its per-mode / per-priority defaults and the values
// runCompression derives from the input (SURVEY.md §8b "outermost contract"; host code, no device work).
//   CCompressorParams, enums            src/colord/params.h:33-110
//   defaults per (mode, priority)       src/colord/arg_parse.cpp:86-376   (compr_<ONT|PBRaw|PBHiFi>_<ratio|balanced|memory>_set_defaults)
//   quality thresholds per -q mode      src/colord/arg_parse.cpp:28-84, :415-451 (adjust_quality_mode_and_thresholds)
//   k / anchor length from the file     src/colord/compression.cpp:41-93   (adjustKmerAndAnchorLen)
//   mean read length, sparse range      src/colord/compression.cpp:443, :501-504
// Checked against what the unmodified reference prints under -v (tests/test_host_presets.py).  Header-only.
#pragma once
#include <cstdint>
#include <ostream>
#include <stdexcept>
#include <string>
#include <vector>
#include "archive_host.h"

namespace clbhost {

enum class CompressionPriority { Ratio, Balanced, Memory };
// how the three stage-3 streams are stored (no counterpart in the reference): Compat = the reference's own streams, an archive the
// unmodified `colord decompress` reads (colord_b200/csrc/stage3_exact.cu); Native = the device's containers; Auto = Compat up to
// compat_max_bases input bases, where it also is the smaller archive, Native above
enum class StreamFormat { Auto, Native, Compat };

struct CCompressorParams {
	DataSource dataSource = DataSource::ONT;
	CompressionPriority priority = CompressionPriority::Memory;
	std::string inputFilePath, outputFilePath;
	uint32_t kmerLen = 0, anchorLen = 0;                   // 0 = adjusted to the input size
	int32_t compressionLevel = 1;
	uint32_t minKmerCount = 4, maxKmerCount = 80, filterHashModulo = 12, maxCandidates = 5;
	double editScriptCostMultiplier = 1.0;
	uint32_t maxRecurence = 3, minPartLenToConsiderAltRead = 64;
	double minFractionOfMmersInEncode = 0.5, minFractionOfMmersInEncodeToAlwaysEncode = 0.9, maxMatchesMultiplier = 10;
	QualityComprMode qualityComprMode = QualityComprMode::QuadAverage;
	std::vector<uint32_t> qualityFwdThresholds, qualityRevThresholds;
	HeaderComprMode headerComprMode = HeaderComprMode::Original;
	ReferenceReadsMode referenceReadsMode = ReferenceReadsMode::Sparse;
	double sparseMode_range_symbols = 1, sparseMode_exponent = 1.0;
	uint32_t minAnchors = 1;
	bool verbose = false;
	uint32_t nThreads = 0;                                 // accepted and printed as the reference does; the device path does not use it
	std::string refGenomePath; bool storeRefGenome = false;
	int device = 0;                                        // CUDA device ordinal (no counterpart in the reference)
	StreamFormat streamFormat = StreamFormat::Auto;
	uint64_t compat_max_bases = 512ull << 20;              // Auto: inputs up to this many bases are written as the reference's own streams
};

inline DataSource dataSourceFromCommand(const std::string& cmd)
{
	if (cmd == "compress-ont") return DataSource::ONT;
	if (cmd == "compress-pbraw") return DataSource::PBRaw;
	if (cmd == "compress-pbhifi") return DataSource::PBHiFi;
	throw std::invalid_argument("unknown compression mode: " + cmd);
}
inline CompressionPriority compressionPriorityFromString(const std::string& s)
{
	if (s == "ratio") return CompressionPriority::Ratio;
	if (s == "balanced") return CompressionPriority::Balanced;
	if (s == "memory") return CompressionPriority::Memory;
	throw std::invalid_argument("unknown priority: " + s);
}
inline const char* compressionPriorityToString(CompressionPriority p) { return p == CompressionPriority::Ratio ? "ratio" : p == CompressionPriority::Balanced ? "balanced" : "memory"; }
inline const char* dataSourceToString(DataSource d) { return d == DataSource::ONT ? "ONT" : d == DataSource::PBRaw ? "PBRaw" : "PBHiFi"; }

// -q names (arg_parse.cpp:476-486) and the thresholds each mode starts from (:28-84): {forward thresholds, values for decompression}
inline QualityComprMode qualityComprModeFromString(const std::string& s)
{
	static const char* names[] = {"org", "5-avg", "4-avg", "2-avg", "5-fix", "4-fix", "2-fix", "avg", "none"};
	for (int i = 0; i < 9; ++i) if (s == names[i]) return static_cast<QualityComprMode>(i);
	throw std::invalid_argument("unknown quality compression mode: " + s);
}
inline const char* qualityComprModeToString(QualityComprMode m)
{
	static const char* names[] = {"org", "5-avg", "4-avg", "2-avg", "5-fix", "4-fix", "2-fix", "avg", "none"};
	return names[static_cast<int>(m)];
}
inline void defaultQualityThresholds(QualityComprMode m, std::vector<uint32_t>& fwd, std::vector<uint32_t>& rev)
{
	switch (m) {
	case QualityComprMode::BinaryThreshold: fwd = {7}; rev = {1, 13}; break;
	case QualityComprMode::QuadThreshold: fwd = {7, 14, 26}; rev = {3, 10, 18, 35}; break;
	case QualityComprMode::QuinaryThreshold: fwd = {7, 14, 26, 93}; rev = {3, 10, 18, 35, 93}; break;
	case QualityComprMode::None: fwd = {}; rev = {0}; break;
	case QualityComprMode::BinaryAverage: fwd = {7}; rev = {}; break;
	case QualityComprMode::QuadAverage: fwd = {7, 14, 26}; rev = {}; break;
	case QualityComprMode::QuinaryAverage: fwd = {7, 14, 26, 93}; rev = {}; break;
	default: fwd = {}; rev = {}; break;                    // org, avg
	}
}

// The nine default sets.  ONT and PBRaw share every number and differ in the quality mode; PBHiFi has its own numbers.
inline CCompressorParams defaultParams(DataSource src, CompressionPriority pri)
{
	CCompressorParams p;
	p.dataSource = src; p.priority = pri;
	const bool hifi = src == DataSource::PBHiFi;
	switch (pri) {
	case CompressionPriority::Ratio:
		p.compressionLevel = 3; p.minKmerCount = 2; p.maxKmerCount = hifi ? 150 : 120; p.filterHashModulo = hifi ? 20 : 8; p.maxCandidates = hifi ? 12 : 10;
		p.maxRecurence = 6; p.minPartLenToConsiderAltRead = 48; p.referenceReadsMode = ReferenceReadsMode::All; p.sparseMode_range_symbols = 1;
		break;
	case CompressionPriority::Balanced:
		p.compressionLevel = 2; p.minKmerCount = 3; p.maxKmerCount = hifi ? 120 : 100; p.filterHashModulo = hifi ? 30 : 9; p.maxCandidates = hifi ? 10 : 8;
		p.maxRecurence = 5; p.minPartLenToConsiderAltRead = 48; p.referenceReadsMode = ReferenceReadsMode::Sparse; p.sparseMode_range_symbols = hifi ? 6 : 2;
		break;
	case CompressionPriority::Memory:
		p.compressionLevel = hifi ? 2 : 1; p.minKmerCount = hifi ? 3 : 4; p.maxKmerCount = hifi ? 100 : 80; p.filterHashModulo = hifi ? 40 : 12; p.maxCandidates = hifi ? 8 : 5;
		p.maxRecurence = hifi ? 5 : 3; p.minPartLenToConsiderAltRead = hifi ? 48 : 64; p.referenceReadsMode = ReferenceReadsMode::Sparse; p.sparseMode_range_symbols = hifi ? 3 : 1;
		break;
	}
	p.qualityComprMode = src == DataSource::ONT ? QualityComprMode::QuadAverage : src == DataSource::PBRaw ? QualityComprMode::None : QualityComprMode::QuinaryAverage;
	defaultQualityThresholds(p.qualityComprMode, p.qualityFwdThresholds, p.qualityRevThresholds);
	return p;
}

// compression.cpp:165-207 (PrintParams): the parameter block the reference prints under -v, same lines and wording
inline void PrintParams(std::ostream& os, const CCompressorParams& p, uint32_t kmerLen, uint32_t anchorLen, unsigned nThreads)
{
	auto list = [](const std::vector<uint32_t>& v) { std::string r; for (uint32_t x : v) r += " " + std::to_string(x); return r; };      // every value preceded by a blank (arg_parse.cpp vec_to_string)
	os << " * * * * * * * * * * * * Parameters * * * * * * * * * * * * \n";
	os << "\tinput file path: " << p.inputFilePath << "\n\toutput file path: " << p.outputFilePath << "\n\tnumber of threads: " << nThreads << "\n";
	os << "\tk-mer length: " << kmerLen << "\n\tanchor length: " << anchorLen << "\n\tdata source type: " << dataSourceToString(p.dataSource) << "\n";
	os << "\tmultipier for predicted cost of storing read part as edit script: " << p.editScriptCostMultiplier << "\n\tfilter modulo: " << p.filterHashModulo << "\n";
	os << "\theader compression mode: " << (p.headerComprMode == HeaderComprMode::Original ? "org" : p.headerComprMode == HeaderComprMode::Main ? "main" : "none") << "\n";
	os << "\tmax candidates: " << p.maxCandidates << "\n\tmin k-mer count: " << p.minKmerCount << "\n\tmax k-mer count: " << p.maxKmerCount << "\n";
	os << "\tmax matches multiplier: " << p.maxMatchesMultiplier << "\n\tmax recurence: " << p.maxRecurence << "\n\tmin anchors: " << p.minAnchors << "\n";
	os << "\tmin fraction of m-mers in encode: " << p.minFractionOfMmersInEncode << "\n\tmin fraction of m-mers in encode to always encode: " << p.minFractionOfMmersInEncodeToAlwaysEncode << "\n";
	os << "\tmin part length to consider alternative reference read: " << p.minPartLenToConsiderAltRead << "\n\tcompression priority: " << compressionPriorityToString(p.priority) << "\n";
	os << "\tquality compression mode: " << qualityComprModeToString(p.qualityComprMode) << "\n\tquality thresholds: " << list(p.qualityFwdThresholds) << "\n\tquality values: " << list(p.qualityRevThresholds) << "\n";
	os << "\treference reads mode: " << (p.referenceReadsMode == ReferenceReadsMode::All ? "all" : "sparse") << "\n\tsparse mode exponent: " << p.sparseMode_exponent << "\n\tsparse mode range: " << p.sparseMode_range_symbols << "\n";
	os << "\tfill factor filtered k-mers: 0.75\n\tfill factor k-mers to reads: 0.8\n";          // the reference's hash-table fill factors; the device tables have their own
}

// compression.cpp:41-93: the estimate of the number of bases from the file size, then the table
inline void adjustKmerAndAnchorLen(uint32_t& kmerLen, uint32_t& anchorLen, bool is_gzip_input, bool is_fastq, uint64_t file_bytes)
{
	if (kmerLen && anchorLen) return;
	const double f = is_gzip_input ? (is_fastq ? 2.08 : 3.98) : (is_fastq ? 0.49 : 0.98);
	const uint64_t base_count = static_cast<uint64_t>(f * file_bytes);
	if (base_count < 1'000'000'000ull) { kmerLen = 20; anchorLen = 16; }
	else if (base_count < 4'000'000'000ull) { kmerLen = 21; anchorLen = 18; }
	else if (base_count < 16'000'000'000ull) { kmerLen = 23; anchorLen = 21; }
	else if (base_count < 48'000'000'000ull) { kmerLen = 24; anchorLen = 22; }
	else if (base_count < 128'000'000'000ull) { kmerLen = 25; anchorLen = 22; }
	else { kmerLen = 26; anchorLen = 23; }
}
// compression.cpp:443
inline uint64_t meanReadLen(uint64_t tot_kmers, uint32_t modulo, uint64_t tot_n_reads, uint32_t kmerLen)
{
	return static_cast<uint64_t>(double(tot_kmers * modulo) / tot_n_reads + kmerLen - 1);
}
// compression.cpp:501-504
inline uint32_t sparseModeRange(double range_symbols, uint64_t n_uniq_counted_kmers, uint32_t modulo, uint64_t mean_read_len)
{
	const uint32_t r = static_cast<uint32_t>((range_symbols * n_uniq_counted_kmers * modulo) / mean_read_len);
	return r ? r : 1;
}

} // namespace clbhost
