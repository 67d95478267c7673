// colord_b200_api.h — library surface for reading archives written by colord-b200: the same classes, member names and
// behaviour as the reference's decompression API (src/API/colord_api.h:27-102, colord_api.cpp:88-105 for Info::ToOstream),
// in namespace colord_b200.  A program written against colord::DecompressionStream (src/API_example/api_example.cpp) ports by
// changing the namespace.  Errors are thrown as std::runtime_error, as the reference's API does (colord_api.cpp:116-123).
// Host code: decoding needs no GPU.  Header-only (C++17); -I the repository root.
#pragma once
#include <ctime>
#include <memory>
#include <ostream>
#include <string>
#include <vector>
#include "../colord_b200/host/decompressor.h"

namespace colord_b200 {

class DecompressionRecord {
	friend class DecompressionStream;
	bool finished = false;
	std::string read_header, read, qual_header, qual;
public:
	operator bool() const { return !finished; }
	const std::string& ReadHeader() const { return read_header; }
	const std::string& Read() const { return read; }
	const std::string& QualHeader() const { return qual_header; }       // empty, or the read header when the '+' line repeated it
	const std::string& Qual() const { return qual; }
};

enum class ReadsSource { ONT, PBRaw, PBHiFi };
enum class QualityCompressionMode { Original, QuinaryAverage, QuadAverage, BinaryAverage, QuinaryThreshold, QuadThreshold, BinaryThreshold, Average, None };
enum class HeaderCompressionMode { Original, Main, None };

struct Info {
	bool isFastq = false;
	uint32_t versionMajor = 0, versionMinor = 0, versionPatch = 0;
	uint64_t totalBytes = 0, totalBases = 0;
	uint32_t totalReads = 0;
	uint64_t time = 0;
	std::string fullCommandLine;
	int32_t compressionLevel{};
	ReadsSource readsSource{};
	QualityCompressionMode qualityCompressionMode{};
	HeaderCompressionMode headerCompressionMode{};
	std::vector<uint32_t> qualityReverseThresholds;

	void ToOstream(std::ostream& oss) const
	{
		static const char* src[] = {"ONT", "PBRaw", "PBHiFi"};
		static const char* qm[] = {"org", "5-avg", "4-avg", "2-avg", "5-fix", "4-fix", "2-fix", "avg", "none"};
		static const char* hm[] = {"org", "main", "none"};
		oss << "is fastq: " << std::boolalpha << isFastq << "\n";
		oss << "colord archive version: " << versionMajor << "." << versionMinor << "." << versionPatch << "\n";
		oss << "total reads: " << totalReads << "\n";
		oss << "colord archive creaton datetime: " << clbhost::CInfo::time_string(time);
		oss << "command line used to create colord archive: " << fullCommandLine << "\n";
		oss << "compression level: " << compressionLevel << "\n";
		auto name = [](const char* const* tab, int n, int v) { return v >= 0 && v < n ? tab[v] : "?"; };
		oss << "reads source: " << name(src, 3, static_cast<int>(readsSource)) << "\n";
		oss << "quality compression mode: " << name(qm, 9, static_cast<int>(qualityCompressionMode)) << "\n";
		oss << "header compression mode: " << name(hm, 3, static_cast<int>(headerCompressionMode)) << "\n";
		oss << "quality reverse thresholds: ";
		for (auto v : qualityReverseThresholds) oss << v << " ";
		oss << "\n";
	}
};

class DecompressionStream {
	std::unique_ptr<clbhost::DecompressedArchive> a;
	uint32_t next = 0;
public:
	explicit DecompressionStream(const std::string& inputFilePath) : a(std::make_unique<clbhost::DecompressedArchive>(inputFilePath)) {}
	// the reference's second constructor (colord_api.h:96): the genome of an archive made with -G and without -s
	DecompressionStream(const std::string& inputFilePath, const std::string& refGenomePath) : a(std::make_unique<clbhost::DecompressedArchive>(inputFilePath, false, refGenomePath)) {}
	Info GetInfo() const
	{
		Info i;
		i.isFastq = a->meta.is_fastq;
		i.versionMajor = a->info.version_major; i.versionMinor = a->info.version_minor; i.versionPatch = a->info.version_patch;
		i.totalBytes = a->info.total_bytes; i.totalBases = a->info.total_bases; i.totalReads = a->info.total_reads; i.time = a->info.time;
		i.fullCommandLine = a->info.full_command_line;
		i.compressionLevel = a->meta.compressionLevel;
		i.readsSource = static_cast<ReadsSource>(a->meta.dataSource);
		i.qualityCompressionMode = static_cast<QualityCompressionMode>(a->meta.qualityComprMode);
		i.headerCompressionMode = static_cast<HeaderCompressionMode>(a->meta.headerComprMode);
		i.qualityReverseThresholds = a->meta.qualityRevThresholds;
		return i;
	}
	DecompressionRecord NextRecord()
	{
		DecompressionRecord r;
		if (next >= a->n_reads()) { r.finished = true; return r; }
		const auto& H = a->headers; const auto& R = a->reads;
		r.read_header.assign(reinterpret_cast<const char*>(H.bytes.data() + H.offsets[next]), H.offsets[next + 1] - H.offsets[next]);
		r.read.assign(reinterpret_cast<const char*>(R.bases.data() + R.offsets[next]), R.offsets[next + 1] - R.offsets[next]);
		if (a->meta.is_fastq) {
			if (H.plus_id[next]) r.qual_header = r.read_header;
			r.qual.assign(reinterpret_cast<const char*>(a->quals.data() + R.offsets[next]), R.offsets[next + 1] - R.offsets[next]);
		}
		++next;
		return r;
	}
};

} // namespace colord_b200

// a program written against the reference's src/API/colord_api.h (namespace colord) compiles unchanged against this header
#ifndef COLORD_B200_NO_COLORD_NAMESPACE
namespace colord = colord_b200;
#endif
