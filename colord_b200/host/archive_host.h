// archive_host.h — host-side mirror of the reference's archive container and of the `meta` / `info` streams
// (pure host code, no device work).  Same public names and argument meaning as
//   CArchive            src/colord/archive.h:31-113, archive.cpp:50-361
//   CInfo               src/colord/utils.h:678-698, utils.cpp:326-363
//   the `meta` record   written at src/colord/compression.cpp:705-779, read at decompression_common.cpp:54-250
// so that a file written here is opened by the reference (`colord info`, CArchive::Open) and a file written by the
// reference is read here, byte for byte (tests/test_host_archive.py rewrites reference archives and compares the bytes).
//
// On-disk layout (all of it restated from archive.cpp):
//   part      = varint(metadata) ++ payload            appended in arrival order, streams interleaved
//   varint(x) = one byte n = number of significant bytes of x, then those n bytes, most significant first (x = 0 -> "00")
//   footer    = varint(n_streams) ++ per stream in id order [ name ++ 00 | varint(n_parts) | varint(raw_size) |
//               per part: varint(offset of the part's varint) varint(payload size) ]
//   trailer   = footer size as 8 bytes, little endian
// Differences in construction (not in format): the reference keeps a FILE* with a 64 MiB stdio buffer behind one mutex; here
// the index lives in plain vectors, parts go out through one pwrite-style append, and readers get the part table itself
// (Parts()) so that several parts can be fetched without walking a cursor.  Header-only.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

namespace clbhost {

// `info` version of archives whose streams are the device's native containers (compressor.h); the reference's own archives are 1.x
constexpr uint32_t B200_VERSION_MAJOR = 201, B200_VERSION_MINOR = 1, B200_VERSION_PATCH = 0;

class CArchive {
public:
	struct Part { uint64_t offset, size; };
private:
	struct Stream { std::string name; uint64_t raw_size = 0, packed_size = 0, packed_data_size = 0; size_t cursor = 0; std::vector<Part> parts; };
	bool input_mode;
	FILE* f = nullptr;
	uint64_t f_offset = 0;
	std::vector<Stream> streams;
	mutable std::mutex mtx;

	static void put_varint(std::vector<uint8_t>& o, uint64_t x)
	{
		int n = 0;
		for (uint64_t t = x; t; t >>= 8) ++n;
		o.push_back(static_cast<uint8_t>(n));
		for (int i = n; i-- > 0;) o.push_back(static_cast<uint8_t>(x >> (8 * i)));
	}
	// returns false at end of data / on a length byte no writer produces
	static bool get_varint(const uint8_t*& p, const uint8_t* end, uint64_t& x)
	{
		if (p >= end) return false;
		const int n = *p++;
		if (n > 8 || end - p < n) return false;
		x = 0;
		for (int i = 0; i < n; ++i) x = (x << 8) | *p++;
		return true;
	}
	bool write_footer()
	{
		std::vector<uint8_t> foot;
		put_varint(foot, streams.size());
		for (Stream& s : streams) {
			const size_t at = foot.size();
			foot.insert(foot.end(), s.name.begin(), s.name.end());
			foot.push_back(0);
			put_varint(foot, s.parts.size());
			put_varint(foot, s.raw_size);
			for (const Part& p : s.parts) { put_varint(foot, p.offset); put_varint(foot, p.size); }
			s.packed_size += foot.size() - at;           // the reference counts a stream's footer bytes into its size (archive.cpp:184)
		}
		const uint64_t n = foot.size();
		for (int b = 0; b < 8; ++b) foot.push_back(static_cast<uint8_t>(n >> (8 * b)));
		return std::fwrite(foot.data(), 1, foot.size(), f) == foot.size();
	}
	bool read_footer()
	{
		if (std::fseek(f, 0, SEEK_END)) return false;
		const long long total = std::ftell(f);
		if (total < 8) return false;
		uint8_t tail[8];
		if (std::fseek(f, (long)(total - 8), SEEK_SET) || std::fread(tail, 1, 8, f) != 8) return false;
		uint64_t n = 0;
		for (int b = 0; b < 8; ++b) n |= static_cast<uint64_t>(tail[b]) << (8 * b);
		if (n > static_cast<uint64_t>(total - 8)) return false;
		std::vector<uint8_t> foot(n);
		if (std::fseek(f, (long)(total - 8 - (long long)n), SEEK_SET) || (n && std::fread(foot.data(), 1, n, f) != n)) return false;
		const uint8_t* p = foot.data(); const uint8_t* end = p + n;
		uint64_t n_streams;
		if (!get_varint(p, end, n_streams)) return false;
		streams.clear();
		for (uint64_t i = 0; i < n_streams; ++i) {
			Stream s;
			const uint8_t* z = static_cast<const uint8_t*>(std::memchr(p, 0, end - p));
			if (!z) return false;
			s.name.assign(reinterpret_cast<const char*>(p), z - p);
			p = z + 1;
			uint64_t n_parts;
			if (!get_varint(p, end, n_parts) || !get_varint(p, end, s.raw_size)) return false;
			s.parts.resize(n_parts);
			for (Part& q : s.parts) {
				if (!get_varint(p, end, q.offset) || !get_varint(p, end, q.size)) return false;
				s.packed_data_size += q.size;
			}
			streams.push_back(std::move(s));
		}
		return true;
	}
	bool append(Part& slot, Stream& s, const uint8_t* data, uint64_t size, uint64_t metadata)
	{
		std::vector<uint8_t> head;
		put_varint(head, metadata);
		slot = Part{f_offset, size};
		if (std::fwrite(head.data(), 1, head.size(), f) != head.size()) return false;
		if (size && std::fwrite(data, 1, size, f) != size) return false;
		f_offset += head.size() + size;
		s.packed_size += head.size() + size;
		s.packed_data_size += size;
		return true;
	}
public:
	explicit CArchive(bool _input_mode) : input_mode(_input_mode) {}
	CArchive(const CArchive&) = delete;
	CArchive& operator=(const CArchive&) = delete;
	~CArchive() { if (f) Close(); }

	bool Open(const std::string& file_name)
	{
		std::lock_guard<std::mutex> lck(mtx);
		if (f) std::fclose(f);
		f = std::fopen(file_name.c_str(), input_mode ? "rb" : "wb");
		if (!f) return false;
		f_offset = 0;
		if (input_mode && !read_footer()) { std::fclose(f); f = nullptr; return false; }
		return true;
	}
	bool Close()
	{
		std::lock_guard<std::mutex> lck(mtx);
		if (!f) return false;
		bool ok = true;
		if (!input_mode) ok = write_footer();
		ok = (std::fclose(f) == 0) && ok;
		f = nullptr;
		return ok;
	}
	bool IsOpen() const { std::lock_guard<std::mutex> lck(mtx); return f != nullptr; }
	// a failed run: close without writing a footer (the caller removes the file)
	void Abandon() { std::lock_guard<std::mutex> lck(mtx); if (f) { std::fclose(f); f = nullptr; } }
	int RegisterStream(const std::string& stream_name)
	{
		std::lock_guard<std::mutex> lck(mtx);
		Stream s; s.name = stream_name;
		streams.push_back(std::move(s));
		return static_cast<int>(streams.size()) - 1;
	}
	int GetStreamId(const std::string& stream_name) const
	{
		std::lock_guard<std::mutex> lck(mtx);
		for (size_t i = 0; i < streams.size(); ++i) if (streams[i].name == stream_name) return static_cast<int>(i);
		return -1;
	}
	size_t GetNoStreams() const { std::lock_guard<std::mutex> lck(mtx); return streams.size(); }
	std::string GetStreamName(int stream_id) const { std::lock_guard<std::mutex> lck(mtx); return valid(stream_id) ? streams[stream_id].name : std::string(); }
	size_t GetStreamPackedSize(int stream_id) const { std::lock_guard<std::mutex> lck(mtx); return valid(stream_id) ? streams[stream_id].packed_size : 0; }
	size_t GetStreamPackedDataSize(int stream_id) const { std::lock_guard<std::mutex> lck(mtx); return valid(stream_id) ? streams[stream_id].packed_data_size : 0; }
	void SetRawSize(int stream_id, size_t raw_size) { std::lock_guard<std::mutex> lck(mtx); if (valid(stream_id)) streams[stream_id].raw_size = raw_size; }
	size_t GetRawSize(int stream_id) const { std::lock_guard<std::mutex> lck(mtx); return valid(stream_id) ? streams[stream_id].raw_size : 0; }
	// the part table of a stream (offset = where the part's metadata varint starts)
	std::vector<Part> Parts(int stream_id) const { std::lock_guard<std::mutex> lck(mtx); return valid(stream_id) ? streams[stream_id].parts : std::vector<Part>(); }

	bool AddPart(int stream_id, const uint8_t* data, size_t size, size_t metadata = 0)
	{
		std::lock_guard<std::mutex> lck(mtx);
		if (!f || input_mode || !valid(stream_id)) return false;
		Stream& s = streams[stream_id];
		s.parts.push_back(Part{0, 0});
		return append(s.parts.back(), s, data, size, metadata);
	}
	bool AddPart(int stream_id, const std::vector<uint8_t>& v_data, size_t metadata = 0) { return AddPart(stream_id, v_data.data(), v_data.size(), metadata); }
	// a slot in the stream's part order now, its bytes later (entr_read.h / entr_qual.h use this to keep part order = pack order)
	int AddPartPrepare(int stream_id)
	{
		std::lock_guard<std::mutex> lck(mtx);
		if (!valid(stream_id)) return -1;
		streams[stream_id].parts.push_back(Part{0, 0});
		return static_cast<int>(streams[stream_id].parts.size()) - 1;
	}
	bool AddPartComplete(int stream_id, int part_id, const std::vector<uint8_t>& v_data, size_t metadata = 0)
	{
		std::lock_guard<std::mutex> lck(mtx);
		if (!f || input_mode || !valid(stream_id) || part_id < 0 || static_cast<size_t>(part_id) >= streams[stream_id].parts.size()) return false;
		Stream& s = streams[stream_id];
		return append(s.parts[part_id], s, v_data.data(), v_data.size(), metadata);
	}
	// part `part_id` of a stream, wherever the cursor is
	bool ReadPart(int stream_id, size_t part_id, std::vector<uint8_t>& v_data, size_t& metadata)
	{
		std::lock_guard<std::mutex> lck(mtx);
		if (!f || !input_mode || !valid(stream_id) || part_id >= streams[stream_id].parts.size()) return false;
		return fetch(streams[stream_id].parts[part_id], v_data, metadata);
	}
	// next part of a stream (the reference's cursor interface); false when the stream is exhausted
	bool GetPart(int stream_id, std::vector<uint8_t>& v_data, size_t& metadata)
	{
		std::lock_guard<std::mutex> lck(mtx);
		if (!f || !input_mode || !valid(stream_id)) return false;
		Stream& s = streams[stream_id];
		if (s.cursor >= s.parts.size()) return false;
		return fetch(s.parts[s.cursor++], v_data, metadata);
	}
private:
	bool valid(int id) const { return id >= 0 && static_cast<size_t>(id) < streams.size(); }
	bool fetch(const Part& p, std::vector<uint8_t>& v_data, size_t& metadata)
	{
		v_data.resize(p.size);
		metadata = 0;
		if (p.size == 0) return true;                     // archive.cpp:340-348: an empty part's metadata is not read back
		if (std::fseek(f, (long)p.offset, SEEK_SET)) return false;
		uint8_t head[9];
		const size_t got = std::fread(head, 1, 9, f);
		const uint8_t* q = head; uint64_t m;
		if (!get_varint(q, head + got, m)) return false;
		metadata = m;
		if (std::fseek(f, (long)(p.offset + (q - head)), SEEK_SET)) return false;
		return std::fread(v_data.data(), 1, p.size, f) == p.size;
	}
};

// ---- little-endian fields of `meta` / `info` (utils.h:484-538; doubles as their raw IEEE bytes, :498-517) ----
namespace le {
template <typename T> inline void put(std::vector<uint8_t>& o, T v) { for (size_t b = 0; b < sizeof(T); ++b) o.push_back(static_cast<uint8_t>(static_cast<uint64_t>(v) >> (8 * b))); }
inline void put_double(std::vector<uint8_t>& o, double v) { uint64_t u; std::memcpy(&u, &v, 8); put<uint64_t>(o, u); }
struct Reader {
	const uint8_t* p; const uint8_t* end;
	template <typename T> T get()
	{
		if (static_cast<size_t>(end - p) < sizeof(T)) throw std::runtime_error("colord_b200: truncated archive record");
		uint64_t v = 0;
		for (size_t b = 0; b < sizeof(T); ++b) v |= static_cast<uint64_t>(*p++) << (8 * b);
		return static_cast<T>(v);
	}
	double get_double() { const uint64_t u = get<uint64_t>(); double d; std::memcpy(&d, &u, 8); return d; }
};
}

// utils.h:678-698
struct CInfo {
	uint32_t version_major = 0, version_minor = 0, version_patch = 0;
	uint64_t total_bytes = 0, total_bases = 0;
	uint32_t total_reads = 0;
	uint64_t time = 0;
	std::string full_command_line;

	// asctime(localtime(&t)) of the reference (info.cpp:51, colord_api.cpp) for a time taken from a file: localtime has no answer for every 64-bit value
	static std::string time_string(uint64_t time)
	{
		const time_t t = static_cast<time_t>(time); struct tm tmv; char buf[64];
		if (!localtime_r(&t, &tmv) || !asctime_r(&tmv, buf)) return "(time out of range)\n";
		return buf;
	}
	std::vector<uint8_t> Serialize() const
	{
		std::vector<uint8_t> r;
		le::put(r, version_major); le::put(r, version_minor); le::put(r, version_patch);
		le::put(r, total_bytes); le::put(r, total_bases); le::put(r, total_reads); le::put(r, time);
		le::put(r, static_cast<uint32_t>(full_command_line.size()));
		r.insert(r.end(), full_command_line.begin(), full_command_line.end());
		return r;
	}
	void Deserialize(const std::vector<uint8_t>& data)
	{
		le::Reader in{data.data(), data.data() + data.size()};
		version_major = in.get<uint32_t>(); version_minor = in.get<uint32_t>(); version_patch = in.get<uint32_t>();
		total_bytes = in.get<uint64_t>(); total_bases = in.get<uint64_t>(); total_reads = in.get<uint32_t>(); time = in.get<uint64_t>();
		const uint32_t n = in.get<uint32_t>();
		if (static_cast<size_t>(in.end - in.p) < n) throw std::runtime_error("colord_b200: truncated info record");
		full_command_line.assign(reinterpret_cast<const char*>(in.p), n);
	}
};

// params.h:33-46 (the values are what the archive stores)
enum class QualityComprMode : uint8_t { Original, QuinaryAverage, QuadAverage, BinaryAverage, QuinaryThreshold, QuadThreshold, BinaryThreshold, Average, None };
enum class HeaderComprMode : uint8_t { Original, Main, None };
enum class ReferenceReadsMode : uint8_t { All, Sparse };
enum class DataSource : uint8_t { ONT, PBRaw, PBHiFi };

// The `meta` record: field order of compression.cpp:705-779 = decompression_common.cpp:54-250.  Whether the quality fields are
// present is not stored in the record: the reader knows it from the presence of a "qual" stream (decompression_common.cpp:44-48).
struct CMeta {
	uint32_t tot_ref_reads = 0, maxCandidates = 0;
	int32_t compressionLevel = 0;
	DataSource dataSource = DataSource::ONT;
	uint64_t approx_stream_size = 0;                       // tot_n_reads * mean_read_len
	bool is_fastq = true;
	QualityComprMode qualityComprMode = QualityComprMode::QuadAverage;
	std::vector<uint32_t> qualityRevThresholds;            // None: 1, Binary/Quad/QuinaryThreshold: 2 / 4 / 5, the others: 0
	HeaderComprMode headerComprMode = HeaderComprMode::Original;
	ReferenceReadsMode referenceReadsMode = ReferenceReadsMode::Sparse;
	uint32_t sparseMode_range = 0; double sparseMode_exponent = 0;     // Sparse only
	bool ref_genome_available = false, storeRefGenome = false;
	uint32_t ref_genome_read_len = 0, ref_genome_overlap_size = 0, n_ref_genome_pseudo_reads = 0;
	std::vector<uint8_t> ref_genome_checksum;              // 16 bytes (MD5), only when the genome is not stored in the archive

	static size_t n_thresholds(QualityComprMode m)
	{
		switch (m) {
		case QualityComprMode::None: return 1;
		case QualityComprMode::BinaryThreshold: return 2;
		case QualityComprMode::QuadThreshold: return 4;
		case QualityComprMode::QuinaryThreshold: return 5;
		default: return 0;
		}
	}
	std::vector<uint8_t> Serialize() const
	{
		std::vector<uint8_t> r;
		le::put(r, tot_ref_reads); le::put(r, maxCandidates); le::put(r, static_cast<uint32_t>(compressionLevel));
		r.push_back(static_cast<uint8_t>(dataSource));
		le::put(r, approx_stream_size);
		if (is_fastq) {
			r.push_back(static_cast<uint8_t>(qualityComprMode));
			if (qualityRevThresholds.size() != n_thresholds(qualityComprMode)) throw std::invalid_argument("colord_b200: qualityRevThresholds do not fit the quality mode");
			for (uint32_t t : qualityRevThresholds) le::put(r, t);
		}
		r.push_back(static_cast<uint8_t>(headerComprMode));
		r.push_back(static_cast<uint8_t>(referenceReadsMode));
		if (referenceReadsMode == ReferenceReadsMode::Sparse) { le::put(r, sparseMode_range); le::put_double(r, sparseMode_exponent); }
		r.push_back(static_cast<uint8_t>(ref_genome_available));
		if (ref_genome_available) {
			r.push_back(static_cast<uint8_t>(storeRefGenome));
			le::put(r, ref_genome_read_len); le::put(r, ref_genome_overlap_size); le::put(r, n_ref_genome_pseudo_reads);
			if (!storeRefGenome) {
				if (ref_genome_checksum.size() != 16) throw std::invalid_argument("colord_b200: the reference genome checksum has 16 bytes");
				r.insert(r.end(), ref_genome_checksum.begin(), ref_genome_checksum.end());
			}
		}
		return r;
	}
	void Deserialize(const std::vector<uint8_t>& data, bool archive_has_qual_stream)
	{
		le::Reader in{data.data(), data.data() + data.size()};
		tot_ref_reads = in.get<uint32_t>(); maxCandidates = in.get<uint32_t>(); compressionLevel = static_cast<int32_t>(in.get<uint32_t>());
		// enum bytes come from the file: anything outside params.h:33-46 is refused here, before a table is indexed with it
		auto enum_byte = [&](unsigned n_values, const char* what) { const uint8_t v = in.get<uint8_t>(); if (v >= n_values) throw std::runtime_error(std::string("colord_b200: bad ") + what + " in the meta record"); return v; };
		dataSource = static_cast<DataSource>(enum_byte(3, "data source"));
		approx_stream_size = in.get<uint64_t>();
		is_fastq = archive_has_qual_stream;
		qualityRevThresholds.clear();
		if (is_fastq) {
			qualityComprMode = static_cast<QualityComprMode>(enum_byte(9, "quality mode"));
			for (size_t i = n_thresholds(qualityComprMode); i; --i) qualityRevThresholds.push_back(in.get<uint32_t>());
		}
		headerComprMode = static_cast<HeaderComprMode>(enum_byte(3, "header mode"));
		referenceReadsMode = static_cast<ReferenceReadsMode>(enum_byte(2, "reference reads mode"));
		sparseMode_range = 0; sparseMode_exponent = 0;
		if (referenceReadsMode == ReferenceReadsMode::Sparse) { sparseMode_range = in.get<uint32_t>(); sparseMode_exponent = in.get_double(); }
		ref_genome_available = in.get<uint8_t>() != 0;
		storeRefGenome = false; ref_genome_read_len = ref_genome_overlap_size = n_ref_genome_pseudo_reads = 0; ref_genome_checksum.clear();
		if (ref_genome_available) {
			storeRefGenome = in.get<uint8_t>() != 0;
			ref_genome_read_len = in.get<uint32_t>(); ref_genome_overlap_size = in.get<uint32_t>(); n_ref_genome_pseudo_reads = in.get<uint32_t>();
			if (!storeRefGenome) for (int i = 0; i < 16; ++i) ref_genome_checksum.push_back(in.get<uint8_t>());
		}
		if (in.p != in.end) throw std::runtime_error("colord_b200: trailing bytes in the meta record");
	}
};

} // namespace clbhost
