// fastq_reader.h — host-side input reader: FASTQ / FASTA / multi-line FASTA, plain or gzipped.
// Mirror of CInputReads (src/colord/in_reads.cpp:24-297) with the same accepted inputs, the same statistics and the same
// refusals, producing the layout the C-ABI takes (include/colord_b200.h: ASCII bases back to back + offsets) instead of
// per-read vectors pushed through queues:
//   * a line ends at '\n' or '\r'; empty lines are skipped, so CRLF files and blank lines parse (in_reads.cpp:189-193)
//   * FASTQ: lines cycle header / read / '+' line / quality (:181-221); the '+' line is empty or repeats the header, anything
//     else is refused (:86-92); a last line without an end-of-line is refused (:278-282)
//   * FASTA: header line, then the read over one or more lines until a line that starts with '>' (:117-176); the last read is
//     closed at the end of the file
//   * only A C G T N inside reads (:31-35, utils.h:469-481) — checked here so that the message is the reference's
//   * read packs close when their reads hold >= 4 MiB counting one guard byte per read, header packs when their headers
//     hold >= 4 MiB (:62-76, :43-48, :95-101; defs.h:45-46)
//   * total_bytes = bytes delivered by the (de)compressor, total_bases, total_symb_header = header + '+' line bytes (:49, :81)
// Construction differs: the file is mapped (gzip: inflated) once and lines are found with memchr, not pushed byte by byte.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

namespace clbhost {

struct InputError : std::runtime_error { using std::runtime_error::runtime_error; };

class CInputReads {
public:
	bool is_fastq = false, is_gzip = false;
	std::vector<uint8_t> bases; std::vector<uint64_t> offsets{0};          // reads back to back, offsets[n + 1]
	std::vector<uint8_t> quals;                                           // same layout as bases (FASTQ only)
	std::vector<uint8_t> headers; std::vector<uint64_t> header_offsets{0}; // headers without their first character
	std::vector<uint8_t> plus_id;                                         // 1: the '+' line repeats the header (qual_header_type::eq_read_header)
	std::vector<uint8_t> has_n;
	std::vector<uint32_t> read_pack_sizes, header_pack_sizes;             // reads per read pack / headers per header pack
	uint64_t total_bytes = 0, total_bases = 0, total_symb_header = 0, file_bytes = 0;

	uint32_t n_reads() const { return static_cast<uint32_t>(offsets.size() - 1); }

	explicit CInputReads(const std::string& path)
	{
		Input in(path, *this);
		total_bytes = in.n;
		if (!in.n) throw InputError("Error: file " + path + " is empty");
		if (in.p[0] != '@' && in.p[0] != '>') throw InputError("Error: unknown file format");
		is_fastq = in.p[0] == '@';
		if (is_fastq) parse_fastq(in.p, in.n); else parse_fasta(in.p, in.n);
		if (cur_reads) read_pack_sizes.push_back(cur_reads);
		if (cur_headers) header_pack_sizes.push_back(cur_headers);
	}

private:
	uint64_t cur_read_bytes = 0, cur_header_bytes = 0; uint32_t cur_reads = 0, cur_headers = 0;

	// The file's bytes: plain files are mapped (no copy, no zero-filled buffer: on a 400 MB file the allocation alone cost more
	// than the parse), gzipped files are inflated into memory.
	struct Input {
		const uint8_t* p = nullptr; uint64_t n = 0;
		void* map = nullptr; uint64_t map_n = 0; std::vector<uint8_t> buf;
		Input(const std::string& path, CInputReads& r)
		{
			const int fd = ::open(path.c_str(), O_RDONLY);
			if (fd < 0) throw InputError("Error: cannot open file: " + path);
			struct stat st{};
			if (::fstat(fd, &st) != 0) { ::close(fd); throw InputError("Error: cannot open file: " + path); }
			r.file_bytes = static_cast<uint64_t>(st.st_size);
			uint8_t magic[2] = {0, 0};
			const ssize_t got = ::pread(fd, magic, 2, 0);
			r.is_gzip = got == 2 && magic[0] == 0x1f && magic[1] == 0x8b;              // utils.cpp izGzipFile
			if (!r.is_gzip) {
				if (r.file_bytes) {
					map = ::mmap(nullptr, r.file_bytes, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
					if (map == MAP_FAILED) { map = nullptr; ::close(fd); throw InputError("Error: cannot read file: " + path); }
					map_n = r.file_bytes;
					::madvise(map, map_n, MADV_SEQUENTIAL);
				}
				::close(fd);
				p = static_cast<const uint8_t*>(map); n = r.file_bytes;
				return;
			}
			::close(fd);
			gzFile gz = gzopen(path.c_str(), "rb");
			if (!gz) throw InputError("Error: cannot open file: " + path);
			gzbuffer(gz, 1u << 20);
			const size_t chunk = 1u << 25;
			for (;;) {
				const size_t at = buf.size();
				buf.resize(at + chunk);
				const int k = gzread(gz, buf.data() + at, static_cast<unsigned>(chunk));
				if (k < 0) { int code; const char* msg = gzerror(gz, &code); std::string m = std::string("zblib error: ") + msg; gzclose(gz); throw InputError(m); }
				buf.resize(at + static_cast<size_t>(k));
				if (static_cast<size_t>(k) < chunk) break;
			}
			gzclose(gz);
			p = buf.data(); n = buf.size();
		}
		~Input() { if (map) ::munmap(map, map_n); }
		Input(const Input&) = delete;
		Input& operator=(const Input&) = delete;
	};
	static const uint8_t* find_eol(const uint8_t* p, const uint8_t* end)
	{
		const uint8_t* e = static_cast<const uint8_t*>(std::memchr(p, '\n', end - p));
		if (!e) e = end;
		const uint8_t* r = static_cast<const uint8_t*>(std::memchr(p, '\r', e - p));
		return r ? r : e;
	}
	void add_header(const uint8_t* s, size_t n, bool plus)
	{
		headers.insert(headers.end(), s, s + n);
		header_offsets.push_back(headers.size());
		plus_id.push_back(plus ? 1 : 0);
		cur_header_bytes += n; ++cur_headers;
		if (cur_header_bytes >= (2u << 21)) { header_pack_sizes.push_back(cur_headers); cur_headers = 0; cur_header_bytes = 0; }
	}
	void add_read(const uint8_t* s, size_t n)
	{
		// branch-free so that the compiler vectorises it: the check must not be what limits the reader
		uint8_t any_n = 0, bad = 0;
		for (size_t i = 0; i < n; ++i) {
			const uint8_t c = s[i];
			const uint8_t is_n = c == 'N';
			bad |= static_cast<uint8_t>(!((c == 'A') | (c == 'C') | (c == 'G') | (c == 'T') | is_n));
			any_n |= is_n;
		}
		if (bad) throw InputError("Only ACGTN symbols supported inside a read");
		bases.insert(bases.end(), s, s + n);
		offsets.push_back(bases.size());
		has_n.push_back(any_n);
		total_bases += n;
		cur_read_bytes += n + 1; ++cur_reads;
		if (cur_read_bytes >= (2u << 21)) { read_pack_sizes.push_back(cur_reads); cur_reads = 0; cur_read_bytes = 0; }
	}
	void parse_fastq(const uint8_t* p, uint64_t size)
	{
		const uint8_t* end = p + size;
		bases.reserve(size / 2 + 16); quals.reserve(size / 2 + 16);
		int where = 0;                                     // 0 header, 1 read, 2 '+' line, 3 quality
		const uint8_t* hdr = nullptr; size_t hdr_n = 0;
		while (p < end) {
			const uint8_t* e = find_eol(p, end);
			if (e == p) { ++p; continue; }                    // empty line / second byte of a CRLF
			if (e == end) throw InputError("Error: something went wrong during input reading");     // no end-of-line after the last line
			const size_t n = static_cast<size_t>(e - p);
			switch (where) {
			case 0: total_symb_header += n; hdr = p + 1; hdr_n = n - 1; break;
			case 1: add_read(p, n); break;
			case 2: {
				total_symb_header += n;
				const bool plus = n > 1;
				if (plus && (n - 1 != hdr_n || std::memcmp(p + 1, hdr, hdr_n) != 0)) throw InputError("Error: quality header not empty but different than read header");
				add_header(hdr, hdr_n, plus);
				break;
			}
			default: quals.insert(quals.end(), p, p + n); break;
			}
			where = (where + 1) & 3;
			p = e + 1;
		}
		if (quals.size() != bases.size() || where != 0) throw InputError("Error: something went wrong during input reading");
	}
	void parse_fasta(const uint8_t* p, uint64_t size)
	{
		const uint8_t* end = p + size;
		std::vector<uint8_t> read;
		bool in_read = false;
		while (p < end) {
			const uint8_t* e = find_eol(p, end);
			if (e == p) { ++p; continue; }
			const size_t n = static_cast<size_t>(e - p);
			if (!in_read) {                                   // header line ('>' checked for the first one by the caller, later ones below)
				total_symb_header += n;
				add_header(p + 1, n - 1, false);
				in_read = true; read.clear();
				// the first line after the header belongs to the read whatever it starts with (in_reads.cpp:134-140)
				p = e < end ? e + 1 : end;
				while (p < end && (*p == '\n' || *p == '\r')) ++p;
				if (p < end) { const uint8_t* e2 = find_eol(p, end); read.insert(read.end(), p, e2); p = e2 < end ? e2 + 1 : end; }
				continue;
			}
			if (*p == '>') { add_read(read.data(), read.size()); in_read = false; continue; }      // re-read this line as a header
			read.insert(read.end(), p, e);
			p = e < end ? e + 1 : end;
		}
		add_read(read.data(), read.size());                   // in_reads.cpp:173-175: the last read closes at the end of the file
	}
};

} // namespace clbhost
