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
// Construction differs.  The file is mapped (gzip: inflated) once and lines are found with memchr.  FASTQ is parsed by several
// threads: the input is cut at record starts (a line that begins with '@' whose second-next line begins with '+' — a quality
// line that begins with '@' is followed by a header and then by a read, which cannot begin with '+'), pass 1 sizes every
// piece, a prefix sum places it, pass 2 checks and copies into the final arrays — each thread first-touches its own part of
// them, which is what the reader's time goes to (page faults).  Whenever the threaded parse meets anything irregular it is
// abandoned and the serial parser, which follows the reference line by line, decides and words the refusal.
// Streaming form ("pinned-memory H2D double buffering", in_reads.cpp:229 — the reference's reader thread feeds
// 32 MiB blocks to its queues while the stages run): plain FASTQ files are cut into pieces at record starts, a pool of threads
// parses the pieces one pass each into reused buffers (a ring of slots: no page faults after the first lap, host memory bounded by
// the ring), and the constructing thread hands every piece, in file order, to a sink — compressor.h's sink is clb_append_reads +
// clb_append_quals, so parsing overlaps the host-to-device copies and stage 1a, and only the headers stay on the host.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

namespace clbhost {

struct InputError : std::runtime_error { using std::runtime_error::runtime_error; };
// the streaming parser met something irregular: the caller starts over with the whole-file reader, which follows the reference line
// by line and words the refusal (or accepts the input)
struct StreamingFallback : std::runtime_error { StreamingFallback() : std::runtime_error("irregular FASTQ: whole-file reader needed") {} };

// A block of T that is NOT value-initialised: the thread that fills a part of it is the one that touches its pages first.
template <typename T>
class Block {
	std::unique_ptr<T[]> p; size_t n = 0, cap = 0;
public:
	void allocate(size_t count) { p.reset(new T[count ? count : 1]); n = count; cap = count ? count : 1; }
	void reserve(size_t count) { if (count > cap) { std::unique_ptr<T[]> q(new T[count]); if (n) std::memcpy(q.get(), p.get(), n * sizeof(T)); p.swap(q); cap = count; } }
	void append(const T* s, size_t count) { if (n + count > cap) reserve(std::max(n + count, cap * 2)); if (count) std::memcpy(p.get() + n, s, count * sizeof(T)); n += count; }
	void push_back(const T& v) { append(&v, 1); }
	T* data() { return p.get(); }
	const T* data() const { return p.get(); }
	size_t size() const { return n; }
	bool empty() const { return n == 0; }
	T& operator[](size_t i) { return p[i]; }
	const T& operator[](size_t i) const { return p[i]; }
	const T* begin() const { return p.get(); }
	const T* end() const { return p.get() + n; }
};

class CInputReads {
public:
	bool is_fastq = false, is_gzip = false;
	Block<uint8_t> bases; Block<uint64_t> offsets;                  // reads back to back, offsets[n + 1]
	Block<uint8_t> quals;                                           // same layout as bases (FASTQ only)
	Block<uint8_t> headers; Block<uint64_t> header_offsets;         // headers without their first character, header_offsets[n + 1]
	Block<uint8_t> plus_id;                                         // 1: the '+' line repeats the header (qual_header_type::eq_read_header)
	Block<uint8_t> has_n;
	std::vector<uint32_t> read_pack_sizes, header_pack_sizes;       // reads per read pack / headers per header pack
	uint64_t total_bytes = 0, total_bases = 0, total_symb_header = 0, file_bytes = 0;
	unsigned threads_used = 1;
	const std::vector<uint32_t>& ReadLengths() const { return read_lens; }      // streaming form only
	bool streamed = false; uint64_t n_reads_streamed = 0;          // streaming form: bases / quals / offsets went to the sink, not into the blocks above

	uint32_t n_reads() const { return streamed ? static_cast<uint32_t>(n_reads_streamed) : static_cast<uint32_t>(offsets.size() - 1); }

	// ---- streaming form ----
	using PieceSink = std::function<void(const uint8_t* bases, const uint8_t* quals, const uint64_t* offsets, uint32_t n_reads)>;
	// piece buffers: plain memory, or whatever the caller's allocator hands out (compressor.h: page-locked memory for large inputs)
	struct HostAlloc { std::function<void*(uint64_t)> alloc; std::function<void(void*)> free; };
	// plain (not gzipped) FASTQ of at least min_bytes: what the streaming form takes; fills file_bytes / is_gzip either way
	static bool streamable(const std::string& path, uint64_t& file_bytes, bool& is_gzip, bool& is_fastq, uint64_t min_bytes = 64u << 20)
	{
		const int fd = ::open(path.c_str(), O_RDONLY);
		if (fd < 0) throw InputError("Error: cannot open file: " + path);
		struct stat st{};
		uint8_t head[2] = {0, 0};
		const bool ok = ::fstat(fd, &st) == 0;
		const ssize_t got = ::pread(fd, head, 2, 0);
		::close(fd);
		if (!ok) throw InputError("Error: cannot open file: " + path);
		file_bytes = static_cast<uint64_t>(st.st_size);
		is_gzip = got == 2 && head[0] == 0x1f && head[1] == 0x8b;
		is_fastq = got >= 1 && head[0] == '@';
		if (is_gzip) {      // the first byte of the inflated data decides
			gzFile gz = gzopen(path.c_str(), "rb");
			uint8_t first = 0;
			if (gz) { if (gzread(gz, &first, 1) == 1) is_fastq = first == '@'; gzclose(gz); }
		}
		if (const char* e = std::getenv("CLB_NO_STREAMING")) if (*e && *e != '0') return false;
		return !is_gzip && is_fastq && file_bytes >= min_bytes;
	}
	// range_begin / range_end: byte offsets into the file; the reader takes the records that START in [first record start at or after
	// range_begin, first record start at or after range_end) — adjacent ranges tile the file (multi-GPU: one range per rank)
	CInputReads(const std::string& path, const PieceSink& sink, unsigned n_threads = 0, uint64_t piece_bytes = 64u << 20, const HostAlloc* host_alloc = nullptr,
		uint64_t range_begin = 0, uint64_t range_end = ~0ull)
	{
		Input in(path, *this);
		if (!in.n || in.p[0] != '@' || is_gzip || (in.p[in.n - 1] != '\n' && in.p[in.n - 1] != '\r')) throw StreamingFallback();
		is_fastq = true; streamed = true;
		if (!n_threads) { const char* e = std::getenv("CLB_READER_THREADS"); n_threads = e ? static_cast<unsigned>(std::atoi(e)) : std::thread::hardware_concurrency(); }
		const uint8_t* end = in.p + in.n;
		const uint8_t* rb = range_begin == 0 ? in.p : range_begin >= in.n ? end : record_start(in.p + range_begin, end);
		const uint8_t* re = range_end >= in.n ? end : record_start(in.p + range_end, end);
		if (!rb || !re) throw StreamingFallback();
		if (re < rb) re = rb;
		total_bytes = static_cast<uint64_t>(re - rb);
		if (re > rb) stream_fastq(rb, total_bytes, std::max(1u, n_threads), std::max<uint64_t>(piece_bytes, 1u << 20), sink, host_alloc);
		else header_offsets.push_back(0);
		pack_sizes_from(read_lens);
	}

	// n_threads 0: CLB_READER_THREADS or the hardware's count; at most one thread per min_piece_bytes of input
	explicit CInputReads(const std::string& path, unsigned n_threads = 0, uint64_t min_piece_bytes = 16u << 20)
	{
		Input in(path, *this);
		total_bytes = in.n;
		if (!in.n) throw InputError("Error: file " + path + " is empty");
		if (in.p[0] != '@' && in.p[0] != '>') throw InputError("Error: unknown file format");
		is_fastq = in.p[0] == '@';
		if (!n_threads) { const char* e = std::getenv("CLB_READER_THREADS"); n_threads = e ? static_cast<unsigned>(std::atoi(e)) : std::thread::hardware_concurrency(); }
		n_threads = static_cast<unsigned>(std::min<uint64_t>(std::max(1u, n_threads), in.n / std::max<uint64_t>(1, min_piece_bytes) + 1));
		if (!is_fastq) parse_fasta(in.p, in.n);
		else if (n_threads < 2 || !parse_fastq_threads(in.p, in.n, n_threads)) { reset(); parse_fastq(in.p, in.n); }
		pack_sizes();
	}

private:
	// The file's bytes: plain files are mapped (no copy, no zero-filled buffer), gzipped files are inflated into memory.
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
					map = ::mmap(nullptr, r.file_bytes, PROT_READ, MAP_PRIVATE, fd, 0);
					if (map == MAP_FAILED) { map = nullptr; ::close(fd); throw InputError("Error: cannot read file: " + path); }
					map_n = r.file_bytes;
					::madvise(map, map_n, MADV_WILLNEED);
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

	void reset()
	{
		bases = Block<uint8_t>(); quals = Block<uint8_t>(); headers = Block<uint8_t>(); plus_id = Block<uint8_t>(); has_n = Block<uint8_t>();
		offsets = Block<uint64_t>(); header_offsets = Block<uint64_t>();
		total_bases = total_symb_header = 0; threads_used = 1;
	}
	static const uint8_t* find_eol(const uint8_t* p, const uint8_t* end)
	{
		const uint8_t* e = static_cast<const uint8_t*>(std::memchr(p, '\n', end - p));
		if (!e) e = end;
		const uint8_t* r = static_cast<const uint8_t*>(std::memchr(p, '\r', e - p));
		return r ? r : e;
	}
	// next non-empty line at or after p: [b, e); false at the end of the data.  A last line without an end-of-line has e == end.
	static bool next_line(const uint8_t*& p, const uint8_t* end, const uint8_t*& b, const uint8_t*& e)
	{
		while (p < end) {
			const uint8_t* q = find_eol(p, end);
			if (q == p) { ++p; continue; }
			b = p; e = q; p = q < end ? q + 1 : end;
			return true;
		}
		return false;
	}
	// branch-free so that the compiler vectorises it; returns 0 ok, 1 ok with N, 2 a symbol outside ACGTN
	static uint8_t classify(const uint8_t* s, size_t n)
	{
		uint8_t any_n = 0, bad = 0;
		for (size_t i = 0; i < n; ++i) {
			const uint8_t c = s[i];
			const uint8_t is_n = c == 'N';
			bad |= static_cast<uint8_t>(!((c == 'A') | (c == 'C') | (c == 'G') | (c == 'T') | is_n));
			any_n |= is_n;
		}
		return bad ? 2 : any_n;
	}
	void pack_sizes()
	{
		const uint32_t n = n_reads();
		uint64_t cur = 0; uint32_t k = 0;
		for (uint32_t i = 0; i < n; ++i) { cur += offsets[i + 1] - offsets[i] + 1; ++k; if (cur >= (2u << 21)) { read_pack_sizes.push_back(k); k = 0; cur = 0; } }
		if (k) read_pack_sizes.push_back(k);
		cur = 0; k = 0;
		for (uint32_t i = 0; i + 1 < header_offsets.size(); ++i) { cur += header_offsets[i + 1] - header_offsets[i]; ++k; if (cur >= (2u << 21)) { header_pack_sizes.push_back(k); k = 0; cur = 0; } }
		if (k) header_pack_sizes.push_back(k);
	}

	// ---- streaming FASTQ parser ----
	std::vector<uint32_t> read_lens;                                 // streaming form: the read lengths (the pack rule and the multi-GPU exchange need them)
	void pack_sizes_from(const std::vector<uint32_t>& lens)
	{
		uint64_t cur = 0; uint32_t k = 0;
		for (uint32_t len : lens) { cur += static_cast<uint64_t>(len) + 1; ++k; if (cur >= (2u << 21)) { read_pack_sizes.push_back(k); k = 0; cur = 0; } }
		if (k) read_pack_sizes.push_back(k);
		cur = 0; k = 0;
		for (uint32_t i = 0; i + 1 < header_offsets.size(); ++i) { cur += header_offsets[i + 1] - header_offsets[i]; ++k; if (cur >= (2u << 21)) { header_pack_sizes.push_back(k); k = 0; cur = 0; } }
		if (k) header_pack_sizes.push_back(k);
	}
	struct Buf {
		uint8_t* p = nullptr; const HostAlloc* A = nullptr; bool own_new = false;
		uint8_t* get() const { return p; }
		void reset(uint64_t bytes, const HostAlloc* a) { drop(); A = a; p = a && a->alloc ? static_cast<uint8_t*>(a->alloc(bytes)) : nullptr; if (!p) { p = new uint8_t[bytes]; own_new = true; } else own_new = false; }
		void drop() { if (!p) return; if (own_new) delete[] p; else A->free(p); p = nullptr; }
		~Buf() { drop(); }
		Buf() = default; Buf(const Buf&) = delete; Buf& operator=(const Buf&) = delete;
	};
	struct Slot {
		Buf bases, quals; uint64_t cap = 0;
		std::vector<uint64_t> offsets; std::vector<uint8_t> hdr, plus; std::vector<uint64_t> hdr_off;
		uint64_t n_bases = 0, symb = 0; bool ok = true;
		uint64_t holds = ~0ull, free_for = 0;                          // piece parsed into the slot / piece that may be parsed into it next
	};
	// one pass over a piece: checks of size_piece + fill_piece, output into the slot
	static void parse_piece(const uint8_t* pb, const uint8_t* pe, Slot& S, const HostAlloc* A)
	{
		const uint64_t need = static_cast<uint64_t>(pe - pb) / 2 + 64;
		if (need > S.cap) { S.cap = need + need / 8; S.bases.reset(S.cap, A); S.quals.reset(S.cap, A); }
		S.offsets.clear(); S.offsets.push_back(0); S.hdr.clear(); S.plus.clear(); S.hdr_off.clear(); S.hdr_off.push_back(0);
		S.n_bases = 0; S.symb = 0; S.ok = true;
		const uint8_t* p = pb; const uint8_t *b, *e; int where = 0;
		const uint8_t* hdr = nullptr; size_t hdr_n = 0, read_n = 0; uint64_t at = 0;
		while (next_line(p, pe, b, e)) {
			const size_t n = static_cast<size_t>(e - b);
			switch (where) {
			case 0: if (*b != '@') { S.ok = false; return; } S.symb += n; hdr = b + 1; hdr_n = n - 1; break;
			case 1:
				if (at + n > S.cap || classify(b, n) == 2) { S.ok = false; return; }
				std::memcpy(S.bases.get() + at, b, n); read_n = n;
				break;
			case 2: {
				if (*b != '+') { S.ok = false; return; }
				S.symb += n;
				const bool plus = n > 1;
				if (plus && (n - 1 != hdr_n || std::memcmp(b + 1, hdr, hdr_n) != 0)) { S.ok = false; return; }
				S.hdr.insert(S.hdr.end(), hdr, hdr + hdr_n); S.hdr_off.push_back(S.hdr.size()); S.plus.push_back(plus ? 1 : 0);
				break;
			}
			default:
				if (n != read_n) { S.ok = false; return; }
				std::memcpy(S.quals.get() + at, b, n); at += n; S.offsets.push_back(at);
				break;
			}
			where = (where + 1) & 3;
		}
		if (where != 0) S.ok = false;
		S.n_bases = at;
	}
	void stream_fastq(const uint8_t* data, uint64_t size, unsigned T, uint64_t piece, const PieceSink& sink, const HostAlloc* A)
	{
		const uint8_t* end = data + size;
		const uint64_t n_pieces = (size + piece - 1) / piece;
		T = static_cast<unsigned>(std::min<uint64_t>(T, n_pieces));
		// the ring: enough slots for every thread + the consumer, and up to 64 pieces (4 GiB) of head start while the consumer is not ready yet
		const uint64_t W = std::min<uint64_t>(n_pieces, std::max<uint64_t>(static_cast<uint64_t>(T) + 2, 64));
		std::vector<Slot> slots(W);      // declared before the threads: destroyed after they are joined
		for (uint64_t i = 0; i < W; ++i) slots[i].free_for = i;
		std::mutex m; std::condition_variable cv; std::atomic<uint64_t> next_piece{0}; bool stop = false;
		auto start_of = [&](uint64_t i) -> const uint8_t* { return i == 0 ? data : i >= n_pieces ? end : record_start(data + i * piece, end); };      // a piece ends at the range's end at the latest: record_start never looks past `end`
		auto worker = [&] {
			for (;;) {
				const uint64_t i = next_piece.fetch_add(1);
				if (i >= n_pieces) return;
				Slot& S = slots[i % W];
				{ std::unique_lock<std::mutex> lk(m); cv.wait(lk, [&] { return stop || S.free_for == i; }); if (stop) return; }
				const uint8_t* pb = start_of(i); const uint8_t* pe = start_of(i + 1);
				if (!pb || !pe) { S.ok = false; S.offsets.assign(1, 0); S.n_bases = 0; }
				else if (pe <= pb) { S.ok = true; S.offsets.assign(1, 0); S.hdr.clear(); S.plus.clear(); S.hdr_off.assign(1, 0); S.n_bases = 0; S.symb = 0; }      // a record longer than a piece
				else parse_piece(pb, pe, S, A);
				{ std::lock_guard<std::mutex> lk(m); S.holds = i; }
				cv.notify_all();
			}
		};
		std::vector<std::thread> th;
		for (unsigned t = 0; t < T; ++t) th.emplace_back(worker);
		auto shut = [&] { { std::lock_guard<std::mutex> lk(m); stop = true; } cv.notify_all(); for (auto& t : th) t.join(); };
		try {
			header_offsets.push_back(0);
			for (uint64_t i = 0; i < n_pieces; ++i) {
				Slot& S = slots[i % W];
				{ std::unique_lock<std::mutex> lk(m); cv.wait(lk, [&] { return S.holds == i; }); }
				if (!S.ok) throw StreamingFallback();
				const uint32_t n = static_cast<uint32_t>(S.offsets.size() - 1);
				if (n_reads_streamed + n >= (1ull << 32)) throw StreamingFallback();
				if (n) sink(S.bases.get(), S.quals.get(), S.offsets.data(), n);
				const uint64_t h0 = headers.size();
				headers.append(S.hdr.data(), S.hdr.size());
				for (uint32_t r = 0; r < n; ++r) { header_offsets.push_back(h0 + S.hdr_off[r + 1]); read_lens.push_back(static_cast<uint32_t>(S.offsets[r + 1] - S.offsets[r])); }
				plus_id.append(S.plus.data(), S.plus.size());
				n_reads_streamed += n; total_bases += S.n_bases; total_symb_header += S.symb;
				{ std::lock_guard<std::mutex> lk(m); S.free_for = i + W; }
				cv.notify_all();
			}
		} catch (...) { shut(); throw; }
		shut();
		threads_used = T;
	}

	// ---- serial parsers: the reference's state machines ----
	void add_header(const uint8_t* s, size_t n, bool plus)
	{
		if (header_offsets.empty()) header_offsets.push_back(0);
		headers.append(s, n);
		header_offsets.push_back(headers.size());
		plus_id.push_back(plus ? 1 : 0);
	}
	void add_read(const uint8_t* s, size_t n)
	{
		const uint8_t cls = classify(s, n);
		if (cls == 2) throw InputError("Only ACGTN symbols supported inside a read");
		if (offsets.empty()) offsets.push_back(0);
		bases.append(s, n);
		offsets.push_back(bases.size());
		has_n.push_back(cls);
		total_bases += n;
	}
	void parse_fastq(const uint8_t* p, uint64_t size)
	{
		const uint8_t* end = p + size;
		bases.reserve(size / 2 + 16); quals.reserve(size / 2 + 16);
		int where = 0;                                     // 0 header, 1 read, 2 '+' line, 3 quality
		const uint8_t* hdr = nullptr; size_t hdr_n = 0;
		const uint8_t *b, *e;
		while (next_line(p, end, b, e)) {
			if (e == end) throw InputError("Error: something went wrong during input reading");     // no end-of-line after the last line
			const size_t n = static_cast<size_t>(e - b);
			switch (where) {
			case 0: total_symb_header += n; hdr = b + 1; hdr_n = n - 1; break;
			case 1: add_read(b, n); break;
			case 2: {
				total_symb_header += n;
				const bool plus = n > 1;
				if (plus && (n - 1 != hdr_n || std::memcmp(b + 1, hdr, hdr_n) != 0)) throw InputError("Error: quality header not empty but different than read header");
				add_header(hdr, hdr_n, plus);
				break;
			}
			default: quals.append(b, n); break;
			}
			where = (where + 1) & 3;
		}
		if (quals.size() != bases.size() || where != 0) throw InputError("Error: something went wrong during input reading");
		if (offsets.empty()) offsets.push_back(0);
		if (header_offsets.empty()) header_offsets.push_back(0);
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

	// ---- threaded FASTQ parser ----
	struct Piece { const uint8_t *b, *e; uint64_t n_reads = 0, n_bases = 0, n_quals = 0, n_hdr = 0, symb = 0; bool ok = true; };
	// first record start after `from`: a line that begins with '@' whose second-next (non-empty) line begins with '+'
	static const uint8_t* record_start(const uint8_t* from, const uint8_t* end)
	{
		const uint8_t* p = find_eol(from, end);                        // skip the line `from` falls into
		if (p < end) ++p;
		const uint8_t *b, *e;
		for (int tries = 0; tries < 16; ++tries) {
			if (!next_line(p, end, b, e)) return end;                    // no record starts between `from` and the end of the data
			if (*b != '@') continue;
			const uint8_t* q = p; const uint8_t *b1, *e1, *b2, *e2;
			if (next_line(q, end, b1, e1) && next_line(q, end, b2, e2) && *b2 == '+') return b;
		}
		return nullptr;
	}
	static void size_piece(Piece& P)
	{
		const uint8_t* p = P.b; const uint8_t *b, *e; int where = 0; size_t read_n = 0;
		while (next_line(p, P.e, b, e)) {
			const size_t n = static_cast<size_t>(e - b);
			switch (where) {
			case 0: if (*b != '@') { P.ok = false; return; } P.symb += n; P.n_hdr += n - 1; break;
			case 1: P.n_bases += n; read_n = n; ++P.n_reads; break;
			case 2: if (*b != '+') { P.ok = false; return; } P.symb += n; break;
			default: if (n != read_n) { P.ok = false; return; } P.n_quals += n; break;
			}
			where = (where + 1) & 3;
		}
		if (where != 0) P.ok = false;
	}
	void fill_piece(Piece& P, uint64_t r0, uint64_t b0, uint64_t h0)
	{
		const uint8_t* p = P.b; const uint8_t *b, *e; int where = 0;
		const uint8_t* hdr = nullptr; size_t hdr_n = 0;
		uint64_t r = r0, at = b0, hat = h0;
		while (next_line(p, P.e, b, e)) {
			const size_t n = static_cast<size_t>(e - b);
			switch (where) {
			case 0: hdr = b + 1; hdr_n = n - 1; break;
			case 1: {
				const uint8_t cls = classify(b, n);
				if (cls == 2) { P.ok = false; return; }
				std::memcpy(bases.data() + at, b, n); has_n[r] = cls; offsets[r + 1] = at + n;
				break;
			}
			case 2: {
				const bool plus = n > 1;
				if (plus && (n - 1 != hdr_n || std::memcmp(b + 1, hdr, hdr_n) != 0)) { P.ok = false; return; }
				std::memcpy(headers.data() + hat, hdr, hdr_n); hat += hdr_n; header_offsets[r + 1] = hat; plus_id[r] = plus ? 1 : 0;
				break;
			}
			default: std::memcpy(quals.data() + at, b, n); at += n; ++r; break;
			}
			where = (where + 1) & 3;
		}
	}
	// false: something irregular — the caller falls back to the serial parser
	bool parse_fastq_threads(const uint8_t* data, uint64_t size, unsigned T)
	{
		const uint8_t* end = data + size;
		if (size == 0 || (end[-1] != '\n' && end[-1] != '\r')) return false;
		std::vector<Piece> pieces;
		const uint8_t* at = data;
		for (unsigned i = 1; i <= T && at < end; ++i) {
			const uint8_t* nxt = i == T ? end : record_start(data + size / T * i, end);
			if (!nxt) return false;
			if (nxt <= at) continue;
			Piece P; P.b = at; P.e = nxt; pieces.push_back(P);
			at = nxt;
		}
		auto run = [&](auto&& fn) {
			std::vector<std::thread> th;
			for (size_t i = 1; i < pieces.size(); ++i) th.emplace_back([&fn, i] { fn(i); });
			fn(0);
			for (auto& t : th) t.join();
		};
		run([&](size_t i) { size_piece(pieces[i]); });
		uint64_t nr = 0, nb = 0, nh = 0;
		std::vector<uint64_t> r0(pieces.size()), b0(pieces.size()), h0(pieces.size());
		for (size_t i = 0; i < pieces.size(); ++i) {
			if (!pieces[i].ok || pieces[i].n_quals != pieces[i].n_bases) return false;
			r0[i] = nr; b0[i] = nb; h0[i] = nh;
			nr += pieces[i].n_reads; nb += pieces[i].n_bases; nh += pieces[i].n_hdr; total_symb_header += pieces[i].symb;
		}
		if (nr >= (1ull << 32)) return false;
		bases.allocate(nb); quals.allocate(nb); headers.allocate(nh);
		offsets.allocate(nr + 1); header_offsets.allocate(nr + 1); plus_id.allocate(nr); has_n.allocate(nr);
		offsets[0] = 0; header_offsets[0] = 0;
		run([&](size_t i) { fill_piece(pieces[i], r0[i], b0[i], h0[i]); });
		for (const Piece& P : pieces) if (!P.ok) return false;
		total_bases = nb; threads_used = static_cast<unsigned>(pieces.size());
		return true;
	}
};

} // namespace clbhost
