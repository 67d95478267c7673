// ref_genome_host.h — reference-genome mode (-G [-s]) on the host side (BASELINE config 5).
// Mirror of CReferenceGenome (src/colord/reference_genome.h:27-89, reference_genome.cpp):
//   :106-233  multi-FASTA (plain or gzipped) -> sequences; header lines skipped, symbols upper-cased, everything but A C G T dropped
//             (addSymb, reference_genome.h:49-54); every sequence packed 4 bases per byte + "symbols in the last byte" (:28-66);
//             MD5 over the packed sequences in order (:205-218) when the genome is not stored in the archive
//   :372-417  pseudo-reads: every sequence cut into reads of read_len = 20 x mean read length that overlap by 10 (k - 1)
//             symbols (compression.cpp:406, :450); they become the first reference reads (reads_sim_graph.cpp:295-322)
//   :281-317  the sequences as a second KMC input: their k-mers are counted with the reads' (compression.cpp:408-430)
//   :319-360  Store(archive): stream "ref-genome", one part, metadata = number of sequences, coded as plain reads by a DNA coder of its
//             own ("level 9") — on the device: clb_xplain_encode; :235-279 the way back
// The MD5 is RFC 1321 written out (the reference wraps a vendored implementation, md5_wrapper.h).  Header-only, no device code.
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>
#include <zlib.h>

namespace clbhost {

class MD5 {
	uint32_t a = 0x67452301u, b = 0xefcdab89u, c = 0x98badcfeu, d = 0x10325476u;
	uint64_t n = 0; uint8_t buf[64]; size_t fill = 0;
	static uint32_t rol(uint32_t x, int s) { return (x << s) | (x >> (32 - s)); }
	void block(const uint8_t* p)
	{
		static const uint32_t K[64] = {
			0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501, 0x698098d8, 0x8b44f7af, 0xffff5bb1, 0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821,
			0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453, 0xd8a1e681, 0xe7d3fbc8, 0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a,
			0xfffa3942, 0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70, 0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05, 0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665,
			0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d, 0x85845dd1, 0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};
		static const int S[64] = {7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 5, 9, 14, 20, 5, 9, 14, 20, 5, 9, 14, 20, 5, 9, 14, 20,
			4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21};
		uint32_t M[16];
		for (int i = 0; i < 16; ++i) M[i] = static_cast<uint32_t>(p[4 * i]) | (static_cast<uint32_t>(p[4 * i + 1]) << 8) | (static_cast<uint32_t>(p[4 * i + 2]) << 16) | (static_cast<uint32_t>(p[4 * i + 3]) << 24);
		uint32_t A = a, B = b, C = c, D = d;
		for (int i = 0; i < 64; ++i) {
			uint32_t F; int g;
			if (i < 16) { F = (B & C) | (~B & D); g = i; }
			else if (i < 32) { F = (D & B) | (~D & C); g = (5 * i + 1) & 15; }
			else if (i < 48) { F = B ^ C ^ D; g = (3 * i + 5) & 15; }
			else { F = C ^ (B | ~D); g = (7 * i) & 15; }
			F += A + K[i] + M[g];
			A = D; D = C; C = B; B += rol(F, S[i]);
		}
		a += A; b += B; c += C; d += D;
	}
public:
	void Update(const uint8_t* p, size_t len)
	{
		n += len;
		while (len) {
			const size_t k = std::min(len, 64 - fill);
			std::memcpy(buf + fill, p, k); fill += k; p += k; len -= k;
			if (fill == 64) { block(buf); fill = 0; }
		}
	}
	std::vector<uint8_t> Get()
	{
		const uint64_t bits = n * 8;
		const uint8_t one = 0x80, zero = 0;
		Update(&one, 1);
		while (fill != 56) Update(&zero, 1);
		uint8_t len[8];
		for (int i = 0; i < 8; ++i) len[i] = static_cast<uint8_t>(bits >> (8 * i));
		Update(len, 8);
		std::vector<uint8_t> out(16);
		const uint32_t w[4] = {a, b, c, d};
		for (int i = 0; i < 16; ++i) out[i] = static_cast<uint8_t>(w[i / 4] >> (8 * (i % 4)));
		return out;
	}
};

class CReferenceGenome {
	std::vector<std::vector<uint8_t>> sequences;      // ASCII A C G T
	uint64_t tot_seqs_len = 0;
	uint32_t overlap_size = 0, read_len = 0;
public:
	// packed form of one sequence (reference_genome.cpp:28-66)
	static std::vector<uint8_t> Pack(const std::vector<uint8_t>& seq)
	{
		auto code = [](uint8_t ch) -> uint8_t { return ch == 'C' ? 1 : ch == 'G' ? 2 : ch == 'T' ? 3 : 0; };
		std::vector<uint8_t> packed((seq.size() + 3) / 4 + 1, 0);
		for (size_t i = 0; i < seq.size(); ++i) packed[i / 4] = static_cast<uint8_t>(packed[i / 4] | (code(seq[i]) << (6 - 2 * (i % 4))));
		packed.back() = static_cast<uint8_t>(seq.size() % 4);
		return packed;
	}
	CReferenceGenome() = default;
	explicit CReferenceGenome(const std::string& path)
	{
		gzFile gz = gzopen(path.c_str(), "rb");
		if (!gz) throw std::runtime_error("Error: cannot open file: " + path);
		gzbuffer(gz, 1u << 20);
		std::vector<uint8_t> buff(1u << 25);
		int got = gzread(gz, buff.data(), static_cast<unsigned>(buff.size()));
		if (got <= 0) { gzclose(gz); throw std::runtime_error(got < 0 ? "zblib error while reading " + path : "Error: file " + path + " is empty"); }
		if (buff[0] != '>') { gzclose(gz); throw std::runtime_error("Error: wrong reference genome file format, multi fasta expected"); }
		enum { header, seq, eol_header, eol_seq } state = header;
		sequences.emplace_back();
		auto add = [&](uint8_t ch) { if (ch >= 'a' && ch <= 'z') ch = static_cast<uint8_t>(ch - 32); if (ch == 'A' || ch == 'C' || ch == 'G' || ch == 'T') sequences.back().push_back(ch); };
		while (got > 0) {
			for (int i = 0; i < got; ++i) {
				const uint8_t ch = buff[i]; const bool eol = ch == '\n' || ch == '\r';
				switch (state) {
				case header: if (eol) state = eol_header; break;
				case seq: if (eol) state = eol_seq; else add(ch); break;
				case eol_seq: if (eol) break; if (ch == '>') { state = header; sequences.emplace_back(); } else { state = seq; add(ch); } break;
				case eol_header: if (eol) break; state = seq; add(ch); break;
				}
			}
			got = gzread(gz, buff.data(), static_cast<unsigned>(buff.size()));
		}
		gzclose(gz);
		for (const auto& s : sequences) tot_seqs_len += s.size();
	}
	// from decoded sequences (symbols as ASCII), e.g. the archive's "ref-genome" stream
	explicit CReferenceGenome(std::vector<std::vector<uint8_t>> seqs) : sequences(std::move(seqs)) { for (const auto& s : sequences) tot_seqs_len += s.size(); }

	std::vector<uint8_t> GetChecksum() const { MD5 md5; for (const auto& s : sequences) { const std::vector<uint8_t> p = Pack(s); md5.Update(p.data(), p.size()); } return md5.Get(); }
	uint64_t GetTotSeqsLen() const { return tot_seqs_len; }
	uint32_t GetTotNSeqs() const { return static_cast<uint32_t>(sequences.size()); }
	void SetReadLen(uint32_t len, uint32_t overlap) { read_len = len; overlap_size = overlap; }
	bool valid_read_len() const { return read_len > overlap_size; }
	uint32_t GetNPseudoReads() const
	{
		uint64_t res = 0;
		for (const auto& s : sequences) res += (s.size() + (read_len - overlap_size) - 1) / (read_len - overlap_size);
		return static_cast<uint32_t>(res);
	}
	// all sequences back to back + offsets: the second counting input / the input of clb_xplain_encode
	void Sequences(std::vector<uint8_t>& bases, std::vector<uint64_t>& offsets) const
	{
		bases.clear(); offsets.assign(1, 0);
		bases.reserve(tot_seqs_len);
		for (const auto& s : sequences) { bases.insert(bases.end(), s.begin(), s.end()); offsets.push_back(bases.size()); }
	}
	// the pseudo-reads back to back + offsets (reference_genome.cpp:386-417)
	void PseudoReads(std::vector<uint8_t>& bases, std::vector<uint64_t>& offsets) const
	{
		bases.clear(); offsets.assign(1, 0);
		for (const auto& s : sequences)
			for (uint64_t start = 0; start < s.size(); start += read_len - overlap_size) {
				const uint64_t end = std::min<uint64_t>(s.size(), start + read_len);
				bases.insert(bases.end(), s.begin() + start, s.begin() + end);
				offsets.push_back(bases.size());
			}
	}
	// native archives keep the packed sequences as they are, one part per sequence (2 bits per base: the reference's level-9 coder
	// reaches ~1.85 on real genomes, but is one serial chain over the whole genome)
	const std::vector<std::vector<uint8_t>>& Raw() const { return sequences; }
	static std::vector<uint8_t> Unpack(const std::vector<uint8_t>& packed)
	{
		if (packed.empty() || packed.back() > 3) throw std::runtime_error("colord-b200: damaged reference-genome stream");
		const size_t in_last = packed.back(), full = packed.size() - 2 + (in_last ? 0 : 1), n = full * 4 + in_last;
		if (packed.size() < 2 && in_last) throw std::runtime_error("colord-b200: damaged reference-genome stream");
		std::vector<uint8_t> seq(n);
		for (size_t i = 0; i < n; ++i) seq[i] = "ACGT"[(packed[i / 4] >> (6 - 2 * (i % 4))) & 3];
		return seq;
	}
};

} // namespace clbhost
