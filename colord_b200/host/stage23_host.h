// stage23_host.h — C++ host-side mirrors of the reference's stage-2 / stage-3 classes over the C-ABI
// (include/colord_b200.h).  Same argument meaning and order of use as
//   CEncoder              src/colord/encoder.h:371-410 (ctor), encoder.cpp:1672-1691 (Encode)
//   CEntrComprQuals       src/colord/entr_qual.h:82-126 (ctor, Compress)
//   CEntrComprReads       src/colord/entr_read.h:35-80 (ctor, Compress)
//   CEntrComprHeaders     src/colord/entr_header.h:30-52, entr_header.cpp:23-46 (Compress)
// The reference runs N CEncoder threads over CCompressPacks and one quality thread over quals packs; here one call covers
// all appended reads and the results stay on the device until they are fetched.  Header-only; link -lcolord_b200.
#pragma once
#include "stage1_host.h"

namespace clbhost {

class CEncoder {
	clb_ctx* ctx; uint32_t n_reads;
	clb_encode_params prm;
public:
	// same order as the reference ctor after the queues: anchor_len, the four doubles, minPartLenToConsiderAltRead, maxRecurence, minAnchors
	CEncoder(CKmerCounter& counter, uint32_t anchor_len, double minFractionOfMmersInEncodeToAlwaysEncode, double minFractionOfMmersInEncode,
		double maxMatchesMultiplier, double editScriptCostMultiplier, uint32_t minPartLenToConsiderAltRead, uint32_t maxRecurence, uint32_t minAnchors)
		: ctx(counter.Context()), n_reads(counter.GetNReads()),
		  prm{anchor_len, minPartLenToConsiderAltRead, maxRecurence, minAnchors, minFractionOfMmersInEncode, minFractionOfMmersInEncodeToAlwaysEncode, maxMatchesMultiplier, editScriptCostMultiplier} {}
	// pack_sizes: reads per read pack in input order (the estimator is reset per pack, encoder.cpp:1677); empty = the reference's pack rule
	void Encode(const std::vector<uint32_t>& pack_sizes = {})
	{
		check(ctx, clb_encode(ctx, &prm, pack_sizes.empty() ? nullptr : pack_sizes.data(), static_cast<uint32_t>(pack_sizes.size())), "clb_encode");
	}
	// es_t of every read (CompactES bytes, utils.h:69-273), what the reference pushes to compressed_queue
	void GetTuples(std::vector<uint64_t>& es_off, std::vector<uint8_t>& es) const
	{
		uint64_t tot = 0; check(ctx, clb_encode_size(ctx, &tot), "clb_encode_size");
		es_off.assign(n_reads + 1, 0); es.assign(tot + 1, 0);
		check(ctx, clb_encode_get(ctx, es_off.data(), es.data(), tot, 0), "clb_encode_get");
		es.resize(tot);
	}
};

class CEntrComprQuals {
	clb_ctx* ctx;
	clb_qual_params prm{};
public:
	// n_bins 2 / 4 / 5 = QualityComprMode Binary/Quad/QuinaryAverage; thresholds = qualityFwdThresholds; level = compressionLevel
	CEntrComprQuals(CKmerCounter& counter, uint32_t n_bins, const std::vector<uint32_t>& qualityFwdThresholds, int32_t compression_level) : ctx(counter.Context())
	{
		prm.n_bins = n_bins; prm.level = static_cast<uint32_t>(compression_level);
		for (size_t i = 0; i < 4 && i < qualityFwdThresholds.size(); ++i) prm.thresholds[i] = qualityFwdThresholds[i];
	}
	// quals: phred+33 of all reads, offsets[n_reads + 1]
	void Compress(const std::vector<uint8_t>& quals, const std::vector<uint64_t>& offsets, const std::vector<uint32_t>& pack_sizes = {})
	{
		check(ctx, clb_qual_encode(ctx, &prm, quals.data(), offsets.data(), 0, pack_sizes.empty() ? nullptr : pack_sizes.data(), static_cast<uint32_t>(pack_sizes.size())), "clb_qual_encode");
	}
	// the same over plain arrays (what the host reader holds)
	void Compress(const uint8_t* quals, const uint64_t* offsets, const std::vector<uint32_t>& pack_sizes = {})
	{
		check(ctx, clb_qual_encode(ctx, &prm, quals, offsets, 0, pack_sizes.empty() ? nullptr : pack_sizes.data(), static_cast<uint32_t>(pack_sizes.size())), "clb_qual_encode");
	}
	void CompressOriginal(uint32_t data_source, const uint8_t* quals, const uint64_t* offsets, const std::vector<uint32_t>& pack_sizes = {})
	{
		check(ctx, clb_qual_encode_original(ctx, data_source, prm.level, quals, offsets, 0, pack_sizes.empty() ? nullptr : pack_sizes.data(), static_cast<uint32_t>(pack_sizes.size())), "clb_qual_encode_original");
	}
	// QualityComprMode::Original (-q org): data_source 0 ONT, 1 PBRaw, 2 PBHiFi (selects the quantiser of the context, quality_coder.cpp:272-505)
	void CompressOriginal(uint32_t data_source, const std::vector<uint8_t>& quals, const std::vector<uint64_t>& offsets, const std::vector<uint32_t>& pack_sizes = {})
	{
		check(ctx, clb_qual_encode_original(ctx, data_source, prm.level, quals.data(), offsets.data(), 0, pack_sizes.empty() ? nullptr : pack_sizes.data(), static_cast<uint32_t>(pack_sizes.size())), "clb_qual_encode_original");
	}
	std::vector<uint8_t> GetStream() const
	{
		uint64_t tot = 0; check(ctx, clb_qual_size(ctx, &tot), "clb_qual_size");
		std::vector<uint8_t> out(tot + 1);
		check(ctx, clb_qual_get(ctx, out.data(), tot, 0), "clb_qual_get");
		out.resize(tot);
		return out;
	}
};

class CEntrComprReads {
	clb_ctx* ctx; uint32_t level;
public:
	// compression_level = compressionLevel (history widths of the DNA model, dna_coder.cpp:1253-1280)
	CEntrComprReads(CKmerCounter& counter, int32_t compression_level) : ctx(counter.Context()), level(static_cast<uint32_t>(compression_level)) {}
	// codes the tuples CEncoder::Encode left on the device; one part per read pack (entr_read.h:66-78)
	void Compress(const std::vector<uint32_t>& pack_sizes = {})
	{
		check(ctx, clb_dna_encode(ctx, level, pack_sizes.empty() ? nullptr : pack_sizes.data(), static_cast<uint32_t>(pack_sizes.size())), "clb_dna_encode");
	}
	std::vector<uint8_t> GetStream() const
	{
		uint64_t tot = 0; check(ctx, clb_dna_size(ctx, &tot, nullptr), "clb_dna_size");
		std::vector<uint8_t> out(tot + 1);
		check(ctx, clb_dna_get(ctx, out.data(), tot, 0), "clb_dna_get");
		out.resize(tot);
		return out;
	}
};

class CEntrComprHeaders {
	clb_ctx* ctx;
public:
	explicit CEntrComprHeaders(CKmerCounter& counter) : ctx(counter.Context()) {}
	// headers: (id, the '+' line repeats the id) in input order = header_pack_t entries (entr_header.cpp:33-34)
	void Compress(const std::vector<std::pair<std::string, bool>>& headers, const std::vector<uint32_t>& pack_sizes = {})
	{
		std::vector<uint8_t> bytes, plus; std::vector<uint64_t> off{0};
		for (const auto& [id, plus_id] : headers) { bytes.insert(bytes.end(), id.begin(), id.end()); off.push_back(bytes.size()); plus.push_back(plus_id); }
		check(ctx, clb_hdr_encode(ctx, bytes.data(), off.data(), plus.data(), headers.size(), 0, pack_sizes.empty() ? nullptr : pack_sizes.data(), static_cast<uint32_t>(pack_sizes.size())), "clb_hdr_encode");
	}
	// the same over headers stored back to back (what the host reader produces)
	void Compress(const uint8_t* bytes, const uint64_t* offsets, const uint8_t* plus_id, uint64_t n_headers, const std::vector<uint32_t>& pack_sizes = {})
	{
		check(ctx, clb_hdr_encode(ctx, bytes, offsets, plus_id, n_headers, 0, pack_sizes.empty() ? nullptr : pack_sizes.data(), static_cast<uint32_t>(pack_sizes.size())), "clb_hdr_encode");
	}
	std::vector<uint8_t> GetStream() const
	{
		uint64_t tot = 0; check(ctx, clb_hdr_size(ctx, &tot, nullptr), "clb_hdr_size");
		std::vector<uint8_t> out(tot + 1);
		check(ctx, clb_hdr_get(ctx, out.data(), tot, 0), "clb_hdr_get");
		out.resize(tot);
		return out;
	}
};

} // namespace clbhost
