// compressor.h — runCompression over the C-ABI: input file -> archive (SURVEY.md §8b "outermost contract", §8f rows 1 and 3).
// Follows src/colord/compression.cpp:344-810 step by step — open the archive and register `meta`, read the input, stage 1a
// (k-mer counts -> filter), derived values (mean read length :443, sparse range :501-504, accepted reference reads :539),
// stage 1b, stage 2, the three stage-3 streams, then the `meta` record (:705-779) and `info` (:781-783) — with every compute
// stage a call into libcolord_b200.so (include/colord_b200.h) instead of a thread over queues.
//
// The stage-3 streams are the device's native containers (DB01 / QB01 / QO01 / HB01), not the reference's
// adaptive range-coder streams, so they are stored under their own names — "dna-b200", "qual-b200", "header-b200", one part
// each — and the archive's `info` carries B200_VERSION_MAJOR as its major version: the reference refuses such an archive at
// its version check (decompression_common.cpp:36-43) instead of misreading it.  Container, `meta` and `info` are the
// reference's formats (archive_host.h), so `colord info` of the reference prints this archive's record.
// Compat mode (--compat, and by default for inputs up to CCompressorParams::compat_max_bases): the three streams are the reference's
// own ("dna", "qual", "header", one part per pack; clb_x*_encode, colord_b200/csrc/stage3_exact.cu) and `info` carries the
// reference's version 1.2.1, so the unmodified `colord decompress` reads the archive and its size is the reference's; every quality
// mode of the reference is available there.
// Reference-genome mode (-G [-s], compression.cpp:401-452, :508-513, :763-779): the genome's sequences are counted with the reads
// (clb_count_sequences), its pseudo-reads enter as always-accepted reference reads in front of the input's reads
// (clb_append_context_reads), `meta` records the pseudo-read geometry and the MD5 of the genome or, with -s, the genome itself goes
// into the archive ("ref-genome": the reference's own coding through clb_xplain_encode in compat archives, packed 2-bit sequences in
// native ones).
// Not covered (refused with a message): the threshold / plain-average quality modes in native containers.
#pragma once
#include <chrono>
#include <cstdio>
#include <ctime>
#include <iostream>
#include <thread>
#include <memory>
#include <exception>
#include "archive_host.h"
#include "fastq_reader.h"
#include "presets.h"
#include "stage23_host.h"
#include "ref_genome_host.h"
#include "stats_report.h"

namespace clbhost {

struct Phase { const char* name; double seconds; };
struct CompressionReport {            // what the reference prints at the end (compression.cpp:795-808)
	bool compat = false, streamed = false; unsigned reader_threads = 1; std::vector<Phase> phases;
	uint64_t dna = 0, qual = 0, header = 0, meta = 0, info = 0, archive = 0;
	uint32_t kmerLen = 0, anchorLen = 0, sparse_range = 0, tot_ref_reads = 0;
	clb_kmer_stats stats{};
	ReadStats read_stats; clb_encode_stats encode_stats{}; bool has_encode_stats = false;      // -v: the reference's statistics block (stats_report.h)
	double seconds = 0;
};

inline void refuse_unsupported(const CCompressorParams& p, bool compat = false)
{
	// limits of the device path, checked before any work is done (the C-ABI would refuse them only after stages 1 and 2)
	if (p.maxCandidates < 1 || p.maxCandidates > 32) throw std::invalid_argument("the number of candidate reads (-c) must be in 1..32 in this build");
	if (p.compressionLevel < 1 || p.compressionLevel > 3) throw std::invalid_argument("the compression level must be in 1..3");
	if (!compat) switch (p.qualityComprMode) {
	case QualityComprMode::Original: case QualityComprMode::QuinaryAverage: case QualityComprMode::QuadAverage: case QualityComprMode::BinaryAverage: case QualityComprMode::None: break;
	default: throw std::invalid_argument(std::string("quality mode '") + qualityComprModeToString(p.qualityComprMode) + "' is available in compat streams only (--compat; inputs up to a few Gbases): the native containers hold org, 2-avg, 4-avg, 5-avg, none");
	}
}

inline CompressionReport runCompressionTo(const CCompressorParams& params, CInfo& info, CArchive& archive, bool allow_streaming);

// The output is opened only once the input has been read and the parameters accepted, and a run that fails removes what it
// wrote: no truncated file is left looking like an archive.
inline CompressionReport runCompression(const CCompressorParams& params, CInfo& info)
{
	refuse_unsupported(params, params.streamFormat != StreamFormat::Native);
	CArchive archive(false);
	try {
		try {
			return runCompressionTo(params, info, archive, true);
		} catch (const StreamingFallback&) {      // the streaming reader met an irregular input: the whole-file reader decides (nothing has been written yet)
			return runCompressionTo(params, info, archive, false);
		}
	} catch (...) {
		if (archive.IsOpen()) { archive.Abandon(); std::remove(params.outputFilePath.c_str()); }
		throw;
	}
}

inline CompressionReport runCompressionTo(const CCompressorParams& params, CInfo& info, CArchive& archive, bool allow_streaming)
{
	const auto t0 = std::chrono::steady_clock::now();
	CompressionReport rep;
	info.version_major = B200_VERSION_MAJOR; info.version_minor = B200_VERSION_MINOR; info.version_patch = B200_VERSION_PATCH;
	std::vector<Phase> phases; auto t_ph = t0;
	auto phase = [&](const char* name) { const auto t = std::chrono::steady_clock::now(); phases.push_back(Phase{name, std::chrono::duration<double>(t - t_ph).count()}); t_ph = t; };

	// Input.  Plain FASTQ files stream: pieces parsed by a pool of threads go to the device as they arrive (clb_append_reads +
	// clb_append_quals: stage 1a and the copies overlap the parsing, the host keeps only the headers).  Everything else (FASTA, gzip,
	// small or irregular files) is read whole first.  The reference reads its input twice, through KMC and through CInputReads.
	uint64_t file_bytes = 0; bool is_gzip = false, fastq_guess = false;
	const bool stream = CInputReads::streamable(params.inputFilePath, file_bytes, is_gzip, fastq_guess) && allow_streaming;
	uint32_t kmerLen = params.kmerLen, anchorLen = params.anchorLen;
	const bool hifi = params.dataSource == DataSource::PBHiFi;
	std::unique_ptr<CInputReads> inp; std::unique_ptr<CKmerCounter> counter;
	// The device context (CUDA start-up, count table) is made by a thread of its own while the input is being read: on a fresh
	// process it costs 1 - 3 s (measured on the B200 boxes, profiles/r02d_*), as much as parsing several GB.  What it needs — k,
	// the table's size — follows from the file's size and first bytes alone (compression.cpp:41-93).
	adjustKmerAndAnchorLen(kmerLen, anchorLen, is_gzip, fastq_guess, file_bytes);
	if (params.verbose) {
		std::cerr << (is_gzip ? "input is gzipped\n" : "input is not gzipped\n");          // compression.cpp:365-371
		PrintParams(std::cerr, params, kmerLen, anchorLen, params.nThreads);
	}
	std::exception_ptr counter_error;
	std::thread make_counter([&] {
		try {
			counter = std::make_unique<CKmerCounter>(kmerLen, params.minKmerCount, params.maxKmerCount, params.filterHashModulo, params.maxCandidates, hifi,
				is_gzip ? file_bytes * 2 : file_bytes / 2, params.device);
		} catch (...) { counter_error = std::current_exception(); }
	});
	struct Joiner { std::thread& t; ~Joiner() { if (t.joinable()) t.join(); } } joiner{make_counter};
	auto wait_counter = [&]() -> clb_ctx* { if (make_counter.joinable()) make_counter.join(); if (counter_error) std::rethrow_exception(counter_error); return counter->Context(); };
	if (stream) {
		// large inputs: the piece buffers are page-locked (full-rate, asynchronous host-to-device copies); below that the time to lock
		// them is not paid back
		const CInputReads::HostAlloc pinned{[](uint64_t bytes) { return clb_host_alloc(bytes); }, [](void* p) { clb_host_free(p); }};
		const bool pin = file_bytes >= (8ull << 30) || std::getenv("CLB_PIN_INPUT") != nullptr;
		clb_ctx* c0 = nullptr;
		inp = std::make_unique<CInputReads>(params.inputFilePath, [&](const uint8_t* b, const uint8_t* q, const uint64_t* off, uint32_t n) {
			if (!c0) { phase("first pieces parsed"); c0 = wait_counter(); phase("device context (rest of it)"); }
			if (params.verbose) rep.read_stats.log_all(off, n);
			check(c0, clb_append_reads(c0, b, off, n, 0), "clb_append_reads");
			check(c0, clb_append_quals(c0, q, off[n], 0), "clb_append_quals");
		}, 0, 64u << 20, pin ? &pinned : nullptr);
		wait_counter();
		phase("read input (streamed to the device, stage 1a inside)");
	} else {
		inp = std::make_unique<CInputReads>(params.inputFilePath);
		phase("read input");
		if (inp->is_fastq != fastq_guess || inp->is_gzip != is_gzip) throw std::runtime_error("Error: the input changed while it was being read");
		clb_ctx* c0 = wait_counter();
		phase("device context (rest of it)");
		if (params.verbose) rep.read_stats.log_all(inp->offsets.data(), inp->n_reads());
		check(c0, clb_append_reads(c0, inp->bases.data(), inp->offsets.data(), inp->n_reads(), 0), "clb_append_reads");
		phase("reads to the device (stage 1a inside)");
	}
	CInputReads& in = *inp; CKmerCounter& kmer_counter = *counter;
	clb_ctx* ctx = kmer_counter.Context();
	rep.kmerLen = kmerLen; rep.anchorLen = anchorLen; rep.streamed = in.streamed; rep.reader_threads = in.threads_used;
	const uint8_t* quals_ptr = in.streamed ? nullptr : in.quals.data();      // streamed: the qualities are on the device already
	const uint64_t* quals_off = in.streamed ? nullptr : in.offsets.data();
	const bool compat = params.streamFormat == StreamFormat::Compat || (params.streamFormat == StreamFormat::Auto && in.total_bases <= params.compat_max_bases);
	refuse_unsupported(params, compat);
	rep.compat = compat;
	if (compat) { info.version_major = 1; info.version_minor = 2; info.version_patch = 1; }      // defs.h:24-26 of the reference: what its decompressor checks
	if (!archive.Open(params.outputFilePath)) throw std::runtime_error("Error: cannot open archive: " + params.outputFilePath);
	const int s_meta = archive.RegisterStream("meta");
	auto add_part = [&](int stream_id, const std::vector<uint8_t>& data, size_t metadata) {
		if (!archive.AddPart(stream_id, data, metadata)) throw std::runtime_error("Error: cannot write to archive: " + params.outputFilePath);
	};
	const bool is_fastq = in.is_fastq;
	info.total_bytes = in.total_bytes; info.total_bases = in.total_bases;

	// reference genome: its sequences are a second counting input (compression.cpp:408-430)
	std::unique_ptr<CReferenceGenome> ref_genome;
	if (!params.refGenomePath.empty()) {
		ref_genome = std::make_unique<CReferenceGenome>(params.refGenomePath);
		std::vector<uint8_t> gb; std::vector<uint64_t> go;
		ref_genome->Sequences(gb, go);
		check(ctx, clb_count_sequences(ctx, gb.data(), go.data(), ref_genome->GetTotNSeqs(), 0), "clb_count_sequences");
		phase("reference genome read and counted");
	}
	const uint32_t tot_n_reads = kmer_counter.GetNReads();
	const uint64_t tot_kmers = kmer_counter.GetTotKmers(), n_uniq_counted_kmers = kmer_counter.GetNUniqueCounted();
	check(ctx, clb_count_finalize(ctx, &rep.stats), "clb_count_finalize");
	if (tot_n_reads == 0) throw std::runtime_error("Error: no reads in the input");
	// KMC's read count includes the genome's sequences; the statistics are then corrected for them (compression.cpp:443-452)
	const uint64_t n_genome_seqs = ref_genome ? ref_genome->GetTotNSeqs() : 0;
	uint64_t mean_read_len = meanReadLen(tot_kmers, params.filterHashModulo, tot_n_reads + n_genome_seqs, kmerLen);
	uint32_t ref_genome_read_len = 0, n_pseudo = 0; const uint32_t ref_genome_overlap_size = (kmerLen - 1) * 10;
	if (ref_genome) {
		mean_read_len = static_cast<uint64_t>(double(mean_read_len * (tot_n_reads + n_genome_seqs) - ref_genome->GetTotSeqsLen()) / tot_n_reads);
		ref_genome_read_len = static_cast<uint32_t>(20 * mean_read_len);
		ref_genome->SetReadLen(ref_genome_read_len, ref_genome_overlap_size);
		if (!ref_genome->valid_read_len()) throw std::runtime_error("Error: the reads are too short for the reference-genome mode (pseudo-read length <= overlap)");
		n_pseudo = ref_genome->GetNPseudoReads();
		std::vector<uint8_t> pb; std::vector<uint64_t> po;
		ref_genome->PseudoReads(pb, po);
		check(ctx, clb_append_context_reads(ctx, pb.data(), po.data(), n_pseudo, 0), "clb_append_context_reads");
		if (params.verbose) std::cerr << "# ref genome pseudo reads: " << n_pseudo << "\n";
		phase("pseudo-reads to the device");
	}
	info.total_reads = tot_n_reads;
	if (params.verbose) std::cerr << "tot k-mers: " << tot_kmers << "\nn uniq counted: " << n_uniq_counted_kmers << "\napprox. avg. read len: " << mean_read_len << "\n";

	// reference reads: all, or the sparse sampler over the range derived from the filtered k-mers
	const uint32_t sparseMode_range = sparseModeRange(params.sparseMode_range_symbols, n_uniq_counted_kmers, params.filterHashModulo, mean_read_len ? mean_read_len : 1);
	const bool sparse = params.referenceReadsMode == ReferenceReadsMode::Sparse;
	CRefReadsAccepter accepter(sparseMode_range, params.sparseMode_exponent, n_pseudo);
	uint32_t tot_ref_reads = sparse ? accepter.GetNAccepted(tot_n_reads + n_pseudo) : tot_n_reads + n_pseudo;      // ref_reads_accepter.h:42-49, compression.cpp:513
	rep.sparse_range = sparseMode_range; rep.tot_ref_reads = tot_ref_reads;

	// stage 1b + 2
	CReadsSimilarityGraph graph(kmer_counter, params.maxCandidates, hifi, sparse, accepter, n_pseudo);
	CEncoder encoder(kmer_counter, anchorLen, params.minFractionOfMmersInEncodeToAlwaysEncode, params.minFractionOfMmersInEncode, params.maxMatchesMultiplier,
		params.editScriptCostMultiplier, params.minPartLenToConsiderAltRead, params.maxRecurence, params.minAnchors);
	if (params.verbose) check(ctx, clb_encode_stats_enable(ctx, 1), "clb_encode_stats_enable");
	encoder.Encode(in.read_pack_sizes);
	if (params.verbose) { check(ctx, clb_encode_stats_get(ctx, &rep.encode_stats), "clb_encode_stats_get"); rep.has_encode_stats = true; }

	phase("stages 1b + 2");
	int s_dna = -1, s_qual = -1, s_header = -1;
	if (ref_genome && params.storeRefGenome) {
		if (compat) {      // CReferenceGenome::Store(archive) (reference_genome.cpp:319-360)
			const int s_gen = archive.RegisterStream("ref-genome");
			std::vector<uint8_t> gb; std::vector<uint64_t> go;
			ref_genome->Sequences(gb, go);
			check(ctx, clb_xplain_encode(ctx, gb.data(), go.data(), ref_genome->GetTotNSeqs(), 9), "clb_xplain_encode");
			uint64_t total = 0; uint32_t n_parts = 0;
			check(ctx, clb_xstream_size(ctx, 3, &total, &n_parts), "clb_xstream_size");
			std::vector<uint8_t> bytes(total + 1); uint64_t sz[2] = {0, 0};
			check(ctx, clb_xstream_get(ctx, 3, bytes.data(), total, sz, 0), "clb_xstream_get");
			bytes.resize(total);
			add_part(s_gen, bytes, ref_genome->GetTotNSeqs());
		} else {
			const int s_gen = archive.RegisterStream("ref-genome-b200");
			for (const auto& seq : ref_genome->Raw()) add_part(s_gen, CReferenceGenome::Pack(seq), 0);
		}
		phase("reference genome stored");
	}
	if (compat) {
		// the reference's own streams: one part per read pack / header pack (entr_read.h:56-80, entr_qual.h:100-126, entr_header.cpp:23-46)
		auto add_parts = [&](int stream_id, uint32_t which, const std::vector<uint32_t>* metadata) {
			uint64_t total = 0; uint32_t n_parts = 0;
			check(ctx, clb_xstream_size(ctx, which, &total, &n_parts), "clb_xstream_size");
			std::vector<uint8_t> bytes(total + 1); std::vector<uint64_t> sizes(n_parts + 1);
			check(ctx, clb_xstream_get(ctx, which, bytes.data(), total, sizes.data(), 0), "clb_xstream_get");
			uint64_t at = 0;
			for (uint32_t p = 0; p < n_parts; ++p) {
				if (!archive.AddPart(stream_id, bytes.data() + at, sizes[p], metadata ? (*metadata)[p] : 0)) throw std::runtime_error("Error: cannot write to archive: " + params.outputFilePath);
				at += sizes[p];
			}
		};
		s_dna = archive.RegisterStream("dna");
		check(ctx, clb_xdna_encode(ctx, static_cast<uint32_t>(params.compressionLevel), in.read_pack_sizes.data(), static_cast<uint32_t>(in.read_pack_sizes.size())), "clb_xdna_encode");
		add_parts(s_dna, 0, &in.read_pack_sizes);
		phase("dna stream");
		if (is_fastq) {
			s_qual = archive.RegisterStream("qual");
			uint32_t thr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
			for (size_t i = 0; i < params.qualityFwdThresholds.size() && i < 8; ++i) thr[i] = params.qualityFwdThresholds[i];
			check(ctx, clb_xqual_encode(ctx, static_cast<uint32_t>(params.qualityComprMode), static_cast<uint32_t>(params.dataSource), static_cast<uint32_t>(params.compressionLevel), thr,
				quals_ptr, quals_off, 0, in.read_pack_sizes.data(), static_cast<uint32_t>(in.read_pack_sizes.size())), "clb_xqual_encode");
			add_parts(s_qual, 1, nullptr);
			phase("quality stream");
		}
		s_header = archive.RegisterStream("header");
		if (params.headerComprMode == HeaderComprMode::Original) {
			check(ctx, clb_xhdr_encode(ctx, in.headers.data(), in.header_offsets.data(), in.plus_id.data(), in.header_offsets.size() - 1, 0,
				in.header_pack_sizes.data(), static_cast<uint32_t>(in.header_pack_sizes.size())), "clb_xhdr_encode");
			add_parts(s_header, 2, &in.header_pack_sizes);
		} else {      // -i none / main: CIDCoder::Encode returns at once (id_coder.cpp:102-110), every part is the coder's 8-byte flush
			const std::vector<uint8_t> flush(8, 0);
			for (uint32_t n_in_pack : in.header_pack_sizes) add_part(s_header, flush, n_in_pack);
		}
		phase("header stream");
	} else {
		// stage 3: one part per stream, metadata = number of reads / headers as in entr_read.h:74, entr_header.cpp:41
		s_dna = archive.RegisterStream("dna-b200");
		{
			CEntrComprReads dna(kmer_counter, params.compressionLevel);
			dna.Compress(in.read_pack_sizes);
			const std::vector<uint8_t> stream = dna.GetStream();
			add_part(s_dna, stream, tot_n_reads);
		}
		if (is_fastq) {
			s_qual = archive.RegisterStream("qual-b200");
			std::vector<uint8_t> stream;
			if (params.qualityComprMode != QualityComprMode::None) {
				const uint32_t n_bins = params.qualityComprMode == QualityComprMode::BinaryAverage ? 2 : params.qualityComprMode == QualityComprMode::QuadAverage ? 4 : 5;
				CEntrComprQuals q(kmer_counter, n_bins, params.qualityFwdThresholds, params.compressionLevel);
				if (params.qualityComprMode == QualityComprMode::Original) q.CompressOriginal(static_cast<uint32_t>(params.dataSource), quals_ptr, quals_off, in.read_pack_sizes);
				else q.Compress(quals_ptr, quals_off, in.read_pack_sizes);
				stream = q.GetStream();
			}
			add_part(s_qual, stream, 0);
		}
		s_header = archive.RegisterStream("header-b200");
		{	// -i none / main store no header bytes, as in the reference (id_coder.cpp:102-110: Encode returns at once for `none`, and
			// compress_instrument — `main` — is an empty function there); the decompressor prints "@" / "" for every read (:393-396, :588-591)
			std::vector<uint8_t> stream;
			if (params.headerComprMode == HeaderComprMode::Original) {
				CEntrComprHeaders h(kmer_counter);
				h.Compress(in.headers.data(), in.header_offsets.data(), in.plus_id.data(), in.header_offsets.size() - 1);
				stream = h.GetStream();
			}
			add_part(s_header, stream, in.header_offsets.size() - 1);
		}

		phase("stage 3 (native containers)");
	}

	// `meta` (compression.cpp:705-779)
	CMeta meta;
	meta.tot_ref_reads = tot_ref_reads; meta.maxCandidates = params.maxCandidates; meta.compressionLevel = params.compressionLevel;
	meta.dataSource = params.dataSource; meta.approx_stream_size = static_cast<uint64_t>(tot_n_reads) * mean_read_len;
	meta.is_fastq = is_fastq; meta.qualityComprMode = params.qualityComprMode;
	meta.qualityRevThresholds = params.qualityRevThresholds;
	meta.qualityRevThresholds.resize(CMeta::n_thresholds(params.qualityComprMode));
	meta.headerComprMode = params.headerComprMode; meta.referenceReadsMode = params.referenceReadsMode;
	meta.sparseMode_range = sparseMode_range; meta.sparseMode_exponent = params.sparseMode_exponent;
	if (ref_genome) {
		meta.ref_genome_available = true; meta.storeRefGenome = params.storeRefGenome;
		meta.ref_genome_read_len = ref_genome_read_len; meta.ref_genome_overlap_size = ref_genome_overlap_size; meta.n_ref_genome_pseudo_reads = n_pseudo;
		if (!params.storeRefGenome) meta.ref_genome_checksum = ref_genome->GetChecksum();
	}
	add_part(s_meta, meta.Serialize(), 0);
	const int s_info = archive.RegisterStream("info");
	info.time = static_cast<uint64_t>(std::time(nullptr));
	add_part(s_info, info.Serialize(), 0);
	if (!archive.Close()) throw std::runtime_error("Error: cannot write to archive: " + params.outputFilePath);

	rep.dna = archive.GetStreamPackedSize(s_dna); rep.qual = s_qual >= 0 ? archive.GetStreamPackedSize(s_qual) : 0; rep.header = archive.GetStreamPackedSize(s_header);
	rep.meta = archive.GetStreamPackedSize(s_meta); rep.info = archive.GetStreamPackedSize(s_info);
	phase("meta, info, close");
	rep.phases = phases;
	rep.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	return rep;
}

} // namespace clbhost
