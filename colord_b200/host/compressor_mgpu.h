// compressor_mgpu.h — `colord-b200 compress-* --gpus N`: one input file on N GPUs of one box (the north star's
// "reads shard by id ... a single NCCL all-gather of the filtered-k-mer table ... per-rank archive segments concatenated on the host").
// One process, one host thread + one clb_ctx per GPU:
//   1. the file is cut into N byte ranges at record starts; rank r streams range r to its GPU (fastq_reader.h streaming form with a
//      range: pieces parsed by the rank's share of the reader threads, clb_append_reads + clb_append_quals) — contiguous read ids
//   2. clb_group_exchange_counts (libcolord_b200_mgpu.so, NCCL): the filtered k-mer set of the WHOLE input on every rank
//   3. the host derives mean read length / sparse range / sampler decisions from the global statistics as for one GPU
//      (compression.cpp:443, :501-504, ref_reads_accepter.h) — the decisions are a function of the GLOBAL read id
//   4. clb_group_exchange_reference_reads: every rank holds the reference reads of the ranks before it as context reads, so the
//      candidates and tuples of a shard are the ones a single GPU computes for those reads (tests/test_gpu_shard.py)
//   5. stages 1b / 2 / 3 per rank, native containers; the archive gets one part per stream and rank ("segments"): the DNA container
//      of rank r names the reference reads of ranks < r as its context reads, which is exactly what the decoder has decoded by then
//      (decompressor.h walks the shards in order).  `meta` / `info` are the whole input's.
// Limits: plain FASTQ input, native containers (the reference's own streams are one chain through the whole file), no -G.
#pragma once
#include <exception>
#include <memory>
#include <thread>
#include "compressor.h"
#include "../../include/colord_b200_mgpu.h"

namespace clbhost {

inline CompressionReport runCompressionMultiGpu(const CCompressorParams& params, CInfo& info, uint32_t n_gpus)
{
	// the exchange buffers of libcolord_b200_mgpu.so come from cudaMalloc after the contexts exist: the contexts must not have taken
	// the devices' free memory as their slabs (csrc/slab.h)
	setenv("CLB_SLAB_GB", "0", 0);
	const auto t0 = std::chrono::steady_clock::now();
	if (n_gpus < 2) throw std::invalid_argument("--gpus takes 2 or more (leave it out for one GPU)");
	if (!params.refGenomePath.empty()) throw std::invalid_argument("--gpus: the reference-genome mode (-G) runs on one GPU");
	if (params.streamFormat == StreamFormat::Compat) throw std::invalid_argument("--gpus writes native containers: the reference's own streams (--compat) are one chain through the whole file");
	refuse_unsupported(params, false);
	uint64_t file_bytes = 0; bool is_gzip = false, is_fastq = false;
	CInputReads::streamable(params.inputFilePath, file_bytes, is_gzip, is_fastq, 0);
	if (is_gzip || !is_fastq) throw std::invalid_argument("--gpus takes a plain (not gzipped) FASTQ file");
	CompressionReport rep; rep.streamed = true;
	std::vector<Phase> phases; auto t_ph = t0;
	auto phase = [&](const char* name) { const auto t = std::chrono::steady_clock::now(); phases.push_back(Phase{name, std::chrono::duration<double>(t - t_ph).count()}); t_ph = t; };
	info.version_major = B200_VERSION_MAJOR; info.version_minor = B200_VERSION_MINOR; info.version_patch = B200_VERSION_PATCH;
	uint32_t kmerLen = params.kmerLen, anchorLen = params.anchorLen;
	adjustKmerAndAnchorLen(kmerLen, anchorLen, false, true, file_bytes);
	rep.kmerLen = kmerLen; rep.anchorLen = anchorLen;
	const bool hifi = params.dataSource == DataSource::PBHiFi;
	if (params.verbose) { std::cerr << "input is not gzipped\n"; PrintParams(std::cerr, params, kmerLen, anchorLen, params.nThreads); std::cerr << "GPUs: " << n_gpus << "\n"; }

	struct Rank {
		std::unique_ptr<CKmerCounter> counter; std::unique_ptr<CInputReads> in; std::exception_ptr err;
		std::vector<uint8_t> dna, qual, hdr; uint32_t n_context = 0;
		clb_encode_stats enc_stats{};
	};
	std::vector<Rank> ranks(n_gpus);
	auto run_all = [&](auto&& fn) {      // fn(rank) on one thread per rank; the first error is rethrown when all are back
		std::vector<std::thread> th;
		for (uint32_t r = 0; r < n_gpus; ++r) th.emplace_back([&, r] { try { fn(r); } catch (...) { ranks[r].err = std::current_exception(); } });
		for (auto& t : th) t.join();
		for (Rank& k : ranks) if (k.err) std::rethrow_exception(k.err);
	};
	const unsigned reader_threads = std::max(1u, std::max(std::thread::hardware_concurrency(), 1u) / n_gpus);
	rep.reader_threads = reader_threads * n_gpus;

	// 1. every rank: device context, its byte range of the file streamed to its GPU (stage 1a inside)
	run_all([&](uint32_t r) {
		Rank& K = ranks[r];
		K.counter = std::make_unique<CKmerCounter>(kmerLen, params.minKmerCount, params.maxKmerCount, params.filterHashModulo, params.maxCandidates, hifi, file_bytes / 2 / n_gpus, params.device + static_cast<int>(r));
		clb_ctx* c = K.counter->Context();
		K.in = std::make_unique<CInputReads>(params.inputFilePath, [c](const uint8_t* b, const uint8_t* q, const uint64_t* off, uint32_t n) {
			check(c, clb_append_reads(c, b, off, n, 0), "clb_append_reads");
			check(c, clb_append_quals(c, q, off[n], 0), "clb_append_quals");
		}, reader_threads, 64u << 20, nullptr, file_bytes / n_gpus * r, r + 1 == n_gpus ? ~0ull : file_bytes / n_gpus * (r + 1));
	});
	phase("device contexts + input streamed to the GPUs (stage 1a inside)");
	uint64_t tot_n_reads64 = 0, total_bases = 0;
	std::vector<uint64_t> first_read(n_gpus + 1, 0);
	for (uint32_t r = 0; r < n_gpus; ++r) { tot_n_reads64 += ranks[r].in->n_reads(); total_bases += ranks[r].in->total_bases; first_read[r + 1] = tot_n_reads64; }
	if (tot_n_reads64 == 0) throw std::runtime_error("Error: no reads in the input");
	if (tot_n_reads64 >= (1ull << 32)) throw std::runtime_error("Error: too many reads");
	const uint32_t tot_n_reads = static_cast<uint32_t>(tot_n_reads64);
	info.total_bytes = file_bytes; info.total_bases = total_bases; info.total_reads = tot_n_reads;

	// 2. the filtered k-mer set of the whole input on every rank
	std::vector<clb_ctx*> ctxs; std::vector<int32_t> devs;
	for (uint32_t r = 0; r < n_gpus; ++r) { ctxs.push_back(ranks[r].counter->Context()); devs.push_back(params.device + static_cast<int>(r)); }
	clb_group* group = nullptr;
	if (clb_group_create(ctxs.data(), devs.data(), n_gpus, &group) != CLB_OK) throw std::runtime_error("clb_group_create: NCCL communicators could not be made");
	struct GroupGuard { clb_group* g; ~GroupGuard() { clb_group_destroy(g); } } guard{group};
	std::vector<clb_kmer_stats> stats(n_gpus);
	run_all([&](uint32_t r) { if (clb_group_exchange_counts(group, r, &stats[r]) != CLB_OK) throw std::runtime_error(std::string("clb_group_exchange_counts: ") + clb_group_last_error(group, r)); });
	phase("k-mer counts exchanged (NCCL all-to-all + all-gather)");
	rep.stats = stats[0];

	// 3. derived values, exactly as for one GPU
	const uint64_t tot_kmers = stats[0].tot_kmers, n_uniq_counted_kmers = stats[0].n_unique_counted;
	const uint64_t mean_read_len = meanReadLen(tot_kmers, params.filterHashModulo, tot_n_reads, kmerLen);
	if (params.verbose) std::cerr << "tot k-mers: " << tot_kmers << "\nn uniq counted: " << n_uniq_counted_kmers << "\napprox. avg. read len: " << mean_read_len << "\n";
	const uint32_t sparseMode_range = sparseModeRange(params.sparseMode_range_symbols, n_uniq_counted_kmers, params.filterHashModulo, mean_read_len ? mean_read_len : 1);
	const bool sparse = params.referenceReadsMode == ReferenceReadsMode::Sparse;
	CRefReadsAccepter accepter(sparseMode_range, params.sparseMode_exponent, 0);
	const std::vector<uint8_t> decisions = sparse ? accepter.Decisions(tot_n_reads) : std::vector<uint8_t>(tot_n_reads, 1);
	uint32_t tot_ref_reads = 0;
	for (uint8_t d : decisions) tot_ref_reads += d;
	rep.sparse_range = sparseMode_range; rep.tot_ref_reads = tot_ref_reads;

	// 4 + 5. reference reads exchanged, then every rank runs stages 1b, 2 and 3 on its shard
	const bool fastq_quals = params.qualityComprMode != QualityComprMode::None;
	run_all([&](uint32_t r) {
		Rank& K = ranks[r]; clb_ctx* c = K.counter->Context(); CInputReads& in = *K.in;
		const uint32_t n_local = in.n_reads();
		const uint8_t* dec = decisions.data() + first_read[r];
		if (clb_group_exchange_reference_reads(group, r, dec, in.ReadLengths().data(), n_local, &K.n_context) != CLB_OK)
			throw std::runtime_error(std::string("clb_group_exchange_reference_reads: ") + clb_group_last_error(group, r));
		check(c, clb_graph_build(c, dec, 0), "clb_graph_build");
		CEncoder encoder(*K.counter, anchorLen, params.minFractionOfMmersInEncodeToAlwaysEncode, params.minFractionOfMmersInEncode, params.maxMatchesMultiplier,
			params.editScriptCostMultiplier, params.minPartLenToConsiderAltRead, params.maxRecurence, params.minAnchors);
		if (params.verbose) check(c, clb_encode_stats_enable(c, 1), "clb_encode_stats_enable");
		encoder.Encode(in.read_pack_sizes);
		if (params.verbose) check(c, clb_encode_stats_get(c, &K.enc_stats), "clb_encode_stats_get");
		{ CEntrComprReads dna(*K.counter, params.compressionLevel); dna.Compress(in.read_pack_sizes); K.dna = dna.GetStream(); }
		if (fastq_quals) {
			const uint32_t n_bins = params.qualityComprMode == QualityComprMode::BinaryAverage ? 2 : params.qualityComprMode == QualityComprMode::QuadAverage ? 4 : 5;
			CEntrComprQuals q(*K.counter, n_bins, params.qualityFwdThresholds, params.compressionLevel);
			if (params.qualityComprMode == QualityComprMode::Original) q.CompressOriginal(static_cast<uint32_t>(params.dataSource), nullptr, nullptr, in.read_pack_sizes);
			else q.Compress(nullptr, nullptr, in.read_pack_sizes);
			K.qual = q.GetStream();
		}
		if (params.headerComprMode == HeaderComprMode::Original) {
			CEntrComprHeaders h(*K.counter);
			h.Compress(in.headers.data(), in.header_offsets.data(), in.plus_id.data(), in.header_offsets.size() - 1);
			K.hdr = h.GetStream();
		}
	});
	phase("reference reads exchanged (NCCL all-gather), stages 1b + 2 + 3 on every GPU");

	// the archive: one part per stream and rank, in rank order
	CArchive archive(false);
	if (!archive.Open(params.outputFilePath)) throw std::runtime_error("Error: cannot open archive: " + params.outputFilePath);
	try {
		auto add_part = [&](int stream_id, const std::vector<uint8_t>& data, size_t metadata) {
			if (!archive.AddPart(stream_id, data, metadata)) throw std::runtime_error("Error: cannot write to archive: " + params.outputFilePath);
		};
		const int s_meta = archive.RegisterStream("meta"), s_dna = archive.RegisterStream("dna-b200"), s_qual = archive.RegisterStream("qual-b200"), s_header = archive.RegisterStream("header-b200");
		for (uint32_t r = 0; r < n_gpus; ++r) {
			add_part(s_dna, ranks[r].dna, ranks[r].in->n_reads());
			add_part(s_qual, ranks[r].qual, 0);
			add_part(s_header, ranks[r].hdr, ranks[r].in->n_reads());
		}
		CMeta meta;
		meta.tot_ref_reads = tot_ref_reads; meta.maxCandidates = params.maxCandidates; meta.compressionLevel = params.compressionLevel;
		meta.dataSource = params.dataSource; meta.approx_stream_size = static_cast<uint64_t>(tot_n_reads) * mean_read_len;
		meta.is_fastq = true; meta.qualityComprMode = params.qualityComprMode;
		meta.qualityRevThresholds = params.qualityRevThresholds;
		meta.qualityRevThresholds.resize(CMeta::n_thresholds(params.qualityComprMode));
		meta.headerComprMode = params.headerComprMode; meta.referenceReadsMode = params.referenceReadsMode;
		meta.sparseMode_range = sparseMode_range; meta.sparseMode_exponent = params.sparseMode_exponent;
		add_part(s_meta, meta.Serialize(), 0);
		const int s_info = archive.RegisterStream("info");
		info.time = static_cast<uint64_t>(std::time(nullptr));
		add_part(s_info, info.Serialize(), 0);
		if (!archive.Close()) throw std::runtime_error("Error: cannot write to archive: " + params.outputFilePath);
		rep.dna = archive.GetStreamPackedSize(s_dna); rep.qual = archive.GetStreamPackedSize(s_qual); rep.header = archive.GetStreamPackedSize(s_header);
		rep.meta = archive.GetStreamPackedSize(s_meta); rep.info = archive.GetStreamPackedSize(s_info);
	} catch (...) {
		if (archive.IsOpen()) { archive.Abandon(); std::remove(params.outputFilePath.c_str()); }
		throw;
	}
	phase("archive written");
	if (params.verbose) {      // the reference's statistics block (stats_report.h): read statistics in file order, the ranks' counters added up —
		// the shards' candidates are those of one GPU (the estimator restarts at a shard's first pack, so a few short-part decisions differ)
		clb_encode_stats& t = rep.encode_stats;
		for (Rank& K : ranks) {
			for (uint32_t len : K.in->ReadLengths()) rep.read_stats.log(len);
			const clb_encode_stats& e = K.enc_stats;
			t.n_not_enough_unique_mmers_in_enc_read += e.n_not_enough_unique_mmers_in_enc_read; t.n_too_many_matches += e.n_too_many_matches; t.n_too_low_anchors += e.n_too_low_anchors;
			t.n_non_rev_choosen += e.n_non_rev_choosen; t.n_rev_choosen += e.n_rev_choosen;
			t.n_plain_reads_tot += e.n_plain_reads_tot; t.n_plain_symb += e.n_plain_symb; t.n_plain_reads_with_n_tot += e.n_plain_reads_with_n_tot; t.n_plain_with_n_symb += e.n_plain_with_n_symb;
			t.n_levels = std::max(t.n_levels, e.n_levels);
			for (uint32_t l = 0; l < CLB_MAX_STAT_LEVELS; ++l) {
				const uint64_t* src = reinterpret_cast<const uint64_t*>(&e.level[l]); uint64_t* dst = reinterpret_cast<uint64_t*>(&t.level[l]);
				for (size_t f = 0; f < sizeof(clb_level_stats) / sizeof(uint64_t); ++f) dst[f] += src[f];
			}
		}
		rep.has_encode_stats = true;
	}
	rep.phases = phases;
	rep.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	return rep;
}

} // namespace clbhost
