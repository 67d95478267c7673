// oracle/ref_stage1_time.cpp — TEST/BENCH INFRASTRUCTURE ONLY (never linked into the product).
//
// Times the reference's OWN stage 1 on the host CPU: the unmodified reference objects are linked as-is
// except compression.o; this runCompression wires the same classes in the same order as
// compression.cpp:432-575 (CKmerCounter -> CKmerFilter -> CInputReads + CReadsSimilarityGraph) and stops
// at the compress_queue, which a null consumer drains (no encoders, no entropy coders).  It prints one
// JSON line with the wall time of each phase.  bench.py --impl reference and the cpu_baseline leg run it.
// With COLORD_TIME_STAGES=12 the N CEncoder threads (compression.cpp:592-625) consume the compress_queue as in the real
// pipeline and a null consumer drains their tuple packs: stage 1 + stage 2, still without the entropy coders.
// With COLORD_TIME_STAGES=12q the quality entropy coder thread (CEntrComprQuals, compression.cpp:654-668) runs as well,
// writing its parts to the archive given on the command line: stages 1 + 2 + the quality stream of stage 3; with
// COLORD_TIME_STAGES=12qd the DNA entropy coder thread (CEntrComprReads, compression.cpp:629-647) runs too (everything but
// the header coder); with COLORD_TIME_STAGES=12qdh the header coder thread (CEntrComprHeaders, compression.cpp:670-689) runs as
// well: all three compute stages of the compression path.
//
// Usage: oracle/_ref/ref_stage1_time compress-ont [flags] -t N in.fastq ignored.out
#include "compression.h"
#include "utils.h"
#include "params.h"
#include "count_kmers.h"
#include "kmer_filter.h"
#include "in_reads.h"
#include "reads_sim_graph.h"
#include "encoder.h"
#include "entr_qual.h"
#include "entr_read.h"
#include "entr_header.h"
#include "archive.h"
#include "reference_reads.h"
#include "ref_reads_accepter.h"
#include "parallel_queue.h"
#include "queues_data.h"
#include <chrono>
#include <filesystem>
#include <thread>
#include <cstdio>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

void runCompression(const CCompressorParams& params, CInfo& info)
{
	if (params.refGenomePath != "") { std::cerr << "ref_stage1_time: -G not supported\n"; exit(1); }
	bool is_gzip_input = izGzipFile(params.inputFilePath);
	bool is_fastq = isFastq(params.inputFilePath);
	int n_compression_threads = std::max(1, (int)params.nThreads - 3) + 2;
	uint32_t kmerLen = params.kmerLen, anchorLen = params.anchorLen;
	if (!kmerLen || !anchorLen)
	{	// compression.cpp:42-94
		uint64_t bytes = std::filesystem::file_size(params.inputFilePath);
		double factor = is_gzip_input ? (is_fastq ? 2.08 : 3.98) : (is_fastq ? 0.49 : 0.98);
		uint64_t bases = static_cast<uint64_t>(factor * bytes);
		struct { uint64_t lim; uint32_t k, a; } tab[] = { {1'000'000'000ull, 20, 16}, {4'000'000'000ull, 21, 18}, {16'000'000'000ull, 23, 21},
			{48'000'000'000ull, 24, 22}, {128'000'000'000ull, 25, 22}, {~0ull, 26, 23} };
		for (auto& t : tab) if (bases < t.lim) { kmerLen = t.k; anchorLen = t.a; break; }
	}
	auto tmp_dir_path = create_tmp_dir(std::filesystem::path(params.outputFilePath).parent_path().string() + "/");
	std::string kmersDbPath = (std::filesystem::path(tmp_dir_path) / "db").string();

	double t0 = now();
	CKmerCounter kmer_counter(kmerLen, params.minKmerCount, params.maxKmerCount, params.nThreads, params.filterHashModulo,
		params.inputFilePath, kmersDbPath, tmp_dir_path, is_fastq, false);
	double t1 = now();
	auto tot_n_reads = kmer_counter.GetNReads();
	auto tot_kmers = kmer_counter.GetTotKmers();
	auto n_uniq = kmer_counter.GetNUniqueCounted();
	uint64_t mean_read_len = static_cast<uint64_t>((double(tot_kmers * params.filterHashModulo) / tot_n_reads + kmerLen - 1));
	CKmerFilter filtered_kmers(kmersDbPath, params.filterHashModulo, kmerLen, n_uniq, params.fillFactorFilteredKmers, false);
	double t2 = now();
	std::error_code ec; std::filesystem::remove_all(tmp_dir_path, ec);

	uint32_t sparse_range = static_cast<uint32_t>((params.sparseMode_range_symbols * n_uniq * params.filterHashModulo) / mean_read_len);
	if (!sparse_range) sparse_range = 1;
	CRefReadsAccepter accepter(sparse_range, params.sparseMode_exponent, 0);
	uint32_t tot_ref_reads = tot_n_reads;
	if (params.referenceReadsMode == ReferenceReadsMode::Sparse)
		tot_ref_reads = accepter.GetNAccepted(tot_n_reads);

	CQueueMonitor qm(std::cerr, false, true);
	CParallelQueue<read_pack_t> reads_queue(reads_queue_size, 1, &qm, 0);
	CParallelQueue<qual_pack_t> quals_queue(quals_queue_size, 1, &qm, 1);
	CParallelQueue<header_pack_t> headers_queue(headers_queue_size, 1, &qm, 2);
	CParallelQueuePopWaiting<CCompressPack> graph_out(compress_queue_size, &qm, 4);
	CReferenceReads reference_reads(tot_ref_reads);

	double t3 = now();
	uint64_t n_links = 0, n_out = 0;
	std::thread reader([&] { CInputReads r(false, params.inputFilePath, reads_queue, quals_queue, headers_queue); });
	const char* stages_env0 = getenv("COLORD_TIME_STAGES");
	const bool hdr_consumed = stages_env0 && std::string(stages_env0) == "12qdh";
	const bool qual_consumed = hdr_consumed || (stages_env0 && (std::string(stages_env0) == "12q" || std::string(stages_env0) == "12qd"));
	std::thread drain_q([&] { if (qual_consumed) return; qual_pack_t p; while (quals_queue.Pop(p)); });
	CArchive archive(false);
	if (qual_consumed && !archive.Open(params.outputFilePath)) { std::cerr << "cannot open " << params.outputFilePath << "\n"; exit(1); }
	std::thread drain_h([&] {
		if (!hdr_consumed) { header_pack_t p; while (headers_queue.Pop(p)); return; }
		CEntrComprHeaders compr{ headers_queue, archive, params.headerComprMode, params.compressionLevel, false };
		compr.Compress();
	});
	std::thread graph([&] {
		CReadsSimilarityGraph g(reads_queue, graph_out, reference_reads, nullptr, filtered_kmers, kmerLen, params.maxCandidates,
			params.maxKmerCount, params.referenceReadsMode, accepter, (double)tot_ref_reads / tot_n_reads, n_compression_threads,
			params.dataSource, params.fillFactorKmersToReads, false);
	});
	const char* stages_env = getenv("COLORD_TIME_STAGES");
	const bool with_dna = stages_env && (std::string(stages_env) == "12qd" || std::string(stages_env) == "12qdh");
	const bool with_qual = with_dna || (stages_env && std::string(stages_env) == "12q");
	const bool with_encoders = with_qual || (stages_env && std::string(stages_env) == "12");
	uint64_t es_bytes = 0;
	if (!with_encoders)
	{
		std::thread sink([&] { CCompressPack pack; while (graph_out.Pop(pack)) for (auto& e : pack.data) { ++n_out; n_links += e.ref_reads.size(); } });
		reader.join(); drain_q.join(); drain_h.join(); graph.join(); sink.join();
	}
	else
	{
		CParallelPriorityQueue<std::vector<es_t>> es_for_qual(2 * n_compression_threads, n_compression_threads, &qm, 3);
		CParallelPriorityQueue<std::vector<es_t>> compressed(2 * n_compression_threads, n_compression_threads, &qm, 5);
		std::vector<std::thread> encoders;
		for (int i = 0; i < n_compression_threads; ++i)
			encoders.emplace_back([&] {
				CEncoder enc(false, graph_out, reference_reads, compressed, es_for_qual, anchorLen,
					params.minFractionOfMmersInEncodeToAlwaysEncode, params.minFractionOfMmersInEncode, params.maxMatchesMultiplier,
					params.editScriptCostMultiplier, params.minPartLenToConsiderAltRead, params.maxRecurence, params.minAnchors,
					is_fastq, params.filterHashModulo, kmerLen, params.dataSource);
				enc.Encode();
			});
		std::thread drain_esq([&] {
			if (!with_qual) { std::vector<es_t> p; while (es_for_qual.Pop(p)); return; }
			CEntrComprQuals compr{ quals_queue, archive, params.qualityComprMode, params.qualityFwdThresholds, params.qualityRevThresholds, false,
				params.compressionLevel, (uint64_t)tot_n_reads * mean_read_len, es_for_qual, params.dataSource };
			compr.Compress();
		});
		std::thread sink([&] {
			if (!with_dna) { std::vector<es_t> pack; while (compressed.Pop(pack)) for (auto& es : pack) { ++n_out; es_bytes += es.size(); } return; }
			CEntrComprReads compr{ compressed, reference_reads, false, params.maxCandidates, params.compressionLevel, (uint64_t)tot_n_reads * mean_read_len, archive, tot_n_reads, 0 };
			compr.Compress();
		});
		reader.join(); drain_q.join(); drain_h.join(); graph.join();
		for (auto& t : encoders) t.join();
		drain_esq.join(); sink.join();
	}
	double t4 = now();
	printf("{\"count_s\": %.4f, \"filter_s\": %.4f, \"graph_s\": %.4f, \"stage1_s\": %.4f, \"k\": %u, \"n_reads\": %u, \"tot_kmers\": %llu, "
		"\"n_unique_counted\": %llu, \"tot_ref_reads\": %u, \"n_links\": %llu, \"threads\": %u, \"stages\": \"%s\", \"anchor_len\": %u, \"es_bytes\": %llu, \"n_out\": %llu}\n",
		t1 - t0, t2 - t1, t4 - t3, (t2 - t0) + (t4 - t3), kmerLen, tot_n_reads, (unsigned long long)tot_kmers, (unsigned long long)n_uniq,
		tot_ref_reads, (unsigned long long)n_links, params.nThreads, hdr_consumed ? "1+2+3" : with_dna ? "1+2+3qd" : with_qual ? "1+2+3q" : with_encoders ? "1+2" : "1", anchorLen, (unsigned long long)es_bytes, (unsigned long long)n_out);
	fflush(stdout);
	_exit(0);     // skip archive/info epilogue of the CLI callback
}
