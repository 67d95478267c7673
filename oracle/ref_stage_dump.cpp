// oracle/ref_stage_dump.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// A tapped build of the *unmodified* reference: every reference object is linked as-is except
// compression.o, whose single entry point `runCompression` (compression.h:22, called from the CLI
// callback in arg_parse.cpp) is provided here.  This version wires the very same reference classes
// in the same order as compression.cpp:344-703 but, instead of entropy-coding into an archive, it
// writes what crosses each stage seam to $COLORD_DUMP_DIR so that the CPU restatement (oracle/*.c)
// and the CUDA path can be pinned per stage:
//
//   params.txt   derived parameters + KMC statistics (compression.cpp:393-542)
//   kmers.bin    the filtered k-mer DB listed through kmc_api, as (kmer u64, count u32), KMC order
//   reads.bin    per read: id, hasN, sampler decision, accepted k-mers (reads_sim_graph.cpp:134-164
//                recomputed here with CKmerWalker + CKmerFilter), candidate ref ids and (HiFi) the
//                shared k-mers exactly as CReadsSimilarityGraph pushed them (queues_data.h:23)
//   es.bin       per read: the CompactES byte string CEncoder emitted (utils.h:69-273)
//
// Usage: COLORD_DUMP_DIR=dir oracle/_ref/ref_stage_dump compress-ont [flags] in.fastq ignored.out
#include "compression.h"
#include "utils.h"
#include "params.h"
#include "count_kmers.h"
#include "kmer_filter.h"
#include "in_reads.h"
#include "reads_sim_graph.h"
#include "encoder.h"
#include "reference_reads.h"
#include "ref_reads_accepter.h"
#include "parallel_queue.h"
#include "queues_data.h"
#include "kmc_file.h"
#include <filesystem>
#include <fstream>
#include <thread>
#include <unordered_set>
#include <cstdlib>

namespace {

template <typename T> void put(std::ofstream& o, T v) { o.write(reinterpret_cast<const char*>(&v), sizeof(T)); }

// Same size-derived defaults as compression.cpp:42-94 (plain FASTQ/FASTA and gz factors).
void derive_k_and_anchor(uint32_t& k, uint32_t& a, bool gz, bool fastq, const std::string& path)
{
	if (k && a) return;
	uint64_t bytes = std::filesystem::file_size(path);
	double factor = gz ? (fastq ? 2.08 : 3.98) : (fastq ? 0.49 : 0.98);
	uint64_t bases = static_cast<uint64_t>(factor * bytes);
	struct { uint64_t lim; uint32_t k, a; } tab[] = {
		{1'000'000'000ull, 20, 16}, {4'000'000'000ull, 21, 18}, {16'000'000'000ull, 23, 21},
		{48'000'000'000ull, 24, 22}, {128'000'000'000ull, 25, 22}, {~0ull, 26, 23} };
	for (auto& t : tab) if (bases < t.lim) { k = t.k; a = t.a; return; }
}

std::vector<uint8_t> es_bytes(es_t& es)
{
	// Re-serialise through the public reader into the documented CompactES layout (utils.h:79-99).
	std::vector<uint8_t> out;
	es.restart_reading();
	tuple_types type; uint32_t v1 = 0, v2 = 0;
	auto p32 = [&](uint32_t v) { out.push_back(v >> 24); out.push_back((v >> 16) & 0xff); out.push_back((v >> 8) & 0xff); out.push_back(v & 0xff); };
	while (es.load(type, v1, v2))
	{
		uint8_t t = static_cast<uint8_t>(type);
		switch (type)
		{
		case tuple_types::insertion: case tuple_types::substitution: case tuple_types::plain:
			out.push_back((t << 4) + v1); break;
		case tuple_types::anchor: case tuple_types::skip:
			out.push_back((t << 4) + (v2 >> 24)); out.push_back((v2 >> 16) & 0xff); out.push_back((v2 >> 8) & 0xff); out.push_back(v2 & 0xff); break;
		case tuple_types::alt_id: case tuple_types::start_es:
			out.push_back((t << 4) + v2); p32(v1); break;
		default:
			out.push_back(t << 4);
		}
	}
	return out;
}

} // namespace

void runCompression(const CCompressorParams& params, CInfo& info)
{
	const char* dd = std::getenv("COLORD_DUMP_DIR");
	if (!dd) { std::cerr << "COLORD_DUMP_DIR not set\n"; exit(1); }
	std::filesystem::path dump_dir(dd);
	std::filesystem::create_directories(dump_dir);
	if (params.refGenomePath != "") { std::cerr << "ref_stage_dump: -G not supported\n"; exit(1); }

	bool is_gzip_input = izGzipFile(params.inputFilePath);
	bool is_fastq = isFastq(params.inputFilePath);
	int n_compression_threads = std::max(1, (int)params.nThreads - 3) + 2;

	uint32_t kmerLen = params.kmerLen, anchorLen = params.anchorLen;
	derive_k_and_anchor(kmerLen, anchorLen, is_gzip_input, is_fastq, params.inputFilePath);

	auto tmp_dir_path = create_tmp_dir(dump_dir.string() + "/");
	std::string kmersDbPath = (std::filesystem::path(tmp_dir_path) / "db").string();

	CKmerCounter kmer_counter(kmerLen, params.minKmerCount, params.maxKmerCount, params.nThreads, params.filterHashModulo,
		params.inputFilePath, kmersDbPath, tmp_dir_path, is_fastq, false);
	auto tot_n_reads = kmer_counter.GetNReads();
	auto tot_kmers = kmer_counter.GetTotKmers();
	auto n_uniq = kmer_counter.GetNUniqueCounted();
	uint64_t mean_read_len = static_cast<uint64_t>((double(tot_kmers * params.filterHashModulo) / tot_n_reads + kmerLen - 1));

	{	// list the DB exactly the way CKmerFilter's ctor does (filter_kmers.cpp:52-83)
		CKMCFile f;
		if (!f.OpenForListing(kmersDbPath)) { std::cerr << "cannot list kmc db\n"; exit(1); }
		CKmerAPI kmer(f.KmerLength());
		uint32_t count; std::vector<uint64> v;
		std::ofstream o(dump_dir / "kmers.bin", std::ios::binary);
		CKMCFileInfo fi; f.Info(fi);
		put<uint64_t>(o, fi.total_kmers);
		while (f.ReadNextKmer(kmer, count)) { kmer.to_long(v); put<uint64_t>(o, v.back()); put<uint32_t>(o, count); }
		f.Close();
	}

	CKmerFilter filtered_kmers(kmersDbPath, params.filterHashModulo, kmerLen, n_uniq, params.fillFactorFilteredKmers, false);
	std::error_code ec; std::filesystem::remove_all(tmp_dir_path, ec);

	uint32_t sparse_range = static_cast<uint32_t>((params.sparseMode_range_symbols * n_uniq * params.filterHashModulo) / mean_read_len);
	if (!sparse_range) sparse_range = 1;
	CRefReadsAccepter accepter(sparse_range, params.sparseMode_exponent, 0);
	CRefReadsAccepter accepter_replay(sparse_range, params.sparseMode_exponent, 0);
	uint32_t tot_ref_reads = tot_n_reads;
	if (params.referenceReadsMode == ReferenceReadsMode::Sparse)
		tot_ref_reads = accepter.GetNAccepted(tot_n_reads);

	{
		std::ofstream p(dump_dir / "params.txt");
		p << "k=" << kmerLen << "\nanchor_len=" << anchorLen << "\nmodulo=" << params.filterHashModulo
			<< "\nmin_count=" << params.minKmerCount << "\nmax_count=" << params.maxKmerCount
			<< "\nmax_candidates=" << params.maxCandidates << "\nlevel=" << params.compressionLevel
			<< "\nsparse=" << (params.referenceReadsMode == ReferenceReadsMode::Sparse)
			<< "\nsparse_range=" << sparse_range << "\nsparse_exponent=" << params.sparseMode_exponent
			<< "\nhifi=" << (params.dataSource == DataSource::PBHiFi)
			<< "\nn_reads=" << tot_n_reads << "\ntot_kmers=" << tot_kmers << "\nn_unique_counted=" << n_uniq
			<< "\ntotal_count_filtered=" << filtered_kmers.GetTotalKmers()
			<< "\nmean_read_len=" << mean_read_len << "\ntot_ref_reads=" << tot_ref_reads
			<< "\nmin_part_len_alt=" << params.minPartLenToConsiderAltRead << "\nmax_recurence=" << params.maxRecurence
			<< "\nmin_anchors=" << params.minAnchors << "\nes_cost_mult=" << params.editScriptCostMultiplier
			<< "\nmin_mmer_frac=" << params.minFractionOfMmersInEncode << "\nmin_mmer_force=" << params.minFractionOfMmersInEncodeToAlwaysEncode
			<< "\nmax_matches_mult=" << params.maxMatchesMultiplier << "\n";
	}

	CQueueMonitor qm(std::cerr, false, true);
	CParallelQueue<read_pack_t> reads_queue(reads_queue_size, 1, &qm, 0);
	CParallelQueue<qual_pack_t> quals_queue(quals_queue_size, 1, &qm, 1);
	CParallelQueue<header_pack_t> headers_queue(headers_queue_size, 1, &qm, 2);
	CParallelQueuePopWaiting<CCompressPack> graph_out(compress_queue_size, &qm, 4);
	CParallelQueuePopWaiting<CCompressPack> encoder_in(compress_queue_size, &qm, 4);
	CParallelPriorityQueue<std::vector<es_t>> es_for_qual(2 * n_compression_threads, n_compression_threads, &qm, 3);
	CParallelPriorityQueue<std::vector<es_t>> compressed(2 * n_compression_threads, n_compression_threads, &qm, 5);
	CReferenceReads reference_reads(tot_ref_reads);

	std::thread reader([&] { CInputReads r(false, params.inputFilePath, reads_queue, quals_queue, headers_queue); });
	std::thread drain_q([&] { qual_pack_t p; while (quals_queue.Pop(p)); });
	std::thread drain_h([&] { header_pack_t p; while (headers_queue.Pop(p)); });
	std::thread graph([&] {
		CReadsSimilarityGraph g(reads_queue, graph_out, reference_reads, nullptr, filtered_kmers, kmerLen, params.maxCandidates,
			params.maxKmerCount, params.referenceReadsMode, accepter, (double)tot_ref_reads / tot_n_reads, n_compression_threads,
			params.dataSource, params.fillFactorKmersToReads, false);
	});

	std::thread tap1([&] {
		std::ofstream o(dump_dir / "reads.bin", std::ios::binary);
		CCompressPack pack;
		while (graph_out.Pop(pack))
		{
			put<uint32_t>(o, 0xFFFFFFFFu); put<uint32_t>(o, pack.id); put<uint32_t>(o, (uint32_t)pack.data.size());
			for (auto& e : pack.data)
			{
				bool is_ref = !e.hasN;
				if (params.referenceReadsMode == ReferenceReadsMode::Sparse)
					is_ref &= accepter_replay.ShouldAddToReference(e.read_id);
				std::vector<kmer_type> acc;
				if (!e.hasN && read_len(e.read) >= kmerLen)
				{
					kmer_type kmer; CKmerWalker w(e.read, kmerLen, kmer);
					std::unordered_set<kmer_type> seen;
					while (w.NextKmer())
					{
						if (!filtered_kmers.Possible(kmer) || !seen.insert(kmer).second) continue;
						if (filtered_kmers.Check(kmer)) acc.push_back(kmer);
					}
				}
				put<uint32_t>(o, e.read_id); put<uint8_t>(o, e.hasN); put<uint8_t>(o, is_ref);
				put<uint32_t>(o, (uint32_t)read_len(e.read));
				put<uint32_t>(o, (uint32_t)acc.size()); for (auto k : acc) put<uint64_t>(o, k);
				put<uint32_t>(o, (uint32_t)e.ref_reads.size()); for (auto r : e.ref_reads) put<uint32_t>(o, r);
				put<uint32_t>(o, (uint32_t)e.common_kmers.size());
				for (auto& v : e.common_kmers) { put<uint32_t>(o, (uint32_t)v.size()); for (auto k : v) put<uint64_t>(o, k); }
			}
			encoder_in.Push(std::move(pack));
		}
		encoder_in.MarkCompleted();
	});

	std::vector<std::thread> encoders;
	for (int i = 0; i < n_compression_threads; ++i)
		encoders.emplace_back([&] {
			CEncoder enc(false, encoder_in, reference_reads, compressed, es_for_qual, anchorLen,
				params.minFractionOfMmersInEncodeToAlwaysEncode, params.minFractionOfMmersInEncode, params.maxMatchesMultiplier,
				params.editScriptCostMultiplier, params.minPartLenToConsiderAltRead, params.maxRecurence, params.minAnchors,
				is_fastq, params.filterHashModulo, kmerLen, params.dataSource);
			enc.Encode();
		});
	std::thread drain_esq([&] { std::vector<es_t> p; while (es_for_qual.Pop(p)); });
	std::thread tap2([&] {
		std::ofstream o(dump_dir / "es.bin", std::ios::binary);
		std::vector<es_t> pack;
		while (compressed.Pop(pack))
		{
			put<uint32_t>(o, 0xFFFFFFFFu); put<uint32_t>(o, (uint32_t)pack.size());
			for (auto& es : pack) { auto b = es_bytes(es); put<uint32_t>(o, (uint32_t)b.size()); o.write((const char*)b.data(), b.size()); }
		}
	});

	reader.join(); drain_q.join(); drain_h.join(); graph.join(); tap1.join();
	for (auto& t : encoders) t.join();
	drain_esq.join(); tap2.join();
	std::cerr << "\nref_stage_dump: wrote " << dump_dir << "\n";
}
