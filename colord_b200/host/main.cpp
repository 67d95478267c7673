// colord-b200 — command-line front end over the host-side mirrors (compressor.h / decompressor.h / archive_host.h).
// Same sub-commands and option letters as the reference CLI (src/colord/arg_parse.cpp:463-545, :820-902):
//   colord-b200 compress-ont | compress-pbhifi | compress-pbraw [options] <input FASTQ/FASTA(.gz)> <archive>
//   colord-b200 decompress <archive> <output>
//   colord-b200 info <archive>
// Exit code 1 with a message on stderr for every error, as the reference does.  Options the device path does not implement
// (-G / -s, the threshold quality modes) are refused with a message instead of being ignored.
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <sstream>
#include <thread>
#include "compressor.h"
#include "compressor_mgpu.h"
#include "decompressor.h"

using namespace clbhost;

static void usage()
{
	std::cerr <<
		"colord-b200 (archive format " << B200_VERSION_MAJOR << "." << B200_VERSION_MINOR << "." << B200_VERSION_PATCH << ")\n"
		"  colord-b200 compress-ont|compress-pbhifi|compress-pbraw [options] input output\n"
		"      -k,--kmer-len N  -a,--anchor-len N  -p,--priority ratio|balanced|memory  -q,--qual org|2-avg|4-avg|5-avg|2-fix|4-fix|5-fix|avg|none\n"
		"      -T,--qual-thresholds a,b,..  -L,--Lowest-count N  -H,--Highest-count N  -f,--filter-modulo N  -c,--max-candidates N\n"
		"      -e,--edit-script-mult X  -r,--max-recurence-level N  --min-to-alt N  --min-mmer-frac X  --min-mmer-force-enc X\n"
		"      --max-matches-mult X  --min-anchors N  -R,--reference-reads-mode all|sparse  -g,--sparse-range X  -x,--sparse-exponent X\n"
		"      -i,--identifier org|main|none  -t,--threads N (accepted, unused)  -v,--verbose  --device N\n"
		"      --gpus N   shard the input over N GPUs of this box (devices --device .. --device + N - 1; plain FASTQ, native containers)\n"
		"      --compat | --native   streams of the archive: the reference's own (readable by `colord decompress`, every -q mode) or the\n"
		"                            device's containers (org, *-avg, none); default: --compat up to --compat-max-mbases N (512) input Mbases\n"
		"      -G,--reference-genome file  -s,--store-reference\n"
		"  colord-b200 decompress [-G,--reference-genome file] archive output\n"
		"  colord-b200 info archive\n";
}

static std::vector<uint32_t> parse_list(const std::string& s)
{
	std::vector<uint32_t> v; std::stringstream ss(s); std::string item;
	while (std::getline(ss, item, ',')) if (!item.empty()) v.push_back(static_cast<uint32_t>(std::stoul(item)));
	return v;
}

static int run_compress(const std::string& cmd, int argc, char** argv, const std::string& full_cmd)
{
	// the priority decides the default set, so it is looked up first (arg_parse.cpp:787-808)
	std::string pri = "memory";
	for (int i = 2; i + 1 < argc; ++i) if (!std::strcmp(argv[i], "-p") || !std::strcmp(argv[i], "--priority")) pri = argv[i + 1];
	CCompressorParams p = defaultParams(dataSourceFromCommand(cmd), compressionPriorityFromString(pri));
	p.nThreads = std::max(std::thread::hardware_concurrency(), 1u);          // arg_parse.cpp:105
	std::vector<std::string> pos;
	bool qual_set = false; std::vector<uint32_t> fwd_user; uint32_t n_gpus = 1;
	for (int i = 2; i < argc; ++i) {
		const std::string a = argv[i];
		auto need = [&]() -> std::string { if (i + 1 >= argc) throw std::invalid_argument("option " + a + " needs a value"); return argv[++i]; };
		if (a == "-p" || a == "--priority") need();
		else if (a == "-k" || a == "--kmer-len") { p.kmerLen = std::stoul(need()); if (p.kmerLen < 15 || p.kmerLen > 28) throw std::invalid_argument("-k must be in 15..28"); }
		else if (a == "-a" || a == "--anchor-len") p.anchorLen = std::stoul(need());
		else if (a == "-t" || a == "--threads") p.nThreads = std::stoul(need());
		else if (a == "-q" || a == "--qual") { p.qualityComprMode = qualityComprModeFromString(need()); qual_set = true; }
		else if (a == "-T" || a == "--qual-thresholds") {      // the reference's form `-T 5 12 20` (arg_parse.cpp:488-500) or one comma-separated list
			fwd_user = parse_list(need());
			while (i + 1 < argc && argv[i + 1][0] && std::string(argv[i + 1]).find_first_not_of("0123456789") == std::string::npos) fwd_user.push_back(static_cast<uint32_t>(std::stoul(argv[++i])));
		}
		else if (a == "-L" || a == "--Lowest-count") p.minKmerCount = std::stoul(need());
		else if (a == "-H" || a == "--Highest-count") p.maxKmerCount = std::stoul(need());
		else if (a == "-f" || a == "--filter-modulo") p.filterHashModulo = std::stoul(need());
		else if (a == "-c" || a == "--max-candidates") p.maxCandidates = std::stoul(need());
		else if (a == "-e" || a == "--edit-script-mult") p.editScriptCostMultiplier = std::stod(need());
		else if (a == "-r" || a == "--max-recurence-level") p.maxRecurence = std::stoul(need());
		else if (a == "--min-to-alt") p.minPartLenToConsiderAltRead = std::stoul(need());
		else if (a == "--min-mmer-frac") p.minFractionOfMmersInEncode = std::stod(need());
		else if (a == "--min-mmer-force-enc") p.minFractionOfMmersInEncodeToAlwaysEncode = std::stod(need());
		else if (a == "--max-matches-mult") p.maxMatchesMultiplier = std::stod(need());
		else if (a == "--min-anchors") p.minAnchors = std::stoul(need());
		else if (a == "-R" || a == "--reference-reads-mode") { const std::string m = need(); if (m != "all" && m != "sparse") throw std::invalid_argument("-R takes all or sparse"); p.referenceReadsMode = m == "all" ? ReferenceReadsMode::All : ReferenceReadsMode::Sparse; }
		else if (a == "-g" || a == "--sparse-range") p.sparseMode_range_symbols = std::stod(need());
		else if (a == "-x" || a == "--sparse-exponent") p.sparseMode_exponent = std::stod(need());
		else if (a == "-i" || a == "--identifier") { const std::string m = need(); if (m != "org" && m != "main" && m != "none") throw std::invalid_argument("-i takes org, main or none"); p.headerComprMode = m == "org" ? HeaderComprMode::Original : m == "main" ? HeaderComprMode::Main : HeaderComprMode::None; }
		else if (a == "-G" || a == "--reference-genome") p.refGenomePath = need();
		else if (a == "-s" || a == "--store-reference") p.storeRefGenome = true;
		else if (a == "-v" || a == "--verbose") p.verbose = true;
		else if (a == "--device") p.device = std::stoi(need());
		else if (a == "--gpus") n_gpus = static_cast<uint32_t>(std::stoul(need()));
		else if (a == "--compat") p.streamFormat = StreamFormat::Compat;
		else if (a == "--native") p.streamFormat = StreamFormat::Native;
		else if (a == "--compat-max-mbases") p.compat_max_bases = std::stoull(need()) << 20;
		else if (!a.empty() && a[0] == '-') throw std::invalid_argument("unknown option " + a);
		else pos.push_back(a);
	}
	if (pos.size() != 2) throw std::invalid_argument("expected: input output");
	p.inputFilePath = pos[0]; p.outputFilePath = pos[1];
	if (qual_set) defaultQualityThresholds(p.qualityComprMode, p.qualityFwdThresholds, p.qualityRevThresholds);
	if (!fwd_user.empty()) { const size_t want = p.qualityFwdThresholds.size(); if (fwd_user.size() < want) throw std::invalid_argument("too few quality thresholds for this mode"); fwd_user.resize(want); p.qualityFwdThresholds = fwd_user; }
	// CUDA initialises every GPU it can see when a process makes its first context (6.6 s on an 8-GPU box against 1.5 - 2 s with one
	// GPU visible): unless the caller has set it, the process is shown only the GPUs it is going to use
	if (!std::getenv("CUDA_VISIBLE_DEVICES")) {
		std::string vis;
		for (uint32_t g = 0; g < std::max(1u, n_gpus); ++g) vis += (g ? "," : "") + std::to_string(p.device + static_cast<int>(g));
		setenv("CUDA_VISIBLE_DEVICES", vis.c_str(), 1);
		p.device = 0;
	}
	CInfo info; info.full_command_line = full_cmd;
	const CompressionReport r = n_gpus > 1 ? runCompressionMultiGpu(p, info, n_gpus) : runCompression(p, info);
	if (p.verbose) { std::cerr << "streams: " << (r.compat ? "compat (the reference's own)" : "native containers") << "\ninput: " << (r.streamed ? "streamed to the device in pieces" : "read whole") << ", " << r.reader_threads << " reader thread(s)\n"; for (const Phase& ph : r.phases) std::cerr << "  phase " << ph.name << ": " << ph.seconds << " s\n"; }
	if (p.verbose && r.has_encode_stats) print_stats_report(std::cerr, r.read_stats, r.encode_stats);      // as the reference does when it exits (stats_collector.cpp:100-146)
	if (p.verbose) std::cerr << "k-mer length: " << r.kmerLen << "\nanchor length: " << r.anchorLen << "\nsparse mode range in reads: " << r.sparse_range << "\nreference reads: " << r.tot_ref_reads << "\n";
	// compression.cpp:802-806
	std::cerr << "DNA size        : " << r.dna << "\nQuality size    : " << r.qual << "\nHeader size     : " << r.header << "\nMeta size       : " << r.meta << "\nInfo size       : " << r.info << "\n";
	std::cerr << "Total time      : " << r.seconds << "s\n";
	return 0;
}

static int run_info(const std::string& path)
{
	CArchive archive(true);
	if (!archive.Open(path)) { std::cerr << "Error: cannot open archive: " << path << "\n"; return 1; }
	const int s_info = archive.GetStreamId("info");
	std::vector<uint8_t> raw; size_t md;
	if (s_info < 0 || !archive.GetPart(s_info, raw, md)) { std::cerr << "Error: the archive has no info record\n"; return 1; }
	CInfo info; info.Deserialize(raw);
	// info.cpp:43-53
	std::cerr << "version major: " << info.version_major << "\nversion minor: " << info.version_minor << "\nversion patch: " << info.version_patch << "\n";
	std::cerr << "total bytes: " << info.total_bytes << "\ntotal bases: " << info.total_bases << "\ntotal reads: " << info.total_reads << "\n";
	std::cerr << "time: " << CInfo::time_string(info.time) << "\n" << "command: " << info.full_command_line << "\n";
	return 0;
}

int main(int argc, char** argv)
{
	if (argc < 2) { usage(); return 1; }
	std::string full_cmd;
	for (int i = 0; i < argc; ++i) { if (i) full_cmd += ' '; full_cmd += argv[i]; }
	const std::string cmd = argv[1];
	try {
		if (cmd == "compress-ont" || cmd == "compress-pbhifi" || cmd == "compress-pbraw") return run_compress(cmd, argc, argv, full_cmd);
		if (cmd == "info") { if (argc != 3) { usage(); return 1; } return run_info(argv[2]); }
		if (cmd == "decompress") {
			std::vector<std::string> pos; bool verbose = false; std::string genome;
			for (int i = 2; i < argc; ++i) { const std::string a = argv[i]; if (a == "-v" || a == "--verbose") verbose = true; else if (a == "-G" || a == "--reference-genome") { if (i + 1 >= argc) throw std::invalid_argument("option -G needs a value"); genome = argv[++i]; } else pos.push_back(a); }
			if (pos.size() != 2) { usage(); return 1; }
			runDecompression(pos[0], pos[1], verbose, genome);
			return 0;
		}
		usage();
		return 1;
	} catch (const std::exception& e) {
		std::cerr << e.what() << "\n";
		return 1;
	}
}
