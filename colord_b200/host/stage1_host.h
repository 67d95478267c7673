// stage1_host.h — C++ host-side mirrors of the reference's stage-1 classes over the C-ABI
// (include/colord_b200.h).  Same names of accessors, argument meaning and order of use as
//   CKmerCounter            src/colord/count_kmers.h:24-34
//   CKmerFilter             src/colord/kmer_filter.h:117-199
//   CRefReadsAccepter       src/colord/ref_reads_accepter.h:27-57
//   CReadsSimilarityGraph   src/colord/reads_sim_graph.h:170 (the ctor is the loop) / CCompressElem queues_data.h:23
// but fed with read packs (the reader thread's output) instead of a file path, and throwing
// std::runtime_error where the reference prints and exit(1)s.  Header-only; link -lcolord_b200.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../include/colord_b200.h"

namespace clbhost {

inline void check(clb_ctx* ctx, clb_status st, const char* what)
{
	if (st != CLB_OK) throw std::runtime_error(std::string(what) + ": " + clb_last_error(ctx));
}

// One read pack as the reader thread produces it: concatenated ASCII bases + offsets (read_pack_t, utils.h:46).
struct ReadPack {
	std::vector<uint8_t> bases;
	std::vector<uint64_t> offsets{0};
	void add(const uint8_t* seq, size_t len) { bases.insert(bases.end(), seq, seq + len); offsets.push_back(bases.size()); }
	uint32_t size() const { return static_cast<uint32_t>(offsets.size() - 1); }
};

class CKmerCounter {
	clb_ctx* ctx = nullptr;
	clb_kmer_stats st{};
	bool finalized = false;
public:
	// (k, ci, cs, modulo) as in the reference ctor; n_threads/paths disappear, max_candidates/is_hifi are needed later by the graph
	CKmerCounter(uint32_t k, uint32_t ci, uint32_t cs, uint32_t modulo, uint32_t max_candidates, bool is_hifi, uint64_t expected_bases = 0, int device = 0)
	{
		clb_params p{k, modulo, ci, cs, max_candidates, is_hifi ? 1u : 0u, expected_bases, device};
		clb_status s = clb_create(&p, &ctx);
		if (s != CLB_OK) throw std::runtime_error(std::string("clb_create: ") + clb_last_error(nullptr));
	}
	~CKmerCounter() { clb_destroy(ctx); }
	CKmerCounter(const CKmerCounter&) = delete;
	CKmerCounter& operator=(const CKmerCounter&) = delete;
	void AddPack(const ReadPack& pack) { check(ctx, clb_append_reads(ctx, pack.bases.data(), pack.offsets.data(), pack.size(), 0), "clb_append_reads"); }
	void Finalize() { if (!finalized) { check(ctx, clb_count_finalize(ctx, &st), "clb_count_finalize"); finalized = true; } }
	uint32_t GetNReads() { Finalize(); return static_cast<uint32_t>(st.n_reads); }
	uint64_t GetTotKmers() { Finalize(); return st.tot_kmers; }
	uint64_t GetNUniqueCounted() { Finalize(); return st.n_unique_counted; }
	clb_ctx* Context() { return ctx; }
};

class CKmerFilter {
	clb_ctx* ctx;
	uint64_t total;
public:
	explicit CKmerFilter(CKmerCounter& counter) : ctx(counter.Context())
	{
		counter.Finalize();
		clb_kmer_stats st; check(ctx, clb_count_finalize(ctx, &st), "clb_count_finalize");
		total = st.total_count_filtered;
	}
	uint64_t GetTotalKmers() const { return total; }            // kmer_filter.h:140
	// batch forms of Possible (:129) and Check (:135)
	void PossibleAndCheck(const std::vector<uint64_t>& kmers, std::vector<uint8_t>& possible, std::vector<uint8_t>& present) const
	{
		possible.assign(kmers.size(), 0); present.assign(kmers.size(), 0);
		check(ctx, clb_filter_check(ctx, kmers.data(), kmers.size(), possible.data(), present.data()), "clb_filter_check");
	}
	void List(std::vector<uint64_t>& kmers, std::vector<uint32_t>& counts) const
	{
		uint64_t n = 0;
		clb_filter_list(ctx, nullptr, nullptr, 0, &n, 0);
		kmers.assign(n, 0); counts.assign(n, 0);
		if (n) check(ctx, clb_filter_list(ctx, kmers.data(), counts.data(), n, &n, 0), "clb_filter_list");
	}
};

class CRefReadsAccepter {
	uint32_t range; double exponent; uint32_t n_pseudo;
public:
	CRefReadsAccepter(uint32_t range, double exponent, uint32_t n_pseudo) : range(range), exponent(exponent), n_pseudo(n_pseudo) {}
	std::vector<uint8_t> Decisions(uint32_t n) const { std::vector<uint8_t> d(n); clb_sampler(range, exponent, n_pseudo, n, d.data()); return d; }
	uint32_t GetNAccepted(uint32_t n) const { uint32_t a = 0; for (auto x : Decisions(n)) a += x; return a; }   // ref_reads_accepter.h:42-49
};

struct CCompressElem {            // queues_data.h:23 without the moved read
	uint32_t read_id;
	std::vector<uint32_t> ref_reads;
	std::vector<std::vector<uint64_t>> common_kmers;   // HiFi only
};

class CReadsSimilarityGraph {
	clb_ctx* ctx; uint32_t max_candidates; bool hifi; uint32_t n_reads;
public:
	// sparse == false is ReferenceReadsMode::All
	CReadsSimilarityGraph(CKmerCounter& counter, uint32_t max_candidates, bool hifi, bool sparse, const CRefReadsAccepter& accepter, uint32_t n_pseudo = 0)
		: ctx(counter.Context()), max_candidates(max_candidates), hifi(hifi), n_reads(counter.GetNReads())
	{
		// the accepter's index runs over pseudo-reads + reads; clb_graph_build takes the decisions of the reads (pseudo-reads are always accepted)
		std::vector<uint8_t> dec(n_reads, 1);
		if (sparse) { const std::vector<uint8_t> all = accepter.Decisions(n_pseudo + n_reads); dec.assign(all.begin() + n_pseudo, all.end()); }
		check(ctx, clb_graph_build(ctx, dec.data(), n_pseudo), "clb_graph_build");
	}
	std::vector<CCompressElem> Elements() const
	{
		std::vector<uint32_t> cand(static_cast<size_t>(n_reads) * max_candidates + 1), cn(n_reads + 1);
		check(ctx, clb_graph_candidates(ctx, cand.data(), cn.data()), "clb_graph_candidates");
		std::vector<uint64_t> coff, ckm; std::vector<uint32_t> ccn;
		if (hifi) {
			uint64_t tot = 0; check(ctx, clb_graph_common_size(ctx, &tot), "clb_graph_common_size");
			coff.assign(cand.size(), 0); ccn.assign(cand.size(), 0); ckm.assign(tot + 1, 0);
			check(ctx, clb_graph_common(ctx, coff.data(), ccn.data(), ckm.data(), tot), "clb_graph_common");
		}
		std::vector<CCompressElem> out(n_reads);
		for (uint32_t i = 0; i < n_reads; ++i) {
			out[i].read_id = i;
			for (uint32_t j = 0; j < cn[i]; ++j) {
				const size_t s = static_cast<size_t>(i) * max_candidates + j;
				out[i].ref_reads.push_back(cand[s]);
				if (hifi) out[i].common_kmers.emplace_back(ckm.begin() + coff[s], ckm.begin() + coff[s] + ccn[s]);
			}
		}
		return out;
	}
};

} // namespace clbhost
