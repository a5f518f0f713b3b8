// TEST INFRASTRUCTURE ONLY -- never linked, imported or executed by the product path.
//
// The reference's own CUDA aligner (LOGAN, loganGPU/*.cuh) recompiled for sm_100 from the sources where they lie under
// $(BELLA_REF), as the second live reference and the performance bar of the "next" row f1.  This TU
// includes the UNMODIFIED loganGPU headers and drives them as the reference's GPU build does:
//   RunPairWiseAlignmentsGPU  include/overlap.hpp:876-972  (strand detection, reverse-complemented copy of the row read,
//                                                            one SeedL / two strings per pair)
//   alignLogan                include/align.hpp:210-255     (batches of BATCH_SIZE = 30000 pairs per GPU, :35)
//   extendSeedL               loganGPU/functions.cuh:410-689 (per batch: host prefix/suffix copies, cudaMalloc, H2D, two
//                                                            kernels of one 32-thread block per alignment, D2H, cudaFree)
// out6[p] = { score, strand char, begH, endH, begV, endV } like bella_ref_align; *seconds = wall time of the extendSeedL
// calls (what the reference's GPU build spends in alignLogan).  Built by oracle/Makefile into oracle/_ref/libbella_logan.so.
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include <omp.h>

#include "logan.cuh"

namespace {
char comp_base(char c)                         // complementbase, include/common/common.h
{
	switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; default: return 'N'; }
}
const int kBatch = 30000;                      // BATCH_SIZE, include/align.hpp:35
}

extern "C" int bella_logan_align(uint64_t n_pairs, const uint32_t* rows, const uint32_t* cols, const unsigned short* posH,
		const unsigned short* posV, const char* seqs, const uint64_t* seq_off, int kmer_len, int xdrop, int32_t* out6, double* seconds)
{
	std::vector<std::string> seq1s(n_pairs), seq2s(n_pairs);
	std::vector<SeedL> seeds(n_pairs);
	std::vector<char> strand(n_pairs);
	// overlap.hpp:890-944
	for (uint64_t p = 0; p < n_pairs; ++p) {
		std::string seq1(seqs + seq_off[rows[p]], seqs + seq_off[rows[p] + 1]);
		seq2s[p].assign(seqs + seq_off[cols[p]], seqs + seq_off[cols[p] + 1]);
		const int i = posH[p], j = posV[p], len1 = (int)seq1.length();
		SeedL seed(i, j, i + kmer_len, j + kmer_len);
		std::string seedH = seq1.substr(i, kmer_len), seedV = seq2s[p].substr(j, kmer_len);
		std::reverse(seedH.begin(), seedH.end());
		std::transform(seedH.begin(), seedH.end(), seedH.begin(), comp_base);
		strand[p] = 'n';
		if (seedH == seedV) {
			strand[p] = 'c';
			std::reverse(seq1.begin(), seq1.end());
			std::transform(seq1.begin(), seq1.end(), seq1.begin(), comp_base);
			setBeginPositionH(seed, len1 - i - kmer_len);
			setBeginPositionV(seed, j);
			setEndPositionH(seed, len1 - i);
			setEndPositionV(seed, j + kmer_len);
		}
		seeds[p] = seed;
		seq1s[p].swap(seq1);
	}
	// align.hpp:214-254
	ScoringSchemeL sscheme(1, -1, -1, -1);
	std::vector<ScoringSchemeL> scoring;
	scoring.push_back(sscheme);
	double t = 0.0;
	for (uint64_t b = 0; b < n_pairs; b += kBatch) {
		const int n = (int)std::min<uint64_t>(kBatch, n_pairs - b);
		std::vector<std::string> target_b(seq1s.begin() + b, seq1s.begin() + b + n);
		std::vector<std::string> query_b(seq2s.begin() + b, seq2s.begin() + b + n);
		std::vector<SeedL> seeds_b(seeds.begin() + b, seeds.begin() + b + n);
		std::vector<int> res(n);
		const auto t0 = std::chrono::high_resolution_clock::now();
		extendSeedL(seeds_b, EXTEND_BOTHL, target_b, query_b, scoring, xdrop, kmer_len, res.data(), n, 1);
		t += std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
		for (int q = 0; q < n; ++q) {
			int32_t* o = out6 + 6 * (b + q);
			o[0] = res[q]; o[1] = strand[b + q];
			o[2] = getBeginPositionH(seeds_b[q]); o[3] = getEndPositionH(seeds_b[q]);
			o[4] = getBeginPositionV(seeds_b[q]); o[5] = getEndPositionV(seeds_b[q]);
		}
	}
	if (seconds) *seconds = t;
	return 0;
}
