// overlap_b200.hpp -- the reference-side binding of the B200 overlap SpGEMM.
//
// Header-only shim with the template signature of BELLA's HashSpGEMM (reference
// include/overlap.hpp:650-652).  #include it in the reference's src/main.cpp after
// "../include/overlap.hpp" and rename the one call at src/main.cpp:499 from HashSpGEMM( to
// HashSpGEMM_b200( -- nothing else changes.  It needs the reference's own types
// (CSC<IT,NT>, readVector_, BELLApars, spmatPtr_, RunPairWiseAlignments), so it only compiles inside
// a translation unit that already includes the reference headers; it links against libbella_b200.so.
//
// What it does, in the order of the reference's HashSpGEMM:
//   estimateFLOP + estimateNNZ_Hash + prefix sums (overlap.hpp:667-679)  -> bella_b200_symbolic
//   the stage loop over column ranges sized by the memory budget (:682-712) -> kept as is
//   LocalSpGEMM + combine step (:719-741)                                   -> bella_b200_numeric, then one
//        spmatType_ per nonzero holding the fold's result (count, the chosen seed, its bin's support and
//        overlap), which is all RunPairWiseAlignments reads through choose()/chain() (:557-590)
//   RunPairWiseAlignments (:748)                                            -> the reference's own, unchanged
// The multiply/add functors are accepted for signature compatibility and never called: their
// arithmetic (chain.hpp:47-150) is what the device implements.  The substring test of checkstrand
// (chain.hpp:35-44) becomes one strand bit per nonzero, computed here from the reads.
//
// Error convention: like the reference (overlap.hpp has no error returns) a failure prints to stderr
// and exit(1)s.  HOPC k-mers (bpars.useHOPC) are rejected: their orientation is not substring
// equality of k characters.
#ifndef BELLA_OVERLAP_B200_HPP_
#define BELLA_OVERLAP_B200_HPP_

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#include "bella_b200.h"

namespace bella_b200_shim {

inline char complement(char c) { return c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : 'A'; }

// 1 iff the window read[p, p+k) is its own canonical form (<= its reverse complement, A<C<G<T).
// Two windows of the same canonical k-mer are equal as strings iff their bits are equal, which is
// all the orientation test of chain.hpp:39-42 asks.  Returns -1 on a character outside ACGT.
inline int strand_bit(const std::string& s, size_t p, size_t k)
{
	if (p + k > s.size()) return -1;
	int decided = -1;
	for (size_t t = 0; t < k; ++t) {
		const char f = s[p + t], b = s[p + k - 1 - t];
		if ((f != 'A' && f != 'C' && f != 'G' && f != 'T')) return -1;
		if (decided < 0) {
			const char r = complement(b);
			if (f != r) decided = f < r ? 1 : 0;
		}
	}
	return decided < 0 ? 1 : decided;      // palindrome: the window is its own reverse complement
}

[[noreturn]] inline void die(bella_b200_handle* h, const char* what, int rc)
{
	std::fprintf(stderr, "bella_b200: %s failed (%d): %s\n", what, rc, h ? bella_b200_last_error(h) : "");
	std::exit(1);
}

} // namespace bella_b200_shim

namespace bella_b200_shim {

// symbolic phase, the reference's stage loop and RunPairWiseAlignments on a handle whose inputs are set
template <typename IT, typename FT>
void run_stages(bella_b200_handle* h, IT n, const readVector_& reads, char* filename, const BELLApars& bpars, const double& ratiophi)
{
	int rc;
	// symbolic phase (overlap.hpp:667-679)
	uint64_t flops64 = 0;
	IT* colptrC = new IT[size_t(n) + 1];
	rc = bella_b200_symbolic(h, &flops64, nullptr, colptrC);
	if (rc) die(h, "bella_b200_symbolic", rc);
	std::string FLOPs = std::to_string(flops64);
	printLog(FLOPs);
	IT nnzc = colptrC[n];
	double compression_ratio = nnzc ? double(flops64) / nnzc : 0.0;

	// stage boundaries exactly as the reference sizes them (overlap.hpp:682-710)
	double free_memory = estimateMemory(bpars);
	uint64_t required_memory = safety_net * nnzc * (sizeof(FT) + sizeof(IT));
	int stages = std::max(1, int(std::ceil(double(required_memory) / free_memory)));
	uint64_t nnzcperstage = uint64_t(free_memory / (safety_net * (sizeof(FT) + sizeof(IT))));
	std::cout << nnzc << std::endl;
	std::string nnzOutput = std::to_string(nnzc);
	std::string CompressionRatio = std::to_string(compression_ratio);
	std::string RequiredStages = std::to_string(stages);
	printLog(nnzOutput);
	printLog(CompressionRatio);
	printLog(RequiredStages);
	IT* colStart = new IT[stages + 1];
	colStart[0] = 0;
	for (int i = 1; i < stages; ++i) {
		auto upper = std::upper_bound(colptrC, colptrC + n + 1, i * nnzcperstage);
		colStart[i] = IT(upper - colptrC - 1);
	}
	colStart[stages] = n;

	for (int b = 0; b < stages; ++b) {
		const IT begnz = colptrC[colStart[b]], endnz = colptrC[colStart[b + 1]];
		const size_t cnt = size_t(endnz - begnz);
		IT* rowids = new IT[cnt ? cnt : 1];
		FT* values = new FT[cnt ? cnt : 1];
		std::vector<uint16_t> count(cnt), posH(cnt), posV(cnt), nbins(cnt), support(cnt), overlap(cnt);
		rc = bella_b200_numeric(h, colStart[b], colStart[b + 1], rowids, count.data(), posH.data(), posV.data());
		if (rc) die(h, "bella_b200_numeric", rc);
		rc = bella_b200_numeric_aux(h, colStart[b], colStart[b + 1], nbins.data(), support.data(), overlap.data());
		if (rc) die(h, "bella_b200_numeric_aux", rc);
		// the fold's result as the reference's value type: one bin = the chosen one (common.h:119-170)
#pragma omp parallel for
		for (int64_t t = 0; t < int64_t(cnt); ++t) {
			spmatPtr_ v(std::make_shared<spmatType_>());
			v->count = count[t];
			v->pos.push_back({std::make_pair(posH[t], posV[t])});
			v->support.push_back(support[t]);
			v->overlap.push_back(overlap[t]);
			values[t] = v;
		}
		auto alignstats = RunPairWiseAlignments(colStart[b], colStart[b + 1], begnz, colptrC, rowids, values, reads, filename, bpars, ratiophi);
		int LinesOutputted = int(std::get<3>(alignstats));
		printLog(LinesOutputted);
		delete[] rowids;
		delete[] values;
	}
	delete[] colptrC;
	delete[] colStart;
}

} // namespace bella_b200_shim

template <typename IT, typename NT, typename FT, typename MultiplyOperation, typename AddOperation>
void HashSpGEMM_b200(const CSC<IT, NT>& A, const CSC<IT, NT>& B, MultiplyOperation, AddOperation, const readVector_& reads,
		FT& getvaluetype, char* filename, const BELLApars& bpars, const double& ratiophi, int device = 0)
{
	static_assert(sizeof(IT) == 4 && sizeof(NT) == 2, "the B200 path is built for CSC<uint32_t, unsigned short> (KMERINDEX = uint32_t)");
	using namespace bella_b200_shim;
	(void)getvaluetype;
	if (bpars.useHOPC) { std::fprintf(stderr, "bella_b200: --hopc is not supported by the B200 overlap path\n"); std::exit(1); }

	const IT n = B.cols;
	// reads -> read lengths + one strand bit per nonzero of B (replaces `reads` inside the multiply)
	std::vector<uint32_t> read_len(n);
	std::vector<uint8_t> strandB((size_t(B.nnz) + 7) / 8 + 8, 0);
	int bad = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(| : bad)
	for (int64_t i = 0; i < int64_t(n); ++i) {
		const std::string& s = reads[i].seq;
		read_len[i] = uint32_t(s.size());
		for (IT j = B.colptr[i]; j < B.colptr[i + 1]; ++j) {
			const int bit = strand_bit(s, B.values[j], bpars.kmerSize);
			if (bit < 0) { bad = 1; continue; }
			if (bit) {
#pragma omp atomic
				strandB[j >> 3] |= uint8_t(1u << (j & 7));
			}
		}
	}
	if (bad) {
		std::fprintf(stderr, "bella_b200: a k-mer window contains a character other than upper-case ACGT (or runs past its read); "
			"the strand-bit form of checkstrand does not cover it\n");
		std::exit(1);
	}

	bella_b200_handle* h = nullptr;
	int rc = bella_b200_create(&h, device);
	if (rc) die(nullptr, "bella_b200_create (no usable sm_100 device; there is no CPU fallback)", rc);
	bella_csc_view vA{A.rows, A.cols, A.nnz, A.colptr, A.rowids, A.values};
	bella_csc_view vB{B.rows, B.cols, B.nnz, B.colptr, B.rowids, B.values};
	rc = bella_b200_set_inputs(h, &vA, &vB, read_len.data(), nullptr, strandB.data(), bpars.kmerSize, bpars.binSize);
	if (rc) die(h, "bella_b200_set_inputs", rc);
	run_stages<IT, FT>(h, n, reads, filename, bpars, ratiophi);
	bella_b200_destroy(h);
}

// "Next" row f2: the same, starting from the tuples BELLA emits (src/main.cpp:393-416), i.e. replacing
//   CSC<IT,NT> transpmat(alltuples, nkmer, numReads, keep-p1, false);  spmat = transpmat.Transpose();  HashSpGEMM(...)
// (src/main.cpp:476-525).  B is built on the device in the reference's MergeDuplicates order (src/CSC.cpp:301-479).
template <typename IT, typename NT, typename FT>
void OverlapFromTuples_b200(const std::vector<std::tuple<IT, IT, NT>>& tuples, IT nkmer, IT nreads, const readVector_& reads,
		FT& getvaluetype, char* filename, const BELLApars& bpars, const double& ratiophi, int device = 0)
{
	static_assert(sizeof(IT) == 4 && sizeof(NT) == 2, "the B200 path is built for CSC<uint32_t, unsigned short> (KMERINDEX = uint32_t)");
	using namespace bella_b200_shim;
	(void)getvaluetype;
	if (bpars.useHOPC) { std::fprintf(stderr, "bella_b200: --hopc is not supported by the B200 overlap path\n"); std::exit(1); }
	const size_t T = tuples.size();
	std::vector<uint32_t> tk(T), tr(T), read_len(nreads);
	std::vector<uint16_t> tp(T);
	std::vector<uint8_t> strand((T + 7) / 8 + 8, 0);
	for (IT i = 0; i < nreads; ++i) read_len[i] = uint32_t(reads[i].seq.size());
	int bad = 0;
#pragma omp parallel for reduction(| : bad)
	for (int64_t t = 0; t < int64_t(T); ++t) {
		tk[t] = std::get<0>(tuples[t]); tr[t] = std::get<1>(tuples[t]); tp[t] = std::get<2>(tuples[t]);
		const int bit = tr[t] < nreads ? strand_bit(reads[tr[t]].seq, tp[t], bpars.kmerSize) : -1;
		if (bit < 0) { bad = 1; continue; }
		if (bit) {
#pragma omp atomic
			strand[t >> 3] |= uint8_t(1u << (t & 7));
		}
	}
	if (bad) { std::fprintf(stderr, "bella_b200: a tuple's k-mer window is outside its read or contains a character other than upper-case ACGT\n"); std::exit(1); }
	bella_b200_handle* h = nullptr;
	int rc = bella_b200_create(&h, device);
	if (rc) die(nullptr, "bella_b200_create (no usable sm_100 device; there is no CPU fallback)", rc);
	rc = bella_b200_set_inputs_tuples(h, nkmer, nreads, T, tk.data(), tr.data(), tp.data(), strand.data(), read_len.data(), bpars.kmerSize, bpars.binSize);
	if (rc) die(h, "bella_b200_set_inputs_tuples", rc);
	run_stages<IT, FT>(h, nreads, reads, filename, bpars, ratiophi);
	bella_b200_destroy(h);
}

#endif // BELLA_OVERLAP_B200_HPP_
