// TEST INFRASTRUCTURE ONLY -- never linked, imported or executed by the product path.
//
// Drop-in check of the reference-side binding bella_b200/csrc/overlap_b200.hpp: this TU includes the
// UNMODIFIED reference headers where they lie under $(BELLA_REF) exactly as src/main.cpp does, then
// runs on the same CSC matrices and reads
//   (1) the reference's own HashSpGEMM        (include/overlap.hpp:650-789, --skip-alignment branch)
//   (2) HashSpGEMM_b200 from the shim header   (same signature, B200 behind the C-ABI)
// with the two lambdas of src/main.cpp:502-524, each writing BELLA's output file.  The test
// (tests/test_shim_dropin.py, -m gpu) compares the two files as sorted line sets: the reference's
// line order is thread-schedule dependent.  Built by oracle/Makefile into
// oracle/_ref/libbella_shim_test.so (needs the reference tree, so it is prebuilt in the container and
// travels to the GPU box like the other built libraries).
#include <iostream>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <istream>
#include <vector>
#include <string>
#include <algorithm>
#include <utility>
#include <array>
#include <typeinfo>
#include <tuple>
#include <queue>
#include <memory>
#include <stack>
#include <functional>
#include <cstring>
#include <numeric>
#include <math.h>
#include <cassert>
#include <ios>
#include <chrono>
#include <thread>
#include <sys/stat.h>
#include <sys/types.h>
#include <map>
#include <unordered_map>
#include <unistd.h>
#include <fcntl.h>
#include <omp.h>

#include "libcuckoo/cuckoohash_map.hh"
#include "include/kmercount.hpp"
#include "include/chain.hpp"
#include "kmercode/hash_funcs.h"
#include "kmercode/Kmer.hpp"
#include "kmercode/Buffer.h"
#include "kmercode/common.h"
#include "kmercode/fq_reader.h"
#include "kmercode/ParallelFASTQ.h"
#include "include/common/utility.h"
#include "include/common/CSC.h"
#include "include/common/common.h"
#include "include/overlap.hpp"
#include "include/align.hpp"

#include "overlap_b200.hpp"      // the binding under test (bella_b200/csrc)

typedef uint32_t IT;
typedef unsigned short NT;
typedef CSC<IT, NT> Mat;

namespace {
struct Quiet {
	int o1, o2, nul;
	Quiet() { fflush(stdout); fflush(stderr); std::cout.flush(); std::cerr.flush(); nul = open("/dev/null", O_WRONLY); o1 = dup(1); o2 = dup(2); dup2(nul, 1); dup2(nul, 2); }
	~Quiet() { fflush(stdout); fflush(stderr); std::cout.flush(); std::cerr.flush(); dup2(o1, 1); dup2(o2, 2); close(o1); close(o2); close(nul); }
};
Mat* make_csc(IT rows, IT cols, IT nnz, const IT* colptr, const IT* rowids, const NT* values)
{
	Mat* M = new Mat(nnz, rows, cols);
	memcpy(M->colptr, colptr, sizeof(IT) * (size_t(cols) + 1));
	memcpy(M->rowids, rowids, sizeof(IT) * size_t(nnz));
	memcpy(M->values, values, sizeof(NT) * size_t(nnz));
	return M;
}
} // namespace

// which: bit 0 = reference HashSpGEMM -> out_ref, bit 1 = HashSpGEMM_b200 -> out_b200,
// bit 2 = OverlapFromTuples_b200 -> out_b200 (tuples = B's entries read by read in position order, what
// src/main.cpp:393-416 emits; the reference arm then also rebuilds B with its own CSC constructor from those
// tuples, src/main.cpp:476-480, so both sides start from the tuples).
// memory_mb sizes the stage loop (overlap.hpp:682-710): a small value forces several stages.
extern "C" int shim_compare(IT n_reads, IT n_kmers, IT nnz, const IT* A_colptr, const IT* A_rowids, const NT* A_values,
		const IT* B_colptr, const IT* B_rowids, const NT* B_values, const char* seqs, const uint64_t* seq_off,
		unsigned short kmer_size, unsigned short bin_size, double memory_mb, const char* out_ref, const char* out_b200, int which)
{
	Quiet q;
	Mat* A = make_csc(n_reads, n_kmers, nnz, A_colptr, A_rowids, A_values);
	Mat* B = make_csc(n_kmers, n_reads, nnz, B_colptr, B_rowids, B_values);
	readVector_ reads(n_reads);
	for (IT i = 0; i < n_reads; ++i) {
		reads[i].readid = i;
		reads[i].nametag = "read" + std::to_string(i);
		reads[i].seq.assign(seqs + seq_off[i], seqs + seq_off[i + 1]);
	}
	BELLApars bpars;
	bpars.kmerSize = kmer_size;
	bpars.binSize = bin_size;
	bpars.skipAlignment = true;
	bpars.userDefMem = true;
	bpars.totalMemory = memory_mb;
	double ratiophi = 0.0;
	spmatPtr_ getvaluetype(make_shared<spmatType_>());
	// src/main.cpp:502-524
	auto multop = [&bpars, &reads] (const unsigned short int& begpH, const unsigned short int& begpV,
			const unsigned int& id1, const unsigned int& id2)
	{
		spmatPtr_ value(make_shared<spmatType_>());
		std::string& read1 = reads[id1].seq;
		std::string& read2 = reads[id2].seq;
		multiop(value, read1, read2, begpH, begpV, bpars.kmerSize);
		return value;
	};
	auto addop = [&bpars, &reads] (spmatPtr_& m1, spmatPtr_& m2, const unsigned int& id1, const unsigned int& id2)
	{
		std::string& readname1 = reads[id1].nametag;
		std::string& readname2 = reads[id2].nametag;
		chainop(m1, m2, bpars, readname1, readname2);
		return m1;
	};
	std::vector<std::tuple<IT, IT, NT>> tuples;
	if (which & 4) {
		tuples.reserve(nnz);
		for (IT i = 0; i < n_reads; ++i) {
			std::vector<std::pair<NT, IT>> col;
			for (IT j = B_colptr[i]; j < B_colptr[i + 1]; ++j) col.emplace_back(B_values[j], B_rowids[j]);
			std::sort(col.begin(), col.end());
			for (auto& e : col) tuples.emplace_back(e.second, i, e.first);
		}
		// the reference's own matrix construction from the same tuples (src/main.cpp:476-489)
		std::vector<std::tuple<IT, IT, NT>> copy(tuples);
		Mat* B2 = new Mat(copy, n_kmers, n_reads, [] (unsigned short int& p1, unsigned short int& p2) { return p1; }, false);
		delete B; B = B2;
		// A = B.Transpose(); transpose.h:35 writes one past its colptr, so call the routine on a padded array
		std::vector<IT> acolptr(size_t(n_kmers) + 2);
		Mat* A2 = new Mat(B->nnz, n_reads, n_kmers);
		csr2csc_atomic_nosort(B->cols, B->rows, B->nnz, B->colptr, B->rowids, B->values, acolptr.data(), A2->rowids, A2->values);
		memcpy(A2->colptr, acolptr.data(), sizeof(IT) * (size_t(n_kmers) + 1));
		delete A; A = A2;
	}
	if (which & 1) {
		remove(out_ref);
		HashSpGEMM(*A, *B, multop, addop, reads, getvaluetype, (char*)out_ref, bpars, ratiophi);
	}
	if (which & 4) {
		remove(out_b200);
		OverlapFromTuples_b200<IT, NT>(tuples, n_kmers, n_reads, reads, getvaluetype, (char*)out_b200, bpars, ratiophi);
	}
	if (which & 2) {
		remove(out_b200);
		HashSpGEMM_b200(*A, *B, multop, addop, reads, getvaluetype, (char*)out_b200, bpars, ratiophi);
	}
	delete A;
	delete B;
	return 0;
}
