// TEST INFRASTRUCTURE ONLY -- never linked, imported or executed by the product path.
//
// Drop-in check of the reference-side binding of the "next" row f1, bella_b200/csrc/align_b200.hpp: this TU includes the
// UNMODIFIED reference headers as src/main.cpp does and writes BELLA's output file twice from the same overlap matrix,
//   (1) reference arm: per nonzero alignSeqAn (include/align.hpp:93-139) + PostAlignDecision (include/overlap.hpp:415-497),
//       the body of RunPairWiseAlignments' loop (:548-577).  alignSeqAn, not the xavierAlign the loop itself calls:
//       overlap.hpp:74-76 defines __SIMD__ unconditionally and xavier is a different, fixed-band algorithm; f1 is the
//       SeqAn/LOGAN recurrence;
//   (2) RunPairWiseAlignments_b200 (same signature as the reference's RunPairWiseAlignments, B200 behind the C-ABI).
// tests/test_xdrop_shim_gpu.py (-m gpu) compares the files as sorted line sets and the returned statistics.
// Built by oracle/Makefile into oracle/_ref/libbella_align_shim_test.so (needs the reference tree).
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

#include "align_b200.hpp"        // the binding under test (bella_b200/csrc)

typedef uint32_t IT;

namespace {
struct Quiet {
	int o1, o2, nul;
	Quiet() { fflush(stdout); fflush(stderr); std::cout.flush(); std::cerr.flush(); nul = open("/dev/null", O_WRONLY); o1 = dup(1); o2 = dup(2); dup2(nul, 1); dup2(nul, 2); }
	~Quiet() { fflush(stdout); fflush(stderr); std::cout.flush(); std::cerr.flush(); dup2(o1, 1); dup2(o2, 2); close(o1); close(o2); close(nul); }
};
} // namespace

// which: bit 0 = reference arm -> out_ref, bit 1 = RunPairWiseAlignments_b200 -> out_b200.
// C = overlap matrix in CSC form over n_reads columns with the fold's result per nonzero (count, chosen seed).
// stats_ref / stats_b200: alignedpairs, alignedbases, totalreadlen, totaloutputt, totsuccbases, totfailbases.
extern "C" int align_shim_compare(IT n_reads, const char* seqs, const uint64_t* seq_off, const IT* colptrC, const IT* rowidsC,
		const unsigned short* count, const unsigned short* posH, const unsigned short* posV, unsigned short kmer_size,
		unsigned short xdrop, double ratiophi, double delta, int fixed_threshold, int paf, const char* out_ref, const char* out_b200,
		int which, uint64_t* stats_ref, uint64_t* stats_b200)
{
	Quiet q;
	readVector_ reads(n_reads);
	for (IT i = 0; i < n_reads; ++i) {
		reads[i].readid = i;
		reads[i].nametag = "read" + std::to_string(i);
		reads[i].seq.assign(seqs + seq_off[i], seqs + seq_off[i + 1]);
	}
	BELLApars bpars;
	bpars.kmerSize = kmer_size; bpars.xDrop = xdrop; bpars.deltaChernoff = delta;
	bpars.fixedThreshold = (short)fixed_threshold; bpars.outputPaf = paf != 0; bpars.skipAlignment = false;
	const IT nnz = colptrC[n_reads];
	std::vector<IT> colptr(colptrC, colptrC + n_reads + 1), rowids(rowidsC, rowidsC + nnz);
	std::vector<spmatPtr_> values(nnz);
	for (IT t = 0; t < nnz; ++t) {
		spmatPtr_ v(std::make_shared<spmatType_>());
		v->count = count[t];
		v->pos.push_back({std::make_pair(posH[t], posV[t])});
		v->support.push_back(count[t]);
		v->overlap.push_back(0);
		values[t] = v;
	}
	if (which & 1) {
		std::stringstream ss;
		size_t outputted = 0, succ = 0, fail = 0, pairs = 0, bases = 0, readlen = 0;
		for (IT j = 0; j < n_reads; ++j)
			for (IT i = colptr[j]; i < colptr[j + 1]; ++i) {
				const std::string& seq1 = reads[rowids[i]].seq;
				const std::string& seq2 = reads[j].seq;
				std::pair<int, int> kmer = values[i]->choose();
				seqAnResult r = alignSeqAn(seq1, seq2, (int)seq1.length(), kmer.first, kmer.second, bpars.xDrop, bpars.kmerSize, false, false, false);
				xavierResult xr;                                   // PostAlignDecision's parameter type in this build
				xr.score = r.score; xr.strand = r.strand;
				xr.seed = SeedX((int)beginPositionH(r.seed), (int)beginPositionV(r.seed), (int)endPositionH(r.seed), (int)endPositionV(r.seed));
				bool passed = false;
				PostAlignDecision(xr, reads[rowids[i]], reads[j], bpars, ratiophi, values[i]->count, ss, outputted, succ, fail, passed, values[i]->chain());
				++pairs; bases += endPositionV(r.seed) - beginPositionV(r.seed);
				readlen += (unsigned short)seq1.length() + (unsigned short)seq2.length();
			}
		remove(out_ref);
		std::ofstream ofs(out_ref, std::ios::binary);
		const std::string text = ss.str();
		ofs.write(text.data(), (std::streamsize)text.size());
		stats_ref[0] = pairs; stats_ref[1] = bases; stats_ref[2] = readlen; stats_ref[3] = outputted; stats_ref[4] = succ; stats_ref[5] = fail;
	}
	if (which & 2) {
		remove(out_b200);
		auto st = RunPairWiseAlignments_b200(IT(0), n_reads, IT(0), colptr.data(), rowids.data(), values.data(), reads, (char*)out_b200, bpars, ratiophi);
		stats_b200[0] = std::get<0>(st); stats_b200[1] = std::get<1>(st); stats_b200[2] = std::get<2>(st);
		stats_b200[3] = std::get<3>(st); stats_b200[4] = std::get<4>(st); stats_b200[5] = std::get<5>(st);
	}
	return 0;
}
