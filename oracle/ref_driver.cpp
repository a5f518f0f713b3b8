// TEST INFRASTRUCTURE ONLY -- never linked, imported or executed by the product path.
//
// Thin C-ABI driver around the UNMODIFIED reference implementation of the overlap SpGEMM.
// This TU #includes the reference's own headers where they lie under $(BELLA_REF)
// (= /root/reference); no reference source is copied into this repository.  It is built by
// oracle/Makefile into oracle/_ref/libbella_ref.so and used
//   * by tests/ to validate oracle/bella_oracle.c (the C restatement) and to generate
//     tests/golden/ fixtures (tests/golden/make_golden.py),
//   * by bench.py as the `cpu_baseline` / `--impl reference` arm (kind = "reference").
//
// What it runs (reference file:line):
//   estimateFLOP        include/overlap.hpp:157-202
//   prefixsum           include/overlap.hpp:110-146
//   estimateNNZ_Hash    include/overlap.hpp:205-276
//   LocalSpGEMM         include/overlap.hpp:281-363
//   the two lambdas of  src/main.cpp:502-524 (multiop / chainop, include/chain.hpp:74-150)
//   spmatType_::choose  include/common/common.h:162-170
//   CSC tuple ctor + MergeDuplicates + Transpose   src/CSC.cpp:289-479, include/common/transpose.h
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

typedef uint32_t IT;             // KMERINDEX, src/main.cpp:60
typedef unsigned short NT;
typedef CSC<IT, NT> Mat;

namespace {

// the reference prints through bare std::cout / printLog(std::cerr); silence fd 1/2 around calls
struct Quiet {
	int o1, o2, nul;
	Quiet() {
		fflush(stdout); fflush(stderr); std::cout.flush(); std::cerr.flush();
		nul = open("/dev/null", O_WRONLY);
		o1 = dup(1); o2 = dup(2);
		dup2(nul, 1); dup2(nul, 2);
	}
	~Quiet() {
		fflush(stdout); fflush(stderr); std::cout.flush(); std::cerr.flush();
		dup2(o1, 1); dup2(o2, 2);
		close(o1); close(o2); close(nul);
	}
};

struct RefHandle {
	Mat* A = nullptr;            // reads x kmers  (spmat,     src/main.cpp:489)
	Mat* B = nullptr;            // kmers x reads  (transpmat, src/main.cpp:476), possibly a column prefix
	readVector_ reads;
	BELLApars bpars;
	IT* flopC = nullptr;
	IT* colptrC = nullptr;
	IT flops = 0;
	double t_flop = 0, t_nnz = 0, t_numeric = 0;
	~RefHandle() { delete A; delete B; delete[] flopC; delete[] colptrC; }
};

Mat* make_csc(IT rows, IT cols, IT nnz, const IT* colptr, const IT* rowids, const NT* values)
{
	Mat* M = new Mat(nnz, rows, cols);     // (nnz, m, n) argument order, include/common/CSC.h:25
	memcpy(M->colptr, colptr, sizeof(IT) * (size_t(cols) + 1));
	memcpy(M->rowids, rowids, sizeof(IT) * size_t(nnz));
	memcpy(M->values, values, sizeof(NT) * size_t(nnz));
	return M;
}

} // namespace

extern "C" {

// Create: copies the matrices, rebuilds `reads` (sequence strings are what the reference's multiply
// needs, chain.hpp:35-44), then runs the symbolic phase exactly as HashSpGEMM does
// (overlap.hpp:667-679).  `ncols_sample` (0 = all) restricts B to its first ncols_sample columns
// (the bounded CPU-baseline sample: output columns [0, ncols_sample) of the same workload).
void* bella_ref_create(IT n_reads, IT n_kmers,
		IT nnzA, const IT* A_colptr, const IT* A_rowids, const NT* A_values,
		IT nnzB, const IT* B_colptr, const IT* B_rowids, const NT* B_values,
		const char* seqs, const uint64_t* seq_off,
		unsigned short kmer_size, unsigned short bin_size,
		IT ncols_sample, int nthreads)
{
	Quiet q;
	if (nthreads > 0) omp_set_num_threads(nthreads);
	RefHandle* h = new RefHandle();
	IT bcols = (ncols_sample && ncols_sample < n_reads) ? ncols_sample : n_reads;
	IT bnnz = B_colptr[bcols];
	if (nnzA == 0 || bnnz == 0) { delete h; return nullptr; }
	h->A = make_csc(n_reads, n_kmers, nnzA, A_colptr, A_rowids, A_values);
	h->B = make_csc(n_kmers, bcols, bnnz, B_colptr, B_rowids, B_values);
	h->reads.resize(n_reads);
	for (IT i = 0; i < n_reads; ++i) {
		h->reads[i].readid = i;
		h->reads[i].nametag = std::to_string(i);
		h->reads[i].seq.assign(seqs + seq_off[i], seqs + seq_off[i + 1]);
	}
	h->bpars.kmerSize = kmer_size;
	h->bpars.binSize = bin_size;
	h->bpars.skipAlignment = true;

	int numThreads = 1;
#pragma omp parallel
	{
		numThreads = omp_get_num_threads();
	}
	double t0 = omp_get_wtime();
	h->flopC = estimateFLOP(*h->A, *h->B, true);
	IT* flopptr = prefixsum<IT>(h->flopC, h->B->cols, numThreads);
	h->flops = flopptr[h->B->cols];
	double t1 = omp_get_wtime();
	IT* colnnzC = estimateNNZ_Hash(*h->A, *h->B, h->flopC, true);
	h->colptrC = prefixsum<IT>(colnnzC, h->B->cols, numThreads);
	double t2 = omp_get_wtime();
	delete[] colnnzC;
	delete[] flopptr;
	h->t_flop = t1 - t0;
	h->t_nnz = t2 - t1;
	return h;
}

IT bella_ref_cols(void* hv) { return ((RefHandle*)hv)->B->cols; }
uint64_t bella_ref_flops(void* hv) { return ((RefHandle*)hv)->flops; }
const IT* bella_ref_flopC(void* hv) { return ((RefHandle*)hv)->flopC; }
const IT* bella_ref_colptrC(void* hv) { return ((RefHandle*)hv)->colptrC; }
void bella_ref_times(void* hv, double* out3)
{
	RefHandle* h = (RefHandle*)hv;
	out3[0] = h->t_flop; out3[1] = h->t_nnz; out3[2] = h->t_numeric;
}

// Numeric phase for output columns [c0, c1): LocalSpGEMM + choose().  Outputs are laid out at
// colptrC[i]-colptrC[c0]; rows are sorted ascending inside each column (the reference's own order
// is hash-slot order and already schedule-dependent through Transpose(), so the canonical order
// for comparisons is (col, row)).  aux (optional, may be NULL) = {nbins, support, overlap} of the
// chosen bin, 3 x u16 per nonzero.
int bella_ref_numeric(void* hv, IT c0, IT c1, IT* rowidsC, NT* count, NT* posH, NT* posV, NT* aux)
{
	Quiet q;
	RefHandle* h = (RefHandle*)hv;
	if (c1 > h->B->cols || c0 > c1) return -1;
	IT ncols = c1 - c0;
	std::vector<IT>* RowIdsofC = new std::vector<IT>[ncols];
	std::vector<spmatPtr_>* ValuesofC = new std::vector<spmatPtr_>[ncols];
	BELLApars& bpars = h->bpars;
	readVector_& reads = h->reads;

	IT s0 = c0, e0 = c1;         // LocalSpGEMM takes non-const lvalue refs (overlap.hpp:281)
	double t0 = omp_get_wtime();
	LocalSpGEMM(s0, e0, *h->A, *h->B,
		// src/main.cpp:502-513
		[&bpars, &reads] (const unsigned short int& begpH, const unsigned short int& begpV,
			const unsigned int& id1, const unsigned int& id2)
		{
			spmatPtr_ value(make_shared<spmatType_>());
			std::string& read1 = reads[id1].seq;
			std::string& read2 = reads[id2].seq;
			multiop(value, read1, read2, begpH, begpV, bpars.kmerSize);
			return value;
		},
		// src/main.cpp:514-524
		[&bpars, &reads] (spmatPtr_& m1, spmatPtr_& m2, const unsigned int& id1,
			const unsigned int& id2)
		{
			std::string& readname1 = reads[id1].nametag;
			std::string& readname2 = reads[id2].nametag;
			chainop(m1, m2, bpars, readname1, readname2);
			return m1;
		},
		RowIdsofC, ValuesofC, h->colptrC, true);
	h->t_numeric = omp_get_wtime() - t0;

	IT base = h->colptrC[c0];
	int rc = 0;
#pragma omp parallel for schedule(dynamic, 64)
	for (IT i = 0; i < ncols; ++i) {
		IT off = h->colptrC[c0 + i] - base;
		IT cnt = RowIdsofC[i].size();
		if (cnt != h->colptrC[c0 + i + 1] - h->colptrC[c0 + i]) { rc = -2; continue; }
		std::vector<IT> perm(cnt);
		std::iota(perm.begin(), perm.end(), 0);
		std::sort(perm.begin(), perm.end(), [&](IT a, IT b) { return RowIdsofC[i][a] < RowIdsofC[i][b]; });
		for (IT t = 0; t < cnt; ++t) {
			IT p = perm[t];
			spmatPtr_& v = ValuesofC[i][p];
			rowidsC[off + t] = RowIdsofC[i][p];
			count[off + t] = v->count;
			NT nb = v->support.size();
			auto seed = v->choose();      // mutates the value; call once, last (common.h:162-170)
			posH[off + t] = seed.first;
			posV[off + t] = seed.second;
			if (aux) {
				aux[3 * size_t(off + t) + 0] = nb;
				aux[3 * size_t(off + t) + 1] = v->support[v->ids[0]];
				aux[3 * size_t(off + t) + 2] = v->overlap[v->ids[0]];
			}
		}
	}
	delete[] RowIdsofC;
	delete[] ValuesofC;
	return rc;
}

void bella_ref_destroy(void* hv) { delete (RefHandle*)hv; }

// Matrix construction exactly as src/main.cpp:476-489: B = CSC(tuples, m, n, keep-p1, needsort=false)
// (counting sort by column + MergeDuplicates, src/CSC.cpp:301-479) and A = B.Transpose()
// (include/common/transpose.h:12-52).  Tuples are (kmer_id, read_id, pos).  The caller passes
// output buffers of capacity ntuples; returns nnz after de-duplication.  A's within-column order
// is schedule dependent in the reference (atomic cursors) -- run with nthreads = 1 for a
// reproducible A.
int64_t bella_ref_build(IT n_kmers, IT n_reads, uint64_t ntuples,
		const IT* t_kmer, const IT* t_read, const NT* t_pos, int nthreads,
		IT* B_colptr, IT* B_rowids, NT* B_values,
		IT* A_colptr, IT* A_rowids, NT* A_values)
{
	Quiet q;
	if (nthreads > 0) omp_set_num_threads(nthreads);
	std::vector<std::tuple<IT, IT, NT>> tuples(ntuples);
	for (uint64_t k = 0; k < ntuples; ++k) tuples[k] = std::make_tuple(t_kmer[k], t_read[k], t_pos[k]);
	Mat transpmat(tuples, n_kmers, n_reads,
		[] (unsigned short int& p1, unsigned short int& p2) { return p1; }, false);
	IT nnz = transpmat.nnz;
	memcpy(B_colptr, transpmat.colptr, sizeof(IT) * (size_t(n_reads) + 1));
	memcpy(B_rowids, transpmat.rowids, sizeof(IT) * size_t(nnz));
	memcpy(B_values, transpmat.values, sizeof(NT) * size_t(nnz));
	// A = B.Transpose() (src/CSC.cpp:289-299) == csr2csc_atomic_nosort(cols, rows, nnz, ...).
	// transpose.h:35 loops `i <= n` and writes cscColPtr[n+1], one past an (n+1)-entry array;
	// call the reference routine directly on an output colptr with one spare slot so the
	// reference code runs unmodified but in bounds.
	std::vector<IT> acolptr(size_t(n_kmers) + 2);
	csr2csc_atomic_nosort(transpmat.cols, transpmat.rows, transpmat.nnz,
		transpmat.colptr, transpmat.rowids, transpmat.values,
		acolptr.data(), A_rowids, A_values);
	memcpy(A_colptr, acolptr.data(), sizeof(IT) * (size_t(n_kmers) + 1));
	return nnz;
}

// "Next" row f1: the reference's CPU seed-and-extend, alignSeqAn (include/align.hpp:93-139) ->
// seqan::extendSeed(..., GappedXDrop) with Score(1,-1,-1).  One call per candidate pair; out6[p] =
// { score, strand char, begH, endH, begV, endV } (H coordinates on the reverse complement when strand == 'c').
int bella_ref_align(uint64_t n_pairs, const IT* rows, const IT* cols, const NT* posH, const NT* posV,
		const char* seqs, const uint64_t* seq_off, int kmer_len, int xdrop, int32_t* out6)
{
	Quiet q;
#pragma omp parallel for schedule(dynamic, 16)
	for (int64_t p = 0; p < (int64_t)n_pairs; ++p) {
		std::string row(seqs + seq_off[rows[p]], seqs + seq_off[rows[p] + 1]);
		std::string col(seqs + seq_off[cols[p]], seqs + seq_off[cols[p] + 1]);
		seqAnResult r = alignSeqAn(row, col, (int)row.length(), posH[p], posV[p], xdrop, kmer_len, false, false, false);
		int32_t* o = out6 + 6 * p;
		o[0] = r.score; o[1] = r.strand[0];
		o[2] = (int32_t)beginPositionH(r.seed); o[3] = (int32_t)endPositionH(r.seed);
		o[4] = (int32_t)beginPositionV(r.seed); o[5] = (int32_t)endPositionV(r.seed);
	}
	return 0;
}

// The reference's accept/reject step behind the alignment, PostAlignDecision (include/overlap.hpp:415-497), called as
// RunPairWiseAlignments does (:574-577).  out8[p] = the six fields of bella_ref_align, then `ov` as the reference PRINTS it
// (column 5 of its output line; -1 when the pair is rejected and nothing is printed), then passed (0/1).
int bella_ref_align_post(uint64_t n_pairs, const IT* rows, const IT* cols, const NT* posH, const NT* posV,
		const char* seqs, const uint64_t* seq_off, int kmer_len, int xdrop, double ratiophi, double delta, int fixed_threshold,
		int32_t* out8)
{
	Quiet q;
	BELLApars bpars;
	bpars.kmerSize = (unsigned short)kmer_len; bpars.xDrop = (unsigned short)xdrop;
	bpars.deltaChernoff = delta; bpars.fixedThreshold = (short)fixed_threshold; bpars.outputPaf = false;
#pragma omp parallel for schedule(dynamic, 16)
	for (int64_t p = 0; p < (int64_t)n_pairs; ++p) {
		readType_ r1, r2;
		r1.nametag = "H"; r2.nametag = "V";
		r1.seq.assign(seqs + seq_off[rows[p]], seqs + seq_off[rows[p] + 1]);
		r2.seq.assign(seqs + seq_off[cols[p]], seqs + seq_off[cols[p] + 1]);
		seqAnResult r = alignSeqAn(r1.seq, r2.seq, (int)r1.seq.length(), posH[p], posV[p], xdrop, kmer_len, false, false, false);
		std::stringstream line;
		size_t outputted = 0, basesTrue = 0, basesFalse = 0;
		bool passed = false;
		// overlap.hpp:74-76 hard-defines __SIMD__, so PostAlignDecision takes a xavierResult: same fields, filled from the SeqAn seed
		xavierResult xr;
		xr.score = r.score; xr.strand = r.strand;
		xr.seed = SeedX((int)beginPositionH(r.seed), (int)beginPositionV(r.seed), (int)endPositionH(r.seed), (int)endPositionV(r.seed));
		PostAlignDecision(xr, r1, r2, bpars, ratiophi, 1, line, outputted, basesTrue, basesFalse, passed, 1);
		int32_t* o = out8 + 8 * p;
		o[0] = r.score; o[1] = r.strand[0];
		o[2] = (int32_t)beginPositionH(r.seed); o[3] = (int32_t)endPositionH(r.seed);
		o[4] = (int32_t)beginPositionV(r.seed); o[5] = (int32_t)endPositionV(r.seed);
		o[6] = -1; o[7] = passed ? 1 : 0;
		if (passed) {
			std::string f; int col = 0; long ov = -1;
			while (std::getline(line, f, '\t')) { if (++col == 5) { ov = atol(f.c_str()); break; } }
			o[6] = (int32_t)ov;
		}
	}
	return 0;
}

// What the reference's DEFAULT CPU build actually calls (overlap.hpp:74-76 defines __SIMD__ unconditionally): xavierAlign
// (include/align.hpp:152-202), a fixed-band SIMD X-drop -- a different algorithm from alignSeqAn / LOGAN, kept here only to
// quantify how far the two reference aligners are from each other.  Same out6 layout as bella_ref_align.
int bella_ref_align_xavier(uint64_t n_pairs, const IT* rows, const IT* cols, const NT* posH, const NT* posV,
		const char* seqs, const uint64_t* seq_off, int kmer_len, int xdrop, int32_t* out6)
{
	Quiet q;
#pragma omp parallel for schedule(dynamic, 16)
	for (int64_t p = 0; p < (int64_t)n_pairs; ++p) {
		std::string row(seqs + seq_off[rows[p]], seqs + seq_off[rows[p] + 1]);
		std::string col(seqs + seq_off[cols[p]], seqs + seq_off[cols[p] + 1]);
		xavierResult r = xavierAlign(row, col, (int)row.length(), posH[p], posV[p], xdrop, kmer_len);
		int32_t* o = out6 + 6 * p;
		o[0] = r.score; o[1] = r.strand[0];
		o[2] = getBeginPositionH(r.seed); o[3] = getEndPositionH(r.seed);
		o[4] = getBeginPositionV(r.seed); o[5] = getEndPositionV(r.seed);
	}
	return 0;
}

// "Next" row f3: the reference's reliable k-mer selection, SplitCount (include/kmercount.hpp:466-677: HyperLogLog estimate
// -> Bloom filter -> cuckoo hash counts -> keep l <= count <= u; what src/main.cpp:299-302 calls), run on a FASTQ file of
// the reads, followed by the tuple emission loop of src/main.cpp:383-416 restated with the reference's own Kmer / rep() /
// dictionary lookup.  K-mer ids are cuckoo iteration order (kmercount.hpp:650-659), i.e. arbitrary, so what is returned is
// the id-free content: every reliable occurrence as (read, pos), in read / position order, and the number of distinct
// reliable k-mers.  One thread: bloom_check_add is not atomic, with several threads two simultaneous first sightings of a
// k-mer can both miss (a race in the reference, not part of its specification).
int bella_ref_reliable_occurrences(const char* fastq_path, uint64_t fastq_size, uint32_t n_reads, const char* seqs, const uint64_t* seq_off,
		int kmer_len, int lower, int upper, uint32_t* out_read, uint16_t* out_pos, uint64_t cap, uint64_t* n_out, uint64_t* n_kmers)
{
	Quiet q;
	const int saved = omp_get_max_threads();
	omp_set_num_threads(1);
	BELLApars bpars;
	bpars.kmerSize = (unsigned short)kmer_len;
	bpars.SplitCount = 1;
	Kmer::set_k(kmer_len);
	std::vector<filedata> allfiles(1);
	strncpy(allfiles[0].filename, fastq_path, MAX_FILE_PATH - 1);
	allfiles[0].filename[MAX_FILE_PATH - 1] = 0;
	allfiles[0].filesize = fastq_size;
	CuckooDict<IT> countsreliable;
	SplitCount(allfiles, countsreliable, lower, upper, (size_t)10000000, bpars);
	omp_set_num_threads(saved);
	*n_kmers = countsreliable.size();
	uint64_t n = 0;
	for (uint32_t r = 0; r < n_reads; ++r) {
		const std::string seq(seqs + seq_off[r], seqs + seq_off[r + 1]);
		const int len = (int)seq.length();
		for (int j = 0; j <= len - kmer_len; ++j) {
			std::string kmerstrfromfastq = seq.substr(j, kmer_len);
			Kmer mykmer(kmerstrfromfastq.c_str(), kmerstrfromfastq.length());
			Kmer lexsmall = mykmer.rep();
			IT idx;
			if (countsreliable.find(lexsmall, idx)) {
				if (n < cap) { out_read[n] = r; out_pos[n] = (uint16_t)j; }
				++n;
			}
		}
	}
	*n_out = n;
	return n <= cap ? 0 : -1;
}

// The reference's own minimizer sampling of one read (include/minimizer.hpp:49-77 on the Kmer objects of every position, exactly
// as src/main.cpp:366-374 builds them).  out: positions, returns their number.
int bella_ref_minimizers(const char* seq, int len, int kmer_len, int window, int* out, int cap)
{
	Kmer::set_k(kmer_len);
	const std::string s(seq, seq + len);
	std::vector<Kmer> seqkmers;
	std::vector<int> seqminimizers;
	for (int j = 0; j <= len - kmer_len; ++j) {
		std::string kmerstrfromfastq = s.substr(j, kmer_len);
		Kmer mykmer(kmerstrfromfastq.c_str(), kmerstrfromfastq.length());
		seqkmers.emplace_back(mykmer);
	}
	getMinimizers((size_t)window, seqkmers, seqminimizers);
	if ((int)seqminimizers.size() > cap) return -1;
	for (size_t t = 0; t < seqminimizers.size(); ++t) out[t] = seqminimizers[t];
	return (int)seqminimizers.size();
}

// MinimizerCount (include/kmercount.hpp:690-832) on a FASTQ of the reads + the minimizer branch of the tuple emission
// (src/main.cpp:363-388); outputs as bella_ref_reliable_occurrences.
int bella_ref_minimizer_occurrences(const char* fastq_path, uint64_t fastq_size, uint32_t n_reads, const char* seqs, const uint64_t* seq_off,
		int kmer_len, int window, int lower, int upper, uint32_t* out_read, uint16_t* out_pos, uint64_t cap, uint64_t* n_out, uint64_t* n_kmers)
{
	Quiet q;
	const int saved = omp_get_max_threads();
	omp_set_num_threads(1);
	BELLApars bpars;
	bpars.kmerSize = (unsigned short)kmer_len;
	bpars.windowLen = (size_t)window;
	bpars.useMinimizer = true;
	Kmer::set_k(kmer_len);
	std::vector<filedata> allfiles(1);
	strncpy(allfiles[0].filename, fastq_path, MAX_FILE_PATH - 1);
	allfiles[0].filename[MAX_FILE_PATH - 1] = 0;
	allfiles[0].filesize = fastq_size;
	CuckooDict<IT> countsreliable;
	MinimizerCount(allfiles, countsreliable, lower, upper, (size_t)10000000, bpars);
	omp_set_num_threads(saved);
	*n_kmers = countsreliable.size();
	uint64_t n = 0;
	for (uint32_t r = 0; r < n_reads; ++r) {
		const std::string seq(seqs + seq_off[r], seqs + seq_off[r + 1]);
		const int len = (int)seq.length();
		std::vector<Kmer> seqkmers;
		std::vector<int> seqminimizers;
		for (int j = 0; j <= len - kmer_len; ++j) {
			std::string kmerstrfromfastq = seq.substr(j, kmer_len);
			Kmer mykmer(kmerstrfromfastq.c_str(), kmerstrfromfastq.length());
			seqkmers.emplace_back(mykmer);
		}
		getMinimizers(bpars.windowLen, seqkmers, seqminimizers);
		for (auto minpos : seqminimizers) {
			std::string strminkmer = seq.substr(minpos, kmer_len);
			Kmer myminkmer(strminkmer.c_str(), strminkmer.length());
			myminkmer = myminkmer.rep();
			IT idx;
			if (countsreliable.find(myminkmer, idx)) {
				if (n < cap) { out_read[n] = r; out_pos[n] = (uint16_t)minpos; }
				++n;
			}
		}
	}
	*n_out = n;
	return n <= cap ? 0 : -1;
}

int bella_ref_max_threads(void) { return omp_get_max_threads(); }

} // extern "C"
