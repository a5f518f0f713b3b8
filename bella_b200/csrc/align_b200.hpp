// align_b200.hpp -- reference-side binding of the "next" row f1 (include/bella_xdrop.h).
//
// Header-only, to be included next to the reference's include/overlap.hpp (it uses the reference's own types readVector_,
// BELLApars, spmatPtr_, so it only compiles inside a BELLA translation unit).  It provides
//
//     RunPairWiseAlignments_b200(start, end, offset, colptrC, rowids, values, reads, filename, bpars, ratiophi)
//
// with the signature, output file format and returned statistics of the reference's RunPairWiseAlignments
// (include/overlap.hpp:499-646): for every nonzero of the block it takes the seed val->choose() (common.h:162-170), runs
// the gapped X-drop seed-and-extend of alignSeqAn (align.hpp:93-139; the algorithm LOGAN ports, loganGPU/functions.cuh) on the
// B200, applies PostAlignDecision (overlap.hpp:415-497) and appends the accepted pairs to `filename` in BELLA's or PAF
// format.  The skip-alignment branch (-z) is forwarded to the reference's own function.  Lines are written in column
// order; the reference's order is thread-schedule dependent, so consumers must not rely on either.
//
// The maintainer's patch: in HashSpGEMM (overlap.hpp:748) / the shim overlap_b200.hpp:126 replace the call
//     RunPairWiseAlignments(colStart[b], colStart[b+1], begnz, colptrC, rowids, values, reads, filename, bpars, ratiophi)
// by RunPairWiseAlignments_b200(...) with the same arguments, and link -lbella_xdrop.
#pragma once

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <string>
#include <tuple>
#include <vector>

#include "bella_xdrop.h"

namespace bella_b200_shim {

[[noreturn]] inline void die_xdrop(bella_xdrop* h, const char* what, int rc)
{
	std::fprintf(stderr, "bella_xdrop: %s failed (%d): %s\n", what, rc, h ? bella_xdrop_last_error(h) : "");
	std::exit(1);
}

// one aligner per process; the read set is uploaded once and stays on the device for every stage
inline bella_xdrop* xdrop_for(const readVector_& reads)
{
	static bella_xdrop* h = nullptr;
	static const readVector_* loaded = nullptr;
	static size_t loaded_n = 0;
	if (!h) {
		h = bella_xdrop_create(0);
		if (!h) { std::fprintf(stderr, "bella_xdrop: no usable B200 (there is no CPU fallback)\n"); std::exit(1); }
	}
	if (loaded != &reads || loaded_n != reads.size()) {
		std::vector<uint64_t> off(reads.size() + 1, 0);
		for (size_t i = 0; i < reads.size(); ++i) off[i + 1] = off[i] + reads[i].seq.size();
		std::string all;
		all.reserve(off.back());
		for (const auto& r : reads) all += r.seq;
		const int rc = bella_xdrop_set_reads(h, all.data(), off.data(), (uint32_t)reads.size());
		if (rc) die_xdrop(h, "bella_xdrop_set_reads", rc);
		loaded = &reads; loaded_n = reads.size();
	}
	return h;
}

} // namespace bella_b200_shim

template <typename IT, typename FT>
auto RunPairWiseAlignments_b200(IT start, IT end, IT offset, IT* colptrC, IT* rowids, FT* values, const readVector_& reads,
	char* filename, const BELLApars& bpars, const double& ratiophi)
{
	if (bpars.skipAlignment)                      // nothing to align: the reference's own writer
		return RunPairWiseAlignments(start, end, offset, colptrC, rowids, values, reads, filename, bpars, ratiophi);

	const size_t n_pairs = size_t(colptrC[end] - colptrC[start]);
	std::vector<uint32_t> rows(n_pairs), cols(n_pairs);
	std::vector<uint16_t> posH(n_pairs), posV(n_pairs);
	for (IT j = start; j < end; ++j)
		for (IT i = colptrC[j]; i < colptrC[j + 1]; ++i) {
			const size_t p = size_t(i - colptrC[start]);
			std::pair<int, int> kmer = values[i - offset]->choose();     // overlap.hpp:560-561
			rows[p] = rowids[i - offset]; cols[p] = j;
			posH[p] = (uint16_t)kmer.first; posV[p] = (uint16_t)kmer.second;
		}
	bella_xdrop* h = bella_b200_shim::xdrop_for(reads);
	int rc = bella_xdrop_set_params(h, bpars.kmerSize, bpars.xDrop, ratiophi, bpars.deltaChernoff, bpars.fixedThreshold);
	if (rc) bella_b200_shim::die_xdrop(h, "bella_xdrop_set_params", rc);
	std::vector<int32_t> out(n_pairs * BELLA_XDROP_OUT_FIELDS);
	rc = bella_xdrop_align(h, n_pairs, rows.data(), cols.data(), posH.data(), posV.data(), out.data());
	if (rc) bella_b200_shim::die_xdrop(h, "bella_xdrop_align", rc);

	size_t alignedpairs = 0, alignedbases = 0, totalreadlen = 0, totaloutputt = 0, totsuccbases = 0, totfailbases = 0;
	std::stringstream ss;
	for (IT j = start; j < end; ++j)
		for (IT i = colptrC[j]; i < colptrC[j + 1]; ++i) {
			const size_t p = size_t(i - colptrC[start]);
			const int32_t* o = out.data() + p * BELLA_XDROP_OUT_FIELDS;
			const readType_& read1 = reads[rows[p]];                      // H
			const readType_& read2 = reads[cols[p]];                      // V
			const unsigned short read1len = read1.seq.length(), read2len = read2.seq.length();
			const int score = o[0];
			int begpH = o[2], endpH = o[3];
			const int begpV = o[4], endpV = o[5];
			++alignedpairs;
			totalreadlen += size_t(read1len) + read2len;
			alignedbases += size_t(endpV - begpV);
			if (!o[7]) { totfailbases += size_t(endpV - begpV); continue; }
			if (!bpars.outputPaf) {                                        // overlap.hpp:470-473
				ss << read2.nametag << '\t' << read1.nametag << '\t' << values[i - offset]->count << '\t' << score << '\t' << o[6] << '\t'
				   << char(o[1]) << '\t' << begpV << '\t' << endpV << '\t' << read2len << '\t' << begpH << '\t' << endpH << '\t' << read1len << '\n';
			} else {                                                       // overlap.hpp:475-488
				const char pafstrand = o[1] == 'n' ? '+' : '-';
				if (pafstrand == '-') { const int tmp = begpH; begpH = read1len - endpH; endpH = read1len - tmp; }
				ss << read2.nametag << '\t' << read2len << '\t' << begpV << '\t' << endpV << '\t' << pafstrand << '\t' << read1.nametag << '\t'
				   << read1len << '\t' << begpH << '\t' << endpH << '\t' << score << '\t' << o[6] << '\t' << 255 << '\n';
			}
			++totaloutputt;
			totsuccbases += size_t(endpV - begpV);
		}
	const double t0 = omp_get_wtime();
	std::ofstream ofs(filename, std::ios::binary | std::ios::app);
	const std::string text = ss.str();
	ofs.write(text.data(), (std::streamsize)text.size());
	ofs.close();
	const double timeoutputt = omp_get_wtime() - t0;
	return std::make_tuple(alignedpairs, alignedbases, totalreadlen, totaloutputt, totsuccbases, totfailbases, timeoutputt);
}
