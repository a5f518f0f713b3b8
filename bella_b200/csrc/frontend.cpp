// Host front end: builds the read x k-mer matrix the overlap SpGEMM consumes.
//
// This is the host-side mirror of the part of the reference that feeds the hot path
// (it is NOT on the GPU path and not part of the timed region):
//   * reliable k-mer selection, multiplicity in [lo,hi] over canonical k-mers
//       -- reference include/kmercount.hpp:466-677 (SplitCount).  The reference uses
//          HyperLogLog + Bloom + cuckoo hash; those are memory optimisations, the selected SET is
//          "every canonical k-mer whose total occurrence count is in [lo,hi]".  Here the same set is
//          found by bucketed sort-and-count.  K-mer ids are arbitrary in the reference (cuckoo
//          iteration order, kmercount.hpp:650-659); here id = rank in (bucket, k-mer) order.
//   * tuple emission (kmer_id, read_id, pos) per reliable occurrence  -- src/main.cpp:339-423
//   * B = Aᵀ (k-mers x reads, CSC): counting sort by read + per-column MergeDuplicates with the
//     reference's slot order ((id*107)&(ht-1), ht = pow2 >= max(16, column nnz before merge),
//     linear probing, insertion in position order, LAST duplicate position wins, compaction in
//     slot order)                                       -- src/CSC.cpp:301-479, src/main.cpp:476-480
//   * A = Bᵀ (reads x k-mers, CSC), rows ascending inside a column (what the reference's
//     csr2csc_atomic_nosort produces at one thread)     -- include/common/transpose.h:12-52
//   * per-nonzero strand bit: 1 iff the k-mer window in the read equals its canonical
//     representative.  oriented(pair) = (strandA == strandB) replaces the substring comparison of
//     include/chain.hpp:35-44; exact for upper-case ACGT reads (anything else is rejected here).
// It also contains the seeded read simulator used by tests and bench.py.
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <omp.h>

namespace {

struct SplitMix64 {
	uint64_t s;
	explicit SplitMix64(uint64_t seed) : s(seed) {}
	inline uint64_t next() {
		uint64_t z = (s += 0x9E3779B97F4A7C15ull);
		z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
		z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
		return z ^ (z >> 31);
	}
	inline double uniform() { return (next() >> 11) * (1.0 / 9007199254740992.0); }
	inline uint64_t below(uint64_t n) { return (uint64_t)(((__uint128_t)next() * n) >> 64); }
};

inline uint64_t mix64(uint64_t z) {
	z = (z ^ (z >> 33)) * 0xff51afd7ed558ccdull;
	z = (z ^ (z >> 33)) * 0xc4ceb9fe1a85ec53ull;
	return z ^ (z >> 33);
}

const char BASES[4] = {'A', 'C', 'G', 'T'};
inline int base_code(char c) {
	switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; default: return -1; }
}

struct Rec {            // one k-mer occurrence
	uint64_t kmer;      // canonical 2-bit code
	uint32_t read;
	uint16_t pos;
	uint16_t strand;    // 1 iff window == canonical
};

struct FrontEnd {
	uint32_t n = 0, m = 0;
	uint64_t nnz = 0, ntuples = 0;
	std::vector<uint32_t> B_colptr, B_rowids, A_colptr, A_rowids, read_len;
	std::vector<uint16_t> B_values, A_values;
	std::vector<uint8_t> B_strand, A_strand;   // bit-packed, array order of the respective matrix
	// raw tuples before de-duplication (for cross-checking the reference's CSC constructor)
	std::vector<uint32_t> t_kmer, t_read;
	std::vector<uint16_t> t_pos;
};

} // namespace

extern "C" {

// Uniform random genome + reads with substitution/insertion/deletion errors.
// Read r is generated from SplitMix64(seed, r) only, so the output does not depend on the thread
// count.  Returns 0 and malloc'd buffers (caller frees with bella_fe_free_buf).
int bella_fe_simulate(uint64_t genome_len, uint32_t n_reads, uint32_t read_len, double err,
		double f_sub, double f_ins, double f_del, uint64_t seed,
		char** seqs_out, uint64_t** offs_out)
{
	if (genome_len < 2 * (uint64_t)read_len + 64 || read_len == 0 || read_len > 65535) return -1;
	std::vector<char> genome(genome_len);
	const uint64_t BLK = 1 << 20;
	const int64_t nblk = (int64_t)((genome_len + BLK - 1) / BLK);
#pragma omp parallel for schedule(dynamic)
	for (int64_t b = 0; b < nblk; ++b) {
		SplitMix64 rng(mix64(seed * 0x100000001b3ull + 0xabcdef) ^ mix64((uint64_t)b));
		uint64_t e = std::min(genome_len, (uint64_t)(b + 1) * BLK);
		for (uint64_t i = (uint64_t)b * BLK; i < e; i += 32) {
			uint64_t w = rng.next();
			for (uint64_t j = i; j < std::min(e, i + 32); ++j, w >>= 2) genome[j] = BASES[w & 3];
		}
	}
	char* seqs = (char*)malloc((size_t)n_reads * read_len + 1);
	uint64_t* offs = (uint64_t*)malloc(sizeof(uint64_t) * ((size_t)n_reads + 1));
	if (!seqs || !offs) { free(seqs); free(offs); return -2; }
	for (uint64_t r = 0; r <= n_reads; ++r) offs[r] = r * read_len;
	const double p_sub = err * f_sub, p_ins = p_sub + err * f_ins, p_del = p_ins + err * f_del;
	const uint64_t span = (uint64_t)(read_len * 1.5) + 64;   // template window that always suffices
#pragma omp parallel for schedule(dynamic, 64)
	for (int64_t r = 0; r < (int64_t)n_reads; ++r) {
		SplitMix64 rng(mix64(seed) ^ mix64(0x5851f42d4c957f2dull * (uint64_t)(r + 1)));
		uint64_t start = rng.below(genome_len - span);
		bool rc = rng.next() & 1;
		char* out = seqs + (size_t)r * read_len;
		uint32_t o = 0;
		uint64_t t = 0;           // template cursor in [0, span)
		while (o < read_len) {
			// template base t on the chosen strand
			char tb;
			if (!rc) tb = genome[start + t];
			else { int c = base_code(genome[start + span - 1 - t]); tb = BASES[3 - c]; }
			double u = rng.uniform();
			if (u < p_sub) { int c = base_code(tb); out[o++] = BASES[(c + 1 + (int)rng.below(3)) & 3]; ++t; }
			else if (u < p_ins) { out[o++] = BASES[rng.below(4)]; }
			else if (u < p_del) { ++t; }
			else { out[o++] = tb; ++t; }
			if (t >= span) t = span - 1;   // cannot happen for err < ~0.3; keeps the access in bounds
		}
	}
	seqs[(size_t)n_reads * read_len] = 0;
	*seqs_out = seqs;
	*offs_out = offs;
	return 0;
}

void bella_fe_free_buf(void* p) { free(p); }

// Build A, B, strand bits and read lengths from reads.  Returns NULL on error (err_out: -1 bad
// arguments, -2 a read contains a character other than upper-case ACGT, -3 read longer than 65535).
// Order of a k-mer under the reference's minimizer sampling (include/minimizer.hpp:23-26): Kmer::rep().hash() = the canonical
// k-mer packed from the most significant end of one 64-bit word (kmercode/Kmer.cpp:206-222), hashed as 8 bytes with
// MurmurHash3_x64_128 (seed 313, first 64 bits: kmercode/hash_funcs.c:135-140; the published algorithm, here for len = 8).
static inline uint64_t rotl64_(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static inline uint64_t fmix64_(uint64_t k) { k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33; return k; }
static inline uint64_t kmer_order(uint64_t canon, int k)
{
	uint64_t h1 = 313, h2 = 313, k1 = canon << (64 - 2 * k);
	k1 *= 0x87c37b91114253d5ull; k1 = rotl64_(k1, 31); k1 *= 0x4cf5ad432745937full; h1 ^= k1;
	h1 ^= 8; h2 ^= 8;
	h1 += h2; h2 += h1;
	h1 = fmix64_(h1); h2 = fmix64_(h2);
	return h1 + h2;
}

void* bella_fe_build_w(const char* seqs, const uint64_t* offs, uint32_t n_reads, int k, int lo, int hi,
		int keep_tuples, int nthreads, int window, int* err_out);

void* bella_fe_build(const char* seqs, const uint64_t* offs, uint32_t n_reads, int k, int lo, int hi,
		int keep_tuples, int nthreads, int* err_out)
{
	return bella_fe_build_w(seqs, offs, n_reads, k, lo, hi, keep_tuples, nthreads, 0, err_out);
}

// window > 0: BELLA's minimizer mode (-w, src/main.cpp:363-388 + MinimizerCount, include/kmercount.hpp:690-832): only the
// positions getMinimizers samples (include/minimizer.hpp:49-77, robust winnowing) are counted and emitted.
void* bella_fe_build_w(const char* seqs, const uint64_t* offs, uint32_t n_reads, int k, int lo, int hi,
		int keep_tuples, int nthreads, int window, int* err_out)
{
	int err_dummy; if (!err_out) err_out = &err_dummy;
	*err_out = 0;
	if (k < 1 || k > 32 || lo < 1 || hi < lo) { *err_out = -1; return nullptr; }
	if (nthreads > 0) omp_set_num_threads(nthreads);
	int T = 1;
#pragma omp parallel
	{
#pragma omp single
		T = omp_get_num_threads();
	}
	const int NB_LOG = 9, NB = 1 << NB_LOG;      // buckets
	const uint64_t kmask = (k == 32) ? ~0ull : ((1ull << (2 * k)) - 1);
	FrontEnd* fe = new FrontEnd();
	fe->n = n_reads;
	fe->read_len.resize(n_reads);
	int bad = 0;
	for (uint32_t r = 0; r < n_reads; ++r) {
		uint64_t len = offs[r + 1] - offs[r];
		if (len > 65535) bad = -3;
		fe->read_len[r] = (uint32_t)len;
	}
	if (bad) { *err_out = bad; delete fe; return nullptr; }

	// pass 1: per-thread bucket histograms (static read partition, reused in pass 2)
	std::vector<std::vector<uint64_t>> hist(T, std::vector<uint64_t>(NB, 0));
	auto scan_read = [&](uint32_t r, auto&& emit) -> int {
		const char* s = seqs + offs[r];
		uint32_t len = fe->read_len[r];
		uint64_t fw = 0, rv = 0;
		uint32_t valid = 0;
		if (window <= 0) {
			for (uint32_t i = 0; i < len; ++i) {
				int c = base_code(s[i]);
				if (c < 0) return -2;
				fw = ((fw << 2) | (uint64_t)c) & kmask;
				rv = (rv >> 2) | ((uint64_t)(3 - c) << (2 * (k - 1)));
				if (++valid >= (uint32_t)k) {
					uint32_t pos = i + 1 - k;
					bool fwd_is_canon = fw <= rv;
					emit(fwd_is_canon ? fw : rv, pos, fwd_is_canon);
				}
			}
			return 0;
		}
		// minimizer mode: the monotone deque of getMinimizers over the orders of the read's k-mers.  The reference compares the
		// front's index with `int(i) - window` where window is a size_t, so the difference wraps for i < window and the deque
		// is emptied: the first `window` k-mers of a read are never sampled.  Reproduced (same unsigned arithmetic).
		if (len < (uint32_t)k) return 0;
		const int total = (int)len - k + 1;
		std::vector<uint64_t> canon(total), ord(total);
		std::vector<uint8_t> isf(total);
		for (uint32_t i = 0; i < len; ++i) {
			int c = base_code(s[i]);
			if (c < 0) return -2;
			fw = ((fw << 2) | (uint64_t)c) & kmask;
			rv = (rv >> 2) | ((uint64_t)(3 - c) << (2 * (k - 1)));
			if (++valid >= (uint32_t)k) {
				const int p = (int)i + 1 - k;
				isf[p] = fw <= rv;
				canon[p] = isf[p] ? fw : rv;
				ord[p] = kmer_order(canon[p], k);
			}
		}
		std::vector<int> dq(total);
		int head = 0, tail = 0, last = -1;
		for (int i = 0; i < total; ++i) {
			while (tail > head && ord[dq[tail - 1]] > ord[i]) --tail;
			dq[tail++] = i;
			while (tail > head && (uint64_t)(int64_t)dq[head] <= (uint64_t)(int64_t)i - (uint64_t)window) {
				while (tail - head > 1 && ord[dq[head]] == ord[dq[head + 1]]) ++head;      // furtherPop (robust winnowing)
				++head;
			}
			if (tail > head && dq[head] != last) { last = dq[head]; emit(canon[last], (uint32_t)last, (bool)isf[last]); }
		}
		return 0;
	};
#pragma omp parallel num_threads(T)
	{
		int t = omp_get_thread_num();
		uint32_t r0 = (uint64_t)n_reads * t / T, r1 = (uint64_t)n_reads * (t + 1) / T;
		auto& h = hist[t];
		for (uint32_t r = r0; r < r1; ++r) {
			int rc = scan_read(r, [&](uint64_t km, uint32_t, bool) { h[mix64(km) >> (64 - NB_LOG)]++; });
			if (rc) {
#pragma omp atomic write
				bad = rc;
			}
		}
	}
	if (bad) { *err_out = bad; delete fe; return nullptr; }
	std::vector<uint64_t> bstart(NB + 1, 0);
	{
		uint64_t acc = 0;
		for (int b = 0; b < NB; ++b) {
			bstart[b] = acc;
			for (int t = 0; t < T; ++t) { uint64_t c = hist[t][b]; hist[t][b] = acc; acc += c; }
		}
		bstart[NB] = acc;
	}
	const uint64_t total = bstart[NB];
	std::vector<Rec> recs(total);
#pragma omp parallel num_threads(T)
	{
		int t = omp_get_thread_num();
		uint32_t r0 = (uint64_t)n_reads * t / T, r1 = (uint64_t)n_reads * (t + 1) / T;
		auto& h = hist[t];
		for (uint32_t r = r0; r < r1; ++r)
			scan_read(r, [&](uint64_t km, uint32_t pos, bool st) {
				Rec& x = recs[h[mix64(km) >> (64 - NB_LOG)]++];
				x.kmer = km; x.read = r; x.pos = (uint16_t)pos; x.strand = st;
			});
	}
	// sort buckets, count reliable k-mers
	std::vector<uint64_t> brel(NB + 1, 0), btup(NB + 1, 0);
#pragma omp parallel for schedule(dynamic, 1)
	for (int b = 0; b < NB; ++b) {
		Rec* p = recs.data() + bstart[b];
		Rec* e = recs.data() + bstart[b + 1];
		std::sort(p, e, [](const Rec& a, const Rec& c) {
			if (a.kmer != c.kmer) return a.kmer < c.kmer;
			if (a.read != c.read) return a.read < c.read;
			return a.pos < c.pos;
		});
		uint64_t nrel = 0, ntup = 0;
		for (Rec* q = p; q < e;) {
			Rec* q2 = q + 1;
			while (q2 < e && q2->kmer == q->kmer) ++q2;
			uint64_t c = q2 - q;
			if (c >= (uint64_t)lo && c <= (uint64_t)hi) { ++nrel; ntup += c; }
			q = q2;
		}
		brel[b + 1] = nrel; btup[b + 1] = ntup;
	}
	for (int b = 0; b < NB; ++b) { brel[b + 1] += brel[b]; btup[b + 1] += btup[b]; }
	if (brel[NB] >= 0xFFFFFFFFull || btup[NB] >= 0xFFFFFFFFull) { *err_out = -4; delete fe; return nullptr; }
	fe->m = (uint32_t)brel[NB];
	fe->ntuples = btup[NB];
	const uint64_t NT_ = fe->ntuples;

	// tuples, bucketed by read via counting sort: colcount -> colptr0
	std::vector<uint32_t> cnt0((size_t)n_reads + 1, 0);
	struct Tup { uint32_t kid; uint16_t pos; uint16_t strand; };
	// first emit (kid, read, pos, strand) compactly in k-mer order
	std::vector<uint32_t> tk(NT_), tr(NT_);
	std::vector<uint16_t> tp(NT_), ts(NT_);
#pragma omp parallel for schedule(dynamic, 1)
	for (int b = 0; b < NB; ++b) {
		Rec* p = recs.data() + bstart[b];
		Rec* e = recs.data() + bstart[b + 1];
		uint64_t id = brel[b], w = btup[b];
		for (Rec* q = p; q < e;) {
			Rec* q2 = q + 1;
			while (q2 < e && q2->kmer == q->kmer) ++q2;
			uint64_t c = q2 - q;
			if (c >= (uint64_t)lo && c <= (uint64_t)hi) {
				for (Rec* x = q; x < q2; ++x, ++w) {
					tk[w] = (uint32_t)id; tr[w] = x->read; tp[w] = x->pos; ts[w] = x->strand;
				}
				++id;
			}
			q = q2;
		}
	}
	std::vector<Rec>().swap(recs);
	for (uint64_t w = 0; w < NT_; ++w) cnt0[tr[w] + 1]++;
	for (uint32_t r = 0; r < n_reads; ++r) cnt0[r + 1] += cnt0[r];
	std::vector<Tup> col((size_t)NT_);
	{
		std::vector<uint32_t> cur(cnt0.begin(), cnt0.end() - 1);
		for (uint64_t w = 0; w < NT_; ++w) { Tup& x = col[cur[tr[w]]++]; x.kid = tk[w]; x.pos = tp[w]; x.strand = ts[w]; }
	}
	if (keep_tuples) {
		// tuples in the reference's emission order: read-major, position ascending (main.cpp:393-416)
		fe->t_kmer.resize(NT_); fe->t_read.resize(NT_); fe->t_pos.resize(NT_);
	}
	std::vector<uint32_t>().swap(tk); std::vector<uint16_t>().swap(tp); std::vector<uint16_t>().swap(ts);
	std::vector<uint32_t>().swap(tr);

	// per read: order by position (emission order), then MergeDuplicates (CSC.cpp:301-420)
	std::vector<uint32_t> newcnt((size_t)n_reads + 1, 0);
#pragma omp parallel
	{
		std::vector<Tup> table;
#pragma omp for schedule(dynamic, 64)
		for (int64_t r = 0; r < (int64_t)n_reads; ++r) {
			Tup* p = col.data() + cnt0[r];
			Tup* e = col.data() + cnt0[r + 1];
			std::sort(p, e, [](const Tup& a, const Tup& c) { return a.pos < c.pos; });
			if (keep_tuples)
				for (Tup* x = p; x < e; ++x) {
					size_t w = x - col.data();
					fe->t_kmer[w] = x->kid; fe->t_read[w] = (uint32_t)r; fe->t_pos[w] = x->pos;
				}
			size_t cn = e - p;
			size_t ht = 16;
			while (ht < cn) ht <<= 1;
			table.assign(ht, Tup{0xFFFFFFFFu, 0, 0});
			for (Tup* x = p; x < e; ++x) {
				uint32_t h = (uint32_t)(x->kid * 107u) & (uint32_t)(ht - 1);   // u32 arithmetic (IT)
				for (;;) {
					if (table[h].kid == x->kid) { table[h].pos = x->pos; table[h].strand = x->strand; break; }  // addop returns p1 = new
					if (table[h].kid == 0xFFFFFFFFu) { table[h] = *x; break; }
					h = (h + 1) & (uint32_t)(ht - 1);
				}
			}
			uint32_t idx = 0;
			for (size_t s = 0; s < ht; ++s) if (table[s].kid != 0xFFFFFFFFu) p[idx++] = table[s];
			newcnt[r + 1] = idx;
		}
	}
	fe->B_colptr.assign((size_t)n_reads + 1, 0);
	for (uint32_t r = 0; r < n_reads; ++r) fe->B_colptr[r + 1] = fe->B_colptr[r] + newcnt[r + 1];
	fe->nnz = fe->B_colptr[n_reads];
	const uint64_t nz = fe->nnz;
	fe->B_rowids.resize(nz); fe->B_values.resize(nz); fe->B_strand.assign((nz + 7) / 8 + 8, 0);
	std::vector<uint8_t> bst(nz);
#pragma omp parallel for schedule(dynamic, 64)
	for (int64_t r = 0; r < (int64_t)n_reads; ++r) {
		const Tup* p = col.data() + cnt0[r];
		uint32_t o = fe->B_colptr[r], c = newcnt[r + 1];
		for (uint32_t i = 0; i < c; ++i) { fe->B_rowids[o + i] = p[i].kid; fe->B_values[o + i] = p[i].pos; bst[o + i] = (uint8_t)p[i].strand; }
	}
	std::vector<Tup>().swap(col);
	for (uint64_t i = 0; i < nz; ++i) if (bst[i]) fe->B_strand[i >> 3] |= (uint8_t)(1u << (i & 7));

	// A = Bᵀ, rows ascending within a column (serial scatter over reads in order)
	fe->A_colptr.assign((size_t)fe->m + 1, 0);
	for (uint64_t i = 0; i < nz; ++i) fe->A_colptr[fe->B_rowids[i] + 1]++;
	for (uint32_t c = 0; c < fe->m; ++c) fe->A_colptr[c + 1] += fe->A_colptr[c];
	fe->A_rowids.resize(nz); fe->A_values.resize(nz); fe->A_strand.assign((nz + 7) / 8 + 8, 0);
	{
		std::vector<uint32_t> cur(fe->A_colptr.begin(), fe->A_colptr.end() - 1);
		for (uint32_t r = 0; r < n_reads; ++r)
			for (uint32_t j = fe->B_colptr[r]; j < fe->B_colptr[r + 1]; ++j) {
				uint32_t d = cur[fe->B_rowids[j]]++;
				fe->A_rowids[d] = r; fe->A_values[d] = fe->B_values[j];
				if (bst[j]) fe->A_strand[d >> 3] |= (uint8_t)(1u << (d & 7));
			}
	}
	return fe;
}

uint32_t bella_fe_n(void* h) { return ((FrontEnd*)h)->n; }
uint32_t bella_fe_m(void* h) { return ((FrontEnd*)h)->m; }
uint64_t bella_fe_nnz(void* h) { return ((FrontEnd*)h)->nnz; }
uint64_t bella_fe_ntuples(void* h) { return ((FrontEnd*)h)->ntuples; }
// which: 0 B_colptr 1 B_rowids 2 B_values 3 B_strand 4 A_colptr 5 A_rowids 6 A_values 7 A_strand
//        8 read_len 9 t_kmer 10 t_read 11 t_pos
const void* bella_fe_array(void* h, int which)
{
	FrontEnd* fe = (FrontEnd*)h;
	switch (which) {
	case 0: return fe->B_colptr.data(); case 1: return fe->B_rowids.data(); case 2: return fe->B_values.data();
	case 3: return fe->B_strand.data(); case 4: return fe->A_colptr.data(); case 5: return fe->A_rowids.data();
	case 6: return fe->A_values.data(); case 7: return fe->A_strand.data(); case 8: return fe->read_len.data();
	case 9: return fe->t_kmer.data(); case 10: return fe->t_read.data(); case 11: return fe->t_pos.data();
	default: return nullptr;
	}
}
void bella_fe_free(void* h) { delete (FrontEnd*)h; }

} // extern "C"
