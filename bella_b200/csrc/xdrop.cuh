// "Next" row f1: batched gapped X-drop seed-and-extend (the step right after the overlap SpGEMM).
//
// What it replaces: the reference's per-pair alignSeqAn -> seqan::extendSeed(.., GappedXDrop()) (include/align.hpp:93-139,
// seqan/seqan/seeds/seeds_extension.h:622-843) and its CUDA port (loganGPU/functions.cuh:223-408, one 32-thread block per
// alignment, anti-diagonals in global memory).  Same recurrence, same trimming rules, same "longest extension" rules, so
// (score, strand, begH, endH, begV, endV) are identical per pair (the tests check it against the CPU restatement oracle_xdrop_align).
//
// Layout of one extension on the device (xd::Ext<G,T>): a group of G lanes owns the live window of the three anti-diagonals
// in REGISTERS.  Column c of the DP matrix lives in slot c mod (G*T) = lane + G*t, so a column never moves between lanes
// while the window slides; the left neighbour (column c-1) is one __shfl away.  The window [minCol-1, maxCol] of every step
// must fit G*T slots -- when it does not (rare: 0.1 % of the extensions at x = 7 with G*T = 32) the extension is handed to
// the wide path (xd::wide_extend: one warp, anti-diagonals in global scratch, any width).  Bases are staged through two
// small shared-memory rings per group (query and database segment), refilled G bytes at a time.
//
// This header is compiled twice: by nvcc into the kernels of bella_xdrop.cu, and by g++ with XD_EMULATE defined into the
// lane-thread emulator of tests/emu/xdrop_emu.cpp (32 host threads per warp, collectives through barriers), which checks
// this very source against the oracle on the CPU.
#pragma once
#include <stdint.h>
#include <limits.h>

#ifdef XD_EMULATE
#define XD_FN inline
#else
#define XD_FN __device__ __forceinline__
#endif

namespace xd {

constexpr int UNDEF = INT_MIN + 1;        // seeds_extension.h:646 (minValue - gapCost, gapCost = -1)
constexpr unsigned FULL = 0xffffffffu;

#ifndef XD_EMULATE
XD_FN int lane_id() { return (int)(threadIdx.x & 31); }
XD_FN int shfl(unsigned mask, int v, int src, int width) { return __shfl_sync(mask, v, src, width); }
XD_FN int rmax(unsigned mask, int v) { return __reduce_max_sync(mask, v); }
XD_FN int rmin(unsigned mask, int v) { return __reduce_min_sync(mask, v); }
XD_FN bool all(unsigned mask, bool p) { return __all_sync(mask, p) != 0; }
XD_FN void wsync(unsigned mask) { __syncwarp(mask); }
XD_FN int atomic_inc(int* p) { return atomicAdd(p, 1); }
XD_FN char ldg(const char* p) { return __ldg(p); }
#else
int lane_id();                              // provided by the emulator
int shfl(unsigned mask, int v, int src, int width);
int rmax(unsigned mask, int v);
int rmin(unsigned mask, int v);
bool all(unsigned mask, bool p);
void wsync(unsigned mask);
int atomic_inc(int* p);
inline char ldg(const char* p) { return *p; }
#endif

XD_FN int imax(int a, int b) { return a > b ? a : b; }
XD_FN int imin(int a, int b) { return a < b ? a : b; }
// Bases are Dna5 codes, as alignSeqAn sees them (seqan::Dna5String, align.hpp:97-100): SeqAn's char -> Dna5 table
// (seqan/basic/alphabet_residue_tabs.h:113-140: A/a 0, C/c 1, G/g 2, T/t/U/u 3, anything else 4 = N), N matches N, the
// reverse complement maps 0<->3, 1<->2, N->N.  The reads are encoded once when they are uploaded.
XD_FN char dna5(char c)
{
	const char l = c | 0x20;
	return l == 'a' ? 0 : l == 'c' ? 1 : l == 'g' ? 2 : (l == 't' || l == 'u') ? 3 : 4;
}
XD_FN char comp(char c) { return c < 4 ? (char)(3 - c) : c; }

// One direction of one pair.  Index t of a segment counts AWAY from the seed: base t of the query (V) segment is
// q[qstep * t], base r of the database (H) segment is d[dstep * r], complemented when the pair is on the reverse strand.
struct Segs {
	const char* q; const char* d;
	int qstep, dstep, dcomp;
	int qlen, dlen;
};

struct Pairs {                              // the candidate pairs, as the overlap SpGEMM emits them
	const uint32_t* rows; const uint32_t* cols;      // H = row read, V = column read; cols == nullptr: CSC form, see colptr
	const uint16_t* posH; const uint16_t* posV;      // seed k-mer
	const char* seqs; const uint64_t* seq_off;       // reads, concatenated, 1 byte per base, Dna5 codes (see dna5())
	int kmer_len, xdrop;
	int n_jobs;                                      // 2 per pair: job 2p = left, 2p+1 = right
	const uint32_t* colptr; int n_cols;              // CSC form: pair p belongs to the column c with colptr[c] <= p < colptr[c+1]
};

XD_FN uint32_t pair_col(const Pairs& P, int p)
{
	if (P.cols) return P.cols[p];
	int lo = 0, hi = P.n_cols;                       // colptr[lo] <= p < colptr[hi]
	while (hi - lo > 1) {
		const int mid = (lo + hi) >> 1;
		if (P.colptr[mid] <= (uint32_t)p) lo = mid; else hi = mid;
	}
	return (uint32_t)lo;
}

// per job: { score, H coordinate, V coordinate, flag }; left: begin positions, flag = reverse strand; right: end positions
struct JobResult { int score, posH, posV, flag; };

// Pair p, direction dir -> segments.  Every lane of the group computes the same values; `lane`/`G` only split the
// seed comparison (twin(seedH) == seedV, align.hpp:110-113).  Returns false when the seed does not fit its reads.
template <int G>
XD_FN bool make_segs(const Pairs& P, int job, unsigned mask, int lane, Segs& s, int& reverse, int& baseH, int& baseV)
{
	const int p = job >> 1, right = job & 1;
	const uint32_t col = pair_col(P, p);
	const uint64_t oh = P.seq_off[P.rows[p]], ov = P.seq_off[col];
	const char* H = P.seqs + oh; const int lenH = (int)(P.seq_off[P.rows[p] + 1] - oh);
	const char* V = P.seqs + ov; const int lenV = (int)(P.seq_off[col + 1] - ov);
	int i = P.posH[p];
	const int j = P.posV[p], k = P.kmer_len;
	if (i + k > lenH || j + k > lenV) return false;
	bool same = true;
	for (int t = lane; t < k; t += G) same = same && (comp(ldg(H + i + k - 1 - t)) == ldg(V + j + t));
	reverse = all(mask, same) ? 1 : 0;
	if (reverse) i = lenH - i - k;
	const int begH = i, endH = i + k, begV = j, endV = j + k;
	s.dcomp = reverse;
	if (!right) {                               // prefixes, EXTEND_LEFT (seeds_extension.h:812-823)
		s.qlen = begV; s.q = V + begV - 1; s.qstep = -1;
		s.dlen = begH;
		if (!reverse) { s.d = H + begH - 1; s.dstep = -1; } else { s.d = H + (lenH - begH); s.dstep = 1; }
		baseH = begH; baseV = begV;
	} else {                                    // suffixes, EXTEND_RIGHT (:825-840)
		s.qlen = lenV - endV; s.q = V + endV; s.qstep = 1;
		s.dlen = lenH - endH;
		if (!reverse) { s.d = H + endH; s.dstep = 1; } else { s.d = H + (lenH - 1 - endH); s.dstep = -1; }
		baseH = endH; baseV = endV;
	}
	return true;
}

XD_FN char load_q(const Segs& s, int t) { return ldg(s.q + (long)s.qstep * t); }
XD_FN char load_d(const Segs& s, int r) { const char c = ldg(s.d + (long)s.dstep * r); return s.dcomp ? comp(c) : c; }

// ------------------------------------------------------------------------------------------------------------------
// Register-resident extension: G lanes, T cells per lane.
// ------------------------------------------------------------------------------------------------------------------
template <int G, int T>
struct Ext {
	static constexpr int GT = G * T;
	static constexpr int RING = 2 * GT;         // >= GT + G - 3, power of two
	static_assert((G & (G - 1)) == 0 && (T & (T - 1)) == 0 && GT >= 4, "G, T powers of two");

	int v1[T], v2[T], v3[T];                    // anti-diagonals n-2, n-1, n at this lane's slots
	int n, minCol, maxCol, best;                // group-uniform state (seeds_extension.h:650-660)
	int off1, n1, off2, n2, off3, n3;
	int qhi, dhi;                               // bases [.. , qhi) / [.., dhi) of the segments are in the rings
	int rows, cols, xdrop;

	XD_FN void init(const Segs& s, int xdrop_, int lane)
	{
		rows = s.dlen + 1; cols = s.qlen + 1; xdrop = xdrop_;
		const int g0 = (1 > xdrop) ? UNDEF : -1;          // _initAntiDiags :462-485
#pragma unroll
		for (int t = 0; t < T; ++t) {
			const int slot = t * G + lane;
			v1[t] = UNDEF;
			v2[t] = slot == 0 ? 0 : UNDEF;
			v3[t] = slot <= 1 ? g0 : UNDEF;
		}
		n1 = 0; n2 = 1; n3 = 2; off1 = off2 = off3 = 0;
		minCol = 1; maxCol = 2; n = 1; best = 0; qhi = dhi = 0;
	}

	XD_FN bool active() const { return minCol < maxCol; }

	// One anti-diagonal (the body of the while loop, seeds_extension.h:662-745).  false: the window no longer fits.
	XD_FN bool step(const Segs& s, unsigned mask, int lane, char* qring, char* dring)
	{
		++n;
#pragma unroll
		for (int t = 0; t < T; ++t) { v1[t] = v2[t]; v2[t] = v3[t]; }
		n1 = n2; n2 = n3; off1 = off2; off2 = off3; off3 = minCol - 1;
		n3 = maxCol + 1 - off3;
		if (n3 > GT) return false;

		if (n - minCol - 1 >= dhi) {                      // database bases up to row n - minCol - 1 are needed
			wsync(mask);
			const int r = dhi + lane;
			if (r < s.dlen) dring[r & (RING - 1)] = load_d(s, r);
			dhi += G;
			wsync(mask);
		}
		if (maxCol - 2 >= qhi) {                          // query bases up to column maxCol - 1
			wsync(mask);
			const int c = qhi + lane;
			if (c < s.qlen) qring[c & (RING - 1)] = load_q(s, c);
			qhi += G;
			wsync(mask);
		}

		const int lim = best - xdrop;
		const bool edge = -n > lim;                       // antiDiagNo * gapCost > best - scoreDropOff (:507)
		int m = UNDEF, lo_fail = INT_MAX, hi_fail = INT_MIN;
#pragma unroll
		for (int t = 0; t < T; ++t) {
			const int slot = t * G + lane;
			const int c = off3 + ((slot - off3) & (GT - 1));          // this slot's column in [minCol-1, minCol-1+GT)
			int s2 = v2[t], s1 = v1[t];
			if (T > 1 && lane == G - 1) { s2 = v2[(t + T - 1) % T]; s1 = v1[(t + T - 1) % T]; }   // what lane 0 reads is slot-1
			const int up2 = shfl(mask, s2, (lane + G - 1) & (G - 1), G);   // a2[c-1]
			const int up1 = shfl(mask, s1, (lane + G - 1) & (G - 1), G);   // a1[c-1]
			int val = UNDEF;
			if (c >= minCol && c < maxCol) {
				const char qc = qring[(c - 1) & (RING - 1)], dc = dring[(n - c - 1) & (RING - 1)];
				int tmp = imax(up2, v2[t]) - 1;
				tmp = imax(tmp, up1 + (qc == dc ? 1 : -1));
				if (tmp >= lim) { val = tmp; m = imax(m, tmp); }
			} else if (edge && ((c == off3 && c == 0) || (c == maxCol && n == maxCol))) {
				val = -n;                                     // first column / first row of the matrix (:509-512)
			}
			v3[t] = val;
			if (c >= minCol && c <= maxCol && !(val == UNDEF && up2 == UNDEF)) lo_fail = imin(lo_fail, c);
			if (c >= off3 && c < maxCol && !(val == UNDEF && v2[t] == UNDEF)) hi_fail = imax(hi_fail, c);
		}
		best = imax(best, rmax(mask, m));
		const int lo = rmin(mask, lo_fail), hi = rmax(mask, hi_fail);
		const int newMin = imin(lo, maxCol + 1);             // :723-727
		const int newMax = (hi == INT_MIN ? off3 : hi + 1) + 1;   // :730-735
		minCol = imax(newMin, n + 2 - rows);                 // :739
		maxCol = imin(newMax, cols);                         // :741
		return true;
	}

	// value of anti-diagonal `v` (whose window started at column `off`) at column col
	XD_FN int at(const int (&v)[T], int col, unsigned mask, int lane) const
	{
		const int want = col & (GT - 1);
		int x = INT_MIN;
#pragma unroll
		for (int t = 0; t < T; ++t) if (t * G + lane == want) x = v[t];
		return rmax(mask, x);
	}

	// the longest extension (:747-843): score and the number of query columns / database rows it covers
	XD_FN int finish(unsigned mask, int lane, int& ext_cols, int& ext_rows) const
	{
		int lcol = n3 + off3 - 2, lrow = n - lcol, lscore = at(v3, lcol, mask, lane);
		if (lscore == UNDEF) {
			const int e2 = at(v2, off2 + n2 - 2, mask, lane);
			if (e2 != UNDEF) { lcol = n2 + off2 - 2; lrow = n - 1 - lcol; lscore = e2; }
			else if (n2 > 2) {
				const int e3 = at(v2, off2 + n2 - 3, mask, lane);
				if (e3 != UNDEF) { lcol = n2 + off2 - 3; lrow = n - 1 - lcol; lscore = e3; }
			}
		}
		if (lscore == UNDEF) {
			int mx = INT_MIN;
			int c1[T];
#pragma unroll
			for (int t = 0; t < T; ++t) {
				c1[t] = off1 + ((t * G + lane - off1) & (GT - 1));
				if (c1[t] - off1 < n1) mx = imax(mx, v1[t]);
			}
			mx = rmax(mask, mx);
			if (mx > lscore) {
				int cc = INT_MAX;
#pragma unroll
				for (int t = 0; t < T; ++t) if (c1[t] - off1 < n1 && v1[t] == mx) cc = imin(cc, c1[t]);
				lcol = rmin(mask, cc); lrow = n - 2 - lcol; lscore = mx;
			}
		}
		ext_cols = 0; ext_rows = 0;
		if (lscore != UNDEF) { ext_cols = lcol; ext_rows = lrow; }
		return lscore;
	}
};

// ------------------------------------------------------------------------------------------------------------------
// Wide path: one warp, anti-diagonals in global scratch (3 arrays of `cap` ints), any window width.
// ------------------------------------------------------------------------------------------------------------------
XD_FN int wide_extend(const Segs& s, int xdrop, int lane, int* scratch, int cap, int& ext_cols, int& ext_rows)
{
	const int cols = s.qlen + 1, rows = s.dlen + 1;
	int *a1 = scratch, *a2 = scratch + cap, *a3 = scratch + 2 * cap;
	int n1 = 0, n2 = 1, n3 = 2, off1 = 0, off2 = 0, off3 = 0;
	int minCol = 1, maxCol = 2, n = 1, best = 0;
	if (lane == 0) { a2[0] = 0; a3[0] = a3[1] = (1 > xdrop) ? UNDEF : -1; }
	wsync(FULL);
	while (minCol < maxCol) {
		++n;
		int* t = a1; a1 = a2; a2 = a3; a3 = t;
		n1 = n2; n2 = n3; off1 = off2; off2 = off3; off3 = minCol - 1;
		n3 = maxCol + 1 - off3;
		const int lim = best - xdrop;
		if (lane == 0) {
			const bool edge = -n > lim;
			a3[0] = (edge && off3 == 0) ? -n : UNDEF;
			a3[maxCol - off3] = (edge && n == maxCol) ? -n : UNDEF;
		}
		int m = UNDEF;
		for (int c = minCol + lane; c < maxCol; c += 32) {
			const char qc = load_q(s, c - 1), dc = load_d(s, n - c - 1);
			int tmp = imax(a2[c - off2 - 1], a2[c - off2]) - 1;
			tmp = imax(tmp, a1[c - off1 - 1] + (qc == dc ? 1 : -1));
			if (tmp < lim) tmp = UNDEF; else m = imax(m, tmp);
			a3[c - off3] = tmp;
		}
		best = imax(best, rmax(FULL, m));
		wsync(FULL);                                        // a3 is complete before anyone trims on it
		while (minCol - off3 < n3 && a3[minCol - off3] == UNDEF && minCol - off2 - 1 < n2 && a2[minCol - off2 - 1] == UNDEF) ++minCol;
		while (maxCol - off3 > 0 && a3[maxCol - off3 - 1] == UNDEF && a2[maxCol - off2 - 1] == UNDEF) --maxCol;
		++maxCol;
		minCol = imax(minCol, n + 2 - rows);
		maxCol = imin(maxCol, cols);
		wsync(FULL);                                        // everyone has read a1 before it is reused as a3
	}
	int lcol = n3 + off3 - 2, lrow = n - lcol, lscore = a3[lcol - off3];
	if (lscore == UNDEF) {
		if (a2[n2 - 2] != UNDEF) { lcol = n2 + off2 - 2; lrow = n - 1 - lcol; lscore = a2[lcol - off2]; }
		else if (n2 > 2 && a2[n2 - 3] != UNDEF) { lcol = n2 + off2 - 3; lrow = n - 1 - lcol; lscore = a2[lcol - off2]; }
	}
	if (lscore == UNDEF) {
		int mx = INT_MIN, cc = INT_MAX;
		for (int i = lane; i < n1; i += 32) mx = imax(mx, a1[i]);
		mx = rmax(FULL, mx);
		if (mx > lscore) {
			for (int i = lane; i < n1; i += 32) if (a1[i] == mx) cc = imin(cc, i + off1);
			lcol = rmin(FULL, cc); lrow = n - 2 - lcol; lscore = mx;
		}
	}
	wsync(FULL);                                            // the scratch may be reused by the next job
	ext_cols = 0; ext_rows = 0;
	if (lscore != UNDEF) { ext_cols = lcol; ext_rows = lrow; }
	return lscore;
}

XD_FN void store_result(JobResult* res, int job, int lane0, int score, int ec, int er, int baseH, int baseV, int reverse)
{
	if (!lane0) return;
	JobResult r;
	r.score = score;
	if (job & 1) { r.posH = baseH + er; r.posV = baseV + ec; r.flag = 0; }
	else { r.posH = baseH - er; r.posV = baseV - ec; r.flag = reverse; }
	res[job] = r;
}

// Per pair: the two halves joined (align.hpp:125-137) and the reference's adaptive-threshold test fused behind it
// (PostAlignDecision, overlap.hpp:415-462, with its unsigned-short arithmetic).
// out8 = { score, strand ('n' / 'c'), begH, endH, begV, endV, estimated overlap `ov`, passed }.
XD_FN void compose(const Pairs& P, const JobResult* res, int p, double ratiophi, double delta, int fixed_threshold, int32_t* out8)
{
	const JobResult L = res[2 * p], R = res[2 * p + 1];
	const int score = L.score + R.score + P.kmer_len;
	const int begH = L.posH, begV = L.posV, endH = R.posH, endV = R.posV;
	const int lenH = (int)(uint16_t)(P.seq_off[P.rows[p] + 1] - P.seq_off[P.rows[p]]);
	const uint32_t col = pair_col(P, p);
	const int lenV = (int)(uint16_t)(P.seq_off[col + 1] - P.seq_off[col]);
	const uint16_t ovV = (uint16_t)(endV - begV), ovH = (uint16_t)(endH - begH);
	const uint16_t minLeft = (uint16_t)imin(begV, begH), minRight = (uint16_t)imin(lenV - endV, lenH - endH);
	const uint16_t ov = (uint16_t)((int)minLeft + (int)minRight + ((int)ovV + (int)ovH) / 2);
	int passed;
	if (fixed_threshold == -1) {
		const float thr = (float)((1 - delta) * (ratiophi * (double)(float)ov));
		passed = (float)score >= thr;
	} else passed = score >= fixed_threshold;
	int32_t* o = out8 + 8 * (long)p;
	o[0] = score; o[1] = L.flag ? 'c' : 'n'; o[2] = begH; o[3] = endH; o[4] = begV; o[5] = endV; o[6] = ov; o[7] = passed;
}

struct Queue {
	int* next;                                  // job counter
	int* wide_count; int* wide_jobs;            // extensions whose window outgrew the registers
	int* bad;                                   // set when a seed does not fit its reads
};

// A warp of the register kernel: 32/G groups, each pulling jobs until the queue is empty.  `rings` = this warp's
// 32/G * 2 * RING bytes of shared memory.
template <int G, int T>
XD_FN void warp_main(const Pairs& P, const Queue& Q, JobResult* res, char* rings)
{
	typedef Ext<G, T> E;
	const int wl = lane_id(), grp = wl / G, lane = wl & (G - 1);
	const unsigned mask = G == 32 ? FULL : (((1u << (G & 31)) - 1u) << (grp * G));
	char* qring = rings + grp * 2 * E::RING;
	char* dring = qring + E::RING;
	E e;
	Segs s;
	int job = 0, reverse = 0, baseH = 0, baseV = 0;
	bool have = false, done = false;
	for (;;) {
		if (!have && !done) {
			int j = 0;
			if (lane == 0) j = atomic_inc(Q.next);
			job = shfl(mask, j, 0, G);
			if (job >= P.n_jobs) done = true;
			else if (!make_segs<G>(P, job, mask, lane, s, reverse, baseH, baseV)) {
				if (lane == 0) *Q.bad = 1;
				store_result(res, job, lane == 0, 0, 0, 0, 0, 0, 0);
			} else if (s.qlen == 0 || s.dlen == 0) {              // :635-636
				store_result(res, job, lane == 0, 0, 0, 0, baseH, baseV, reverse);
			} else {
				e.init(s, P.xdrop, lane);
				have = true;
			}
		}
		if (all(FULL, done)) break;
		if (have) {
			if (e.active()) {
				if (!e.step(s, mask, lane, qring, dring)) {
					if (lane == 0) Q.wide_jobs[atomic_inc(Q.wide_count)] = job;
					have = false;
				}
			} else {
				int ec, er;
				const int score = e.finish(mask, lane, ec, er);
				store_result(res, job, lane == 0, score, ec, er, baseH, baseV, reverse);
				have = false;
			}
		}
	}
}

// ------------------------------------------------------------------------------------------------------------------
// Thread-per-extension path, for small x (BELLA's default x = 7: the live window is ~9 columns, far too narrow to feed
// a warp).  Every lane runs its OWN extension with the plain sequential recurrence; 32 extensions advance per warp
// instruction and there is no cross-lane traffic at all.  Per thread, in shared memory (slot-major, thread-minor, so a
// warp's accesses never conflict on the score arrays): two anti-diagonals of W columns as rings indexed by column
// (a3 is written over a1 in place, going down the columns: a3[c] needs a1[c-1], which is overwritten only afterwards) and
// W bases of each segment.  What the final "longest extension" rule needs of a1 (its maximum and first position,
// seeds_extension.h:783-790) is tracked while the anti-diagonal is produced.  A window wider than W goes to the wide path.
// ------------------------------------------------------------------------------------------------------------------
template <int W>
struct ThreadExt {
	int* sc; char* ch; int nt, tid;              // sc: 2 * W * nt ints, ch: 2 * W * nt bytes of shared memory
	XD_FN int& S(int arr, int c) const { return sc[(arr * W + (c & (W - 1))) * nt + tid]; }
	XD_FN char& QR(int t) const { return ch[(t & (W - 1)) * nt + tid]; }
	XD_FN char& DR(int r) const { return ch[(W + (r & (W - 1))) * nt + tid]; }

	int ia, ib;                                  // arrays holding a1 (becomes a3) and a2
	int n, minCol, maxCol, best, off2, n2, off3, n3, qhi, dhi, rows, cols, xdrop;
	int m1, c1, m2, c2, m3, c3;                  // maximum / its first column of anti-diagonals n-2, n-1, n

	XD_FN void init(const Segs& s, int xdrop_)
	{
		rows = s.dlen + 1; cols = s.qlen + 1; xdrop = xdrop_;
		const int g0 = (1 > xdrop) ? UNDEF : -1;
		ia = 1; ib = 0;                                    // swapped at the start of the first step
		S(0, 0) = 0; S(1, 0) = g0; S(1, 1) = g0;
		m1 = UNDEF; c1 = 0; m2 = 0; c2 = 0; m3 = g0; c3 = 0;
		n2 = 1; n3 = 2; off2 = off3 = 0;
		minCol = 1; maxCol = 2; n = 1; best = 0; qhi = dhi = 0;
	}

	XD_FN bool active() const { return minCol < maxCol; }

	XD_FN bool step(const Segs& s)
	{
		++n;
		{ const int t = ia; ia = ib; ib = t; }
		m1 = m2; c1 = c2; m2 = m3; c2 = c3;
		n2 = n3; off2 = off3; off3 = minCol - 1;
		n3 = maxCol + 1 - off3;
		if (n3 > W) return false;
		while (dhi <= n - minCol - 1) { DR(dhi) = load_d(s, dhi); ++dhi; }
		while (qhi <= maxCol - 2) { QR(qhi) = load_q(s, qhi); ++qhi; }
		const int lim = best - xdrop;
		const bool edge = -n > lim;
		int v = (edge && n == maxCol) ? -n : UNDEF;        // first row of the matrix (:511)
		S(ia, maxCol) = v;
		m3 = v; c3 = maxCol;
		int a2c = S(ib, maxCol - 1);
		for (int c = maxCol - 1; c >= minCol; --c) {
			const int a2l = S(ib, c - 1), a1l = S(ia, c - 1);
			int tmp = imax(a2l, a2c) - 1;
			tmp = imax(tmp, a1l + (QR(c - 1) == DR(n - c - 1) ? 1 : -1));
			v = tmp < lim ? UNDEF : tmp;
			S(ia, c) = v;
			if (v >= m3) { m3 = v; c3 = c; }
			a2c = a2l;
		}
		v = (edge && off3 == 0) ? -n : UNDEF;              // first column of the matrix (:509)
		S(ia, off3) = v;
		if (v >= m3) { m3 = v; c3 = off3; }
		best = imax(best, m3);
		while (minCol - off3 < n3 && S(ia, minCol) == UNDEF && minCol - off2 - 1 < n2 && S(ib, minCol - 1) == UNDEF) ++minCol;
		while (maxCol - off3 > 0 && S(ia, maxCol - 1) == UNDEF && S(ib, maxCol - 1) == UNDEF) --maxCol;
		++maxCol;
		minCol = imax(minCol, n + 2 - rows);
		maxCol = imin(maxCol, cols);
		return true;
	}

	XD_FN int finish(int& ext_cols, int& ext_rows) const
	{
		int lcol = n3 + off3 - 2, lrow = n - lcol, lscore = S(ia, lcol);
		if (lscore == UNDEF) {
			const int e2 = S(ib, off2 + n2 - 2);
			if (e2 != UNDEF) { lcol = n2 + off2 - 2; lrow = n - 1 - lcol; lscore = e2; }
			else if (n2 > 2) {
				const int e3 = S(ib, off2 + n2 - 3);
				if (e3 != UNDEF) { lcol = n2 + off2 - 3; lrow = n - 1 - lcol; lscore = e3; }
			}
		}
		if (lscore == UNDEF && m1 > lscore) { lscore = m1; lcol = c1; lrow = n - 2 - lcol; }
		ext_cols = 0; ext_rows = 0;
		if (lscore != UNDEF) { ext_cols = lcol; ext_rows = lrow; }
		return lscore;
	}
};

// Second form of the thread-per-extension path (opt-in until measured): every cell is ONE 32-bit word
//     score << 6 | q code << 3 | d code
// that carries the two bases the cell was computed from -- the query base of its column and the database base of its
// row.  The cell (c, row) of anti-diagonal n reads a2[c-1] (same row: its d code) and a2[c] (same column: its q code), so
// the bases travel through the recurrence and the inner loop has no base loads at all: 2 shared loads + 1 store per
// cell.  Only the two window ends fetch a new base from the reads per anti-diagonal.  The thread count NT is a template
// constant (a power of two), so a ring slot's address is (c * NT) & (W * NT - 1).
template <int W, int NT>
struct ThreadExtP {
	static constexpr int PUNDEF = -(1 << 24);   // below every real score, far from overflow when shifted by 6
	int* sc; int tid;                            // sc: 2 * W * NT ints of shared memory
	XD_FN int& S(int arr, int c) const { return sc[arr * (W * NT) + ((c * NT) & (W * NT - 1)) + tid]; }
	static XD_FN int pack(int score, int q, int d) { return (int)(((unsigned)score << 6) | (unsigned)(q << 3) | (unsigned)d); }

	int ia, ib;
	int n, minCol, maxCol, best, off2, n2, off3, n3, rows, cols, xdrop;
	int m1, c1, m2, c2, m3, c3;

	XD_FN void init(const Segs& s, int xdrop_)
	{
		rows = s.dlen + 1; cols = s.qlen + 1; xdrop = xdrop_;
		const int g0 = (1 > xdrop) ? PUNDEF : -1;
		ia = 1; ib = 0;
		const int q0 = load_q(s, 0), d0 = load_d(s, 0);   // both segments are non-empty here
		S(0, 0) = pack(0, 0, 0);                         // cell (0,0)
		S(1, 0) = pack(g0, 0, d0);                       // cell (0,1): row 1 -> d[0]
		S(1, 1) = pack(g0, q0, 0);                       // cell (1,0): column 1 -> q[0]
		m1 = PUNDEF; c1 = 0; m2 = 0; c2 = 0; m3 = g0; c3 = 0;
		n2 = 1; n3 = 2; off2 = off3 = 0;
		minCol = 1; maxCol = 2; n = 1; best = 0;
	}

	XD_FN bool active() const { return minCol < maxCol; }

	XD_FN bool step(const Segs& s)
	{
		++n;
		{ const int t = ia; ia = ib; ib = t; }
		m1 = m2; c1 = c2; m2 = m3; c2 = c3;
		n2 = n3; off2 = off3; off3 = minCol - 1;
		n3 = maxCol + 1 - off3;
		if (n3 > W) return false;
		const int lim = best - xdrop;
		const bool edge = -n > lim;
		// running slot offset instead of one address computation per access: slot(c-1) = (slot(c) - NT) & (W*NT - 1)
		int* const A = sc + ia * (W * NT) + tid;
		const int* const B = sc + ib * (W * NT) + tid;
		int o = ((maxCol - 1) * NT) & (W * NT - 1);
		int a2c = B[o];                                  // cell (maxCol-1, n-maxCol): the row of the new top cell
		{
			const int qn = maxCol - 1 < s.qlen ? load_q(s, maxCol - 1) : 0;
			const int v = (edge && n == maxCol) ? -n : PUNDEF;
			A[(o + NT) & (W * NT - 1)] = pack(v, qn, a2c & 7);
			m3 = v; c3 = maxCol;
		}
		int s2c = a2c >> 6;
		for (int c = maxCol - 1; c >= minCol; --c) {
			const int oc = o;
			o = (o - NT) & (W * NT - 1);
			const int a2l = B[o], a1l = A[o];
			const int q = (a2c >> 3) & 7, d = a2l & 7, s2l = a2l >> 6;
			int tmp = imax(s2l, s2c) - 1;
			tmp = imax(tmp, (a1l >> 6) + (q == d ? 1 : -1));
			const int v = tmp < lim ? PUNDEF : tmp;
			A[oc] = pack(v, q, d);
			if (v >= m3) { m3 = v; c3 = c; }
			a2c = a2l; s2c = s2l;
		}
		{
			const int dn = n - minCol < s.dlen ? load_d(s, n - minCol) : 0;     // the new bottom row n - off3
			const int v = (edge && off3 == 0) ? -n : PUNDEF;
			A[o] = pack(v, (a2c >> 3) & 7, dn);             // o is the slot of off3 and a2c is a2[off3]: same column
			if (v >= m3) { m3 = v; c3 = off3; }
		}
		best = imax(best, m3);
		while (minCol - off3 < n3 && (S(ia, minCol) >> 6) == PUNDEF && minCol - off2 - 1 < n2 && (S(ib, minCol - 1) >> 6) == PUNDEF) ++minCol;
		while (maxCol - off3 > 0 && (S(ia, maxCol - 1) >> 6) == PUNDEF && (S(ib, maxCol - 1) >> 6) == PUNDEF) --maxCol;
		++maxCol;
		minCol = imax(minCol, n + 2 - rows);
		maxCol = imin(maxCol, cols);
		return true;
	}

	XD_FN int finish(int& ext_cols, int& ext_rows) const
	{
		int lcol = n3 + off3 - 2, lrow = n - lcol, lscore = S(ia, lcol) >> 6;
		if (lscore == PUNDEF) {
			const int e2 = S(ib, off2 + n2 - 2) >> 6;
			if (e2 != PUNDEF) { lcol = n2 + off2 - 2; lrow = n - 1 - lcol; lscore = e2; }
			else if (n2 > 2) {
				const int e3 = S(ib, off2 + n2 - 3) >> 6;
				if (e3 != PUNDEF) { lcol = n2 + off2 - 3; lrow = n - 1 - lcol; lscore = e3; }
			}
		}
		if (lscore == PUNDEF && m1 > lscore) { lscore = m1; lcol = c1; lrow = n - 2 - lcol; }
		ext_cols = 0; ext_rows = 0;
		if (lscore == PUNDEF) return UNDEF;
		ext_cols = lcol; ext_rows = lrow;
		return lscore;
	}
};

// Third form (opt-in until measured): ONE 32-bit word per column slot holds BOTH live anti-diagonals of that column,
//     [31:22] score field of half 1 | [21:12] score field of half 0 | [11:9] q code of the column | [8:6] d code of half 1 |
//     [5:3] d code of half 0
// so a cell costs one shared load and one store, and a thread needs W words (256 B at W = 64: 768 threads per SM).
// Scores are stored relative to the drop-off limit of their own anti-diagonal: field = score - lim + 4, field 0 =
// undefined (a defined score is >= lim when it is written, seeds_extension.h:704-710, and <= lim + x + 1).  At step n,
// a2 lives in half n & 1 and a3 is written over a1 in the other half.  Decoding with the limit of the CURRENT step is
// one add: an undefined cell decodes to <= -4, which no +1 brings back to >= 0, so it needs no special case.
template <int W, int NT>
struct ThreadExtQ {
	int* sc; int tid;                            // sc: W * NT ints of shared memory
	XD_FN int& S(int c) const { return sc[((c * NT) & (W * NT - 1)) + tid]; }

	int n, minCol, maxCol, best, off2, n2, off3, n3, rows, cols, xdrop;
	int lim2, lim3;                              // limits the anti-diagonals n-1 and n were encoded with
	int m1, c1, m2, c2, m3, c3;                  // absolute maximum / its first column of anti-diagonals n-2, n-1, n

	XD_FN void init(const Segs& s, int xdrop_)
	{
		rows = s.dlen + 1; cols = s.qlen + 1; xdrop = xdrop_;
		const int lim = -xdrop;                          // best = 0 for the two initial anti-diagonals
		const int g0f = (1 > xdrop) ? 0 : (-1 - lim + 4);
		const int q0 = load_q(s, 0), d0 = load_d(s, 0);
		// anti-diagonal 0 = cell (0,0) in half 1 (it is a1 at step 2); anti-diagonal 1 = cells (0,1), (1,0) in half 0
		S(0) = ((0 - lim + 4) << 22) | (g0f << 12) | (d0 << 3);
		S(1) = (g0f << 12) | (q0 << 9);
		lim2 = lim; lim3 = lim;
		m1 = UNDEF; c1 = 0; m2 = 0; c2 = 0; m3 = (1 > xdrop) ? UNDEF : -1; c3 = 0;
		n2 = 1; n3 = 2; off2 = off3 = 0;
		minCol = 1; maxCol = 2; n = 1; best = 0;
	}

	XD_FN bool active() const { return minCol < maxCol; }

	template <int P>                             // P = n & 1: the half that holds a2
	XD_FN void cells(const Segs& s, int lim)
	{
		constexpr int SH2 = P ? 22 : 12, SH3 = P ? 12 : 22;          // score fields of a2 and of a1/a3
		constexpr int DS2 = P ? 6 : 3, DS3 = P ? 3 : 6;              // their d codes
		constexpr unsigned KEEP = (1023u << SH2) | (7u << DS2) | (7u << 9);   // what a3's store leaves alone: a2's half and q
		const int dec2 = 4 + (lim - lim3), dec1 = 4 + (lim - lim2); // lim3 / lim2 still describe anti-diagonals n-1 / n-2 here
		const bool edge = -n > lim;
		int o = ((maxCol - 1) * NT) & (W * NT - 1);
		int* const A = sc + tid;
		unsigned wc = (unsigned)A[o];                   // column maxCol-1
		{
			const unsigned top = (unsigned)A[(o + NT) & (W * NT - 1)];
			const int qn = maxCol - 1 < s.qlen ? load_q(s, maxCol - 1) : 0;
			const int v = (edge && n == maxCol) ? -n : UNDEF;
			const unsigned f = v == UNDEF ? 0u : (unsigned)(v - lim + 4);
			A[(o + NT) & (W * NT - 1)] = (int)((top & (1023u << SH2)) | ((top & (7u << DS2))) | ((unsigned)qn << 9) | (f << SH3) | (((wc >> DS2) & 7u) << DS3));
			m3 = v; c3 = maxCol;
		}
		int x2c = (int)((wc >> SH2) & 1023u) - dec2;
		for (int c = maxCol - 1; c >= minCol; --c) {
			const int oc = o;
			o = (o - NT) & (W * NT - 1);
			const unsigned w = (unsigned)A[o];          // column c-1: a2[c-1] and a1[c-1]
			const int x2l = (int)((w >> SH2) & 1023u) - dec2, x1l = (int)((w >> SH3) & 1023u) - dec1;
			const unsigned q = (wc >> 9) & 7u, d = (w >> DS2) & 7u;
			int tmp = imax(x2l, x2c) - 1;
			tmp = imax(tmp, x1l + (q == d ? 1 : -1));
			const unsigned f = tmp < 0 ? 0u : (unsigned)(tmp + 4);
			A[oc] = (int)((wc & KEEP) | (f << SH3) | (d << DS3));
			const int v = tmp < 0 ? UNDEF : tmp + lim;
			if (v >= m3) { m3 = v; c3 = c; }
			wc = w; x2c = x2l;
		}
		{
			const int dn = n - minCol < s.dlen ? load_d(s, n - minCol) : 0;
			const int v = (edge && off3 == 0) ? -n : UNDEF;
			const unsigned f = v == UNDEF ? 0u : (unsigned)(v - lim + 4);
			A[o] = (int)((wc & KEEP) | (f << SH3) | ((unsigned)dn << DS3));     // o is the slot of off3, wc its word
			if (v >= m3) { m3 = v; c3 = off3; }
		}
	}

	XD_FN bool undef3(int c) const { return ((((unsigned)S(c)) >> ((n & 1) ? 12 : 22)) & 1023u) == 0; }   // a3[c]
	XD_FN bool undef2(int c) const { return ((((unsigned)S(c)) >> ((n & 1) ? 22 : 12)) & 1023u) == 0; }   // a2[c]

	XD_FN bool step(const Segs& s)
	{
		++n;
		m1 = m2; c1 = c2; m2 = m3; c2 = c3;
		n2 = n3; off2 = off3; off3 = minCol - 1;
		n3 = maxCol + 1 - off3;
		if (n3 > W) return false;
		const int lim = best - xdrop;
		if (n & 1) cells<1>(s, lim); else cells<0>(s, lim);
		lim2 = lim3; lim3 = lim;
		best = imax(best, m3);
		while (minCol - off3 < n3 && undef3(minCol) && minCol - off2 - 1 < n2 && undef2(minCol - 1)) ++minCol;
		while (maxCol - off3 > 0 && undef3(maxCol - 1) && undef2(maxCol - 1)) --maxCol;
		++maxCol;
		minCol = imax(minCol, n + 2 - rows);
		maxCol = imin(maxCol, cols);
		return true;
	}

	XD_FN int score3(int c) const { const unsigned f = (((unsigned)S(c)) >> ((n & 1) ? 12 : 22)) & 1023u; return f ? (int)f - 4 + lim3 : UNDEF; }
	XD_FN int score2(int c) const { const unsigned f = (((unsigned)S(c)) >> ((n & 1) ? 22 : 12)) & 1023u; return f ? (int)f - 4 + lim2 : UNDEF; }

	XD_FN int finish(int& ext_cols, int& ext_rows) const
	{
		int lcol = n3 + off3 - 2, lrow = n - lcol, lscore = score3(lcol);
		if (lscore == UNDEF) {
			const int e2 = score2(off2 + n2 - 2);
			if (e2 != UNDEF) { lcol = n2 + off2 - 2; lrow = n - 1 - lcol; lscore = e2; }
			else if (n2 > 2) {
				const int e3 = score2(off2 + n2 - 3);
				if (e3 != UNDEF) { lcol = n2 + off2 - 3; lrow = n - 1 - lcol; lscore = e3; }
			}
		}
		if (lscore == UNDEF && m1 > lscore) { lscore = m1; lcol = c1; lrow = n - 2 - lcol; }
		ext_cols = 0; ext_rows = 0;
		if (lscore != UNDEF) { ext_cols = lcol; ext_rows = lrow; }
		return lscore;
	}
};

// make_segs for a single thread (no group to split the seed comparison over)
XD_FN bool make_segs_thread(const Pairs& P, int job, Segs& s, int& reverse, int& baseH, int& baseV)
{
	const int p = job >> 1, right = job & 1;
	const uint32_t col = pair_col(P, p);
	const uint64_t oh = P.seq_off[P.rows[p]], ov = P.seq_off[col];
	const char* H = P.seqs + oh; const int lenH = (int)(P.seq_off[P.rows[p] + 1] - oh);
	const char* V = P.seqs + ov; const int lenV = (int)(P.seq_off[col + 1] - ov);
	int i = P.posH[p];
	const int j = P.posV[p], k = P.kmer_len;
	if (i + k > lenH || j + k > lenV) return false;
	reverse = 1;
	for (int t = 0; t < k; ++t) if (comp(ldg(H + i + k - 1 - t)) != ldg(V + j + t)) { reverse = 0; break; }
	if (reverse) i = lenH - i - k;
	const int begH = i, endH = i + k, begV = j, endV = j + k;
	s.dcomp = reverse;
	if (!right) {
		s.qlen = begV; s.q = V + begV - 1; s.qstep = -1;
		s.dlen = begH;
		if (!reverse) { s.d = H + begH - 1; s.dstep = -1; } else { s.d = H + (lenH - begH); s.dstep = 1; }
		baseH = begH; baseV = begV;
	} else {
		s.qlen = lenV - endV; s.q = V + endV; s.qstep = 1;
		s.dlen = lenH - endH;
		if (!reverse) { s.d = H + endH; s.dstep = 1; } else { s.d = H + (lenH - 1 - endH); s.dstep = -1; }
		baseH = endH; baseV = endV;
	}
	return true;
}

// A thread of the thread-per-extension kernels; the only warp-wide operation is the "everybody done" vote.
// order != nullptr: the jobs are taken in that order (longest expected extension first, see k_xdrop_estimate).
template <class E>
XD_FN void thread_loop(E& e, const Pairs& P, const Queue& Q, JobResult* res, const int* order)
{
	Segs s;
	int job = 0, reverse = 0, baseH = 0, baseV = 0;
	bool have = false, done = false;
	for (;;) {
		if (!have && !done) {
			job = atomic_inc(Q.next);
			if (job >= P.n_jobs) done = true;
			else {
				if (order) job = order[job];
				if (!make_segs_thread(P, job, s, reverse, baseH, baseV)) {
					*Q.bad = 1;
					store_result(res, job, true, 0, 0, 0, 0, 0, 0);
				} else if (s.qlen == 0 || s.dlen == 0) {
					store_result(res, job, true, 0, 0, 0, baseH, baseV, reverse);
				} else {
					e.init(s, P.xdrop);
					have = true;
				}
			}
		}
		if (all(FULL, done)) break;
		if (have) {
			if (e.active()) {
				if (!e.step(s)) { Q.wide_jobs[atomic_inc(Q.wide_count)] = job; have = false; }
			} else {
				int ec, er;
				const int score = e.finish(ec, er);
				store_result(res, job, true, score, ec, er, baseH, baseV, reverse);
				have = false;
			}
		}
	}
}

template <int W>
XD_FN void thread_main(const Pairs& P, const Queue& Q, JobResult* res, int* sc, char* ch, int nt, int tid, const int* order = nullptr)
{
	ThreadExt<W> e;
	e.sc = sc; e.ch = ch; e.nt = nt; e.tid = tid;
	thread_loop(e, P, Q, res, order);
}

template <int W, int NT>
XD_FN void thread_main_packed(const Pairs& P, const Queue& Q, JobResult* res, int* sc, int tid, const int* order = nullptr)
{
	ThreadExtP<W, NT> e;
	e.sc = sc; e.tid = tid;
	thread_loop(e, P, Q, res, order);
}

template <int W, int NT>
XD_FN void thread_main_two(const Pairs& P, const Queue& Q, JobResult* res, int* sc, int tid, const int* order = nullptr)
{
	ThreadExtQ<W, NT> e;
	e.sc = sc; e.tid = tid;
	thread_loop(e, P, Q, res, order);
}

// Expected length of a job = min(query segment, database segment): the sort key of the longest-first schedule.
XD_FN int job_estimate(const Pairs& P, int job)
{
	Segs s;
	int reverse, baseH, baseV;
	if (!make_segs_thread(P, job, s, reverse, baseH, baseV)) return 0;
	return imin(s.qlen, s.dlen);
}

// A warp of the wide kernel: jobs from the overflow list (or, with list == nullptr, every job).
XD_FN void wide_main(const Pairs& P, const int* list, int n_list, int* next, JobResult* res, int* scratch, int cap, int* bad)
{
	const int lane = lane_id();
	for (;;) {
		int j = 0;
		if (lane == 0) j = atomic_inc(next);
		j = shfl(FULL, j, 0, 32);
		if (j >= n_list) break;
		const int job = list ? list[j] : j;
		Segs s;
		int reverse, baseH, baseV, ec = 0, er = 0, score = 0;
		if (!make_segs<32>(P, job, FULL, lane, s, reverse, baseH, baseV)) {
			if (lane == 0) *bad = 1;
			store_result(res, job, lane == 0, 0, 0, 0, 0, 0, 0);
			continue;
		}
		if (s.qlen > 0 && s.dlen > 0) score = wide_extend(s, P.xdrop, lane, scratch, cap, ec, er);
		store_result(res, job, lane == 0, score, ec, er, baseH, baseV, reverse);
	}
}

}  // namespace xd
