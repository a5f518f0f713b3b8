// bella_b200.cu -- B200 (sm_100a) overlap-detection SpGEMM  C = A * A^T  with BELLA's binning
// semiring, and the C-ABI of include/bella_b200.h.  See DESIGN.md for the data layout and the
// kernel-by-kernel roofline.  There is no CPU fallback in this file.
//
// Reference behaviour reproduced (bit-exact), by reference file:line:
//   estimateFLOP        include/overlap.hpp:157-202   -> k_pack_B (per-column kept-product count)
//   estimateNNZ_Hash    include/overlap.hpp:205-276   -> k_expand (distinct rows > col)
//   LocalSpGEMM         include/overlap.hpp:281-363   -> k_expand (group products by pair, in B-column
//                                                        order) + k_fold
//   multiop/overlapop   include/chain.hpp:47-86       -> overlap_estimate()
//   chainop             include/chain.hpp:100-150     -> fold_pair()
//   choose()            include/common/common.h:162-170 -> end of fold_pair()
#include "bella_b200.h"

#include <cuda_runtime.h>
#include <cub/cub.cuh>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace {

constexpr uint32_t EMPTY = 0xFFFFFFFFu;
constexpr int N_CLASSES = 4;                    // expand size classes, by products per column
constexpr uint32_t CLASS_LIMIT[3] = {2048, 4096, 8192};
constexpr int EXPAND_THREADS = 128;

// ---------------------------------------------------------------------------------------------
// Device data layout (DESIGN.md "Data layout in HBM")
//   Aent[k]  : u64 = row(32) | pos(16)<<32 | strand(1)<<48, A's columns sorted by row ascending
//   Bent[j]  : u64 = aoff(32) | pos(16)<<32 | cnt(15)<<48 | strand(1)<<63
//              aoff/cnt = the suffix of A's (sorted) column rowB[j] holding rows > this column,
//              i.e. exactly the products LocalSpGEMM keeps (overlap.hpp:315)
//   prod[q]  : u64 = h(16) | v(16)<<16 | jrank(16)<<32 | oriented(1)<<48   (before the fold)
//              u64 = h(16) | v(16)<<16 | bin_overlap(16)<<32 | label(16)<<48 (fold state)
// ---------------------------------------------------------------------------------------------

struct Params {
	uint32_t n, m;             // reads, k-mers
	uint32_t lo, hi;           // output column range of this handle
	uint32_t K, BIN;
	const uint32_t* B_colptr;
	const uint32_t* A_colptr;
	const uint32_t* read_len;
	const uint64_t* Aent;
	const uint64_t* Bent;
	uint32_t* flopC;           // [hi-lo+1]
	uint64_t* flopptr;         // [hi-lo+1]
	uint32_t* nnzC;            // [hi-lo+1]
	uint32_t* colptrC;         // [hi-lo+1]
	uint64_t* prod;            // [F]
	uint32_t* prow;            // [F] pair row id, at flopptr[col]+p
	uint2* pdesc;              // [F] {start within column region, length}
	uint32_t* rowsC;
	uint16_t* countC;
	uint16_t* posH;
	uint16_t* posV;
	uint16_t* aux;             // 3 per nnz
	int* err;
};

struct Meta {                  // small device->host record read after the symbolic kernels
	unsigned long long flops;
	unsigned int class_count[N_CLASSES];
	unsigned int max_flop;
	unsigned int pad;
};

__device__ __forceinline__ void set_err(int* err, int code) { atomicCAS(err, 0, code); }

// ---- layout kernels -------------------------------------------------------------------------

__global__ void k_count_deg(const uint32_t* __restrict__ Brow, uint64_t nnz, uint32_t* __restrict__ deg)
{
	for (uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j < nnz; j += (uint64_t)gridDim.x * blockDim.x)
		atomicAdd(&deg[Brow[j]], 1u);
}

__device__ __forceinline__ uint32_t getbit(const uint8_t* __restrict__ bits, uint64_t i)
{
	return (bits[i >> 3] >> (i & 7)) & 1u;
}

// A = B^T on the device: one warp per column (read) of B scatters its nonzeros into A's columns.
__global__ void k_transpose_fill(uint32_t n, const uint32_t* __restrict__ Bcolptr, const uint32_t* __restrict__ Brow,
		const uint16_t* __restrict__ Bval, const uint8_t* __restrict__ Bstrand,
		const uint32_t* __restrict__ Acolptr, uint32_t* __restrict__ cursor, uint64_t* __restrict__ Aent)
{
	uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
	for (uint32_t i = warp; i < n; i += nwarps) {
		uint32_t j1 = Bcolptr[i + 1];
		for (uint32_t j = Bcolptr[i] + lane; j < j1; j += 32) {
			uint32_t c = Brow[j];
			uint32_t slot = atomicAdd(&cursor[c], 1u);
			Aent[Acolptr[c] + slot] = (uint64_t)i | ((uint64_t)Bval[j] << 32) | ((uint64_t)getbit(Bstrand, j) << 48);
		}
	}
}

__global__ void k_pack_A(uint64_t nnz, const uint32_t* __restrict__ Arow, const uint16_t* __restrict__ Aval,
		const uint8_t* __restrict__ Astrand, uint64_t* __restrict__ Aent)
{
	for (uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; k < nnz; k += (uint64_t)gridDim.x * blockDim.x)
		Aent[k] = (uint64_t)Arow[k] | ((uint64_t)Aval[k] << 32) | ((uint64_t)getbit(Astrand, k) << 48);
}

// sort every column of A by row id (thread per column; columns hold 2..u entries, u = 8 by default)
__global__ void k_sort_A(uint32_t m, const uint32_t* __restrict__ Acolptr, uint64_t* __restrict__ Aent)
{
	for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < m; c += gridDim.x * blockDim.x) {
		uint32_t s = Acolptr[c], e = Acolptr[c + 1];
		for (uint32_t a = s + 1; a < e; ++a) {
			uint64_t x = Aent[a];
			uint32_t b = a;
			while (b > s) {
				uint64_t y = Aent[b - 1];
				if ((uint32_t)y <= (uint32_t)x) break;
				Aent[b] = y;
				--b;
			}
			if (b != a) Aent[b] = x;
		}
	}
}

// Bent + per-column kept-product count (== estimateFLOP, overlap.hpp:157-202): one warp per column.
__global__ void k_pack_B(uint32_t lo, uint32_t hi, const uint32_t* __restrict__ Bcolptr,
		const uint32_t* __restrict__ Brow, const uint16_t* __restrict__ Bval, const uint8_t* __restrict__ Bstrand,
		const uint32_t* __restrict__ Acolptr, const uint64_t* __restrict__ Aent,
		uint64_t* __restrict__ Bent, uint32_t* __restrict__ flopC, unsigned long long* __restrict__ flop64, int* err)
{
	uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
	for (uint32_t i = lo + warp; i < hi; i += nwarps) {
		uint32_t j0 = Bcolptr[i], j1 = Bcolptr[i + 1];
		if (j1 - j0 > 65535u) { if (lane == 0) set_err(err, BELLA_B200_ERR_RANGE); }
		unsigned long long f = 0;
		for (uint32_t j = j0 + lane; j < j1; j += 32) {
			uint32_t c = Brow[j];
			uint32_t s = Acolptr[c], e = Acolptr[c + 1];
			// upper_bound(row <= i) in the sorted column; the first 8 rows are fetched with independent
			// loads (one round trip covers every column when u <= 8)
			uint32_t a;
			{
				uint32_t le = 0;
#pragma unroll
				for (int q = 0; q < 8; ++q) {
					uint32_t r = (s + q < e) ? (uint32_t)Aent[s + q] : 0xFFFFFFFFu;
					le += (r <= i);
				}
				a = s + le;
				if (le == 8 && e - s > 8) {
					uint32_t b = e;
					while (a < b) { uint32_t mid = (a + b) >> 1; if ((uint32_t)Aent[mid] <= i) a = mid + 1; else b = mid; }
				}
			}
			uint32_t cnt = e - a;
			if (cnt > 32767u) { set_err(err, BELLA_B200_ERR_RANGE); cnt = 32767u; }
			f += cnt;
			Bent[j] = (uint64_t)a | ((uint64_t)Bval[j] << 32) | ((uint64_t)cnt << 48) | ((uint64_t)getbit(Bstrand, j) << 63);
		}
		for (int o = 16; o; o >>= 1) f += __shfl_xor_sync(0xFFFFFFFFu, f, o);
		if (lane == 0) {
			if (f > 0xFFFFFFFFull) { set_err(err, BELLA_B200_ERR_RANGE); f = 0xFFFFFFFFull; }
			flopC[i - lo] = (uint32_t)f;
			flop64[i - lo] = f;
		}
	}
}

__global__ void k_classify(uint32_t ncols, const uint32_t* __restrict__ flopC, uint32_t* __restrict__ lists, Meta* meta)
{
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < ncols; i += gridDim.x * blockDim.x) {
		uint32_t f = flopC[i];
		if (f == 0) continue;
		int c = f <= CLASS_LIMIT[0] ? 0 : f <= CLASS_LIMIT[1] ? 1 : f <= CLASS_LIMIT[2] ? 2 : 3;
		uint32_t idx = atomicAdd(&meta->class_count[c], 1u);
		lists[(size_t)c * ncols + idx] = i;
		if (c == 3) atomicMax(&meta->max_flop, f);
	}
}

__global__ void k_zero_u32(uint32_t* p, uint64_t n)
{
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = 0;
}

__global__ void k_set_total(Meta* meta, const uint64_t* flopptr, uint32_t ncols) { meta->flops = flopptr[ncols]; }

// ---- expand: gather, hash, group by pair ----------------------------------------------------

__device__ __forceinline__ uint32_t ht_insert(uint32_t* keys, uint32_t mask, int shift, uint32_t key)
{
	uint32_t h = (key * 0x9E3779B1u) >> shift;
	for (;;) {
		uint32_t k = *(volatile uint32_t*)(keys + h);
		if (k == EMPTY) {
			k = atomicCAS(keys + h, EMPTY, key);
			if (k == EMPTY) return h;
		}
		if (k == key) return h;
		h = (h + 1) & mask;
	}
}

__device__ __forceinline__ uint32_t ht_find(const uint32_t* keys, uint32_t mask, int shift, uint32_t key)
{
	uint32_t h = (key * 0x9E3779B1u) >> shift;
	while (keys[h] != key) h = (h + 1) & mask;
	return h;
}

// exclusive scan of a[0..n) in place (block-wide), returns nothing; s_tmp needs 33 words
__device__ void block_excl_scan(uint32_t* a, uint32_t n, uint32_t* s_tmp)
{
	const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
	if (tid == 0) s_tmp[32] = 0;
	__syncthreads();
	for (uint32_t base = 0; base < n; base += blockDim.x) {
		uint32_t idx = base + tid;
		uint32_t x = idx < n ? a[idx] : 0, v = x;
		for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xFFFFFFFFu, v, o); if (lane >= o) v += y; }
		if (lane == 31) s_tmp[wid] = v;
		__syncthreads();
		if (wid == 0) {
			uint32_t w = lane < nw ? s_tmp[lane] : 0, ws = w;
			for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xFFFFFFFFu, ws, o); if (lane >= o) ws += y; }
			s_tmp[lane] = ws - w;              // exclusive warp offsets
			if (lane == 31) s_tmp[33] = ws;    // tile total
		}
		__syncthreads();
		uint32_t carry = s_tmp[32];
		if (idx < n) a[idx] = v - x + s_tmp[wid] + carry;
		__syncthreads();
		if (tid == 0) s_tmp[32] = carry + s_tmp[33];
		__syncthreads();
	}
}

// One CTA per output column.  HTMAX = table capacity of the class; GLOBAL = tables live in a global
// slab (columns with more than CLASS_LIMIT[2] products) instead of shared memory.
template <bool GLOBAL>
__global__ void __launch_bounds__(256) k_expand(Params P, const uint32_t* __restrict__ list, uint32_t count,
		uint32_t htmax, uint32_t* __restrict__ slab)
{
	extern __shared__ uint32_t smem[];
	__shared__ uint32_t s_z;
	__shared__ uint32_t s_tmp[34];
	const uint32_t tid = threadIdx.x, nt = blockDim.x;
	uint32_t* tbl = GLOBAL ? slab + (size_t)blockIdx.x * 5 * htmax : smem;
	uint32_t* keys = tbl;
	uint32_t* val = tbl + htmax;
	uint32_t* skeys = tbl + 2 * (size_t)htmax;
	uint32_t* poff = tbl + 3 * (size_t)htmax;
	uint32_t* cursor = tbl + 4 * (size_t)htmax;

	for (uint32_t it = blockIdx.x; it < count; it += gridDim.x) {
		const uint32_t li = list[it];
		const uint32_t i = P.lo + li;
		const uint32_t j0 = P.B_colptr[i], j1 = P.B_colptr[i + 1];
		const uint64_t base = P.flopptr[li];
		const uint32_t Fi = P.flopC[li];
		uint32_t ht = 32; int shift = 27;
		while (ht < Fi) { ht <<= 1; --shift; }
		const uint32_t mask = ht - 1;
		for (uint32_t s = tid; s < ht; s += nt) { keys[s] = EMPTY; val[s] = 0; }
		if (tid == 0) s_z = 0;
		__syncthreads();

		// pass 1: distinct rows + products per row
		for (uint32_t j = j0 + tid; j < j1; j += nt) {
			uint64_t be = P.Bent[j];
			uint32_t aoff = (uint32_t)be, cnt = (uint32_t)(be >> 48) & 0x7FFFu;
			for (uint32_t e = 0; e < cnt; ++e) {
				uint32_t key = (uint32_t)P.Aent[aoff + e];
				uint32_t slot = ht_insert(keys, mask, shift, key);
				atomicAdd(&val[slot], 1u);
			}
		}
		__syncthreads();
		// compact the distinct rows
		for (uint32_t s = tid; s < ht; s += nt) {
			uint32_t k = keys[s];
			if (k != EMPTY) skeys[atomicAdd(&s_z, 1u)] = k;
		}
		__syncthreads();
		const uint32_t Z = s_z;
		uint32_t Zp = 1;
		while (Zp < Z) Zp <<= 1;
		for (uint32_t s = Z + tid; s < Zp; s += nt) skeys[s] = EMPTY;
		__syncthreads();
		// bitonic sort ascending
		for (uint32_t k = 2; k <= Zp; k <<= 1)
			for (uint32_t j = k >> 1; j > 0; j >>= 1) {
				for (uint32_t x = tid; x < Zp; x += nt) {
					uint32_t y = x ^ j;
					if (y > x) {
						uint32_t a = skeys[x], b = skeys[y];
						bool up = (x & k) == 0;
						if ((a > b) == up) { skeys[x] = b; skeys[y] = a; }
					}
				}
				__syncthreads();
			}
		// pair p = rank of its row; per-pair product counts in rank order
		for (uint32_t p = tid; p < Z; p += nt) {
			uint32_t slot = ht_find(keys, mask, shift, skeys[p]);
			poff[p] = val[slot];
			val[slot] = p;
			cursor[p] = 0;
		}
		__syncthreads();
		for (uint32_t p = tid; p < Z; p += nt) P.pdesc[base + p].y = poff[p];
		block_excl_scan(poff, Z, s_tmp);
		for (uint32_t p = tid; p < Z; p += nt) {
			P.prow[base + p] = skeys[p];
			P.pdesc[base + p].x = poff[p];
		}
		if (tid == 0) P.nnzC[li] = Z;
		// pass 2: place every product into its pair's list (unordered inside the pair; the fold
		// orders by jrank)
		for (uint32_t j = j0 + tid; j < j1; j += nt) {
			uint64_t be = P.Bent[j];
			uint32_t aoff = (uint32_t)be, cnt = (uint32_t)(be >> 48) & 0x7FFFu;
			uint32_t v = (uint32_t)(be >> 32) & 0xFFFFu, sB = (uint32_t)(be >> 63);
			for (uint32_t e = 0; e < cnt; ++e) {
				uint64_t ae = P.Aent[aoff + e];
				uint32_t slot = ht_find(keys, mask, shift, (uint32_t)ae);
				uint32_t p = val[slot];
				uint32_t pos = poff[p] + atomicAdd(&cursor[p], 1u);
				uint32_t h = (uint32_t)(ae >> 32) & 0xFFFFu, sA = (uint32_t)(ae >> 48) & 1u;
				P.prod[base + pos] = (uint64_t)h | ((uint64_t)v << 16) | ((uint64_t)(j - j0) << 32) | ((uint64_t)(sA == sB) << 48);
			}
		}
		__syncthreads();
	}
}

// ---- fold: the semiring -------------------------------------------------------------------------

// multiop -> overlapop (chain.hpp:47-71), checkstrand replaced by the strand-bit comparison.
__device__ __forceinline__ uint32_t overlap_estimate(int lenH, int lenV, uint32_t h, uint32_t v, uint32_t oriented, uint32_t K)
{
	uint32_t hh = oriented ? h : ((uint32_t)lenH - h - K) & 0xFFFFu;   // unsigned short begpH, wraps
	uint32_t endH = (hh + K) & 0xFFFFu, endV = (v + K) & 0xFFFFu;
	int m1 = (int)min(hh, v);
	int m2 = min(lenH - (int)endH, lenV - (int)endV);
	return (uint32_t)(m1 + m2 + (int)K) & 0xFFFFu;                      // stored into vector<unsigned short>
}

// Sequential fold of one pair's products, in place.  State per processed product s:
//   bin_overlap = overlap value of the bin s currently belongs to, label = index of the product that
//   created that bin (0xFFFF = dropped).  chainop (chain.hpp:100-150) with m1 = the fresh one-k-mer
//   value: every stored k-mer of a bin within binSize of the new overlap either moves to the new
//   bin (if farther than K on both axes) or is dropped; other bins are untouched; the new bin goes
//   to the front, so bins are ordered by creator index descending.
__device__ void fold_pair(uint64_t* rec, uint32_t np, int lenH, int lenV, uint32_t K, int BIN,
		uint32_t& out_count, uint32_t& out_h, uint32_t& out_v, uint32_t& out_nbins, uint32_t& out_sup, uint32_t& out_ov)
{
	if (np == 1) {
		uint64_t r = rec[0];
		out_h = (uint32_t)r & 0xFFFFu; out_v = (uint32_t)(r >> 16) & 0xFFFFu;
		out_count = 1; out_nbins = 1; out_sup = 1;
		out_ov = overlap_estimate(lenH, lenV, out_h, out_v, (uint32_t)(r >> 48) & 1u, K);
		return;
	}
	uint32_t count = 0;
	for (uint32_t t = 0; t < np; ++t) {
		// next product in B-column order (smallest jrank among the unprocessed)
		uint64_t rb = rec[t];
		uint32_t qb = (uint32_t)(rb >> 32) & 0xFFFFu, best = t;
		for (uint32_t s = t + 1; s < np; ++s) {
			uint64_t r = rec[s];
			uint32_t q = (uint32_t)(r >> 32) & 0xFFFFu;
			if (q < qb) { qb = q; best = s; rb = r; }
		}
		if (best != t) rec[best] = rec[t];
		const uint32_t h = (uint32_t)rb & 0xFFFFu, v = (uint32_t)(rb >> 16) & 0xFFFFu;
		const uint32_t ov = overlap_estimate(lenH, lenV, h, v, (uint32_t)(rb >> 48) & 1u, K);
		uint32_t nrel = 0;
		for (uint32_t s = 0; s < t; ++s) {
			uint64_t r = rec[s];
			uint32_t lab = (uint32_t)(r >> 48);
			if (lab == 0xFFFFu) continue;
			int bo = (int)((uint32_t)(r >> 32) & 0xFFFFu);
			if (abs(bo - (int)ov) < BIN) {                                   // chain.hpp:114
				int hs = (int)((uint32_t)r & 0xFFFFu), vs = (int)((uint32_t)(r >> 16) & 0xFFFFu);
				if (abs((int)h - hs) > (int)K && abs((int)v - vs) > (int)K) { // chain.hpp:121
					rec[s] = (r & 0xFFFFFFFFull) | ((uint64_t)ov << 32) | ((uint64_t)t << 48);
					++nrel;
				} else {
					rec[s] = r | (0xFFFFull << 48);
				}
			}
		}
		count = t == 0 ? 1u : (((1u + count) & 0xFFFFu) + nrel) & 0xFFFFu;   // chain.hpp:105,140
		rec[t] = (uint64_t)h | ((uint64_t)v << 16) | ((uint64_t)ov << 32) | ((uint64_t)t << 48);
	}
	// choose(): most supported bin, ties -> lowest bin index = most recent creator (common.h:162-170)
	uint32_t best_sup = 0, best_c = 0, nbins = 0;
	for (uint32_t c = np; c-- > 0;) {
		uint64_t r = rec[c];
		if ((uint32_t)(r >> 48) != c) continue;
		++nbins;
		uint32_t sup = 0;
		for (uint32_t s = 0; s <= c; ++s) sup += ((uint32_t)(rec[s] >> 48) == c);
		if (sup > best_sup) { best_sup = sup; best_c = c; }
	}
	uint64_t r = rec[best_c];
	out_count = count; out_h = (uint32_t)r & 0xFFFFu; out_v = (uint32_t)(r >> 16) & 0xFFFFu;
	out_nbins = nbins; out_sup = best_sup & 0xFFFFu; out_ov = (uint32_t)(r >> 32) & 0xFFFFu;
}

// ---- fold, restated without mutable per-k-mer state ---------------------------------------------
// chainop only ever merges whole bins: whether bin b is absorbed at step t depends on the overlap
// values alone (|ov_b - ov_t| < binSize, chain.hpp:114), never on the k-mers.  So the bins form a
// forest: parent[b] = the first later product whose overlap is within binSize of bin b's overlap
// (bin b's overlap is the overlap of the product that created it).  A k-mer s then meets exactly
// its ancestors, in order, and is dropped at the first ancestor a with |dh| <= K or |dv| <= K
// (chain.hpp:121).  With c_s = number of ancestors s passes:
//     count   = (P + sum_s c_s) mod 2^16            (chain.hpp:105,140)
//     bins    = roots of the forest; support(root) = number of k-mers that reach it (the creator included)
//     choose  = root with the largest support, ties -> the most recent one (bin order = newest first)
// When every consecutive pair of overlaps is within binSize (the common case) the forest is the
// chain t -> t+1 and the whole fold is an all-pairs test with no sequential dependency.

constexpr int NBUCKETS = 7;            // P==1 | 2..4 | 5..8 | 9..16 | 17..32 | 33..256 (warp) | >256 (in-place)
constexpr uint32_t NONE16 = 0xFFFFu;

struct FDesc { uint32_t row, col; unsigned long long off_len; };   // off(48) | len(16)<<48

struct FoldMeta { unsigned int count[8]; };

__device__ __forceinline__ int bucket_of(uint32_t len)
{
	return len == 1 ? 0 : len <= 4 ? 1 : len <= 8 ? 2 : len <= 16 ? 3 : len <= 32 ? 4 : len <= 256 ? 5 : 6;
}

// one warp per column: flat pair descriptors at their final output index + per-bucket work lists
__global__ void k_flatten(Params P, FDesc* __restrict__ fdesc, uint32_t* __restrict__ lists, uint64_t list_cap, FoldMeta* meta)
{
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
	const uint32_t ncols = P.hi - P.lo;
	for (uint32_t li = warp; li < ncols; li += nwarps) {
		const uint32_t Z = P.nnzC[li];
		if (!Z) continue;
		const uint64_t base = P.flopptr[li];
		const uint32_t out0 = P.colptrC[li];
		for (uint32_t p0 = 0; p0 < Z; p0 += 32) {
			uint32_t p = p0 + lane;
			int b = -1;
			uint32_t g = out0 + p;
			if (p < Z) {
				uint2 d = P.pdesc[base + p];
				FDesc f;
				f.row = P.prow[base + p]; f.col = P.lo + li;
				f.off_len = (base + d.x) | ((unsigned long long)d.y << 48);
				fdesc[g] = f;
				b = bucket_of(d.y);
			}
			for (int k = 0; k < NBUCKETS; ++k) {
				uint32_t m = __ballot_sync(0xFFFFFFFFu, b == k);
				if (!m) continue;
				uint32_t start = 0;
				if (lane == (uint32_t)(__ffs(m) - 1)) start = atomicAdd(&meta->count[k], (unsigned)__popc(m));
				start = __shfl_sync(0xFFFFFFFFu, start, __ffs(m) - 1);
				if (b == k) lists[(uint64_t)k * list_cap + start + __popc(m & ((1u << lane) - 1))] = g;
			}
		}
	}
}

__device__ __forceinline__ void store_result(const Params& P, uint32_t g, uint32_t row, uint32_t cnt, uint32_t hv,
		uint32_t nb, uint32_t sup, uint32_t ov)
{
	P.rowsC[g] = row;
	P.countC[g] = (uint16_t)cnt;
	P.posH[g] = (uint16_t)(hv & 0xFFFFu);
	P.posV[g] = (uint16_t)(hv >> 16);
	P.aux[3 * (size_t)g + 0] = (uint16_t)nb;
	P.aux[3 * (size_t)g + 1] = (uint16_t)sup;
	P.aux[3 * (size_t)g + 2] = (uint16_t)ov;
}

__device__ __forceinline__ bool is_far(uint32_t x, uint32_t A, uint32_t B, uint32_t K2)
{
	// |h_t - h_s| > K  <=>  (unsigned)(h_t - h_s + K) > 2K ; A = K - h_s, B = K - v_s
	return ((x & 0xFFFFu) + A) > K2 && ((x >> 16) + B) > K2;
}

// thread per pair, P <= CAP
template <int CAP>
__global__ void __launch_bounds__(128) k_fold_short(Params P, const FDesc* __restrict__ fdesc, const uint32_t* __restrict__ list,
		const unsigned int* __restrict__ count_ptr)
{
	const uint32_t count = *count_ptr;
	const uint32_t K = P.K, K2 = 2 * P.K;
	const int BIN = (int)P.BIN;
	for (uint32_t it = blockIdx.x * blockDim.x + threadIdx.x; it < count; it += gridDim.x * blockDim.x) {
		const uint32_t g = list[it];
		const FDesc f = fdesc[g];
		const uint32_t np = (uint32_t)(f.off_len >> 48);
		const uint64_t* rec = P.prod + (f.off_len & 0xFFFFFFFFFFFFull);
		const int lenH = (int)P.read_len[f.row], lenV = (int)P.read_len[f.col];
		if (CAP == 1) {
			uint64_t r = rec[0];
			uint32_t hv = (uint32_t)r;
			store_result(P, g, f.row, 1, hv, 1, 1, overlap_estimate(lenH, lenV, hv & 0xFFFFu, hv >> 16, (uint32_t)(r >> 48) & 1u, K));
			continue;
		}
		uint32_t hv[CAP];
		uint16_t ov[CAP], jr[CAP];
		for (uint32_t a = 0; a < np; ++a) {          // insertion sort into B-column order
			uint64_t r = rec[a];
			uint32_t x = (uint32_t)r;
			uint16_t key = (uint16_t)(r >> 32);
			uint16_t o = (uint16_t)overlap_estimate(lenH, lenV, x & 0xFFFFu, x >> 16, (uint32_t)(r >> 48) & 1u, K);
			uint32_t b = a;
			while (b > 0 && jr[b - 1] > key) { jr[b] = jr[b - 1]; hv[b] = hv[b - 1]; ov[b] = ov[b - 1]; --b; }
			jr[b] = key; hv[b] = x; ov[b] = o;
		}
		bool linear = true;
		for (uint32_t t = 1; t < np; ++t) linear &= abs((int)ov[t] - (int)ov[t - 1]) < BIN;
		uint32_t csum = 0;
		if (linear) {
			uint32_t surv = 0;
			for (uint32_t s = 0; s < np; ++s) {
				uint32_t x = hv[s], A = K - (x & 0xFFFFu), B = K - (x >> 16), t = s + 1;
				while (t < np && is_far(hv[t], A, B, K2)) ++t;
				csum += t - s - 1;
				surv += (t == np);
			}
			store_result(P, g, f.row, (np + csum) & 0xFFFFu, hv[np - 1], 1, surv, ov[np - 1]);
		} else {
			uint16_t par[CAP], sup[CAP];
			uint32_t nlive = 0;                       // jr[] is free now: reuse it as the live-bin list
			for (uint32_t t = 0; t < np; ++t) {
				for (uint32_t q = 0; q < nlive;) {
					uint32_t b = jr[q];
					if (abs((int)ov[b] - (int)ov[t]) < BIN) { par[b] = (uint16_t)t; jr[q] = jr[--nlive]; } else ++q;
				}
				jr[nlive++] = (uint16_t)t;
				sup[t] = 0;
			}
			for (uint32_t q = 0; q < nlive; ++q) par[jr[q]] = NONE16;
			for (uint32_t s = 0; s < np; ++s) {
				uint32_t x = hv[s], A = K - (x & 0xFFFFu), B = K - (x >> 16);
				uint32_t a = par[s], last = s;
				while (a != NONE16 && is_far(hv[a], A, B, K2)) { ++csum; last = a; a = par[a]; }
				if (a == NONE16) ++sup[last];
			}
			uint32_t best = 0, bt = 0;
			for (uint32_t t = 0; t < np; ++t) if (par[t] == NONE16 && sup[t] >= best) { best = sup[t]; bt = t; }
			store_result(P, g, f.row, (np + csum) & 0xFFFFu, hv[bt], nlive, best, ov[bt]);
		}
	}
}

// warp per pair, 33 <= P <= 256
constexpr int WARP_FOLD_MAX = 256;
constexpr int WARP_FOLD_WARPS = 8;

__global__ void __launch_bounds__(WARP_FOLD_WARPS * 32) k_fold_long(Params P, const FDesc* __restrict__ fdesc,
		const uint32_t* __restrict__ list, const unsigned int* __restrict__ count_ptr)
{
	__shared__ uint32_t s_hv[WARP_FOLD_WARPS][WARP_FOLD_MAX];
	__shared__ uint32_t s_sup[WARP_FOLD_WARPS][WARP_FOLD_MAX];
	__shared__ uint16_t s_ov[WARP_FOLD_WARPS][WARP_FOLD_MAX];
	__shared__ uint16_t s_jr[WARP_FOLD_WARPS][WARP_FOLD_MAX];
	__shared__ uint16_t s_par[WARP_FOLD_WARPS][WARP_FOLD_MAX];
	const uint32_t FULL = 0xFFFFFFFFu;
	const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	uint32_t* shv = s_hv[w]; uint32_t* ssup = s_sup[w];
	uint16_t* sov = s_ov[w]; uint16_t* sjr = s_jr[w]; uint16_t* spar = s_par[w];
	const uint32_t count = *count_ptr;
	const uint32_t K = P.K, K2 = 2 * P.K;
	const int BIN = (int)P.BIN;
	for (uint32_t it = blockIdx.x * WARP_FOLD_WARPS + w; it < count; it += gridDim.x * WARP_FOLD_WARPS) {
		const uint32_t g = list[it];
		const FDesc f = fdesc[g];
		const uint32_t np = (uint32_t)(f.off_len >> 48);
		uint64_t* rec = P.prod + (f.off_len & 0xFFFFFFFFFFFFull);
		const int lenH = (int)P.read_len[f.row], lenV = (int)P.read_len[f.col];
		const uint32_t R = (np + 31) >> 5;
		__syncwarp();
		// load, rank by jrank (counting), scatter into B-column order
		uint32_t myhv[WARP_FOLD_MAX / 32], myjr[WARP_FOLD_MAX / 32], myov[WARP_FOLD_MAX / 32], rank[WARP_FOLD_MAX / 32];
#pragma unroll
		for (int r = 0; r < WARP_FOLD_MAX / 32; ++r) {
			uint32_t idx = lane + 32 * r;
			rank[r] = 0; myjr[r] = 0xFFFFFFFFu; myhv[r] = 0; myov[r] = 0;
			if (r < (int)R && idx < np) {
				uint64_t x = rec[idx];
				myhv[r] = (uint32_t)x; myjr[r] = (uint32_t)(x >> 32) & 0xFFFFu;
				myov[r] = overlap_estimate(lenH, lenV, myhv[r] & 0xFFFFu, myhv[r] >> 16, (uint32_t)(x >> 48) & 1u, K);
				sjr[idx] = (uint16_t)myjr[r];
			}
		}
		__syncwarp();
		for (uint32_t b = 0; b < np; ++b) {
			uint32_t x = sjr[b];
#pragma unroll
			for (int r = 0; r < WARP_FOLD_MAX / 32; ++r) rank[r] += (x < myjr[r]);
		}
#pragma unroll
		for (int r = 0; r < WARP_FOLD_MAX / 32; ++r)
			if (r < (int)R && lane + 32 * r < np) { shv[rank[r]] = myhv[r]; sov[rank[r]] = (uint16_t)myov[r]; }
		__syncwarp();
		// is the bin forest the chain t -> t+1 ?
		bool lin = true;
		for (uint32_t idx = lane + 1; idx < np; idx += 32) lin &= abs((int)sov[idx] - (int)sov[idx - 1]) < BIN;
		const bool linear = __all_sync(FULL, lin);
		bool fallback = false;
		if (!linear) {
			// phase A: parents of the bin forest; each lane keeps one live bin
			uint32_t lb = NONE16;
			for (uint32_t idx = lane; idx < np; idx += 32) ssup[idx] = 0;
			for (uint32_t t = 0; t < np; ++t) {
				int ot = (int)sov[t];
				if (lb != NONE16 && abs((int)sov[lb] - ot) < BIN) { spar[lb] = (uint16_t)t; lb = NONE16; }
				uint32_t freem = __ballot_sync(FULL, lb == NONE16);
				if (!freem) { fallback = true; break; }
				if (lane == (uint32_t)(__ffs(freem) - 1)) lb = t;
			}
			if (lb != NONE16) spar[lb] = NONE16;
			__syncwarp();
		}
		if (fallback) {     // more than 32 simultaneous bins: sequential in-place fold by one lane
			if (lane == 0) {
				uint32_t cnt, h, v, nb, sup, ov;
				fold_pair(rec, np, lenH, lenV, K, BIN, cnt, h, v, nb, sup, ov);
				store_result(P, g, f.row, cnt, h | (v << 16), nb, sup, ov);
			}
			continue;
		}
		// phase B: every k-mer walks its ancestors; lanes take s from alternating ends for balance
		uint32_t csum = 0, surv = 0, r = 0, s = 0, t = 0, A = 0, B = 0, last = 0;
		bool active = false;
		auto advance = [&]() {
			active = false;
			while (r < R) {
				s = (r & 1) ? 32 * r + 31 - lane : 32 * r + lane;
				++r;
				if (s < np) {
					uint32_t x = shv[s];
					A = K - (x & 0xFFFFu); B = K - (x >> 16);
					t = linear ? s + 1 : spar[s];
					last = s; active = true;
					return;
				}
			}
		};
		advance();
		while (__any_sync(FULL, active)) {
			if (active) {
				if (t >= np) {                       // reached a root alive
					if (linear) ++surv; else atomicAdd(&ssup[last], 1u);
					advance();
				} else if (is_far(shv[t], A, B, K2)) {
					++csum; last = t;
					t = linear ? t + 1 : spar[t];
				} else {
					advance();
				}
			}
		}
		for (int o = 16; o; o >>= 1) { csum += __shfl_xor_sync(FULL, csum, o); surv += __shfl_xor_sync(FULL, surv, o); }
		uint32_t root = np - 1, nb = 1, sup = surv;
		if (!linear) {
			__syncwarp();
			uint32_t best = 0, nroots = 0;
			for (uint32_t idx = lane; idx < np; idx += 32)
				if (spar[idx] == NONE16) { ++nroots; uint32_t c = (ssup[idx] << 16) | idx; best = max(best, c); }
			for (int o = 16; o; o >>= 1) { best = max(best, __shfl_xor_sync(FULL, best, o)); nroots += __shfl_xor_sync(FULL, nroots, o); }
			root = best & 0xFFFFu; sup = best >> 16; nb = nroots;
		}
		if (lane == 0) store_result(P, g, f.row, (np + csum) & 0xFFFFu, shv[root], nb, sup, sov[root]);
	}
}

// P > 256: sequential in-place fold, thread per pair (rare)
__global__ void k_fold_huge(Params P, const FDesc* __restrict__ fdesc, const uint32_t* __restrict__ list,
		const unsigned int* __restrict__ count_ptr)
{
	const uint32_t count = *count_ptr;
	for (uint32_t it = blockIdx.x * blockDim.x + threadIdx.x; it < count; it += gridDim.x * blockDim.x) {
		const uint32_t g = list[it];
		const FDesc f = fdesc[g];
		uint32_t cnt, h, v, nb, sup, ov;
		fold_pair(P.prod + (f.off_len & 0xFFFFFFFFFFFFull), (uint32_t)(f.off_len >> 48), (int)P.read_len[f.row], (int)P.read_len[f.col],
			P.K, (int)P.BIN, cnt, h, v, nb, sup, ov);
		store_result(P, g, f.row, cnt, h | (v << 16), nb, sup, ov);
	}
}

// ---- host side ------------------------------------------------------------------------------

struct DevBuf {
	void* p = nullptr;
	size_t cap = 0;
	int ensure(size_t bytes)
	{
		if (bytes <= cap) return 0;
		if (p) cudaFree(p);
		p = nullptr; cap = 0;
		size_t want = bytes + bytes / 8 + 256;
		if (cudaMalloc(&p, want) != cudaSuccess) { cudaGetLastError(); if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); p = nullptr; return -1; } want = bytes; }
		cap = want;
		return 0;
	}
	void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
	template <class T> T* as() const { return (T*)p; }
};

} // namespace

struct bella_b200_handle {
	int device = 0;
	cudaStream_t stream = nullptr;
	bool own_stream = true;
	std::string err;
	// problem
	uint32_t n = 0, m = 0, lo = 0, hi = 0;
	uint64_t nnzA = 0, nnzB = 0;
	uint16_t K = 17, BIN = 500;
	bool have_inputs = false, have_A = false, layout_done = false, symbolic_done = false, numeric_done = false;
	// input device pointers (borrowed or pointing into the owned buffers below)
	const uint32_t *dA_colptr = nullptr, *dA_rowids = nullptr, *dB_colptr = nullptr, *dB_rowids = nullptr, *d_len = nullptr;
	const uint16_t *dA_values = nullptr, *dB_values = nullptr;
	const uint8_t *dA_strand = nullptr, *dB_strand = nullptr;
	DevBuf oA_colptr, oA_rowids, oA_values, oA_strand, oB_colptr, oB_rowids, oB_values, oB_strand, o_len;
	// layout + work
	DevBuf Aent, Bent, tA_colptr, tcursor, flopC, flop64, flopptr, nnzC, colptrC, lists, meta, errflag, cubtmp, slab;
	DevBuf prod, prow, pdesc, rowsC, countC, posH, posV, aux, fdesc, flists, fmeta;
	Meta hmeta{};
	uint64_t flops = 0, Z = 0;
	cudaEvent_t ev[8]{};
	float t_ms[8]{};
	int launches = 0;
};

namespace {

int fail(bella_b200_handle* h, int code, const char* fmt, ...)
{
	char buf[512];
	va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
	if (h) h->err = buf;
	return code;
}

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cudaGetLastError(); \
	return fail(h, BELLA_B200_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } } while (0)
#define ENSURE(buf, bytes) do { if ((buf).ensure(bytes)) return fail(h, BELLA_B200_ERR_OOM, "device allocation of %zu bytes failed (%s)", (size_t)(bytes), #buf); } while (0)
#define LAUNCHED() do { ++h->launches; CK(cudaGetLastError()); } while (0)

inline int grid_for(uint64_t work, int threads, int cap = 148 * 16)
{
	uint64_t g = (work + threads - 1) / threads;
	if (g < 1) g = 1;
	if (g > (uint64_t)cap) g = cap;
	return (int)g;
}

template <class In, class Out>
int exclusive_scan(bella_b200_handle* h, In in, Out out, uint32_t count)
{
	size_t bytes = 0;
	CK(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, count, h->stream));
	ENSURE(h->cubtmp, bytes);
	CK(cub::DeviceScan::ExclusiveSum(h->cubtmp.p, bytes, in, out, count, h->stream));
	++h->launches;
	return 0;
}

int check_device_error(bella_b200_handle* h)
{
	int e = 0;
	CK(cudaMemcpyAsync(&e, h->errflag.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
	CK(cudaStreamSynchronize(h->stream));
	if (e == BELLA_B200_ERR_RANGE)
		return fail(h, e, "a column of B has more than 65535 nonzeros, a k-mer occurs in more than 32767 reads, or a column needs more than 2^32-1 products");
	if (e) return fail(h, e, "device-side error %d", e);
	return 0;
}

// layout: Aent (sorted columns) and Bent + flopC for the handle's column range
int run_layout(bella_b200_handle* h)
{
	const uint32_t n = h->n, m = h->m;
	const uint64_t nnz = h->nnzB;
	const uint32_t ncols = h->hi - h->lo;
	ENSURE(h->Aent, sizeof(uint64_t) * (nnz + 1));
	ENSURE(h->Bent, sizeof(uint64_t) * (nnz + 1));
	ENSURE(h->flopC, sizeof(uint32_t) * ((size_t)ncols + 1));
	ENSURE(h->flop64, sizeof(uint64_t) * ((size_t)ncols + 1));
	ENSURE(h->errflag, sizeof(int));
	CK(cudaMemsetAsync(h->errflag.p, 0, sizeof(int), h->stream));
	const uint32_t* Acolptr;
	if (h->have_A) {
		Acolptr = h->dA_colptr;
		k_pack_A<<<grid_for(nnz, 256), 256, 0, h->stream>>>(nnz, h->dA_rowids, h->dA_values, h->dA_strand, h->Aent.as<uint64_t>());
		LAUNCHED();
	} else {
		ENSURE(h->tA_colptr, sizeof(uint32_t) * ((size_t)m + 2));
		ENSURE(h->tcursor, sizeof(uint32_t) * ((size_t)m + 2));
		CK(cudaMemsetAsync(h->tcursor.p, 0, sizeof(uint32_t) * ((size_t)m + 2), h->stream));
		k_count_deg<<<grid_for(nnz, 256), 256, 0, h->stream>>>(h->dB_rowids, nnz, h->tcursor.as<uint32_t>());
		LAUNCHED();
		if (int rc = exclusive_scan(h, h->tcursor.as<uint32_t>(), h->tA_colptr.as<uint32_t>(), m + 1)) return rc;
		CK(cudaMemsetAsync(h->tcursor.p, 0, sizeof(uint32_t) * ((size_t)m + 2), h->stream));
		k_transpose_fill<<<grid_for((uint64_t)n * 32, 256), 256, 0, h->stream>>>(n, h->dB_colptr, h->dB_rowids, h->dB_values,
			h->dB_strand, h->tA_colptr.as<uint32_t>(), h->tcursor.as<uint32_t>(), h->Aent.as<uint64_t>());
		LAUNCHED();
		Acolptr = h->tA_colptr.as<uint32_t>();
		h->dA_colptr = Acolptr;
	}
	k_sort_A<<<grid_for(m, 256), 256, 0, h->stream>>>(m, Acolptr, h->Aent.as<uint64_t>());
	LAUNCHED();
	CK(cudaMemsetAsync(h->flopC.p, 0, sizeof(uint32_t) * ((size_t)ncols + 1), h->stream));
	CK(cudaMemsetAsync(h->flop64.p, 0, sizeof(uint64_t) * ((size_t)ncols + 1), h->stream));
	k_pack_B<<<grid_for((uint64_t)ncols * 32, 256), 256, 0, h->stream>>>(h->lo, h->hi, h->dB_colptr, h->dB_rowids, h->dB_values,
		h->dB_strand, Acolptr, h->Aent.as<uint64_t>(), h->Bent.as<uint64_t>(), h->flopC.as<uint32_t>(), h->flop64.as<unsigned long long>(), h->errflag.as<int>());
	LAUNCHED();
	h->layout_done = true;
	return 0;
}

Params make_params(bella_b200_handle* h)
{
	Params P{};
	P.n = h->n; P.m = h->m; P.lo = h->lo; P.hi = h->hi; P.K = h->K; P.BIN = h->BIN;
	P.B_colptr = h->dB_colptr; P.A_colptr = h->dA_colptr; P.read_len = h->d_len;
	P.Aent = h->Aent.as<uint64_t>(); P.Bent = h->Bent.as<uint64_t>();
	P.flopC = h->flopC.as<uint32_t>(); P.flopptr = h->flopptr.as<uint64_t>();
	P.nnzC = h->nnzC.as<uint32_t>(); P.colptrC = h->colptrC.as<uint32_t>();
	P.prod = h->prod.as<uint64_t>(); P.prow = h->prow.as<uint32_t>(); P.pdesc = h->pdesc.as<uint2>();
	P.rowsC = h->rowsC.as<uint32_t>(); P.countC = h->countC.as<uint16_t>(); P.posH = h->posH.as<uint16_t>();
	P.posV = h->posV.as<uint16_t>(); P.aux = h->aux.as<uint16_t>(); P.err = h->errflag.as<int>();
	return P;
}

// symbolic: flop scan, size classes, expand (group products by pair), nnz scan
int run_symbolic(bella_b200_handle* h)
{
	const uint32_t ncols = h->hi - h->lo;
	ENSURE(h->flopptr, sizeof(uint64_t) * ((size_t)ncols + 1));
	ENSURE(h->nnzC, sizeof(uint32_t) * ((size_t)ncols + 1));
	ENSURE(h->colptrC, sizeof(uint32_t) * ((size_t)ncols + 1));
	ENSURE(h->lists, sizeof(uint32_t) * (size_t)N_CLASSES * (ncols + 1));
	ENSURE(h->meta, sizeof(Meta));
	if (int rc = exclusive_scan(h, h->flop64.as<unsigned long long>(), h->flopptr.as<unsigned long long>(), ncols + 1)) return rc;
	CK(cudaMemsetAsync(h->meta.p, 0, sizeof(Meta), h->stream));
	CK(cudaMemsetAsync(h->nnzC.p, 0, sizeof(uint32_t) * ((size_t)ncols + 1), h->stream));
	k_classify<<<grid_for(ncols, 256), 256, 0, h->stream>>>(ncols, h->flopC.as<uint32_t>(), h->lists.as<uint32_t>(), h->meta.as<Meta>());
	LAUNCHED();
	k_set_total<<<1, 1, 0, h->stream>>>(h->meta.as<Meta>(), h->flopptr.as<uint64_t>(), ncols);
	LAUNCHED();
	CK(cudaMemcpyAsync(&h->hmeta, h->meta.p, sizeof(Meta), cudaMemcpyDeviceToHost, h->stream));
	if (int rc = check_device_error(h)) return rc;     // synchronises
	h->flops = h->hmeta.flops;
	const uint64_t F = h->flops;
	ENSURE(h->prod, sizeof(uint64_t) * (F + 1));
	ENSURE(h->prow, sizeof(uint32_t) * (F + 1));
	ENSURE(h->pdesc, sizeof(uint2) * (F + 1));
	Params P = make_params(h);
	CK(cudaEventRecord(h->ev[6], h->stream));
	for (int c = 0; c < 3; ++c) {
		uint32_t cnt = h->hmeta.class_count[c];
		if (!cnt) continue;
		uint32_t htmax = CLASS_LIMIT[c];
		size_t smem = (size_t)5 * htmax * sizeof(uint32_t);
		CK(cudaFuncSetAttribute(k_expand<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(5 * CLASS_LIMIT[2] * sizeof(uint32_t))));
		int threads = c == 0 ? EXPAND_THREADS : 256;
		k_expand<false><<<cnt, threads, smem, h->stream>>>(P, h->lists.as<uint32_t>() + (size_t)c * ncols, cnt, htmax, nullptr);
		LAUNCHED();
	}
	if (uint32_t cnt = h->hmeta.class_count[3]) {
		uint32_t htmax = 32;
		while (htmax < h->hmeta.max_flop) htmax <<= 1;
		uint32_t ctas = cnt < 148 ? cnt : 148;
		while (ctas > 1 && (size_t)ctas * 5 * htmax * sizeof(uint32_t) > ((size_t)8 << 30)) ctas >>= 1;
		ENSURE(h->slab, (size_t)ctas * 5 * htmax * sizeof(uint32_t));
		k_expand<true><<<ctas, 256, 0, h->stream>>>(P, h->lists.as<uint32_t>() + (size_t)3 * ncols, cnt, htmax, h->slab.as<uint32_t>());
		LAUNCHED();
	}
	CK(cudaEventRecord(h->ev[7], h->stream));
	if (int rc = exclusive_scan(h, h->nnzC.as<uint32_t>(), h->colptrC.as<uint32_t>(), ncols + 1)) return rc;
	// total nnz (u32 like the reference's IT; a wrapped total is detected through the 64-bit sum below)
	uint32_t z32 = 0;
	CK(cudaMemcpyAsync(&z32, h->colptrC.as<uint32_t>() + ncols, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
	if (int rc = check_device_error(h)) return rc;
	h->Z = z32;
	h->symbolic_done = true;
	h->numeric_done = false;
	return 0;
}

int run_numeric(bella_b200_handle* h)
{
	const uint32_t ncols = h->hi - h->lo;
	const uint64_t Z = h->Z;
	ENSURE(h->rowsC, sizeof(uint32_t) * (Z + 1));
	ENSURE(h->countC, sizeof(uint16_t) * (Z + 1));
	ENSURE(h->posH, sizeof(uint16_t) * (Z + 1));
	ENSURE(h->posV, sizeof(uint16_t) * (Z + 1));
	ENSURE(h->aux, sizeof(uint16_t) * 3 * (Z + 1));
	ENSURE(h->fdesc, sizeof(FDesc) * (Z + 1));
	ENSURE(h->flists, sizeof(uint32_t) * NBUCKETS * (Z + 1));
	ENSURE(h->fmeta, sizeof(FoldMeta));
	Params P = make_params(h);
	if (!ncols || !Z) { h->numeric_done = true; return 0; }
	CK(cudaMemsetAsync(h->fmeta.p, 0, sizeof(FoldMeta), h->stream));
	FDesc* fd = h->fdesc.as<FDesc>();
	uint32_t* lists = h->flists.as<uint32_t>();
	FoldMeta* fm = h->fmeta.as<FoldMeta>();
	const uint64_t cap = Z + 1;
	k_flatten<<<grid_for((uint64_t)ncols * 32, 256), 256, 0, h->stream>>>(P, fd, lists, cap, fm);
	LAUNCHED();
	const int g = 148 * 8;
	k_fold_short<1><<<g, 128, 0, h->stream>>>(P, fd, lists + 0 * cap, &fm->count[0]); LAUNCHED();
	k_fold_short<4><<<g, 128, 0, h->stream>>>(P, fd, lists + 1 * cap, &fm->count[1]); LAUNCHED();
	k_fold_short<8><<<g, 128, 0, h->stream>>>(P, fd, lists + 2 * cap, &fm->count[2]); LAUNCHED();
	k_fold_short<16><<<g, 128, 0, h->stream>>>(P, fd, lists + 3 * cap, &fm->count[3]); LAUNCHED();
	k_fold_short<32><<<g, 128, 0, h->stream>>>(P, fd, lists + 4 * cap, &fm->count[4]); LAUNCHED();
	k_fold_long<<<148 * 4, WARP_FOLD_WARPS * 32, 0, h->stream>>>(P, fd, lists + 5 * cap, &fm->count[5]); LAUNCHED();
	k_fold_huge<<<148, 128, 0, h->stream>>>(P, fd, lists + 6 * cap, &fm->count[6]); LAUNCHED();
	h->numeric_done = true;
	return 0;
}

int copy_in(bella_b200_handle* h, DevBuf& buf, const void* src, size_t bytes, const void** dst)
{
	ENSURE(buf, bytes + 16);
	CK(cudaMemcpyAsync(buf.p, src, bytes, cudaMemcpyHostToDevice, h->stream));
	*dst = buf.p;
	return 0;
}

int validate_views(bella_b200_handle* h, const bella_csc_view* A, const bella_csc_view* B, const uint32_t* read_len,
		const uint8_t* sA, const uint8_t* sB)
{
	if (!h) return BELLA_B200_ERR_ARG;
	if (!B || !B->colptr || !read_len || !sB) return fail(h, BELLA_B200_ERR_ARG, "B, read_len and strand_B are required");
	if (B->nnz && (!B->rowids || !B->values)) return fail(h, BELLA_B200_ERR_ARG, "B.rowids/B.values missing");
	if (A) {
		if (!A->colptr || !sA || (A->nnz && (!A->rowids || !A->values))) return fail(h, BELLA_B200_ERR_ARG, "A view incomplete (or strand_A missing)");
		if (A->rows != B->cols || A->cols != B->rows) return fail(h, BELLA_B200_ERR_ARG, "A is %ux%u but B is %ux%u", A->rows, A->cols, B->rows, B->cols);
	}
	return 0;
}

void reset_problem(bella_b200_handle* h, const bella_csc_view* A, const bella_csc_view* B, uint16_t K, uint16_t BIN)
{
	h->n = B->cols; h->m = B->rows; h->nnzB = B->nnz; h->nnzA = A ? A->nnz : B->nnz;
	h->lo = 0; h->hi = h->n; h->K = K; h->BIN = BIN;
	h->have_A = A != nullptr;
	h->have_inputs = true; h->layout_done = h->symbolic_done = h->numeric_done = false;
	h->flops = h->Z = 0;
}

} // namespace

extern "C" {

int bella_b200_create(bella_b200_handle** out, int device)
{
	if (!out) return BELLA_B200_ERR_ARG;
	*out = nullptr;
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return BELLA_B200_ERR_CUDA; }
	if (device < 0 || device >= ndev) return BELLA_B200_ERR_ARG;
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return BELLA_B200_ERR_CUDA;
	if (prop.major != 10) return BELLA_B200_ERR_CUDA;             // sm_100a only, no other code path
	if (cudaSetDevice(device) != cudaSuccess) return BELLA_B200_ERR_CUDA;
	bella_b200_handle* h = new bella_b200_handle();
	h->device = device;
	if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return BELLA_B200_ERR_CUDA; }
	for (auto& e : h->ev) cudaEventCreate(&e);
	*out = h;
	return BELLA_B200_OK;
}

int bella_b200_destroy(bella_b200_handle* h)
{
	if (!h) return BELLA_B200_ERR_ARG;
	cudaSetDevice(h->device);
	cudaStreamSynchronize(h->stream);
	DevBuf* bufs[] = {&h->oA_colptr, &h->oA_rowids, &h->oA_values, &h->oA_strand, &h->oB_colptr, &h->oB_rowids, &h->oB_values,
		&h->oB_strand, &h->o_len, &h->Aent, &h->Bent, &h->tA_colptr, &h->tcursor, &h->flopC, &h->flop64, &h->flopptr, &h->nnzC, &h->colptrC,
		&h->lists, &h->meta, &h->errflag, &h->cubtmp, &h->slab, &h->prod, &h->prow, &h->pdesc, &h->rowsC, &h->countC, &h->posH,
		&h->posV, &h->aux, &h->fdesc, &h->flists, &h->fmeta};
	for (DevBuf* b : bufs) b->release();
	for (auto& e : h->ev) if (e) cudaEventDestroy(e);
	if (h->own_stream) cudaStreamDestroy(h->stream);
	delete h;
	return BELLA_B200_OK;
}

const char* bella_b200_last_error(const bella_b200_handle* h) { return h ? h->err.c_str() : "null handle"; }

int bella_b200_set_inputs(bella_b200_handle* h, const bella_csc_view* A, const bella_csc_view* B,
		const uint32_t* read_len, const uint8_t* strand_A, const uint8_t* strand_B, uint16_t kmer_size, uint16_t bin_size)
{
	if (int rc = validate_views(h, A, B, read_len, strand_A, strand_B)) return rc;
	CK(cudaSetDevice(h->device));
	reset_problem(h, A, B, kmer_size, bin_size);
	CK(cudaEventRecord(h->ev[0], h->stream));
	const void* p;
	if (int rc = copy_in(h, h->oB_colptr, B->colptr, sizeof(uint32_t) * ((size_t)B->cols + 1), &p)) return rc; h->dB_colptr = (const uint32_t*)p;
	if (int rc = copy_in(h, h->oB_rowids, B->rowids, sizeof(uint32_t) * (size_t)B->nnz, &p)) return rc; h->dB_rowids = (const uint32_t*)p;
	if (int rc = copy_in(h, h->oB_values, B->values, sizeof(uint16_t) * (size_t)B->nnz, &p)) return rc; h->dB_values = (const uint16_t*)p;
	if (int rc = copy_in(h, h->oB_strand, strand_B, ((size_t)B->nnz + 7) / 8, &p)) return rc; h->dB_strand = (const uint8_t*)p;
	if (int rc = copy_in(h, h->o_len, read_len, sizeof(uint32_t) * (size_t)B->cols, &p)) return rc; h->d_len = (const uint32_t*)p;
	if (A) {
		if (int rc = copy_in(h, h->oA_colptr, A->colptr, sizeof(uint32_t) * ((size_t)A->cols + 1), &p)) return rc; h->dA_colptr = (const uint32_t*)p;
		if (int rc = copy_in(h, h->oA_rowids, A->rowids, sizeof(uint32_t) * (size_t)A->nnz, &p)) return rc; h->dA_rowids = (const uint32_t*)p;
		if (int rc = copy_in(h, h->oA_values, A->values, sizeof(uint16_t) * (size_t)A->nnz, &p)) return rc; h->dA_values = (const uint16_t*)p;
		if (int rc = copy_in(h, h->oA_strand, strand_A, ((size_t)A->nnz + 7) / 8, &p)) return rc; h->dA_strand = (const uint8_t*)p;
	}
	CK(cudaEventRecord(h->ev[1], h->stream));
	CK(cudaStreamSynchronize(h->stream));
	CK(cudaEventElapsedTime(&h->t_ms[3], h->ev[0], h->ev[1]));
	return BELLA_B200_OK;
}

int bella_b200_set_inputs_device(bella_b200_handle* h, const bella_csc_view* A, const bella_csc_view* B,
		const uint32_t* read_len, const uint8_t* strand_A, const uint8_t* strand_B, uint16_t kmer_size, uint16_t bin_size)
{
	if (int rc = validate_views(h, A, B, read_len, strand_A, strand_B)) return rc;
	CK(cudaSetDevice(h->device));
	reset_problem(h, A, B, kmer_size, bin_size);
	h->dB_colptr = B->colptr; h->dB_rowids = B->rowids; h->dB_values = B->values; h->dB_strand = strand_B; h->d_len = read_len;
	if (A) { h->dA_colptr = A->colptr; h->dA_rowids = A->rowids; h->dA_values = A->values; h->dA_strand = strand_A; }
	h->t_ms[3] = 0;
	return BELLA_B200_OK;
}

int bella_b200_set_column_range(bella_b200_handle* h, uint32_t col_lo, uint32_t col_hi)
{
	if (!h || !h->have_inputs) return fail(h, BELLA_B200_ERR_ARG, "set_inputs first");
	if (col_lo > col_hi || col_hi > h->n) return fail(h, BELLA_B200_ERR_ARG, "column range [%u,%u) outside [0,%u)", col_lo, col_hi, h->n);
	h->lo = col_lo; h->hi = col_hi;
	h->layout_done = h->symbolic_done = h->numeric_done = false;
	return BELLA_B200_OK;
}

static int do_symbolic(bella_b200_handle* h)
{
	CK(cudaSetDevice(h->device));
	h->launches = 0;
	CK(cudaEventRecord(h->ev[0], h->stream));
	if (int rc = run_layout(h)) return rc;
	CK(cudaEventRecord(h->ev[1], h->stream));
	if (int rc = run_symbolic(h)) return rc;
	CK(cudaEventRecord(h->ev[2], h->stream));
	CK(cudaStreamSynchronize(h->stream));
	CK(cudaEventElapsedTime(&h->t_ms[0], h->ev[0], h->ev[1]));
	CK(cudaEventElapsedTime(&h->t_ms[1], h->ev[1], h->ev[2]));
	CK(cudaEventElapsedTime(&h->t_ms[6], h->ev[6], h->ev[7]));
	return 0;
}

int bella_b200_symbolic(bella_b200_handle* h, uint64_t* flops, uint32_t* flopC, uint32_t* colptrC)
{
	if (!h || !h->have_inputs) return fail(h, BELLA_B200_ERR_ARG, "set_inputs first");
	if (int rc = do_symbolic(h)) return rc;
	const uint32_t ncols = h->hi - h->lo;
	CK(cudaEventRecord(h->ev[3], h->stream));
	if (flopC && ncols) CK(cudaMemcpyAsync(flopC, h->flopC.p, sizeof(uint32_t) * ncols, cudaMemcpyDeviceToHost, h->stream));
	if (colptrC) CK(cudaMemcpyAsync(colptrC, h->colptrC.p, sizeof(uint32_t) * ((size_t)ncols + 1), cudaMemcpyDeviceToHost, h->stream));
	CK(cudaEventRecord(h->ev[4], h->stream));
	CK(cudaStreamSynchronize(h->stream));
	CK(cudaEventElapsedTime(&h->t_ms[4], h->ev[3], h->ev[4]));
	if (flops) *flops = h->flops;
	return BELLA_B200_OK;
}

int bella_b200_numeric_device(bella_b200_handle* h)
{
	if (!h || !h->symbolic_done) return fail(h, BELLA_B200_ERR_ARG, "bella_b200_symbolic first");
	if (h->numeric_done) return BELLA_B200_OK;
	CK(cudaSetDevice(h->device));
	CK(cudaEventRecord(h->ev[2], h->stream));
	if (int rc = run_numeric(h)) return rc;
	CK(cudaEventRecord(h->ev[3], h->stream));
	if (int rc = check_device_error(h)) return rc;
	CK(cudaEventElapsedTime(&h->t_ms[2], h->ev[2], h->ev[3]));
	return BELLA_B200_OK;
}

static int copy_range(bella_b200_handle* h, uint32_t c0, uint32_t c1, uint64_t* off, uint64_t* cnt)
{
	if (c0 < h->lo || c1 > h->hi || c0 > c1) return fail(h, BELLA_B200_ERR_ARG, "columns [%u,%u) outside the handle's range [%u,%u)", c0, c1, h->lo, h->hi);
	uint32_t ends[2];
	CK(cudaMemcpyAsync(&ends[0], h->colptrC.as<uint32_t>() + (c0 - h->lo), sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
	CK(cudaMemcpyAsync(&ends[1], h->colptrC.as<uint32_t>() + (c1 - h->lo), sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
	CK(cudaStreamSynchronize(h->stream));
	*off = ends[0]; *cnt = ends[1] - ends[0];
	return 0;
}

int bella_b200_numeric(bella_b200_handle* h, uint32_t col_begin, uint32_t col_end,
		uint32_t* rowidsC, uint16_t* count, uint16_t* posH, uint16_t* posV)
{
	if (int rc = bella_b200_numeric_device(h)) return rc;
	uint64_t off, cnt;
	if (int rc = copy_range(h, col_begin, col_end, &off, &cnt)) return rc;
	CK(cudaEventRecord(h->ev[4], h->stream));
	if (cnt) {
		if (rowidsC) CK(cudaMemcpyAsync(rowidsC, h->rowsC.as<uint32_t>() + off, sizeof(uint32_t) * cnt, cudaMemcpyDeviceToHost, h->stream));
		if (count) CK(cudaMemcpyAsync(count, h->countC.as<uint16_t>() + off, sizeof(uint16_t) * cnt, cudaMemcpyDeviceToHost, h->stream));
		if (posH) CK(cudaMemcpyAsync(posH, h->posH.as<uint16_t>() + off, sizeof(uint16_t) * cnt, cudaMemcpyDeviceToHost, h->stream));
		if (posV) CK(cudaMemcpyAsync(posV, h->posV.as<uint16_t>() + off, sizeof(uint16_t) * cnt, cudaMemcpyDeviceToHost, h->stream));
	}
	CK(cudaEventRecord(h->ev[5], h->stream));
	CK(cudaStreamSynchronize(h->stream));
	float t; CK(cudaEventElapsedTime(&t, h->ev[4], h->ev[5]));
	h->t_ms[4] += t;
	return BELLA_B200_OK;
}

int bella_b200_numeric_aux(bella_b200_handle* h, uint32_t col_begin, uint32_t col_end,
		uint16_t* nbins, uint16_t* support, uint16_t* overlap)
{
	if (int rc = bella_b200_numeric_device(h)) return rc;
	uint64_t off, cnt;
	if (int rc = copy_range(h, col_begin, col_end, &off, &cnt)) return rc;
	if (!cnt) return BELLA_B200_OK;
	std::vector<uint16_t> tmp(3 * cnt);
	CK(cudaMemcpyAsync(tmp.data(), h->aux.as<uint16_t>() + 3 * off, sizeof(uint16_t) * 3 * cnt, cudaMemcpyDeviceToHost, h->stream));
	CK(cudaStreamSynchronize(h->stream));
	for (uint64_t i = 0; i < cnt; ++i) {
		if (nbins) nbins[i] = tmp[3 * i];
		if (support) support[i] = tmp[3 * i + 1];
		if (overlap) overlap[i] = tmp[3 * i + 2];
	}
	return BELLA_B200_OK;
}

int bella_b200_result_device(bella_b200_handle* h, const uint32_t** colptrC, const uint32_t** rowidsC,
		const uint16_t** count, const uint16_t** posH, const uint16_t** posV, uint64_t* nnzC)
{
	if (!h || !h->numeric_done) return fail(h, BELLA_B200_ERR_ARG, "numeric phase has not run");
	if (colptrC) *colptrC = h->colptrC.as<uint32_t>();
	if (rowidsC) *rowidsC = h->rowsC.as<uint32_t>();
	if (count) *count = h->countC.as<uint16_t>();
	if (posH) *posH = h->posH.as<uint16_t>();
	if (posV) *posV = h->posV.as<uint16_t>();
	if (nnzC) *nnzC = h->Z;
	return BELLA_B200_OK;
}

int bella_b200_run_resident(bella_b200_handle* h, uint64_t* nnzC_out, uint64_t* flops_out)
{
	if (!h || !h->have_inputs) return fail(h, BELLA_B200_ERR_ARG, "set_inputs first");
	h->layout_done = h->symbolic_done = h->numeric_done = false;
	if (int rc = do_symbolic(h)) return rc;
	int launches = h->launches;
	if (int rc = bella_b200_numeric_device(h)) return rc;
	(void)launches;
	if (nnzC_out) *nnzC_out = h->Z;
	if (flops_out) *flops_out = h->flops;
	return BELLA_B200_OK;
}

int bella_b200_get_timings(bella_b200_handle* h, float* ms8)
{
	if (!h || !ms8) return BELLA_B200_ERR_ARG;
	memcpy(ms8, h->t_ms, sizeof(float) * 8);
	ms8[5] = (float)h->launches;
	return BELLA_B200_OK;
}

void* bella_b200_stream(bella_b200_handle* h) { return h ? (void*)h->stream : nullptr; }

int bella_b200_set_stream(bella_b200_handle* h, void* stream)
{
	if (!h) return BELLA_B200_ERR_ARG;
	cudaSetDevice(h->device);
	cudaStreamSynchronize(h->stream);
	if (h->own_stream) cudaStreamDestroy(h->stream);
	h->stream = (cudaStream_t)stream;
	h->own_stream = false;
	return BELLA_B200_OK;
}

} // extern "C"
