// kernels.cuh -- device code of the B200 overlap SpGEMM (included once, by bella_b200.cu).
//
// Pipeline (DESIGN.md has the traffic model of every stage):
//   layout   k_build_A      A's columns (k-mers) as packed entries sorted by read id, plus the per
//                           output-column product count == estimateFLOP (overlap.hpp:157-202)
//            k_count_deg / k_transpose_fill   A = B^T on the device when the caller passes only B
//   scatter  k_scatter      outer-product expansion on the A side: k-mer column (r0<r1<..) emits the
//                           kept products (col r_a, row r_b), a<b, into column r_a's private region.
//                           Streams A once; the write frontier (one cursor per output column) lives
//                           in L2, so no random DRAM gathers.
//   group    k_group        one CTA per output column: shared-memory hash of the column's k-mers
//                           (k-mer id -> position in B's column = fold order) and of the row ids
//                           (== estimateNNZ_Hash, overlap.hpp:205-276); groups the products by pair
//                           in B-column order (== LocalSpGEMM's visiting order, overlap.hpp:306-341)
//            k_expand_gather  fallback for columns that do not fit shared memory (gather formulation)
//   fold     k_flatten, k_fold_short<>, k_fold_long, k_fold_huge   the semiring (chain.hpp:74-150)
//                           + choose() (common.h:162-170), bucketed by products per pair
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bk {

constexpr uint32_t EMPTY = 0xFFFFFFFFu;
constexpr uint32_t NONE16 = 0xFFFFu;
constexpr int N_CLASSES = 4;          // group classes: S, M (shared memory), gather-smem, gather-global
constexpr int NBUCKETS = 7;           // fold buckets: P==1 | 2..4 | 5..8 | 9..16 | 17..32 | 33..256 | >256
constexpr uint32_t GATHER_SMEM_LIMIT = 8192;

// Packed formats (64-bit unless noted)
//   Aent  : row(32) | pos(16)<<32 | strand(1)<<48            A's columns, rows ascending
//   Bent  : aoff(32) | pos(16)<<32 | cnt(15)<<48 | strand<<63 only for gather-fallback columns
//   raw   : uint4 per product    { row | oriented<<31,  h | v<<16,  k-mer id, 0 }  (one 16-byte store)
//   prod  : h(16) | v(16)<<16 | overlap(16)<<32 [| fold state label(16)<<48]   grouped by pair, in fold order

struct Params {
	uint32_t n, m, lo, hi, K, BIN;
	const uint32_t* B_colptr;
	const uint32_t* B_rowids;
	const uint32_t* A_colptr;
	const uint32_t* read_len;
	const uint64_t* Aent;
	const uint64_t* Bent;
	unsigned long long* flop64;   // [ncols+1] products per column
	uint64_t* flopptr;            // [ncols+1] exclusive scan
	uint32_t* cursor;             // [ncols]   scatter cursors
	uint4* raw;                   // [F]
	uint32_t* nnzC;               // [ncols+1]
	uint32_t* colptrC;            // [ncols+1]
	uint32_t* bcount;             // [NBUCKETS*ncols+1] pairs per fold bucket per column, then scanned in place -> boffs
	uint64_t* prod;               // [2F]  (second half: ordered output of the gather fallback)
	uint64_t prod_half;           // F
	uint32_t* prow;               // [F]   pair row id at flopptr[col]+p
	uint2* pdesc;                 // [F]   {absolute start in prod (low 32 bits of offset from region base), length}
	uint32_t* rowsC;
	uint16_t* countC;
	uint16_t* posH;
	uint16_t* posV;
	uint16_t* aux;
	int* err;
};

struct Meta {
	unsigned long long flops;
	unsigned int class_count[N_CLASSES];
	unsigned int max_flop;
	unsigned int pad;
};

__device__ __forceinline__ void set_err(int* err, int code) { atomicCAS(err, 0, code); }
__device__ __forceinline__ uint32_t getbit(const uint8_t* __restrict__ bits, uint64_t i) { return (bits[i >> 3] >> (i & 7)) & 1u; }

// ================================ layout ====================================================

__global__ void k_count_deg(const uint32_t* __restrict__ Brow, uint64_t nnz, uint32_t* __restrict__ deg)
{
	for (uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j < nnz; j += (uint64_t)gridDim.x * blockDim.x)
		atomicAdd(&deg[Brow[j]], 1u);
}

// A = B^T: one warp per column (read) of B scatters its nonzeros into A's columns (unsorted).
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

// Thread per k-mer column: pack (FROM_ENT = false: from the caller's CSC arrays) or re-read
// (FROM_ENT = true: after k_transpose_fill) the column, sort it by read id, write it back, and add
// each entry's kept-product count (the entries after it) to its read's column counter.
template <bool FROM_ENT>
__global__ void __launch_bounds__(256) k_build_A(uint32_t m, uint32_t lo, uint32_t hi, const uint32_t* __restrict__ Acolptr,
		const uint32_t* __restrict__ Arow, const uint16_t* __restrict__ Aval, const uint8_t* __restrict__ Astrand,
		uint64_t* __restrict__ Aent, unsigned long long* __restrict__ flop64, int* err)
{
	constexpr int LOCAL = 16;
	for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < m; c += gridDim.x * blockDim.x) {
		const uint32_t s = Acolptr[c], e = Acolptr[c + 1], d = e - s;
		if (d == 0) continue;
		if (d > 32768u) { set_err(err, -4); continue; }
		if (d <= LOCAL) {
			uint64_t ent[LOCAL];
#pragma unroll
			for (int q = 0; q < LOCAL; ++q) {
				if (q < (int)d) {
					uint64_t x;
					if (FROM_ENT) x = Aent[s + q];
					else x = (uint64_t)Arow[s + q] | ((uint64_t)Aval[s + q] << 32) | ((uint64_t)getbit(Astrand, (uint64_t)s + q) << 48);
					// insertion into the sorted prefix (static indices keep ent[] in registers)
					ent[q] = x;
#pragma unroll
					for (int b = q; b > 0; --b) {
						if ((uint32_t)ent[b - 1] > (uint32_t)ent[b]) { uint64_t t = ent[b - 1]; ent[b - 1] = ent[b]; ent[b] = t; }
					}
				}
			}
#pragma unroll
			for (int q = 0; q < LOCAL; ++q) {
				if (q < (int)d) {
					Aent[s + q] = ent[q];
					uint32_t r = (uint32_t)ent[q];
					if (q + 1 < (int)d && r >= lo && r < hi) atomicAdd(&flop64[r - lo], (unsigned long long)(d - 1 - q));
				}
			}
		} else {
			if (!FROM_ENT)
				for (uint32_t q = s; q < e; ++q)
					Aent[q] = (uint64_t)Arow[q] | ((uint64_t)Aval[q] << 32) | ((uint64_t)getbit(Astrand, q) << 48);
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
			for (uint32_t q = s; q + 1 < e; ++q) {
				uint32_t r = (uint32_t)Aent[q];
				if (r >= lo && r < hi) atomicAdd(&flop64[r - lo], (unsigned long long)(e - 1 - q));
			}
		}
	}
}

// Outer-product expansion.  Thread per k-mer column (r_0 < r_1 < ...): entry a owns the run of
// products (col r_a, row r_b), b > a; the run is written contiguously into column r_a's region.
__global__ void __launch_bounds__(256) k_scatter(uint32_t m, uint32_t lo, uint32_t hi, const uint32_t* __restrict__ Acolptr,
		const uint64_t* __restrict__ Aent, const uint64_t* __restrict__ flopptr, uint32_t* __restrict__ cursor,
		uint4* __restrict__ raw)
{
	for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < m; c += gridDim.x * blockDim.x) {
		const uint32_t s = Acolptr[c], e = Acolptr[c + 1];
		for (uint32_t a = s; a + 1 < e; ++a) {
			const uint64_t ea = Aent[a];
			const uint32_t ra = (uint32_t)ea;
			if (ra < lo || ra >= hi) continue;
			const uint32_t run = e - 1 - a;
			const uint32_t v = (uint32_t)(ea >> 32) & 0xFFFFu, sa = (uint32_t)(ea >> 48) & 1u;
			uint64_t q = flopptr[ra - lo] + atomicAdd(&cursor[ra - lo], run);
			for (uint32_t b = a + 1; b < e; ++b, ++q) {
				const uint64_t eb = Aent[b];
				const uint32_t h = (uint32_t)(eb >> 32) & 0xFFFFu, sb = (uint32_t)(eb >> 48) & 1u;
				raw[q] = make_uint4((uint32_t)eb | ((sa == sb) << 31), h | (v << 16), c, 0u);
			}
		}
	}
}

// Bent for the gather-fallback columns only (list of local column ids)
__global__ void k_pack_B_list(uint32_t lo, const uint32_t* __restrict__ list, uint32_t count, const uint32_t* __restrict__ Bcolptr,
		const uint32_t* __restrict__ Brow, const uint16_t* __restrict__ Bval, const uint8_t* __restrict__ Bstrand,
		const uint32_t* __restrict__ Acolptr, const uint64_t* __restrict__ Aent, uint64_t* __restrict__ Bent, int* err)
{
	uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
	for (uint32_t it = warp; it < count; it += nwarps) {
		const uint32_t i = lo + list[it];
		uint32_t j0 = Bcolptr[i], j1 = Bcolptr[i + 1];
		for (uint32_t j = j0 + lane; j < j1; j += 32) {
			uint32_t c = Brow[j];
			uint32_t s = Acolptr[c], e = Acolptr[c + 1];
			uint32_t a = s, b = e;      // upper_bound(row <= i) in the sorted column
			while (a < b) { uint32_t mid = (a + b) >> 1; if ((uint32_t)Aent[mid] <= i) a = mid + 1; else b = mid; }
			uint32_t cnt = e - a;
			if (cnt > 32767u) { set_err(err, -4); cnt = 32767u; }
			Bent[j] = (uint64_t)a | ((uint64_t)Bval[j] << 32) | ((uint64_t)cnt << 48) | ((uint64_t)getbit(Bstrand, j) << 63);
		}
	}
}

constexpr uint32_t CLASS_F[2] = {2048, 4096};       // product capacity of the shared-memory group classes
constexpr uint32_t CLASS_L[2] = {3072, 6144};       // B-column length capacity (k-mer hash = 4/3 of it, pow2)

__global__ void k_classify(uint32_t lo, uint32_t ncols, const unsigned long long* __restrict__ flop64,
		const uint32_t* __restrict__ Bcolptr, uint32_t* __restrict__ lists, Meta* meta, int* err)
{
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < ncols; i += gridDim.x * blockDim.x) {
		unsigned long long f = flop64[i];
		if (f == 0) continue;
		uint32_t L = Bcolptr[lo + i + 1] - Bcolptr[lo + i];
		if (f > 0xFFFFFFFFull || L > 65535u) { set_err(err, -4); continue; }
		int c = (f <= CLASS_F[0] && L <= CLASS_L[0]) ? 0 : (f <= CLASS_F[1] && L <= CLASS_L[1]) ? 1 : f <= GATHER_SMEM_LIMIT ? 2 : 3;
		uint32_t idx = atomicAdd(&meta->class_count[c], 1u);
		lists[(size_t)c * ncols + idx] = i;
		if (c == 3) atomicMax(&meta->max_flop, (unsigned int)f);
	}
}

__global__ void k_set_total(Meta* meta, const uint64_t* flopptr, uint32_t ncols) { meta->flops = flopptr[ncols]; }

// ================================ hashing helpers ===========================================

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

// bounded probe: returns EMPTY when the key is absent (internal consistency check)
__device__ __forceinline__ uint32_t ht_find_checked(const uint32_t* keys, uint32_t mask, int shift, uint32_t key)
{
	uint32_t h = (key * 0x9E3779B1u) >> shift;
	for (uint32_t probes = 0; probes <= mask; ++probes) {
		uint32_t k = keys[h];
		if (k == key) return h;
		if (k == EMPTY) return EMPTY;
		h = (h + 1) & mask;
	}
	return EMPTY;
}

// exclusive scan of a[0..n) in place, block-wide; s_tmp needs 34 words; also leaves the total in a[n]
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
			s_tmp[lane] = ws - w;
			if (lane == 31) s_tmp[33] = ws;
		}
		__syncthreads();
		uint32_t carry = s_tmp[32];
		if (idx < n) a[idx] = v - x + s_tmp[wid] + carry;
		__syncthreads();
		if (tid == 0) s_tmp[32] = carry + s_tmp[33];
		__syncthreads();
	}
	if (tid == 0) a[n] = s_tmp[32];
	__syncthreads();
}

__device__ __forceinline__ void block_bitonic_sort(uint32_t* a, uint32_t np2)
{
	const uint32_t tid = threadIdx.x, nt = blockDim.x;
	for (uint32_t k = 2; k <= np2; k <<= 1)
		for (uint32_t j = k >> 1; j > 0; j >>= 1) {
			for (uint32_t x = tid; x < np2; x += nt) {
				uint32_t y = x ^ j;
				if (y > x) {
					uint32_t u = a[x], w = a[y];
					bool up = (x & k) == 0;
					if ((u > w) == up) { a[x] = w; a[y] = u; }
				}
			}
			__syncthreads();
		}
}

__device__ __forceinline__ int bucket_of(uint32_t len)
{
	return len == 1 ? 0 : len <= 4 ? 1 : len <= 8 ? 2 : len <= 16 ? 3 : len <= 32 ? 4 : len <= 256 ? 5 : 6;
}

// multiop -> overlapop (chain.hpp:47-71), checkstrand replaced by the strand-bit comparison.
__device__ __forceinline__ uint32_t overlap_estimate(int lenH, int lenV, uint32_t h, uint32_t v, uint32_t oriented, uint32_t K)
{
	uint32_t hh = oriented ? h : ((uint32_t)lenH - h - K) & 0xFFFFu;   // unsigned short begpH, wraps
	uint32_t endH = (hh + K) & 0xFFFFu, endV = (v + K) & 0xFFFFu;
	int m1 = (int)min(hh, v);
	int m2 = min(lenH - (int)endH, lenV - (int)endV);
	return (uint32_t)(m1 + m2 + (int)K) & 0xFFFFu;                      // stored into vector<unsigned short>
}

// ================================ group =====================================================
// One CTA per output column i.  Shared memory (FCAP products, KHT k-mer slots):
//   prodS  u64[FCAP]     h | v<<16 | jrank<<32 | oriented<<47 | slot-then-pair<<48
//   pkeys  u32[FCAP+2]   row-id hash keys; after the pairs are numbered the region holds poff[] (u32)
//   pcnt   u16[FCAP]     per slot: product count, then pair index (16-bit halves, packed atomics)
//   X      phase 1: kkeys u32[KHT] + kjr u16[KHT]    k-mer id -> position in B's column
//          phase 2: skeys u32[FCAP] + cursor u16[FCAP] + grp u16[FCAP] + jrs u16[FCAP]
template <int FCAP, int KHT>
struct GroupSmem {
	static constexpr size_t X1 = (size_t)KHT * 6;
	static constexpr size_t X2 = (size_t)FCAP * 10;
	static constexpr size_t X = X1 > X2 ? X1 : X2;
	static constexpr size_t BYTES = (size_t)FCAP * 8 + (size_t)(FCAP + 2) * 4 + (size_t)FCAP * 2 + X;
};

// atomicAdd on a 16-bit counter packed two per 32-bit word (no carry: counts stay < 65536)
__device__ __forceinline__ uint32_t atomic_add16(uint16_t* base, uint32_t idx, uint32_t v)
{
	uint32_t sh = (idx & 1u) * 16u;
	uint32_t old = atomicAdd((uint32_t*)base + (idx >> 1), v << sh);
	return (old >> sh) & 0xFFFFu;
}

template <int FCAP, int KHT>
__global__ void __launch_bounds__(256) k_group(Params P, const uint32_t* __restrict__ list, uint32_t count)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	__shared__ uint32_t s_z;
	__shared__ uint32_t s_tmp[34];
	__shared__ uint32_t s_bc[NBUCKETS];
	uint64_t* prodS = (uint64_t*)smem_raw;
	uint32_t* pkeys = (uint32_t*)(prodS + FCAP);
	uint32_t* poff = pkeys;                                  // alias, valid after the pairs are numbered
	uint16_t* pcnt = (uint16_t*)(pkeys + FCAP + 2);
	unsigned char* X = (unsigned char*)(pcnt + FCAP);
	uint32_t* kkeys = (uint32_t*)X;
	uint16_t* kjr = (uint16_t*)(kkeys + KHT);
	uint32_t* skeys = (uint32_t*)X;
	uint16_t* cursor = (uint16_t*)(skeys + FCAP);
	uint16_t* grp = cursor + FCAP;
	uint16_t* jrs = grp + FCAP;
	const uint32_t tid = threadIdx.x, nt = blockDim.x;
	const uint32_t ncols = P.hi - P.lo;

	for (uint32_t it = blockIdx.x; it < count; it += gridDim.x) {
		const uint32_t li = list[it];
		const uint32_t i = P.lo + li;
		const uint32_t j0 = P.B_colptr[i], L = P.B_colptr[i + 1] - j0;
		const uint64_t base = P.flopptr[li];
		const uint32_t Fi = (uint32_t)(P.flopptr[li + 1] - base);
		uint32_t ht = 32; int shift = 27;
		while (ht < Fi) { ht <<= 1; --shift; }
		uint32_t kht = 32; int kshift = 27;
		while (kht * 3 < L * 4) { kht <<= 1; --kshift; }
		const uint32_t mask = ht - 1, kmask = kht - 1;
		for (uint32_t s = tid; s < ht; s += nt) pkeys[s] = EMPTY;
		for (uint32_t s = tid; s < (ht >> 1); s += nt) ((uint32_t*)pcnt)[s] = 0;
		for (uint32_t s = tid; s < kht; s += nt) kkeys[s] = EMPTY;
		if (tid < NBUCKETS) s_bc[tid] = 0;
		if (tid == 0) s_z = 0;
		__syncthreads();
		// the column's k-mers: id -> position in B's column (the fold order)
		for (uint32_t j = tid; j < L; j += nt) {
			uint32_t slot = ht_insert(kkeys, kmask, kshift, P.B_rowids[j0 + j]);
			kjr[slot] = (uint16_t)j;
		}
		__syncthreads();
		// products: k-mer -> jrank, row -> pair slot, count
		for (uint32_t x = tid; x < Fi; x += nt) {
			const uint4 r = P.raw[base + x];
			uint32_t ks = ht_find_checked(kkeys, kmask, kshift, r.z);
			uint32_t jr = 0;
			if (ks == EMPTY) set_err(P.err, -5); else jr = kjr[ks];
			uint32_t slot = ht_insert(pkeys, mask, shift, r.x & 0x7FFFFFFFu);
			atomic_add16(pcnt, slot, 1u);
			prodS[x] = (uint64_t)r.y | ((uint64_t)jr << 32) | ((uint64_t)(r.x >> 31) << 47) | ((uint64_t)slot << 48);
		}
		__syncthreads();
		// --- phase 2: the k-mer hash is dead, X is reused ---
		for (uint32_t s = tid; s < ht; s += nt) {
			uint32_t k = pkeys[s];
			if (k != EMPTY) skeys[atomicAdd(&s_z, 1u)] = k;
		}
		__syncthreads();
		const uint32_t Z = s_z;
		uint32_t Zp = 1;
		while (Zp < Z) Zp <<= 1;
		for (uint32_t s = Z + tid; s < Zp; s += nt) skeys[s] = EMPTY;
		__syncthreads();
		block_bitonic_sort(skeys, Zp);
		// number the pairs by ascending row; per-pair product counts (kept in registers across the
		// barrier, because poff[] overwrites the hash keys they were found through)
		uint32_t mylen[FCAP / 256];
#pragma unroll
		for (int q = 0; q < FCAP / 256; ++q) {
			uint32_t p = tid + q * 256;
			mylen[q] = 0;
			if (p < Z) {
				uint32_t row = skeys[p];
				uint32_t slot = ht_find(pkeys, mask, shift, row);
				mylen[q] = pcnt[slot];
				pcnt[slot] = (uint16_t)p;
				P.prow[base + p] = row;
				atomicAdd(&s_bc[bucket_of(mylen[q])], 1u);
			}
		}
		__syncthreads();
#pragma unroll
		for (int q = 0; q < FCAP / 256; ++q) {
			uint32_t p = tid + q * 256;
			if (p < Z) poff[p] = mylen[q];
		}
		for (uint32_t s = tid; s < ((Z + 1) >> 1); s += nt) ((uint32_t*)cursor)[s] = 0;
		__syncthreads();
		block_excl_scan(poff, Z, s_tmp);            // poff[Z] = Fi
#pragma unroll
		for (int q = 0; q < FCAP / 256; ++q) {
			uint32_t p = tid + q * 256;
			if (p < Z) P.pdesc[base + p] = make_uint2(poff[p], mylen[q]);
		}
		if (tid == 0) P.nnzC[li] = Z;
		if (tid < NBUCKETS) P.bcount[(size_t)tid * ncols + li] = s_bc[tid];
		// unordered membership lists + compact jrank array in pair order
		for (uint32_t x = tid; x < Fi; x += nt) {
			uint64_t r = prodS[x];
			uint32_t p = pcnt[(uint32_t)(r >> 48)];
			uint32_t y = poff[p] + atomic_add16(cursor, p, 1u);
			grp[y] = (uint16_t)x;
			jrs[y] = (uint16_t)((uint32_t)(r >> 32) & 0x7FFFu);
			prodS[x] = (r & 0x0000FFFFFFFFFFFFull) | ((uint64_t)p << 48);
		}
		__syncthreads();
		// rank inside the pair by position in B's column, overlap estimate, ordered write
		const int lenV = (int)P.read_len[i];
		for (uint32_t y = tid; y < Fi; y += nt) {
			uint64_t r = prodS[grp[y]];
			uint32_t p = (uint32_t)(r >> 48), jr = jrs[y];
			uint32_t s0 = poff[p], s1 = poff[p + 1], rank = 0;
			if (s1 - s0 > 1) {
#pragma unroll 4
				for (uint32_t z = s0; z < s1; ++z) rank += (jrs[z] < jr);
			}
			uint32_t hv = (uint32_t)r;
			uint32_t ov = overlap_estimate((int)P.read_len[skeys[p]], lenV, hv & 0xFFFFu, hv >> 16, (uint32_t)(r >> 47) & 1u, P.K);
			P.prod[base + s0 + rank] = (uint64_t)hv | ((uint64_t)ov << 32);
		}
		__syncthreads();
	}
}

// Gather formulation for columns that exceed the shared-memory classes: products are fetched from
// A through Bent (aoff,cnt) twice; GLOBAL = tables in a global slab.  Ordered output goes to the
// second half of prod.
template <bool GLOBAL>
__global__ void __launch_bounds__(256) k_expand_gather(Params P, const uint32_t* __restrict__ list, uint32_t count,
		uint32_t htmax, uint32_t* __restrict__ slab)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	__shared__ uint32_t s_z;
	__shared__ uint32_t s_tmp[34];
	__shared__ uint32_t s_bc[NBUCKETS];
	const uint32_t tid = threadIdx.x, nt = blockDim.x;
	const uint32_t ncols = P.hi - P.lo;
	uint32_t* tbl = GLOBAL ? slab + (size_t)blockIdx.x * 5 * ((size_t)htmax + 1) : (uint32_t*)smem_raw;
	uint32_t* keys = tbl;
	uint32_t* val = keys + (htmax + 1);
	uint32_t* skeys = val + (htmax + 1);
	uint32_t* poff = skeys + (htmax + 1);
	uint32_t* cursor = poff + (htmax + 1);

	for (uint32_t it = blockIdx.x; it < count; it += gridDim.x) {
		const uint32_t li = list[it];
		const uint32_t i = P.lo + li;
		const uint32_t j0 = P.B_colptr[i], j1 = P.B_colptr[i + 1];
		const uint64_t base = P.flopptr[li];
		const uint32_t Fi = (uint32_t)(P.flopptr[li + 1] - base);
		uint32_t ht = 32; int shift = 27;
		while (ht < Fi) { ht <<= 1; --shift; }
		const uint32_t mask = ht - 1;
		for (uint32_t s = tid; s < ht; s += nt) { keys[s] = EMPTY; val[s] = 0; }
		if (tid == 0) s_z = 0;
		if (tid < NBUCKETS) s_bc[tid] = 0;
		__syncthreads();
		for (uint32_t j = j0 + tid; j < j1; j += nt) {
			uint64_t be = P.Bent[j];
			uint32_t aoff = (uint32_t)be, cnt = (uint32_t)(be >> 48) & 0x7FFFu;
			for (uint32_t e = 0; e < cnt; ++e) {
				uint32_t slot = ht_insert(keys, mask, shift, (uint32_t)P.Aent[aoff + e]);
				atomicAdd(&val[slot], 1u);
			}
		}
		__syncthreads();
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
		block_bitonic_sort(skeys, Zp);
		for (uint32_t p = tid; p < Z; p += nt) {
			uint32_t slot = ht_find(keys, mask, shift, skeys[p]);
			uint32_t len = val[slot];
			poff[p] = len;
			val[slot] = p;
			cursor[p] = 0;
			P.prow[base + p] = skeys[p];
			atomicAdd(&s_bc[bucket_of(len)], 1u);
		}
		__syncthreads();
		block_excl_scan(poff, Z, s_tmp);
		// descriptors point at the ordered copy in the second half of prod
		for (uint32_t p = tid; p < Z; p += nt) P.pdesc[base + p] = make_uint2(poff[p], (poff[p + 1] - poff[p]) | 0x80000000u);
		if (tid == 0) P.nnzC[li] = Z;
		if (tid < NBUCKETS) P.bcount[(size_t)tid * ncols + li] = s_bc[tid];
		// unordered placement, tagged with the position in B's column
		for (uint32_t j = j0 + tid; j < j1; j += nt) {
			uint64_t be = P.Bent[j];
			uint32_t aoff = (uint32_t)be, cnt = (uint32_t)(be >> 48) & 0x7FFFu;
			uint32_t v = (uint32_t)(be >> 32) & 0xFFFFu, sB = (uint32_t)(be >> 63);
			for (uint32_t e = 0; e < cnt; ++e) {
				uint64_t ae = P.Aent[aoff + e];
				uint32_t p = val[ht_find(keys, mask, shift, (uint32_t)ae)];
				uint32_t pos = poff[p] + atomicAdd(&cursor[p], 1u);
				uint32_t h = (uint32_t)(ae >> 32) & 0xFFFFu, sA = (uint32_t)(ae >> 48) & 1u;
				P.prod[base + pos] = (uint64_t)h | ((uint64_t)v << 16) | ((uint64_t)(j - j0) << 32) | ((uint64_t)(sA == sB) << 48) | ((uint64_t)p << 49);
			}
		}
		__threadfence_block();
		__syncthreads();
		// order inside each pair -> second half
		const int lenV = (int)P.read_len[i];
		for (uint32_t y = tid; y < Fi; y += nt) {
			uint64_t r = P.prod[base + y];
			uint32_t p = (uint32_t)(r >> 49), jr = (uint32_t)(r >> 32) & 0xFFFFu;
			uint32_t s0 = poff[p], s1 = poff[p + 1], rank = 0;
			for (uint32_t z = s0; z < s1; ++z) rank += (((uint32_t)(P.prod[base + z] >> 32) & 0xFFFFu) < jr);
			uint32_t hv = (uint32_t)r;
			uint32_t ov = overlap_estimate((int)P.read_len[skeys[p]], lenV, hv & 0xFFFFu, hv >> 16, (uint32_t)(r >> 48) & 1u, P.K);
			P.prod[P.prod_half + base + s0 + rank] = (uint64_t)hv | ((uint64_t)ov << 32);
		}
		__syncthreads();
	}
}

// ================================ fold ======================================================
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

struct FDesc { uint32_t row, col; unsigned long long off_len; };   // off(48) | len(16)<<48

// one warp per column: flat pair descriptors at their final output index + per-bucket work lists
// (list positions come from the scanned per-column bucket counts: no atomics, deterministic)
__global__ void k_flatten(Params P, FDesc* __restrict__ fdesc, uint32_t* __restrict__ flist)
{
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
	const uint32_t ncols = P.hi - P.lo;
	for (uint32_t li = warp; li < ncols; li += nwarps) {
		const uint32_t Z = P.nnzC[li];
		if (!Z) continue;
		const uint64_t base = P.flopptr[li];
		const uint32_t out0 = P.colptrC[li];
		uint32_t run = (lane < NBUCKETS) ? P.bcount[(size_t)lane * ncols + li] : 0;   // scanned: list offset of bucket `lane`
		for (uint32_t p0 = 0; p0 < Z; p0 += 32) {
			uint32_t p = p0 + lane;
			int b = -1;
			uint32_t g = out0 + p;
			if (p < Z) {
				uint2 d = P.pdesc[base + p];
				uint32_t len = d.y & 0x7FFFFFFFu;
				FDesc f;
				f.row = P.prow[base + p]; f.col = P.lo + li;
				f.off_len = (base + d.x + ((d.y >> 31) ? P.prod_half : 0ull)) | ((unsigned long long)len << 48);
				fdesc[g] = f;
				b = bucket_of(len);
			}
#pragma unroll
			for (int k = 0; k < NBUCKETS; ++k) {
				uint32_t m = __ballot_sync(0xFFFFFFFFu, b == k);
				uint32_t start = __shfl_sync(0xFFFFFFFFu, run, k);
				if (b == k) flist[start + __popc(m & ((1u << lane) - 1))] = g;
				if (lane == (uint32_t)k) run += __popc(m);
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

// thread per pair, P <= CAP, records already in fold order
template <int CAP>
__global__ void __launch_bounds__(128) k_fold_short(Params P, const FDesc* __restrict__ fdesc, const uint32_t* __restrict__ flist,
		int bucket)
{
	const uint32_t ncols = P.hi - P.lo;
	const uint32_t start = P.bcount[(size_t)bucket * ncols], count = P.bcount[(size_t)(bucket + 1) * ncols] - start;
	const uint32_t* list = flist + start;
	const uint32_t K = P.K, K2 = 2 * P.K;
	const int BIN = (int)P.BIN;
	for (uint32_t it = blockIdx.x * blockDim.x + threadIdx.x; it < count; it += gridDim.x * blockDim.x) {
		const uint32_t g = list[it];
		const FDesc f = fdesc[g];
		const uint32_t np = (uint32_t)(f.off_len >> 48);
		const uint64_t* rec = P.prod + (f.off_len & 0xFFFFFFFFFFFFull);
		if (CAP == 1) {
			uint64_t r = rec[0];
			store_result(P, g, f.row, 1, (uint32_t)r, 1, 1, (uint32_t)(r >> 32) & 0xFFFFu);
			continue;
		}
		uint32_t hv[CAP];
		uint16_t ov[CAP];
		bool linear = true;
		for (uint32_t a = 0; a < np; ++a) {
			uint64_t r = rec[a];
			hv[a] = (uint32_t)r; ov[a] = (uint16_t)(r >> 32);
			if (a) linear &= abs((int)ov[a] - (int)ov[a - 1]) < BIN;
		}
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
			uint16_t par[CAP], sup[CAP], live[CAP];
			uint32_t nlive = 0;
			for (uint32_t t = 0; t < np; ++t) {
				for (uint32_t q = 0; q < nlive;) {
					uint32_t b = live[q];
					if (abs((int)ov[b] - (int)ov[t]) < BIN) { par[b] = (uint16_t)t; live[q] = live[--nlive]; } else ++q;
				}
				live[nlive++] = (uint16_t)t;
				sup[t] = 0;
			}
			for (uint32_t q = 0; q < nlive; ++q) par[live[q]] = NONE16;
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

// Sequential in-place fold (any P): state per processed product s is bin overlap (bits 32..47) and
// label = creator index of its bin (bits 48..63, 0xFFFF = dropped).  Literal chainop.
__device__ void fold_pair_inplace(uint64_t* rec, uint32_t np, uint32_t K, int BIN,
		uint32_t& out_count, uint32_t& out_hv, uint32_t& out_nbins, uint32_t& out_sup, uint32_t& out_ov)
{
	uint32_t count = 0;
	for (uint32_t t = 0; t < np; ++t) {
		const uint64_t rt = rec[t];
		const uint32_t h = (uint32_t)rt & 0xFFFFu, v = (uint32_t)(rt >> 16) & 0xFFFFu, ov = (uint32_t)(rt >> 32) & 0xFFFFu;
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
		rec[t] = (rt & 0xFFFFFFFFFFFFull) | ((uint64_t)t << 48);
	}
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
	out_count = count; out_hv = (uint32_t)r; out_nbins = nbins; out_sup = best_sup & 0xFFFFu; out_ov = (uint32_t)(r >> 32) & 0xFFFFu;
}

// warp per pair, 33 <= P <= 256
constexpr int WARP_FOLD_MAX = 256;
constexpr int WARP_FOLD_WARPS = 8;

__global__ void __launch_bounds__(WARP_FOLD_WARPS * 32) k_fold_long(Params P, const FDesc* __restrict__ fdesc,
		const uint32_t* __restrict__ flist, int bucket)
{
	__shared__ uint32_t s_hv[WARP_FOLD_WARPS][WARP_FOLD_MAX];
	__shared__ uint32_t s_sup[WARP_FOLD_WARPS][WARP_FOLD_MAX];
	__shared__ uint16_t s_ov[WARP_FOLD_WARPS][WARP_FOLD_MAX];
	__shared__ uint16_t s_par[WARP_FOLD_WARPS][WARP_FOLD_MAX];
	const uint32_t FULL = 0xFFFFFFFFu;
	const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	uint32_t* shv = s_hv[w]; uint32_t* ssup = s_sup[w];
	uint16_t* sov = s_ov[w]; uint16_t* spar = s_par[w];
	const uint32_t ncols = P.hi - P.lo;
	const uint32_t start = P.bcount[(size_t)bucket * ncols], count = P.bcount[(size_t)(bucket + 1) * ncols] - start;
	const uint32_t* list = flist + start;
	const uint32_t K = P.K, K2 = 2 * P.K;
	const int BIN = (int)P.BIN;
	for (uint32_t it = blockIdx.x * WARP_FOLD_WARPS + w; it < count; it += gridDim.x * WARP_FOLD_WARPS) {
		const uint32_t g = list[it];
		const FDesc f = fdesc[g];
		const uint32_t np = (uint32_t)(f.off_len >> 48);
		uint64_t* rec = P.prod + (f.off_len & 0xFFFFFFFFFFFFull);
		const uint32_t R = (np + 31) >> 5;
		__syncwarp();
		for (uint32_t idx = lane; idx < np; idx += 32) {
			uint64_t x = rec[idx];
			shv[idx] = (uint32_t)x; sov[idx] = (uint16_t)(x >> 32);
		}
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
				uint32_t cnt, hv, nb, sup, ov;
				fold_pair_inplace(rec, np, K, BIN, cnt, hv, nb, sup, ov);
				store_result(P, g, f.row, cnt, hv, nb, sup, ov);
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
__global__ void k_fold_huge(Params P, const FDesc* __restrict__ fdesc, const uint32_t* __restrict__ flist, int bucket)
{
	const uint32_t ncols = P.hi - P.lo;
	const uint32_t start = P.bcount[(size_t)bucket * ncols], count = P.bcount[(size_t)(bucket + 1) * ncols] - start;
	const uint32_t* list = flist + start;
	for (uint32_t it = blockIdx.x * blockDim.x + threadIdx.x; it < count; it += gridDim.x * blockDim.x) {
		const uint32_t g = list[it];
		const FDesc f = fdesc[g];
		uint32_t cnt, hv, nb, sup, ov;
		fold_pair_inplace(P.prod + (f.off_len & 0xFFFFFFFFFFFFull), (uint32_t)(f.off_len >> 48), P.K, (int)P.BIN, cnt, hv, nb, sup, ov);
		store_result(P, g, f.row, cnt, hv, nb, sup, ov);
	}
}

} // namespace bk
