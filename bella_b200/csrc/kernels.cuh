// kernels.cuh -- device code of the B200 overlap SpGEMM (included once, by bella_b200.cu).
//
// The whole path is a chain of "partition into contiguous regions, then finish each region in
// shared memory" stages, so that every HBM access is either a coalesced stream or a small store
// to one of a few thousand write frontiers that live in L2 (DESIGN.md has the traffic model):
//
//   transpose  k_partition                   B's nonzeros (read-major) -> fixed-capacity buckets of consecutive k-mer ids
//              k_bucket                      one CTA per bucket: counting sort by k-mer, columns sorted by
//                                            read id, written as packed entries `Aent` + A's colptr, and the
//                                            per-output-column product count == estimateFLOP
//                                            (overlap.hpp:157-202)
//   plan       k_plan / k_units_init / k_count_units / k_classify_units
//                                            output columns -> units (a column, or a row range of a heavy
//                                            column) of at most UNIT_CAP products
//   scatter    k_scatter                     outer-product expansion on the A side: k-mer column
//                                            (r0<r1<..) emits the kept products (col r_a, row r_b), a<b,
//                                            into the unit's region (one 8-byte record per product)
//   group+fold k_group_fold<CAP,NT,EXACT>    one CTA per unit: region staged into shared memory with one
//                                            bulk async copy (TMA 1-D), distinct rows through a bitmap (two
//                                            levels when the unit spans more rows than CAP words cover;
//                                            == estimateNNZ_Hash, overlap.hpp:205-276), products
//                                            grouped by pair in B-column order (== LocalSpGEMM's visiting
//                                            order, overlap.hpp:306-341), then the semiring fold
//                                            (chain.hpp:74-150) and choose() (common.h:162-170) without
//                                            leaving shared memory
//              k_huge_pair                   a single pair with more products than fit in shared memory
//   output     k_colptr / k_compact          per-unit results -> C in CSC order, rows ascending
//   multi-GPU  k_mg_colinfo / k_mg_sum_counts / k_regroup   send-buffer cursors, summed counts, received segments -> unit regions
//   build      k_tuple_runs / k_tuple_counts / k_merge_duplicates / k_compact_B   tuples -> B in the reference's
//                                            MergeDuplicates order (src/CSC.cpp:301-479; "next" row f2)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bk {

constexpr uint32_t NONE16 = 0xFFFFu;
constexpr uint32_t FULL = 0xFFFFFFFFu;
constexpr int NCLASS = 3;
constexpr uint32_t CLASS_CAP[NCLASS] = {2048, 4096, 8192};
constexpr uint32_t UNIT_CAP = 8192;        // products per unit (largest shared-memory class)
constexpr uint32_t BUCKET_CAP = 4096;      // entries per transpose bucket
constexpr uint32_t BUCKET_WMAX = 2048;     // k-mers per transpose bucket
constexpr uint32_t MAX_SPAN_SHIFT = 22;    // a unit covers at most 2^22 rows (two-level bitmap: 4097 words)
constexpr uint32_t SHORT_FOLD = 8;         // pairs up to this many products are folded by one thread
constexpr int GF_THREADS = 256;
constexpr size_t BUCKET_SMEM = (size_t)BUCKET_CAP * 12 + ((size_t)BUCKET_WMAX + 2) * 4;

// Packed formats (64-bit)
//   Aent : row(31) | strand<<31 | pos(16)<<32 | jrank(16)<<48     A's columns, rows ascending;
//          jrank = position of the k-mer inside B's column of that read (the fold order)
//   raw  : row(31) | strandH<<31 | h(16)<<32 | jrank(16)<<48      one per kept product, in the unit's region;
//          row/strandH/h come from the row read's entry, jrank from the column read's entry
//   fin  : h(16) | v(16)<<16 | overlap(16)<<32                     grouped by pair, in fold order (shared memory)
//   out  : uint4 { row, count | nbins<<16, h | v<<16, support | overlap<<16 }  per pair, per unit

struct ColInfo { uint32_t ubase; uint32_t sh; };

struct Params {
	uint32_t n, m, lo, hi, K, BIN;
	const uint32_t* B_colptr;
	const uint32_t* B_rowids;
	const uint16_t* B_values;
	const uint8_t* B_strand;      // bit-packed, or NULL: then the strand bit is bit 31 of B_rowids (panel format)
	const uint32_t* read_len;
	const uint32_t* A_colptr;     // [m+1]
	const uint64_t* Aent;         // [nnz]
	const ColInfo* colinfo;       // [ncols]
	const uint32_t* ucol;         // [U]  local column of the unit
	const uint32_t* ucount;       // [U]  products of the unit
	const uint64_t* uptr;         // [U+1] region start (even-padded sizes)
	unsigned long long* ucur;     // [U]  scatter cursors (start at uptr)
	uint64_t* raw;                // [F + U]
	uint4* out;                   // [F + U] per-unit pair records at uptr[u] + p
	uint32_t* unnz;               // [U+1] pairs per unit
	int* err;
};

struct Meta {
	unsigned long long flops;
	unsigned int class_count[NCLASS + 1];   // + overflow
	unsigned int n_heavy_cols;
	unsigned int n_units;
	unsigned int n_refine;
};

// -DBELLA_PHASE_CLOCKS: per-phase SM cycles of the group + fold and bucket kernels, summed over CTAs (thread 0's clock
// between barriers); read back through bella_b200_debug_phases.  Profiling builds only.
#ifdef BELLA_PHASE_CLOCKS
__device__ unsigned long long g_phase[32];
#define PHASE_BEGIN() long long ph_t0_ = clock64()
#define PHASE(i) do { if (threadIdx.x == 0) { const long long t_ = clock64(); atomicAdd(&g_phase[i], (unsigned long long)(t_ - ph_t0_)); ph_t0_ = t_; } } while (0)
#else
#define PHASE_BEGIN() do {} while (0)
#define PHASE(i) do {} while (0)
#endif

__device__ __forceinline__ void set_err(int* err, int code) { atomicCAS(err, 0, code); }
__device__ __forceinline__ uint32_t getbit(const uint8_t* __restrict__ bits, uint64_t i) { return (bits[i >> 3] >> (i & 7)) & 1u; }
__device__ __forceinline__ uint32_t ent_row(uint64_t e) { return (uint32_t)e & 0x7FFFFFFFu; }
__device__ __forceinline__ uint32_t strand_of(const uint8_t* __restrict__ bits, const uint32_t* __restrict__ rowids, uint64_t j)
{
	return bits ? getbit(bits, j) : rowids[j] >> 31;
}

// ================================ block helpers =============================================

// exclusive scan of a[0..n) in place (T = uint16_t or uint32_t), block-wide; total left in a[n].
// s_tmp needs 34 words.  Ends with a barrier.
template <class T>
__device__ void block_excl_scan(T* a, uint32_t n, uint32_t* s_tmp)
{
	const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
	if (tid == 0) s_tmp[32] = 0;
	__syncthreads();
	for (uint32_t base = 0; base < n; base += blockDim.x) {
		uint32_t idx = base + tid;
		uint32_t x = idx < n ? (uint32_t)a[idx] : 0, v = x;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(FULL, v, o); if (lane >= (uint32_t)o) v += y; }
		if (lane == 31) s_tmp[wid] = v;
		__syncthreads();
		if (wid == 0) {
			uint32_t w = lane < nw ? s_tmp[lane] : 0, ws = w;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(FULL, ws, o); if (lane >= (uint32_t)o) ws += y; }
			s_tmp[lane] = ws - w;
			if (lane == 31) s_tmp[33] = ws;
		}
		__syncthreads();
		uint32_t carry = s_tmp[32];
		if (idx < n) a[idx] = (T)(v - x + s_tmp[wid] + carry);
		__syncthreads();
		if (tid == 0) s_tmp[32] = carry + s_tmp[33];
		__syncthreads();
	}
	if (tid == 0) a[n] = (T)s_tmp[32];
	__syncthreads();
}

// pre[i] = sum_{j<i} popc(bits[j]), pre[n] = total.  Ends with a barrier.
template <class T>
__device__ void block_popc_scan(const uint32_t* bits, T* pre, uint32_t n, uint32_t* s_tmp)
{
	for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) pre[i] = (T)__popc(bits[i]);
	__syncthreads();
	block_excl_scan<T>(pre, n, s_tmp);
}

// atomicAdd on a 16-bit counter packed two per 32-bit word (no carry: counts stay < 65536)
__device__ __forceinline__ uint32_t atomic_add16(uint16_t* base, uint32_t idx, uint32_t v)
{
	uint32_t sh = (idx & 1u) * 16u;
	uint32_t old = atomicAdd((uint32_t*)base + (idx >> 1), v << sh);
	return (old >> sh) & 0xFFFFu;
}

// ---- 1-D bulk async copy global -> shared (TMA), completion on an mbarrier ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
	asm volatile(
		"{\n .reg .pred p;\n"
		"WAIT_%=:\n"
		" mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
		" @p bra DONE_%=;\n"
		" bra WAIT_%=;\n"
		"DONE_%=:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ================================ transpose =================================================

// One warp per read (column of B): its nonzeros go to their k-mer bucket (W consecutive k-mer ids,
// a fixed-capacity region of BUCKET_CAP 16-byte records {k-mer id, -, entry}).  The write frontier is
// one open sector per bucket, so the small stores merge in L2.  Rows below lo never matter.
__global__ void __launch_bounds__(256) k_partition(uint32_t n, uint32_t lo, uint32_t klo, uint32_t khi, const uint32_t* __restrict__ Bcolptr,   // reads [lo, n)
		const uint32_t* __restrict__ Brow, const uint16_t* __restrict__ Bval, const uint8_t* __restrict__ Bstrand,
		uint32_t W, uint32_t* __restrict__ bcnt, uint4* __restrict__ part, int* err)
{
	const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
	for (uint32_t i = lo + warp; i < n; i += nwarps) {
		const uint32_t j0 = Bcolptr[i], j1 = Bcolptr[i + 1];
		if (j1 - j0 > 65536u) { if (lane == 0) set_err(err, -4); continue; }
		for (uint32_t jb = j0; jb < j1; jb += 128) {
			uint32_t c[4], b[4], q[4];
			bool mine[4];                                      // k-mers outside [klo, khi) belong to another GPU's transpose
#pragma unroll
			for (int u = 0; u < 4; ++u) {
				uint32_t j = jb + u * 32 + lane;
				mine[u] = false;
				if (j < j1) {
					c[u] = Brow[j];
					const uint32_t kid = Bstrand ? c[u] : c[u] & 0x7FFFFFFFu;
					mine[u] = kid >= klo && kid < khi;
					if (mine[u]) { b[u] = (kid - klo) / W; q[u] = atomicAdd(&bcnt[b[u]], 1u); }
				}
			}
#pragma unroll
			for (int u = 0; u < 4; ++u) {
				uint32_t j = jb + u * 32 + lane;
				if (mine[u]) {
					if (q[u] >= BUCKET_CAP) { set_err(err, -6); continue; }
					const uint32_t st = Bstrand ? getbit(Bstrand, j) : c[u] >> 31;
					const uint64_t e = (uint64_t)i | ((uint64_t)st << 31) | ((uint64_t)Bval[j] << 32) | ((uint64_t)(j - j0) << 48);
					part[(size_t)b[u] * BUCKET_CAP + q[u]] = make_uint4(Bstrand ? c[u] : c[u] & 0x7FFFFFFFu, 0u, (uint32_t)e, (uint32_t)(e >> 32));
				}
			}
		}
	}
}

// One CTA per bucket: counting sort by k-mer in shared memory, each column sorted by read id,
// coalesced write of Aent and A's colptr, product counts per output column.
__global__ void __launch_bounds__(256) k_bucket(uint32_t klo, uint32_t m, uint32_t lo, uint32_t hi, uint32_t W, uint32_t nb,
		const uint32_t* __restrict__ boff, const uint4* __restrict__ part,
		uint32_t* __restrict__ Acolptr, uint64_t* __restrict__ Aent, uint32_t* __restrict__ flop32, const int* err)
{
	extern __shared__ __align__(16) unsigned char bsm[];
	uint64_t* E = (uint64_t*)bsm;                              // [BUCKET_CAP]
	uint32_t* tmp = (uint32_t*)(E + BUCKET_CAP);               // [BUCKET_CAP]
	uint32_t* off = tmp + BUCKET_CAP;                          // [BUCKET_WMAX + 2]
	__shared__ uint32_t s_tmp[34];
	const uint32_t tid = threadIdx.x, nt = blockDim.x;
	if (*err != 0) return;                                     // a bucket overflowed: the host retries with narrower buckets
	for (uint32_t b = blockIdx.x; b < nb; b += gridDim.x) {
		const uint32_t o0 = boff[b], size = boff[b + 1] - o0;
		const uint32_t kbase = b * W, kw = min(W, m - kbase);      // k-mer ids are klo + kbase + k; A's colptr is local to [klo, klo + m)
		const uint4* src = part + (size_t)b * BUCKET_CAP;
		PHASE_BEGIN();
		for (uint32_t k = tid; k <= kw; k += nt) off[k] = 0;
		__syncthreads();
		for (uint32_t x = tid; x < size; x += nt) {
			const uint4 r = src[x];
			const uint32_t k = r.x - klo - kbase;
			const uint32_t arr = atomicAdd(&off[k], 1u);
			tmp[x] = k | (arr << 12);
			E[x] = (uint64_t)r.z | ((uint64_t)r.w << 32);
		}
		__syncthreads();
		PHASE(16);
		block_excl_scan<uint32_t>(off, kw, s_tmp);
		PHASE(17);
		// permute in place through registers: every thread first reads its elements, then all write
		uint64_t ev[BUCKET_CAP / 256];
#pragma unroll
		for (int q = 0; q < (int)(BUCKET_CAP / 256); ++q) { uint32_t x = tid + q * 256; ev[q] = x < size ? E[x] : 0; }
		__syncthreads();
#pragma unroll
		for (int q = 0; q < (int)(BUCKET_CAP / 256); ++q) {
			uint32_t x = tid + q * 256;
			if (x < size) { uint32_t t = tmp[x]; E[off[t & 0xFFFu] + (t >> 12)] = ev[q]; }
		}
		__syncthreads();
		PHASE(18);
		for (uint32_t k = tid; k < kw; k += nt) {
			const uint32_t s = off[k], e = off[k + 1];
			Acolptr[kbase + k] = o0 + s;
			for (uint32_t a = s + 1; a < e; ++a) {          // insertion sort by read id (columns are 2..8 long)
				uint64_t x = E[a];
				uint32_t p = a;
				while (p > s && ent_row(E[p - 1]) > ent_row(x)) { E[p] = E[p - 1]; --p; }
				E[p] = x;
			}
			for (uint32_t a = s; a + 1 < e; ++a) {
				uint32_t r = ent_row(E[a]);
				if (r >= lo && r < hi) atomicAdd(&flop32[r - lo], e - 1 - a);
			}
		}
		if (b == nb - 1 && tid == 0) Acolptr[m] = o0 + size;
		__syncthreads();
		PHASE(19);
		for (uint32_t x = tid; x < size; x += nt) Aent[o0 + x] = E[x];
		__syncthreads();
		PHASE(20);
	}
}

// ================================ plan ======================================================

// Units of a column: row buckets of width 2^sh over (i, n).  Light columns: one unit (sh = 31).
__global__ void k_plan(uint32_t n, uint32_t lo, uint32_t ncols, const uint32_t* __restrict__ flop32,
		const uint8_t* __restrict__ refine, uint32_t* __restrict__ nunits, uint8_t* __restrict__ shv, Meta* meta)
{
	unsigned long long fsum = 0;
	uint32_t heavy = 0;
	for (uint32_t li = blockIdx.x * blockDim.x + threadIdx.x; li < ncols; li += gridDim.x * blockDim.x) {
		const uint32_t f = flop32[li], i = lo + li;
		fsum += f;
		uint32_t sh = 31, nb = f ? 1u : 0u;
		const uint32_t span = n - 1 - i;                     // rows i+1 .. n-1
		if (f && (f > UNIT_CAP || span > (1u << MAX_SPAN_SHIFT))) {
			uint32_t parts = (uint32_t)min((unsigned long long)span, ((unsigned long long)f * 8 + UNIT_CAP - 1) / UNIT_CAP);
			uint32_t width = max(1u, span / max(parts, 1u));
			sh = 31 - __clz(width);
			sh = min(sh, MAX_SPAN_SHIFT);
			uint32_t fine = refine[li];                        // each refinement round: 8x narrower row buckets
			sh = sh > 3 * fine ? sh - 3 * fine : 0;
			nb = ((n - 1) >> sh) - ((i + 1) >> sh) + 1;
			++heavy;
		}
		nunits[li] = nb;
		shv[li] = (uint8_t)sh;
	}
	for (int o = 16; o; o >>= 1) { fsum += __shfl_xor_sync(FULL, fsum, o); heavy += __shfl_xor_sync(FULL, heavy, o); }
	if ((threadIdx.x & 31) == 0) {
		if (fsum) atomicAdd(&meta->flops, fsum);
		if (heavy) atomicAdd(&meta->n_heavy_cols, heavy);
	}
}

__global__ void k_units_init(uint32_t ncols, uint32_t ucap, const uint32_t* __restrict__ flop32, const uint32_t* __restrict__ ubase,
		const uint8_t* __restrict__ shv, ColInfo* __restrict__ colinfo, uint32_t* __restrict__ ucol, uint32_t* __restrict__ ucount,
		Meta* meta, int* err)
{
	const uint32_t U = ubase[ncols];
	if (blockIdx.x == 0 && threadIdx.x == 0) meta->n_units = U;
	if (U > ucap) { if (blockIdx.x == 0 && threadIdx.x == 0) set_err(err, -7); return; }
	for (uint32_t li = blockIdx.x * blockDim.x + threadIdx.x; li < ncols; li += gridDim.x * blockDim.x) {
		const uint32_t u0 = ubase[li], u1 = ubase[li + 1];
		colinfo[li] = ColInfo{u0, shv[li]};
		if (shv[li] == 31) { if (u1 > u0) { ucol[u0] = li; ucount[u0] = flop32[li]; } }
		else for (uint32_t u = u0; u < u1; ++u) { ucol[u] = li; ucount[u] = 0; }
	}
}

__device__ __forceinline__ uint32_t unit_of(const ColInfo ci, uint32_t i, uint32_t row)
{
	return ci.sh == 31 ? ci.ubase : ci.ubase + (row >> ci.sh) - ((i + 1) >> ci.sh);
}

// products per unit of the heavy columns (light columns were filled by k_units_init)
__global__ void __launch_bounds__(256) k_count_units(uint32_t m, uint32_t lo, uint32_t hi, const uint32_t* __restrict__ Acolptr,
		const uint64_t* __restrict__ Aent, const ColInfo* __restrict__ colinfo, uint32_t* __restrict__ ucount, const Meta* meta,
		const int* err)
{
	if (meta->n_heavy_cols == 0 || *err != 0) return;
	for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < m; c += gridDim.x * blockDim.x) {
		const uint32_t s = Acolptr[c], e = Acolptr[c + 1];
		for (uint32_t a = s; a + 1 < e; ++a) {
			const uint32_t ra = ent_row(Aent[a]);
			if (ra < lo || ra >= hi) continue;
			const ColInfo ci = colinfo[ra - lo];
			if (ci.sh == 31) continue;
			for (uint32_t b = a + 1; b < e; ++b) atomicAdd(&ucount[unit_of(ci, ra, ent_row(Aent[b]))], 1u);
		}
	}
}

// classes by unit size; units that do not fit ask for a finer split of their column (refine) unless
// they already are a single row (sh == 0): those go to the huge-pair list (class NCLASS).
__global__ void k_classify_units(uint32_t ucap, const uint32_t* __restrict__ ucol, const uint32_t* __restrict__ ucount,
		const ColInfo* __restrict__ colinfo, const uint64_t* __restrict__ uptr, unsigned long long* __restrict__ ucur,
		uint32_t* __restrict__ lists, uint8_t* __restrict__ refine, uint32_t round, Meta* meta, const int* err)
{
	if (*err != 0) return;
	const uint32_t U = meta->n_units;
	for (uint32_t u = blockIdx.x * blockDim.x + threadIdx.x; u < U; u += gridDim.x * blockDim.x) {
		const uint32_t f = ucount[u];
		ucur[u] = uptr[u];
		if (!f) continue;
		int c = f <= CLASS_CAP[0] ? 0 : f <= CLASS_CAP[1] ? 1 : f <= CLASS_CAP[2] ? 2 : 3;
		if (c == 3) {
			const uint32_t li = ucol[u];
			if (colinfo[li].sh != 0) { refine[li] = (uint8_t)(round + 1); atomicAdd(&meta->n_refine, 1u); continue; }
		}
		uint32_t idx = atomicAdd(&meta->class_count[c], 1u);
		lists[(size_t)c * ucap + idx] = u;
	}
}

// ================================ scatter ===================================================
// Thread per k-mer column (r_0 < r_1 < ...): entry a owns the run of products (col r_a, row r_b), b > a.
// A light column's run is written contiguously into the column's region after one cursor atomic.
__global__ void __launch_bounds__(256) k_scatter(uint32_t m, uint32_t lo, uint32_t hi, const uint32_t* __restrict__ Acolptr,
		const uint64_t* __restrict__ Aent, const ColInfo* __restrict__ colinfo, unsigned long long* __restrict__ ucur,
		uint64_t* __restrict__ raw)
{
	constexpr int D = 8;
	constexpr uint64_t LOW48 = 0x0000FFFFFFFFFFFFull;
	for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < m; c += gridDim.x * blockDim.x) {
		const uint32_t s = Acolptr[c], e = Acolptr[c + 1], d = e - s;
		if (d < 2) continue;
		if (d <= D) {
			uint64_t ent[D];
			unsigned long long q[D];
			uint32_t heavy = 0;
#pragma unroll
			for (int a = 0; a < D; ++a) ent[a] = a < (int)d ? Aent[s + a] : 0;
#pragma unroll
			for (int a = 0; a < D - 1; ++a) {
				q[a] = ~0ull;
				if (a + 1 < (int)d) {
					const uint32_t ra = ent_row(ent[a]);
					if (ra >= lo && ra < hi) {
						const ColInfo ci = colinfo[ra - lo];
						if (ci.sh == 31) q[a] = atomicAdd(&ucur[ci.ubase], (unsigned long long)(d - 1 - a));
						else heavy |= 1u << a;
					}
				}
			}
#pragma unroll
			for (int a = 0; a < D - 1; ++a) {
				if (q[a] != ~0ull) {
					const uint64_t top = ent[a] & ~LOW48;
					unsigned long long p = q[a];
#pragma unroll
					for (int b = a + 1; b < D; ++b)
						if (b < (int)d) raw[p++] = (ent[b] & LOW48) | top;
				}
			}
			if (heavy) {
#pragma unroll
				for (int a = 0; a < D - 1; ++a) {
					if (heavy >> a & 1u) {
						const uint32_t ra = ent_row(ent[a]);
						const ColInfo ci = colinfo[ra - lo];
						const uint64_t top = ent[a] & ~LOW48;
#pragma unroll
						for (int b = a + 1; b < D; ++b)
							if (b < (int)d) raw[atomicAdd(&ucur[unit_of(ci, ra, ent_row(ent[b]))], 1ull)] = (ent[b] & LOW48) | top;
					}
				}
			}
		} else {
			for (uint32_t a = s; a + 1 < e; ++a) {
				const uint64_t ea = Aent[a];
				const uint32_t ra = ent_row(ea);
				if (ra < lo || ra >= hi) continue;
				const ColInfo ci = colinfo[ra - lo];
				const uint64_t top = ea & ~LOW48;
				if (ci.sh == 31) {
					unsigned long long p = atomicAdd(&ucur[ci.ubase], (unsigned long long)(e - 1 - a));
					for (uint32_t b = a + 1; b < e; ++b) raw[p++] = (Aent[b] & LOW48) | top;
				} else {
					for (uint32_t b = a + 1; b < e; ++b) {
						const uint64_t eb = Aent[b];
						raw[atomicAdd(&ucur[unit_of(ci, ra, ent_row(eb))], 1ull)] = (eb & LOW48) | top;
					}
				}
			}
		}
	}
}

// ================================ the semiring ==============================================

// multiop -> overlapop (chain.hpp:47-71), checkstrand replaced by the strand-bit comparison.
__device__ __forceinline__ uint32_t overlap_estimate(int lenH, int lenV, uint32_t h, uint32_t v, uint32_t oriented, uint32_t K)
{
	uint32_t hh = oriented ? h : ((uint32_t)lenH - h - K) & 0xFFFFu;   // unsigned short begpH, wraps
	uint32_t endH = (hh + K) & 0xFFFFu, endV = (v + K) & 0xFFFFu;
	int m1 = (int)min(hh, v);
	int m2 = min(lenH - (int)endH, lenV - (int)endV);
	return (uint32_t)(m1 + m2 + (int)K) & 0xFFFFu;                      // stored into vector<unsigned short>
}

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

// Far test (chain.hpp:121 through operator>, chain.hpp:93-97): |h_t - h_s| > K && |v_t - v_s| > K.
// Packed form: with the key (K - h_s | K - v_s) per 16-bit half, x + key per half is d + K mod 2^16, and
// far <=> both halves > 2K <=> min(x + key, 2K+1) == 2K+1 per half: one VIADDMNMX.U16x2 and one compare.
// It is exact whenever every position is <= 65535 - K (k-mer start positions of reads shorter than
// 64 Ki always are); EXACT = true is the 32-bit form, which holds for any u16 values.
struct FarKey { uint32_t a, b; };
template <bool EXACT> __device__ __forceinline__ FarKey far_key(uint32_t x, uint32_t K)
{
	FarKey k;
	if (EXACT) { k.a = K - (x & 0xFFFFu); k.b = K - (x >> 16); }
	else { k.a = ((K - (x & 0xFFFFu)) & 0xFFFFu) | ((K - (x >> 16)) << 16); k.b = 0; }
	return k;
}
template <bool EXACT> __device__ __forceinline__ uint32_t far_limit(uint32_t K) { return EXACT ? 2 * K : (2 * K + 1) * 0x10001u; }
template <bool EXACT> __device__ __forceinline__ bool is_far(uint32_t x, const FarKey k, uint32_t lim)
{
	if (EXACT) return ((x & 0xFFFFu) + k.a) > lim && ((x >> 16) + k.b) > lim;
	return __vminu2(__vadd2(x, k.a), lim) == lim;
}

struct PairResult { uint32_t count, hv, nbins, sup, ov; };

__device__ __forceinline__ uint4 pack_result(uint32_t row, const PairResult& r)
{
	return make_uint4(row, (r.count & 0xFFFFu) | (r.nbins << 16), r.hv, (r.sup & 0xFFFFu) | (r.ov << 16));
}

// the rare case of fold_short: consecutive overlap estimates further apart than binSize (several bins).
// Out of line and rolled: it keeps the hot kernel's instruction footprint small.
template <bool EXACT>
__device__ __noinline__ PairResult fold_short_forest(const uint32_t (&hvr)[SHORT_FOLD], const uint16_t (&ovr)[SHORT_FOLD], uint32_t np, uint32_t K, int BIN)
{
	constexpr int CAP = SHORT_FOLD;
	const uint32_t lim = far_limit<EXACT>(K);
	uint32_t hv[CAP];
	uint16_t ov[CAP];
	uint8_t par[CAP], sup[CAP];
#pragma unroll
	for (int a = 0; a < CAP; ++a) { hv[a] = hvr[a]; ov[a] = ovr[a]; }
#pragma unroll 1
	for (uint32_t b = 0; b < np; ++b) {
		uint32_t t = b + 1;
		while (t < np && abs((int)ov[t] - (int)ov[b]) >= BIN) ++t;
		par[b] = t < np ? (uint8_t)t : (uint8_t)0xFF;
		sup[b] = 0;
	}
	uint32_t csum = 0;
#pragma unroll 1
	for (uint32_t s = 0; s < np; ++s) {
		const FarKey key = far_key<EXACT>(hv[s], K);
		uint32_t a = par[s], last = s;
		while (a != 0xFF && is_far<EXACT>(hv[a], key, lim)) { ++csum; last = a; a = par[a]; }
		if (a == 0xFF) ++sup[last];
	}
	uint32_t best = 0, bt = 0, nb = 0;
#pragma unroll 1
	for (uint32_t t = 0; t < np; ++t)
		if (par[t] == 0xFF) { ++nb; if (sup[t] >= best) { best = sup[t]; bt = t; } }
	PairResult R;
	R.count = (np + csum) & 0xFFFFu; R.hv = hv[bt]; R.nbins = nb; R.sup = best; R.ov = ov[bt];
	return R;
}

// one thread, P <= SHORT_FOLD, products in fold order (hv = h | v<<16, ov = overlap estimate)
template <bool EXACT>
__device__ __forceinline__ PairResult fold_short(const uint32_t (&hv)[SHORT_FOLD], const uint16_t (&ov)[SHORT_FOLD], uint32_t np, uint32_t K, int BIN)
{
	constexpr int CAP = SHORT_FOLD;
	const uint32_t lim = far_limit<EXACT>(K);
	bool linear = true;
#pragma unroll
	for (int a = 1; a < CAP; ++a)
		if (a < (int)np) linear &= abs((int)ov[a] - (int)ov[a - 1]) < BIN;
	PairResult R;
	uint32_t csum = 0;
	if (np == 1) { R.count = 1; R.hv = hv[0]; R.nbins = 1; R.sup = 1; R.ov = ov[0]; return R; }
	if (linear) {
		uint32_t surv = 0, last_hv = 0, last_ov = 0;
#pragma unroll
		for (int s = 0; s < CAP; ++s) {
			if (s < (int)np) {
				const FarKey key = far_key<EXACT>(hv[s], K);
				bool alive = true;
#pragma unroll
				for (int t = s + 1; t < CAP; ++t) {
					if (t < (int)np) { alive = alive && is_far<EXACT>(hv[t], key, lim); csum += alive; }
				}
				surv += alive;
				last_hv = hv[s]; last_ov = ov[s];
			}
		}
		R.count = (np + csum) & 0xFFFFu; R.hv = last_hv; R.nbins = 1; R.sup = surv; R.ov = last_ov;
		return R;
	}
	return fold_short_forest<EXACT>(hv, ov, np, K, BIN);
}

// The forest part shared by the cooperative folds: parents, ancestor walks, best root.
// Threads tid0, tid0+stride, ... of the group take the products; sync() separates the phases.
template <class Sync>
__device__ __noinline__ void fold_forest(const uint32_t* hv, const uint16_t* ov, uint16_t* par, uint32_t* sup, uint32_t P, uint32_t K,
		int BIN, uint32_t tid0, uint32_t stride, Sync sync, uint32_t& csum, uint32_t& nroots, uint32_t& best)
{
	const uint32_t lim = far_limit<true>(K);
	for (uint32_t b = tid0; b < P; b += stride) {
		const int ob = (int)ov[b];
		uint32_t t = b + 1;
		while (t < P && abs((int)ov[t] - ob) >= BIN) ++t;
		par[b] = t < P ? (uint16_t)t : (uint16_t)NONE16;
		sup[b] = 0;
	}
	sync();
	for (uint32_t s = tid0; s < P; s += stride) {
		const FarKey key = far_key<true>(hv[s], K);
		uint32_t a = par[s], last = s;
		while (a != NONE16 && is_far<true>(hv[a], key, lim)) { ++csum; last = a; a = par[a]; }
		if (a == NONE16) atomicAdd(&sup[last], 1u);
	}
	sync();
	for (uint32_t idx = tid0; idx < P; idx += stride)
		if (par[idx] == NONE16) { ++nroots; best = max(best, (sup[idx] << 16) | idx); }
	for (int o = 16; o; o >>= 1) {
		csum += __shfl_xor_sync(FULL, csum, o); nroots += __shfl_xor_sync(FULL, nroots, o);
		best = max(best, __shfl_xor_sync(FULL, best, o));
	}
}

// linear case, one warp's share: rounds r0 = first, first + step, ... of 32 products s (one per lane).
// Every later product t is broadcast from memory and tested against the 32 lanes at once; per block
// of 32 t's a lane collects the "near" bits and its first near t ends its walk:
//     c_s = (first near t > s, or P) - s - 1,   s survives iff there is none.
template <bool EXACT>
__device__ __forceinline__ void fold_linear_rounds(const uint32_t* hv, uint32_t P, uint32_t K, uint32_t first, uint32_t step,
		uint32_t& csum, uint32_t& surv)
{
	const uint32_t lane = threadIdx.x & 31, lim = far_limit<EXACT>(K);
	for (uint32_t r0 = first; r0 < P; r0 += step) {
		const uint32_t s = r0 + lane;
		const bool valid = s < P;
		const FarKey key = far_key<EXACT>(valid ? hv[s] : 0u, K);
		bool alive = valid;
		uint32_t tfirst = P;
		for (uint32_t tb = r0; tb < P; tb += 32) {
			uint32_t near = 0;
			if (tb + 32 <= P) {
#pragma unroll 1
				for (uint32_t c0 = 0; c0 < 32; c0 += 8) {
					uint32_t m = 0;
#pragma unroll
					for (int c = 0; c < 8; ++c)
						if (!is_far<EXACT>(hv[tb + c0 + c], key, lim)) m |= 1u << c;
					near |= m << c0;
				}
			} else {
				for (uint32_t c = 0; tb + c < P; ++c)
					if (!is_far<EXACT>(hv[tb + c], key, lim)) near |= 1u << c;
			}
			if (tb == r0) near &= ~((2u << lane) - 1u);              // only t > s
			if (alive && near) { tfirst = tb + __ffs(near) - 1; alive = false; }
			if (!__any_sync(FULL, alive)) break;
		}
		if (valid) { csum += tfirst - s - 1; surv += alive; }
	}
	for (int o = 16; o; o >>= 1) { csum += __shfl_xor_sync(FULL, csum, o); surv += __shfl_xor_sync(FULL, surv, o); }
}

// Whole-CTA fold of one pair (hv/ov/par/sup in shared or global memory, hv 16-byte aligned at
// index 0).  part[0..2] (shared, zeroed by the caller) combines the warps.  All threads return the result.
__device__ __noinline__ PairResult fold_cta(const uint32_t* hv, const uint16_t* ov, uint16_t* par, uint32_t* sup, uint32_t P, uint32_t K, int BIN,
		uint32_t* part)
{
	const uint32_t tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nw = blockDim.x >> 5;
	bool lin = true;
#pragma unroll 1
	for (uint32_t t = 1 + tid; t < P; t += blockDim.x) lin &= abs((int)ov[t] - (int)ov[t - 1]) < BIN;
	const bool linear = (bool)__syncthreads_and(lin);
	uint32_t csum = 0;
	PairResult R;
	if (linear) {
		uint32_t surv = 0;
		fold_linear_rounds<true>(hv, P, K, 32 * w, 32 * nw, csum, surv);
		if (lane == 0) { atomicAdd(&part[0], csum); atomicAdd(&part[1], surv); }
		__syncthreads();
		R.count = (P + part[0]) & 0xFFFFu; R.hv = hv[P - 1]; R.nbins = 1; R.sup = part[1]; R.ov = ov[P - 1];
		return R;
	}
	uint32_t nroots = 0, best = 0;
	fold_forest(hv, ov, par, sup, P, K, BIN, tid, blockDim.x, [] { __syncthreads(); }, csum, nroots, best);
	if (lane == 0) { atomicAdd(&part[0], csum); atomicAdd(&part[1], nroots); atomicMax(&part[2], best); }
	__syncthreads();
	best = part[2];
	R.count = (P + part[0]) & 0xFFFFu; R.hv = hv[best & 0xFFFFu]; R.nbins = part[1]; R.sup = best >> 16; R.ov = ov[best & 0xFFFFu];
	return R;
}

constexpr uint32_t WSCR_WORDS = 192;       // per-warp scratch: position bitmap [128] + its prefix u16[128]
constexpr uint32_t JR_BITMAP_MAX = 4096;   // columns of B up to this length rank through the bitmap

// One warp folds a pair of P > SHORT_FOLD products (hv/ov in fold order; hv + P0 is the pair's slice of a
// 16-byte aligned array, see fold_linear_rounds).  `own` is the pair's own 8*P-byte scratch.
template <bool EXACT>
__device__ __forceinline__ PairResult warp_fold_pair(const uint32_t* hv, const uint16_t* ov, uint64_t* own, uint32_t P, uint32_t K, int BIN,
		uint32_t lane)
{
	bool lin = true;
#pragma unroll 1
	for (uint32_t t = 1 + lane; t < P; t += 32) lin &= abs((int)ov[t] - (int)ov[t - 1]) < BIN;
	const bool linear = __all_sync(FULL, lin);
	uint32_t csum = 0;
	PairResult R;
	if (linear) {
		uint32_t surv = 0;
		fold_linear_rounds<EXACT>(hv, P, K, 0, 32, csum, surv);
		R.count = (P + csum) & 0xFFFFu; R.hv = hv[P - 1]; R.nbins = 1; R.sup = surv; R.ov = ov[P - 1];
		return R;
	}
	uint16_t* par = (uint16_t*)own;
	uint32_t* sup = (uint32_t*)own + ((P + 1) >> 1);
	uint32_t nroots = 0, best = 0;
	fold_forest(hv, ov, par, sup, P, K, BIN, lane, 32, [] { __syncwarp(); }, csum, nroots, best);
	R.count = (P + csum) & 0xFFFFu; R.hv = hv[best & 0xFFFFu]; R.nbins = nroots; R.sup = best >> 16; R.ov = ov[best & 0xFFFFu];
	return R;
}

// ================================ group + fold ==============================================

// exclusive scan by contiguous per-thread chunks: out[i] = sum_{j<i} val(j), out[n] = total (returned).
// Three barriers whatever n is; `out` may alias the array val() reads (each element is read before
// it is written, by the same thread).  s_tmp needs 33 words.
template <class T, class F>
__device__ __forceinline__ uint32_t block_scan_chunked(T* out, uint32_t n, F val, uint32_t* s_tmp)
{
	const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
	const uint32_t per = ((n + blockDim.x - 1) / blockDim.x) | 1u;     // odd stride: no bank conflicts
	const uint32_t i0 = min(tid * per, n), i1 = min(i0 + per, n);
	uint32_t sum = 0;
	#pragma unroll 1
	for (uint32_t i = i0; i < i1; ++i) sum += val(i);
	uint32_t incl = sum;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(FULL, incl, o); if (lane >= (uint32_t)o) incl += y; }
	if (lane == 31) s_tmp[wid] = incl;
	__syncthreads();
	if (wid == 0) {
		uint32_t w = lane < nw ? s_tmp[lane] : 0, ws = w;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(FULL, ws, o); if (lane >= (uint32_t)o) ws += y; }
		s_tmp[lane] = ws - w;
		if (lane == 31) s_tmp[32] = ws;
	}
	__syncthreads();
	uint32_t run = s_tmp[wid] + incl - sum;
	#pragma unroll 1
	for (uint32_t i = i0; i < i1; ++i) { const uint32_t v = val(i); out[i] = (T)run; run += v; }
	const uint32_t total = s_tmp[32];
	if (tid == 0) out[n] = (T)total;
	__syncthreads();
	return total;
}

// sorting network for 8 keys (19 compare-exchanges)
__device__ __forceinline__ void cex(uint64_t& a, uint64_t& b) { const uint64_t lo = min(a, b), hi = max(a, b); a = lo; b = hi; }
__device__ __forceinline__ void sort8(uint64_t (&k)[8])
{
	cex(k[0], k[1]); cex(k[2], k[3]); cex(k[4], k[5]); cex(k[6], k[7]);
	cex(k[0], k[2]); cex(k[1], k[3]); cex(k[4], k[6]); cex(k[5], k[7]);
	cex(k[1], k[2]); cex(k[5], k[6]); cex(k[0], k[4]); cex(k[3], k[7]);
	cex(k[1], k[5]); cex(k[2], k[6]);
	cex(k[1], k[4]); cex(k[3], k[6]);
	cex(k[2], k[4]); cex(k[3], k[5]);
	cex(k[3], k[4]);
}

template <int CAP>
struct GF {
	static constexpr size_t PROD = 0;                                  // u64[CAP]  raw -> packed (h, jr, pair, slot); later hv u32[CAP] + ov u16[CAP] of the long pairs
	static constexpr size_t REC = PROD + 8 * (size_t)CAP;              // u64[CAP]  bits u32[CAP] + pre u16[CAP+2], then the products sorted by pair
	static constexpr size_t CNT = REC + 8 * (size_t)CAP + 16;          // u16[CAP+2] per pair count -> offsets
	static constexpr size_t ROW = CNT + 2 * (size_t)CAP + 16;          // u32[CAP]  row id of the pair
	static constexpr size_t LST = ROW + 4 * (size_t)CAP;               // u16[CAP/8] pairs longer than SHORT_FOLD
	static constexpr size_t L1 = LST + 2 * ((size_t)CAP / 8);          // u32[l1cap+1] level-1 bitmap + u32[l1cap+2] prefix (two-level units only)
	static size_t bytes(uint32_t l1cap) { return L1 + 8 * ((size_t)l1cap + 2); }
};

// multiply of one product (overlapop) -> key ordered by the position in B's column:
// jrank(16)<<48 | overlap(16)<<32 | v(16)<<16 | h(16)
__device__ __forceinline__ uint64_t product_key(const Params& P, uint64_t rec, uint32_t j0, int lenH, int lenV, bool& wide)
{
	const uint32_t h = (uint32_t)rec & 0xFFFFu, jr = ((uint32_t)rec >> 16), sH = (uint32_t)(rec >> 32) & 1u;
	const uint32_t jg = j0 + jr;
	const uint32_t v = P.B_values[jg], sV = strand_of(P.B_strand, P.B_rowids, jg);
	const uint32_t ov = overlap_estimate(lenH, lenV, h, v, sH == sV, P.K);
	wide |= max(h, v) > 65535u - P.K;                              // positions this large need the 32-bit far test
	return ((uint64_t)jr << 48) | ((uint64_t)ov << 32) | (uint64_t)(h | (v << 16));
}

// One warp: multiply the products of a long pair (rec, arrival order) and store them in fold order
// (position in B's column) as hv/ov.  The positions of one pair are distinct, so a bitmap over
// them ranks in O(len + L/32).  Returns (warp-uniform) whether a position needs the exact far test.
__device__ __forceinline__ bool warp_prepare_pair(const Params& P, const uint64_t* rec, uint32_t* hv, uint16_t* ov, uint32_t len, uint32_t j0,
		uint32_t L, int lenH, int lenV, uint32_t* scr, uint32_t lane)
{
	bool wide = false;
	if (L <= JR_BITMAP_MAX) {
		const uint32_t Lw = (L + 31) >> 5;
		uint16_t* jpre = (uint16_t*)(scr + 128);
		#pragma unroll 1
		for (uint32_t w = lane; w < Lw; w += 32) scr[w] = 0;
		__syncwarp();
		#pragma unroll 1
		for (uint32_t y = lane; y < len; y += 32) {
			const uint32_t jr = ((uint32_t)rec[y] >> 16);
			atomicOr(&scr[jr >> 5], 1u << (jr & 31));
		}
		__syncwarp();
		uint32_t carry = 0;
		#pragma unroll 1
		for (uint32_t w0 = 0; w0 < Lw; w0 += 32) {
			const uint32_t w = w0 + lane;
			const uint32_t c = w < Lw ? __popc(scr[w]) : 0;
			uint32_t v = c;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(FULL, v, o); if (lane >= (uint32_t)o) v += y; }
			if (w < Lw) jpre[w] = (uint16_t)(carry + v - c);
			carry += __shfl_sync(FULL, v, 31);
		}
		__syncwarp();
		#pragma unroll 1
		for (uint32_t y = lane; y < len; y += 32) {
			const uint64_t k = product_key(P, rec[y], j0, lenH, lenV, wide);
			const uint32_t jr = (uint32_t)(k >> 48);
			const uint32_t rank = jpre[jr >> 5] + __popc(scr[jr >> 5] & ((1u << (jr & 31)) - 1u));
			hv[rank] = (uint32_t)k; ov[rank] = (uint16_t)(k >> 32);
		}
	} else {
		#pragma unroll 1
		for (uint32_t y = lane; y < len; y += 32) {
			const uint64_t k = product_key(P, rec[y], j0, lenH, lenV, wide);
			const uint32_t jr = (uint32_t)(k >> 48);
			uint32_t rank = 0;
			#pragma unroll 1
			for (uint32_t z = 0; z < len; ++z) rank += (((uint32_t)rec[z] >> 16) < jr);
			hv[rank] = (uint32_t)k; ov[rank] = (uint16_t)(k >> 32);
		}
	}
	__syncwarp();
	return __any_sync(FULL, wide);
}

// EXACT = false is the fast kernel (packed 16-bit far test); a unit in which a position is too large for
// it is appended to `redo` and done again by the EXACT = true instance (launched with redo as its list).
template <int CAP, int NT, bool EXACT>
__global__ void __launch_bounds__(NT) k_group_fold(Params P, const uint32_t* __restrict__ list, const uint32_t* __restrict__ count_ptr,
		uint32_t l1cap, uint32_t* __restrict__ redo, uint32_t* __restrict__ redo_count)
{
	const uint32_t count = *count_ptr;
	extern __shared__ __align__(128) unsigned char smem[];
	__shared__ __align__(8) uint64_t s_bar;
	__shared__ uint32_t s_tmp[34];
	__shared__ uint32_t s_nlong, s_nhuge, s_next, s_wide;
	__shared__ uint32_t s_part[3];
	__shared__ uint16_t hugelist[8];
	__shared__ uint32_t wscr[(NT / 32) * WSCR_WORDS];
	uint64_t* prodS = (uint64_t*)(smem + GF<CAP>::PROD);
	uint32_t* hvL = (uint32_t*)(smem + GF<CAP>::PROD);
	uint16_t* ovL = (uint16_t*)(smem + GF<CAP>::PROD + 4 * (size_t)CAP);
	uint64_t* sorted = (uint64_t*)(smem + GF<CAP>::REC);
	uint32_t* bits = (uint32_t*)(smem + GF<CAP>::REC);
	uint16_t* pre = (uint16_t*)(smem + GF<CAP>::REC + 4 * (size_t)CAP);
	uint16_t* cnt = (uint16_t*)(smem + GF<CAP>::CNT);
	uint32_t* rowS = (uint32_t*)(smem + GF<CAP>::ROW);
	uint16_t* longlist = (uint16_t*)(smem + GF<CAP>::LST);
	uint32_t* l1 = (uint32_t*)(smem + GF<CAP>::L1);
	uint32_t* l1pre = l1 + l1cap + 1;
	const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	const uint32_t K = P.K;
	const int BIN = (int)P.BIN;

	if (tid == 0) mbar_init(&s_bar, 1);
	__syncthreads();
	uint32_t phase = 0;

	#pragma unroll 1
	for (uint32_t it = blockIdx.x; it < count; it += gridDim.x) {
		PHASE_BEGIN();
		const uint32_t u = list[it];
		const uint32_t li = P.ucol[u], i = P.lo + li;
		const uint64_t base = P.uptr[u];
		const uint32_t Fi = P.ucount[u];
		const ColInfo ci = P.colinfo[li];
		// rows covered by this unit: (i, n) for a light column, else one 2^sh-aligned bucket of it
		uint32_t r0 = i + 1, r1 = P.n;
		if (ci.sh != 31) {
			const uint32_t bk = ((i + 1) >> ci.sh) + (u - ci.ubase);
			r0 = max(r0, bk << ci.sh);
			r1 = min(r1, (bk + 1) << ci.sh);
		}
		const uint32_t rbase = r0 & ~31u;
		const uint32_t words = (r1 - rbase + 31) >> 5;              // 32-row words the unit spans
		const bool two = words > (uint32_t)CAP;                     // too many for a direct bitmap: index the occupied words
		const uint32_t l1w = (words + 31) >> 5;
		if (Fi > (uint32_t)CAP || (two && l1w > l1cap)) { if (tid == 0) { set_err(P.err, -5); P.unnz[u] = 0; } continue; }

		// --- stage the unit's products: one bulk async copy, overlapped with clearing the tables ---
		if (tid == 0) {
			fence_proxy_async();
			bulk_load(prodS, P.raw + base, ((Fi + 1) & ~1u) * 8u, &s_bar);
		}
		if (two) { for (uint32_t s = tid; s <= l1w; s += NT) l1[s] = 0; }
		const uint32_t nclear = two ? min(Fi, (uint32_t)CAP) : words;
		#pragma unroll 1
		for (uint32_t s = tid; s < nclear; s += NT) bits[s] = 0;
		#pragma unroll 1
		for (uint32_t s = tid; s < ((Fi + 3) >> 1); s += NT) ((uint32_t*)cnt)[s] = 0;
		if (tid == 0) { s_nlong = 0; s_nhuge = 0; s_next = 0; s_wide = 0; }
		PHASE(15);
		mbar_wait(&s_bar, phase);
		phase ^= 1;
		__syncthreads();                                           // tables cleared by all threads before anyone sets a bit
		PHASE(0);

		// --- distinct rows (== estimateNNZ_Hash) and the pair index, rows ascending ---
		uint32_t nwords = words;
		if (two) {
			#pragma unroll 1
			for (uint32_t x = tid; x < Fi; x += NT) {
				const uint32_t rel = ent_row(prodS[x]) - rbase;
				atomicOr(&l1[rel >> 10], 1u << ((rel >> 5) & 31));
			}
			__syncthreads();
			nwords = block_scan_chunked<uint32_t>(l1pre, l1w, [&](uint32_t w) { return (uint32_t)__popc(l1[w]); }, s_tmp);
		}
		auto word_of = [&](uint32_t rel) {
			const uint32_t w = rel >> 5;
			return two ? l1pre[w >> 5] + __popc(l1[w >> 5] & ((1u << (w & 31)) - 1u)) : w;
		};
		#pragma unroll 1
		for (uint32_t x = tid; x < Fi; x += NT) {
			const uint32_t rel = ent_row(prodS[x]) - rbase;
			atomicOr(&bits[word_of(rel)], 1u << (rel & 31));
		}
		__syncthreads();
		PHASE(1);
		const uint32_t Z = block_scan_chunked<uint16_t>(pre, nwords, [&](uint32_t w) { return (uint32_t)__popc(bits[w]); }, s_tmp);
		PHASE(2);
		#pragma unroll 1
		for (uint32_t x = tid; x < Fi; x += NT) {
			const uint64_t r = prodS[x];
			const uint32_t row = ent_row(r), rel = row - rbase, q = word_of(rel);
			const uint32_t p = pre[q] + __popc(bits[q] & ((1u << (rel & 31)) - 1u));
			const uint32_t a = atomic_add16(cnt, p, 1u);
			rowS[p] = row;
			// h(16) | jrank(16)<<16 | strandH<<32 | pair(14)<<34 | slot(14)<<48
			prodS[x] = ((r >> 32) & 0xFFFFFFFFull) | ((uint64_t)((uint32_t)r >> 31) << 32) | ((uint64_t)p << 34) | ((uint64_t)a << 48);
		}
		__syncthreads();
		PHASE(3);
		block_scan_chunked<uint16_t>(cnt, Z, [&](uint32_t p) { return (uint32_t)cnt[p]; }, s_tmp);     // counts -> offsets, poff[Z] = Fi
		const uint16_t* poff = cnt;
		PHASE(4);
		#pragma unroll 1
		for (uint32_t x = tid; x < Fi; x += NT) {
			const uint64_t t = prodS[x];
			sorted[poff[(uint32_t)(t >> 34) & 0x3FFFu] + ((uint32_t)(t >> 48) & 0x3FFFu)] = t & 0x1FFFFFFFFull;
		}
		__syncthreads();
		PHASE(5);

		// --- fold.  Queue: first the pairs longer than SHORT_FOLD (one warp each), then tiles of 32 pairs whose
		//     short members are multiplied, ordered (position in B's column) and folded by one thread each ---
		const uint32_t j0 = P.B_colptr[i];
		const int lenV = (int)P.read_len[i];
		uint4* out = P.out + base;
		#pragma unroll 1
		for (uint32_t p = tid; p < Z; p += NT)
			if ((uint32_t)(poff[p + 1] - poff[p]) > SHORT_FOLD) longlist[atomicAdd(&s_nlong, 1u)] = (uint16_t)p;
		__syncthreads();
		const uint32_t nlong = s_nlong, nitems = nlong + ((Z + 31) >> 5);
		PHASE(6);
		const uint32_t Lcol = P.B_colptr[i + 1] - j0;
		uint32_t* scr = wscr + wid * WSCR_WORDS;
		bool wide = false;
		for (;;) {
			uint32_t q = 0;
			if (lane == 0) q = atomicAdd(&s_next, 1u);
			q = __shfl_sync(FULL, q, 0);
			if (q >= nitems) break;
#if BELLA_PHASE_CLOCKS > 1
			const long long wq0_ = clock64();
#endif
			if (q < nlong) {
				const uint32_t p = longlist[q], s0 = poff[p], len = poff[p + 1] - s0;
				if (len > 1024) { if (lane == 0) hugelist[atomicAdd(&s_nhuge, 1u)] = (uint16_t)p; continue; }
				const uint32_t row = rowS[p];
				wide |= warp_prepare_pair(P, sorted + s0, hvL + s0, ovL + s0, len, j0, Lcol, (int)P.read_len[row], lenV, scr, lane);
				PairResult R = warp_fold_pair<EXACT>(hvL + s0, ovL + s0, sorted + s0, len, K, BIN, lane);
				if (lane == 0) out[p] = pack_result(row, R);
#if BELLA_PHASE_CLOCKS > 1
				if (lane == 0) { atomicAdd(&g_phase[9], (unsigned long long)(clock64() - wq0_)); atomicAdd(&g_phase[11], (unsigned long long)len); atomicAdd(&g_phase[13], 1ull); }
#endif
			} else {
				const uint32_t p = ((q - nlong) << 5) + lane;
				if (p >= Z) continue;
				const uint32_t s0 = poff[p], len = poff[p + 1] - s0;
				if (len > SHORT_FOLD) continue;
				const uint32_t row = rowS[p];
				const int lenH = (int)P.read_len[row];
				uint64_t key[SHORT_FOLD];
#pragma unroll 1
				for (uint32_t k = 0; k < len; ++k) sorted[s0 + k] = product_key(P, sorted[s0 + k], j0, lenH, lenV, wide);   // in place (own slots)
#pragma unroll
				for (int k = 0; k < (int)SHORT_FOLD; ++k) key[k] = k < (int)len ? sorted[s0 + k] : ~0ull;
				if (len > 1) sort8(key);
				uint32_t hv[SHORT_FOLD];
				uint16_t ov[SHORT_FOLD];
#pragma unroll
				for (int k = 0; k < (int)SHORT_FOLD; ++k) { hv[k] = (uint32_t)key[k]; ov[k] = (uint16_t)(key[k] >> 32); }
				out[p] = pack_result(row, fold_short<EXACT>(hv, ov, len, K, BIN));
#if BELLA_PHASE_CLOCKS > 1
				atomicAdd(&g_phase[12], (unsigned long long)len); atomicAdd(&g_phase[14], 1ull);
				if (lane == (uint32_t)__ffs(__activemask()) - 1u) atomicAdd(&g_phase[10], (unsigned long long)(clock64() - wq0_));
#endif
			}
		}
		if (!EXACT && (wide || K > 16383u)) s_wide = 1;
		__syncthreads();
		PHASE(7);
		const uint32_t nhuge = s_nhuge;                             // at most CAP/1024 pairs: the whole CTA takes each
		#pragma unroll 1
		for (uint32_t q = 0; q < nhuge; ++q) {
			const uint32_t p = hugelist[q], s0 = poff[p], len = poff[p + 1] - s0;
			const uint32_t row = rowS[p];
			const int lenH = (int)P.read_len[row];
			#pragma unroll 1
			for (uint32_t y = tid; y < len; y += NT) {
				bool w2 = false;
				const uint64_t k = product_key(P, sorted[s0 + y], j0, lenH, lenV, w2);
				if (!EXACT && w2) s_wide = 1;
				const uint32_t jr = (uint32_t)(k >> 48);
				uint32_t rank = 0;
				#pragma unroll 1
				for (uint32_t z = s0; z < s0 + len; ++z) rank += (((uint32_t)sorted[z] >> 16) < jr);
				hvL[s0 + rank] = (uint32_t)k; ovL[s0 + rank] = (uint16_t)(k >> 32);
			}
			if (tid < 3) s_part[tid] = 0;
			__syncthreads();
			uint16_t* par = (uint16_t*)(sorted + s0);
			uint32_t* sup = (uint32_t*)(sorted + s0) + ((len + 1) >> 1);
			PairResult R = fold_cta(hvL + s0, ovL + s0, par, sup, len, K, BIN, s_part);
			if (tid == 0) out[p] = pack_result(row, R);
			__syncthreads();
		}
		if (tid == 0) {
			P.unnz[u] = Z;
			if (!EXACT && s_wide) redo[atomicAdd(redo_count, 1u)] = u;
		}
		__syncthreads();
		PHASE(8);
	}
}

// A single pair with more than UNIT_CAP products (a unit of one row): its products have distinct
// positions in B's column, so a bitmap over those positions ranks them.  One CTA per unit, the
// ordered list and the fold's scratch live in global memory (the `out` region of the unit is big enough:
// 16 bytes per product).
__global__ void __launch_bounds__(1024) k_huge_pair(Params P, const uint32_t* __restrict__ list, uint32_t count)
{
	__shared__ uint32_t jbits[2048];       // 65536 positions
	__shared__ uint32_t jpre[2050];
	__shared__ uint32_t s_tmp[34];
	__shared__ uint32_t s_part[3];
	const uint32_t tid = threadIdx.x, NT = blockDim.x;
	for (uint32_t it = blockIdx.x; it < count; it += gridDim.x) {
		const uint32_t u = list[it];
		const uint32_t li = P.ucol[u], i = P.lo + li;
		const uint64_t base = P.uptr[u];
		const uint32_t Fi = P.ucount[u];
		const uint64_t* raw = P.raw + base;
		uint32_t* hv = (uint32_t*)(P.out + base + 1);              // out[0] is the result; 16 B/product region: hv 4, ov 2, par 2, sup 4
		uint16_t* ov = (uint16_t*)(hv + ((Fi + 3) & ~3u));
		uint16_t* par = ov + ((Fi + 1) & ~1u);
		uint32_t* sup = (uint32_t*)(par + ((Fi + 1) & ~1u));
		const uint32_t row = ent_row(raw[0]);
		if (Fi > 65535u) { if (tid == 0) { set_err(P.err, -4); P.unnz[u] = 0; } continue; }
		for (uint32_t s = tid; s < 2048; s += NT) jbits[s] = 0;
		__syncthreads();
		for (uint32_t x = tid; x < Fi; x += NT) {
			const uint32_t jr = (uint32_t)(raw[x] >> 48);
			atomicOr(&jbits[jr >> 5], 1u << (jr & 31));
		}
		__syncthreads();
		block_popc_scan<uint32_t>(jbits, jpre, 2048, s_tmp);
		if (jpre[2048] != Fi) { if (tid == 0) { set_err(P.err, -5); P.unnz[u] = 0; } continue; }
		const uint32_t j0 = P.B_colptr[i];
		const int lenV = (int)P.read_len[i], lenH = (int)P.read_len[row];
		for (uint32_t x = tid; x < Fi; x += NT) {
			const uint64_t r = raw[x];
			const uint32_t jr = (uint32_t)(r >> 48), h = (uint32_t)(r >> 32) & 0xFFFFu, sH = ((uint32_t)r >> 31);
			const uint32_t rank = jpre[jr >> 5] + __popc(jbits[jr >> 5] & ((1u << (jr & 31)) - 1u));
			const uint32_t v = P.B_values[j0 + jr], sV = strand_of(P.B_strand, P.B_rowids, j0 + jr);
			hv[rank] = h | (v << 16);
			ov[rank] = (uint16_t)overlap_estimate(lenH, lenV, h, v, sH == sV, P.K);
		}
		if (tid < 3) s_part[tid] = 0;
		__threadfence_block();
		__syncthreads();
		PairResult R = fold_cta(hv, ov, par, sup, Fi, P.K, (int)P.BIN, s_part);
		__syncthreads();
		if (tid == 0) { P.out[base] = pack_result(row, R); P.unnz[u] = 1; }
		__syncthreads();
	}
}

// ================================ multi-GPU product exchange ================================
// Each GPU transposes a k-mer range and expands its products for ALL output columns into a send
// buffer ordered by column; after the all-to-all a GPU holds, for each of its columns, one segment
// per source GPU.  k_regroup moves the segments into the unit regions the group kernel expects
// (COUNT = true only counts the products per unit of the heavy columns, for the planner).

__global__ void k_mg_colinfo(uint32_t n, const uint64_t* __restrict__ sendoff, ColInfo* __restrict__ colinfo, unsigned long long* __restrict__ ucur)
{
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		colinfo[i] = ColInfo{i, 31u};
		ucur[i] = sendoff[i];
	}
}

__global__ void k_mg_sum_counts(uint32_t n, uint32_t lo, uint32_t ncols, uint32_t world, const uint32_t* __restrict__ counts_all,
		uint32_t* __restrict__ flop32, int* err)
{
	for (uint32_t li = blockIdx.x * blockDim.x + threadIdx.x; li < ncols; li += gridDim.x * blockDim.x) {
		unsigned long long f = 0;
		for (uint32_t s = 0; s < world; ++s) f += counts_all[(size_t)s * n + lo + li];
		if (f > 0xFFFFFFFFull) { set_err(err, -4); f = 0; }
		flop32[li] = (uint32_t)f;
	}
}

template <bool COUNT>
__global__ void __launch_bounds__(256) k_regroup(uint32_t n, uint32_t lo, uint32_t ncols, uint32_t world, const uint32_t* __restrict__ counts_all,
		const uint64_t* __restrict__ segoff, const uint64_t* __restrict__ recvbase, const uint64_t* __restrict__ recv,
		const ColInfo* __restrict__ colinfo, uint32_t* __restrict__ ucount, unsigned long long* __restrict__ ucur, uint64_t* __restrict__ raw,
		const Meta* meta, const int* err)
{
	if (*err != 0) return;
	if (COUNT && meta->n_heavy_cols == 0) return;
	const uint32_t lane = threadIdx.x & 31;
	const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
	const uint64_t nseg = (uint64_t)world * ncols;
	for (uint64_t sg = warp; sg < nseg; sg += nwarps) {
		const uint32_t s = (uint32_t)(sg / ncols), li = (uint32_t)(sg % ncols);
		const uint32_t cnt = counts_all[(size_t)s * n + lo + li];
		if (!cnt) continue;
		const ColInfo ci = colinfo[li];
		const uint64_t* src = recv + recvbase[s] + segoff[(size_t)s * (ncols + 1) + li];
		if (ci.sh == 31) {
			if (COUNT) continue;
			unsigned long long q = 0;
			if (lane == 0) q = atomicAdd(&ucur[ci.ubase], (unsigned long long)cnt);
			q = __shfl_sync(FULL, q, 0);
			for (uint32_t t = lane; t < cnt; t += 32) raw[q + t] = src[t];
		} else {
			const uint32_t i = lo + li;
			for (uint32_t t = lane; t < cnt; t += 32) {
				const uint64_t r = src[t];
				const uint32_t u = unit_of(ci, i, ent_row(r));
				if (COUNT) atomicAdd(&ucount[u], 1u);
				else raw[atomicAdd(&ucur[u], 1ull)] = r;
			}
		}
	}
}

// ---- multi-GPU "route" mode: instead of all-gathering B, every GPU sends each of its nonzeros to the GPU that
// transposes that k-mer range (12-byte records {k-mer id | strand<<31, read id, pos | jrank<<16}) ----

// One CTA per contiguous range of this GPU's reads; warp per read, lanes with the same destination vote together and
// the CTA keeps its per-destination counters in shared memory, so the few global counters (one per destination GPU)
// see one atomic per CTA instead of one per warp.
__device__ __forceinline__ void route_count_cta(uint32_t i0, uint32_t i1, const uint32_t* __restrict__ colptr, const uint32_t* __restrict__ rowids,
		uint32_t kpr, uint32_t world, uint32_t* s_cnt)
{
	const uint32_t w = threadIdx.x >> 5, nw = blockDim.x >> 5, lane = threadIdx.x & 31;
	for (uint32_t i = i0 + w; i < i1; i += nw) {
		const uint32_t j0 = colptr[i], j1 = colptr[i + 1];
		for (uint32_t jb = j0; jb < j1; jb += 32) {
			const uint32_t j = jb + lane;
			const bool have = j < j1;
			const uint32_t dest = have ? min((rowids[j] & 0x7FFFFFFFu) / kpr, world - 1) : 0xFFFFFFFFu;
			const uint32_t act = __ballot_sync(FULL, have);
			if (have) {
				const uint32_t same = __match_any_sync(act, dest);
				if ((uint32_t)(__ffs(same) - 1) == lane) atomicAdd(&s_cnt[dest], (uint32_t)__popc(same));
			}
		}
	}
}

__global__ void __launch_bounds__(256) k_route_count(uint32_t n_local, const uint32_t* __restrict__ colptr, const uint32_t* __restrict__ rowids,
		uint32_t kpr, uint32_t world, unsigned long long* __restrict__ counts)
{
	__shared__ uint32_t s_cnt[64];
	if (threadIdx.x < 64) s_cnt[threadIdx.x] = 0;
	__syncthreads();
	const uint32_t i0 = (uint32_t)((uint64_t)n_local * blockIdx.x / gridDim.x), i1 = (uint32_t)((uint64_t)n_local * (blockIdx.x + 1) / gridDim.x);
	route_count_cta(i0, i1, colptr, rowids, kpr, world, s_cnt);
	__syncthreads();
	if (threadIdx.x < world && s_cnt[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
}

__global__ void __launch_bounds__(256) k_route_fill(uint32_t n_local, uint32_t read_base, const uint32_t* __restrict__ colptr,
		const uint32_t* __restrict__ rowids, const uint16_t* __restrict__ values, uint32_t kpr, uint32_t world,
		unsigned long long* __restrict__ cursor, uint32_t* __restrict__ send)
{
	__shared__ uint32_t s_cnt[64];
	__shared__ unsigned long long s_base[64];
	if (threadIdx.x < 64) s_cnt[threadIdx.x] = 0;
	__syncthreads();
	const uint32_t i0 = (uint32_t)((uint64_t)n_local * blockIdx.x / gridDim.x), i1 = (uint32_t)((uint64_t)n_local * (blockIdx.x + 1) / gridDim.x);
	route_count_cta(i0, i1, colptr, rowids, kpr, world, s_cnt);           // how much this CTA sends to every destination
	__syncthreads();
	if (threadIdx.x < world) {
		s_base[threadIdx.x] = s_cnt[threadIdx.x] ? atomicAdd(&cursor[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]) : 0ull;
		s_cnt[threadIdx.x] = 0;                                             // becomes the CTA's cursor inside its reservation
	}
	__syncthreads();
	const uint32_t w = threadIdx.x >> 5, nw = blockDim.x >> 5, lane = threadIdx.x & 31;
	for (uint32_t i = i0 + w; i < i1; i += nw) {
		const uint32_t j0 = colptr[i], j1 = colptr[i + 1];
		for (uint32_t jb = j0; jb < j1; jb += 32) {
			const uint32_t j = jb + lane;
			const bool have = j < j1;
			const uint32_t c = have ? rowids[j] : 0;
			const uint32_t dest = have ? min((c & 0x7FFFFFFFu) / kpr, world - 1) : 0xFFFFFFFFu;
			const uint32_t act = __ballot_sync(FULL, have);
			if (have) {
				const uint32_t same = __match_any_sync(act, dest);
				const uint32_t leader = __ffs(same) - 1;
				uint32_t off = 0;
				if (leader == lane) off = atomicAdd(&s_cnt[dest], (uint32_t)__popc(same));
				off = __shfl_sync(same, off, leader);
				const unsigned long long q = s_base[dest] + off + __popc(same & ((1u << lane) - 1u));
				send[3 * q + 0] = c;
				send[3 * q + 1] = read_base + i;
				send[3 * q + 2] = (uint32_t)values[j] | ((j - j0) << 16);
			}
		}
	}
}

// received records -> k-mer buckets (the same fixed-capacity layout k_partition fills)
__global__ void __launch_bounds__(256) k_partition_rec(uint64_t nrec, const uint32_t* __restrict__ rec, uint32_t klo, uint32_t khi, uint32_t W,
		uint32_t* __restrict__ bcnt, uint4* __restrict__ part, int* err)
{
	for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < nrec; t += (uint64_t)gridDim.x * blockDim.x) {
		const uint32_t c = rec[3 * t], kid = c & 0x7FFFFFFFu, rd = rec[3 * t + 1], pj = rec[3 * t + 2];
		if (kid < klo || kid >= khi) { set_err(err, -5); continue; }
		const uint32_t b = (kid - klo) / W;
		const uint32_t q = atomicAdd(&bcnt[b], 1u);
		if (q >= BUCKET_CAP) { set_err(err, -6); continue; }
		const uint64_t e = (uint64_t)rd | ((uint64_t)(c >> 31) << 31) | ((uint64_t)(pj & 0xFFFFu) << 32) | ((uint64_t)(pj >> 16) << 48);
		part[(size_t)b * BUCKET_CAP + q] = make_uint4(kid, 0u, (uint32_t)e, (uint32_t)(e >> 32));
	}
}

// ================================ matrix construction (tuples -> B) =========================
// The reference builds B = CSC(tuples (k-mer id, read id, position), ..., keep-p1, needsort = false)
// (src/main.cpp:476-480): a stable counting sort of the tuples by read (src/CSC.cpp:432-475), then
// MergeDuplicates per column (src/CSC.cpp:301-420): table of ht = pow2 >= max(16, column nnz) slots,
// slot = (key * 107) & (ht - 1) in 32-bit arithmetic, linear probing, tuples inserted in order, a repeated
// k-mer keeps the LAST position, and the column is the table read out in slot order.  That order is the
// fold order of the SpGEMM, so it is reproduced exactly: one warp per read, lane 0 replays the insertions
// in shared memory (the slot of a key depends on every earlier insertion; there is nothing to parallelise
// inside a read), all lanes compact.  BELLA emits the tuples of a read contiguously and in position order
// (src/main.cpp:393-416), which is what the stable sort preserves; the runs are used in place.

// run boundaries: rs[read] = first tuple of the read's run, re[read] = one past its last, nruns[read] counts runs
__global__ void k_tuple_runs(uint64_t T, const uint32_t* __restrict__ t_read, uint32_t n, uint32_t* __restrict__ rs, uint32_t* __restrict__ re,
		uint32_t* __restrict__ nruns, int* err)
{
	for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < T; t += (uint64_t)gridDim.x * blockDim.x) {
		const uint32_t r = t_read[t];
		if (r >= n) { set_err(err, -1); continue; }
		if (t == 0 || t_read[t - 1] != r) { rs[r] = (uint32_t)t; atomicAdd(&nruns[r], 1u); }
		if (t + 1 == T || t_read[t + 1] != r) re[r] = (uint32_t)(t + 1);
	}
}

__global__ void k_tuple_counts(uint32_t n, const uint32_t* __restrict__ rs, const uint32_t* __restrict__ re, const uint32_t* __restrict__ nruns,
		uint32_t* __restrict__ cnt, int* err)
{
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const uint32_t k = nruns[i];
		if (k > 1) set_err(err, -8);                               // the tuples of a read are not contiguous
		const uint32_t c = k ? re[i] - rs[i] : 0;
		cnt[i] = c;
		if (c > 65536u) set_err(err, -4);
		if (c > 2048u) atomicMax(err + 1, (int)c);                  // err[1]: the longest read, when one exceeds the shared-memory table
	}
}

// HT = slots of the shared-memory table per warp (reads with more tuples use the global slab: slab != nullptr)
template <int HT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_merge_duplicates(uint32_t n, const uint32_t* __restrict__ rs, const uint32_t* __restrict__ cnt,
		const uint32_t* __restrict__ cp, const uint32_t* __restrict__ t_kmer, const uint16_t* __restrict__ t_pos, const uint8_t* __restrict__ t_strand,
		uint32_t* __restrict__ tmpK, uint16_t* __restrict__ tmpV, uint32_t* __restrict__ merged, bool big, uint32_t* __restrict__ slab, uint32_t slab_ht)
{
	extern __shared__ __align__(16) uint32_t msm[];
	const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	const uint32_t gw = blockIdx.x * WARPS + w, nw = gridDim.x * WARPS;
	uint32_t* keys = big ? slab + (size_t)gw * 2 * slab_ht : msm + (size_t)w * 2 * HT;
	uint32_t* vals = keys + (big ? slab_ht : (uint32_t)HT);
	for (uint32_t i = gw; i < n; i += nw) {
		const uint32_t c = cnt[i];
		uint32_t ht = 16;
		while (ht < c) ht <<= 1;
		if ((ht > (uint32_t)HT) != big) continue;                   // the other launch takes this read
		if (c == 0) { if (lane == 0) merged[i] = 0; continue; }
		for (uint32_t s = lane; s < ht; s += 32) keys[s] = 0xFFFFFFFFu;
		__syncwarp();
		const uint32_t t0 = rs[i], mask = ht - 1;
		// 32 tuples at a time.  Every lane probes read-only for its slot (the first empty slot at or after its home,
		// or the slot already holding its key); slots found by one lane only cannot lie on another lane's probe
		// path (a path crosses occupied slots only), so the lanes before the first one that shares its slot with an
		// earlier lane commit together and give exactly the sequential result; the rest probe again.
		for (uint32_t b = 0; b < c; b += 32) {
			const uint32_t t = t0 + b + lane;
			const bool have = b + lane < c;
			uint32_t key = 0, val = 0;
			if (have) { key = t_kmer[t]; val = (uint32_t)t_pos[t] | (getbit(t_strand, t) << 16); }
			uint32_t start = 0;
			const uint32_t m = min(32u, c - b);
			while (start < m) {
				const bool act = lane >= start && lane < m;
				uint32_t h = (key * 107u) & mask;
				if (act) {
					for (;;) {
						const uint32_t cur = keys[h];
						if (cur == key || cur == 0xFFFFFFFFu) break;
						h = (h + 1) & mask;
					}
				}
				const uint32_t actmask = __ballot_sync(FULL, act);
				uint32_t same = act ? __match_any_sync(actmask, h) : 0u;
				const bool loser = act && (uint32_t)(__ffs(same) - 1) != lane;
				const uint32_t losers = __ballot_sync(FULL, loser);
				const uint32_t stop = losers ? (uint32_t)(__ffs(losers) - 1) : m;
				if (act && lane < stop) { keys[h] = key; vals[h] = val; }          // a repeated k-mer keeps the LAST position (addop returns the new value)
				__syncwarp();
				start = stop;
			}
		}
		if (big) __threadfence_block();
		__syncwarp();
		// compaction in slot order into the read's pre-merge region
		uint32_t outp = cp[i], total = 0;
		for (uint32_t s0 = 0; s0 < ht; s0 += 32) {
			const uint32_t key = keys[s0 + lane];
			const uint32_t bal = __ballot_sync(FULL, key != 0xFFFFFFFFu);
			if (key != 0xFFFFFFFFu) {
				const uint32_t o = outp + total + __popc(bal & ((1u << lane) - 1u));
				const uint32_t v = vals[s0 + lane];
				tmpK[o] = key | ((v >> 16) << 31);                        // strand bit rides in bit 31 (panel format)
				tmpV[o] = (uint16_t)v;
			}
			total += __popc(bal);
		}
		if (lane == 0) merged[i] = total;
		__syncwarp();
	}
}

// warp per read: pre-merge region -> final CSC position
__global__ void __launch_bounds__(256) k_compact_B(uint32_t n, const uint32_t* __restrict__ cp, const uint32_t* __restrict__ Bcolptr,
		const uint32_t* __restrict__ tmpK, const uint16_t* __restrict__ tmpV, uint32_t* __restrict__ Brow, uint16_t* __restrict__ Bval)
{
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
	for (uint32_t i = warp; i < n; i += nwarps) {
		const uint32_t src = cp[i], dst = Bcolptr[i], c = Bcolptr[i + 1] - dst;
		for (uint32_t x = lane; x < c; x += 32) { Brow[dst + x] = tmpK[src + x]; Bval[dst + x] = tmpV[src + x]; }
	}
}

// ================================ output ====================================================

__global__ void k_colptr(uint32_t ncols, const uint32_t* __restrict__ ubase, const uint32_t* __restrict__ uoff, uint32_t* __restrict__ colptrC)
{
	for (uint32_t li = blockIdx.x * blockDim.x + threadIdx.x; li <= ncols; li += gridDim.x * blockDim.x) colptrC[li] = uoff[ubase[li]];
}

// one warp per unit: per-unit pair records -> C (SoA), at the unit's final offset
__global__ void __launch_bounds__(256) k_compact(uint32_t U, const uint64_t* __restrict__ uptr, const uint32_t* __restrict__ uoff,
		const uint4* __restrict__ out, uint32_t* __restrict__ rowsC, uint16_t* __restrict__ countC, uint16_t* __restrict__ posH,
		uint16_t* __restrict__ posV, uint16_t* __restrict__ aux)
{
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
	for (uint32_t u = warp; u < U; u += nwarps) {
		const uint32_t g0 = uoff[u], Z = uoff[u + 1] - g0;
		const uint4* src = out + uptr[u];
		for (uint32_t p = lane; p < Z; p += 32) {
			const uint4 r = src[p];
			const size_t g = (size_t)g0 + p;
			rowsC[g] = r.x;
			countC[g] = (uint16_t)(r.y & 0xFFFFu);
			posH[g] = (uint16_t)(r.z & 0xFFFFu);
			posV[g] = (uint16_t)(r.z >> 16);
			aux[3 * g + 0] = (uint16_t)(r.y >> 16);
			aux[3 * g + 1] = (uint16_t)(r.w & 0xFFFFu);
			aux[3 * g + 2] = (uint16_t)(r.w >> 16);
		}
	}
}

} // namespace bk
