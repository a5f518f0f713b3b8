// kernels.cuh -- device code of the B200 overlap SpGEMM (included once, by bella_b200.cu).
//
// The whole path is a chain of "partition into contiguous regions, then finish each region in shared memory" stages.  A
// scattered L2 transaction (a cursor atomic with return, a partial-sector store) costs about 12 ps on this chip whatever it
// is, so wherever the fan-out allows a CTA stages a tile in shared memory, ordered by destination, and global
// memory sees one cursor atomic per (tile, destination) and coalesced runs:
//
//   transpose  k_rp1 / k_rp2                  B's nonzeros (read-major) -> coarse buckets of 2^shift1 k-mers -> fine buckets of
//                                            2^wshift k-mers (two-level CTA-staged radix partition)
//              k_bucket                      one CTA per fine bucket (TMA 1-D bulk load): counting sort by k-mer, every entry
//                                            ranked by read id inside its column, written as packed entries `Aent` (+ `Ainfo`,
//                                            A's colptr) and the per-output-column product count == estimateFLOP
//                                            (overlap.hpp:157-202)
//   plan       k_plan / k_units_init / k_count_units / k_classify_units
//                                            output columns -> units (a column, or a row range of a heavy column) of at most
//                                            UNIT_CAP products; column ranges of the scatter | group pipeline
//   scatter    k_scatter                     outer-product expansion on the A side: thread per entry; in its k-mer column
//                                            (r0<r1<..) entry a emits the kept products (col r_a, row r_b), a<b, into the
//                                            unit's region (one 8-byte record per product)
//   group+fold k_group_fold<CAP,NT,EXACT>    one CTA per unit: region staged into shared memory with one bulk async copy (TMA
//                                            1-D), distinct rows through a bitmap (two levels when the unit spans more rows
//                                            than CAP words cover; == estimateNNZ_Hash, overlap.hpp:205-276), products grouped
//                                            by pair, multiplied (overlapop), put in B-column order (== LocalSpGEMM's visiting
//                                            order, overlap.hpp:306-341) and folded (chain.hpp:74-150, choose():
//                                            common.h:162-170) without leaving shared memory: short pairs one thread per
//                                            product, longer pairs one warp per pair
//              k_huge_pair                   a single pair with more products than fit in shared memory
//   output     k_uoff_range / k_colptr / k_compact   per-unit results -> C in CSC order, rows ascending
//   multi-GPU  k_rp1 (route) / k_rp_post / k_rp1b / k_mg_post / k_mg_plan / k_mg_push / k_regroup   stores into the owner's
//                                            (peer-mapped) memory over NVLink, exchange planned on the device
//   build      k_tuple_runs / k_tuple_counts / k_merge_duplicates / k_compact_B   tuples -> B in the reference's
//                                            MergeDuplicates order (src/CSC.cpp:301-479; "next" row f2)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bk {

constexpr uint32_t NONE16 = 0xFFFFu;
constexpr uint32_t FULL = 0xFFFFFFFFu;
constexpr int NCLASS = 3;
constexpr int NRANGE = 4;                  // column ranges of the scatter | group + fold pipeline
constexpr uint32_t CLASS_CAP[NCLASS] = {2048, 4096, 8192};
constexpr uint32_t UNIT_CAP = 8192;        // products per unit (largest shared-memory class)
constexpr uint32_t BUCKET_CAP = 4096;      // entries per transpose bucket
constexpr uint32_t BUCKET_WMAX = 1024;     // k-mers per transpose bucket (a power of two)
constexpr uint32_t MAX_SPAN_SHIFT = 22;    // a unit covers at most 2^22 rows (two-level bitmap: 4097 words)
constexpr int GF_THREADS = 256;

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
	uint32_t ep_max;              // pairs up to this many products are ranked and folded one thread per product, longer ones one warp per pair
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
	unsigned int class_count[NRANGE][NCLASS + 1];   // per column range (the scatter | group pipeline) and capacity class (+ overflow)
	unsigned int range_col[NRANGE + 1];      // first local column of every range (ncols when the range is empty)
	unsigned int range_unit[NRANGE + 1];     // first unit of every range (n_units when the range is empty)
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

// B's nonzeros (read-major) -> buckets of 2^wshift consecutive k-mer ids (a fixed-capacity region of BUCKET_CAP records
// per bucket: the packed entry in partE, the k-mer id inside the bucket in partK), as a two-level radix partition.  A
// direct partition costs one cursor atomic with return and one scattered store per nonzero, and the chip sustains only
// about 83 G such scattered transactions per second (profiles/README.md); here a CTA stages a tile of nonzeros in shared
// memory, ordered by bucket, so that global memory sees one cursor atomic per (tile, bucket) and coalesced runs:
//   k_rp1   level 1: reads -> nb1 coarse buckets of 2^shift1 k-mers (records: entry 8 B + k-mer id 4 B)
//   k_rp2   level 2: one coarse bucket at a time -> its 2^(shift1 - wshift) fine buckets (entry 8 B + id inside the bucket 2 B)
// Every global cursor has its own 32-byte sector (BCNT_STRIDE words).
constexpr uint32_t BCNT_STRIDE = 8;
constexpr uint32_t CCUR_STRIDE = 4;        // the scatter's per-column cursors: 8-byte words, one per 32-byte sector
constexpr unsigned long long CCUR_HEAVY = 1ull << 63;
constexpr uint32_t AINFO_ESC = 255;        // Ainfo: entries behind this one in its column of A; 255 = that many or more (the count then comes from A's colptr)
#ifndef BELLA_RP_THREADS
#define BELLA_RP_THREADS 512
#endif
constexpr int RP_THREADS = BELLA_RP_THREADS;
constexpr int RP_ITEMS = 8;                // nonzeros per thread and tile
constexpr int RP_TILE = RP_THREADS * RP_ITEMS;
constexpr uint32_t RP_NBMAX = 1024;        // buckets of one level
constexpr size_t RP_SMEM = (size_t)RP_TILE * 12 + ((size_t)RP_NBMAX + 2) * 12;

// The tile's histogram -> bucket starts inside the tile (off[0..nb], by warp 0 alone: nb <= 1024 bins, 32 per lane) and, at the
// same time, one global cursor atomic per non-empty bucket (every thread takes a bucket: the round trip of the atomic hides
// behind the scan).  hist is left zeroed for the next tile.  One barrier, at the end.  gstride: distance of the global cursors in words.
__device__ __forceinline__ void rp_reserve(uint32_t* hist, uint32_t* off, uint32_t* gbase, uint32_t nb, uint32_t* gcur, uint32_t gstride, uint32_t cap, int* err)
{
	const uint32_t tid = threadIdx.x, lane = tid & 31;
	uint32_t g[(RP_NBMAX + RP_THREADS - 1) / RP_THREADS], c[(RP_NBMAX + RP_THREADS - 1) / RP_THREADS];
#pragma unroll
	for (int q = 0; q < (int)((RP_NBMAX + RP_THREADS - 1) / RP_THREADS); ++q) {
		const uint32_t b = tid + q * RP_THREADS;
		c[q] = b < nb ? hist[b] : 0;
		if (c[q]) g[q] = atomicAdd(&gcur[(size_t)b * gstride], c[q]);
	}
	if (tid < 32) {
		const uint32_t per = (nb + 31) >> 5, b0 = lane * per, b1 = min(b0 + per, nb);
		uint32_t sum = 0;
		for (uint32_t b = b0; b < b1; ++b) sum += hist[b];
		uint32_t incl = sum;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(FULL, incl, o); if (lane >= (uint32_t)o) incl += y; }
		uint32_t run = incl - sum;
		for (uint32_t b = b0; b < b1; ++b) { off[b] = run; run += hist[b]; }
		if (lane == 31) off[nb] = incl;
	}
#pragma unroll
	for (int q = 0; q < (int)((RP_NBMAX + RP_THREADS - 1) / RP_THREADS); ++q) {
		const uint32_t b = tid + q * RP_THREADS;
		if (c[q]) { if (g[q] + c[q] > cap) set_err(err, -6); gbase[b] = g[q]; }
	}
	__syncthreads();
	for (uint32_t b = tid; b < nb; b += RP_THREADS) hist[b] = 0;   // the next tile's atomics come after at least one more barrier
}

// Where level 1 writes.  A coarse bucket is a region of `groups` sub-regions of cap1 records, one per writer: on one GPU there
// is one writer (groups == 1); on several, coarse bucket b belongs to the rank b / nb1_loc that transposes its k-mers, the
// regions are in that rank's (peer-mapped, NVLink) memory and every rank fills its own sub-region `me`, so the remote
// stores need no remote cursor -- the cursors stay local and are posted to the owners afterwards (k_rp_post).  Measured on
// 4 GPUs, runs of 18 records (220 coarse buckets per tile) reach a third of the NVLink rate; with route = 1 the pass only
// splits by destination rank (runs of a thousand records) and the owner runs the coarse level itself (k_rp1b).
constexpr int MG_MAXW = 8;
struct RpOut {
	uint32_t groups, me, nb1_loc, kpr;     // kpr = k-mers per rank = nb1_loc << shift1
	uint32_t route;                        // 1: the buckets of this pass are the RANKS (bucket = k-mer id / kpr): long runs for the NVLink stores
	uint64_t* E[MG_MAXW];
	uint32_t* K[MG_MAXW];
};

__global__ void __launch_bounds__(RP_THREADS, 2) k_rp1(uint32_t n, uint32_t lo, uint32_t klo, uint32_t khi, const uint32_t* __restrict__ Bcolptr,   // reads [lo, n)
		const uint32_t* __restrict__ Brow, const uint16_t* __restrict__ Bval, const uint8_t* __restrict__ Bstrand,
		uint32_t shift1, uint32_t nb1, uint32_t cap1, uint32_t* __restrict__ gcur1, const RpOut O, int* err)
{
	extern __shared__ __align__(16) unsigned char rsm[];
	uint64_t* SE = (uint64_t*)rsm;                             // [RP_TILE]
	uint32_t* SK = (uint32_t*)(SE + RP_TILE);                  // [RP_TILE]
	uint32_t* hist = SK + RP_TILE;                             // [RP_NBMAX + 2]
	uint32_t* gbase = hist + RP_NBMAX + 2;                     // [RP_NBMAX + 2]
	uint32_t* off = gbase + RP_NBMAX + 2;                      // [RP_NBMAX + 2]
	const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nwc = RP_THREADS / 32;
	const uint32_t wstride = gridDim.x * nwc;
	for (uint32_t b = tid; b <= nb1; b += RP_THREADS) hist[b] = 0;
	__syncthreads();
	uint32_t i = lo + blockIdx.x * nwc + wid;                  // this warp's current read
	uint32_t j0 = 0, j1 = 0, jb = 0;
	if (i < n) { j0 = Bcolptr[i]; j1 = Bcolptr[i + 1]; jb = j0; if (j1 - j0 > 65536u) { if (lane == 0) set_err(err, -4); jb = j1; } }
	for (;;) {
		// a warp that has finished its read moves to the next one (reads without k-mers are skipped)
		while (i < n && jb >= j1) {
			i += wstride;
			if (i < n) { j0 = Bcolptr[i]; j1 = Bcolptr[i + 1]; jb = j0; if (j1 - j0 > 65536u) { if (lane == 0) set_err(err, -4); jb = j1; } }
		}
		if (!__syncthreads_or(i < n)) break;
		uint64_t ev[RP_ITEMS];
		uint32_t kv[RP_ITEMS], sl[RP_ITEMS];
		// all the loads of the tile first (k-mer id, position, strand byte: three independent streams), then the arithmetic:
		// the kernel is bound by memory latency, so every load this warp will need is in flight at once
		uint32_t cw[RP_ITEMS];
		uint16_t pw[RP_ITEMS];
		uint8_t sw[RP_ITEMS];
#pragma unroll
		for (int u = 0; u < RP_ITEMS; ++u) {
			const uint32_t j = jb + u * 32 + lane;
			const bool have = i < n && j < j1;
			cw[u] = have ? Brow[j] : 0xFFFFFFFFu;
			pw[u] = have ? Bval[j] : (uint16_t)0;
			sw[u] = have && Bstrand ? Bstrand[j >> 3] : (uint8_t)0;
		}
#pragma unroll
		for (int u = 0; u < RP_ITEMS; ++u) {
			const uint32_t j = jb + u * 32 + lane;
			sl[u] = 0xFFFFFFFFu;
			if (i < n && j < j1) {
				const uint32_t c = cw[u];
				const uint32_t kid = Bstrand ? c : c & 0x7FFFFFFFu;
				if (kid >= klo && kid < khi) {                        // k-mers outside [klo, khi) belong to another GPU's transpose
					const uint32_t st = Bstrand ? (sw[u] >> (j & 7)) & 1u : c >> 31;
					kv[u] = kid - klo;
					ev[u] = (uint64_t)i | ((uint64_t)st << 31) | ((uint64_t)pw[u] << 32) | ((uint64_t)(j - j0) << 48);
					sl[u] = atomicAdd(&hist[O.route ? kv[u] / O.kpr : kv[u] >> shift1], 1u);
				}
			}
		}
		if (i < n) jb += 32 * RP_ITEMS;
		__syncthreads();
		rp_reserve(hist, off, gbase, nb1, gcur1, BCNT_STRIDE, cap1, err);
#pragma unroll
		for (int u = 0; u < RP_ITEMS; ++u)
			if (sl[u] != 0xFFFFFFFFu) { const uint32_t at = off[O.route ? kv[u] / O.kpr : kv[u] >> shift1] + sl[u]; SE[at] = ev[u]; SK[at] = kv[u]; }
		__syncthreads();
		const uint32_t total = off[nb1];
		for (uint32_t t = tid; t < total; t += RP_THREADS) {
			const uint32_t k = SK[t], b = O.route ? k / O.kpr : k >> shift1;
			const uint32_t g = gbase[b] + (t - off[b]);
			if (g < cap1) {
				const uint32_t d = O.route ? b : O.groups > 1 ? b / O.nb1_loc : 0u, bl = O.route ? 0u : b - d * O.nb1_loc;
				const size_t at = ((size_t)bl * O.groups + O.me) * cap1 + g;
				O.E[d][at] = SE[t]; O.K[d][at] = k - d * O.kpr;      // k-mer id relative to the owner's range
			}
		}
	}
}

// several GPUs: every rank tells the owner of a coarse bucket how many records it has put into its sub-region
struct RpPost { uint32_t* cnt[MG_MAXW]; };
__global__ void k_rp_post(uint32_t nb1, uint32_t cap1, uint32_t world, uint32_t me, uint32_t nb1_loc, const uint32_t* __restrict__ gcur1, const RpPost P)
{
	for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < nb1; b += gridDim.x * blockDim.x) {
		const uint32_t d = b / nb1_loc, bl = b - d * nb1_loc;
		P.cnt[d][((size_t)bl * world + me) * BCNT_STRIDE] = min(gcur1[(size_t)b * BCNT_STRIDE], cap1);
	}
}

// tiles of the coarse sub-buckets for level 2: tstart[sb] = first tile of sub-bucket sb = c * groups + g, tstart[nsb] = number of tiles
__global__ void __launch_bounds__(1024) k_rp_tiles(uint32_t nsb, uint32_t cap1, const uint32_t* __restrict__ gcur1, uint32_t* __restrict__ tstart)
{
	__shared__ uint32_t s_tmp[34];
	for (uint32_t c = threadIdx.x; c < nsb; c += blockDim.x) tstart[c] = (min(gcur1[(size_t)c * BCNT_STRIDE], cap1) + RP_TILE - 1) / RP_TILE;
	__syncthreads();
	block_excl_scan<uint32_t>(tstart, nsb, s_tmp);
}

// several GPUs, after the route pass: the records this rank received (one sub-bucket per source rank; k-mer ids relative to
// this rank's range) -> its coarse buckets.  Same tile loop as k_rp2; the output has level 1's format.
__global__ void __launch_bounds__(RP_THREADS) k_rp1b(uint32_t shift1, uint32_t nb1, uint32_t nsb, uint32_t cap_in, const uint32_t* __restrict__ cnt_in,
		const uint32_t* __restrict__ tstart, const uint64_t* __restrict__ Ein, const uint32_t* __restrict__ Kin,
		uint32_t cap1, uint32_t* __restrict__ gcur1, uint64_t* __restrict__ E1, uint32_t* __restrict__ K1, int* err)
{
	extern __shared__ __align__(16) unsigned char rsm[];
	uint64_t* SE = (uint64_t*)rsm;
	uint32_t* SK = (uint32_t*)(SE + RP_TILE);
	uint32_t* hist = SK + RP_TILE;
	uint32_t* gbase = hist + RP_NBMAX + 2;
	uint32_t* off = gbase + RP_NBMAX + 2;
	__shared__ uint32_t s_c;
	const uint32_t tid = threadIdx.x;
	const uint32_t ntiles = tstart[nsb];
	for (uint32_t b = tid; b <= nb1; b += RP_THREADS) hist[b] = 0;
	for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
		if (tid == 0) {
			uint32_t a = 0, b = nsb;
			while (b - a > 1) { const uint32_t c = (a + b) >> 1; if (tstart[c] <= tile) a = c; else b = c; }
			s_c = a;
		}
		__syncthreads();
		const uint32_t sb = s_c;
		const uint32_t cnt = min(cnt_in[(size_t)sb * BCNT_STRIDE], cap_in), first = (tile - tstart[sb]) * RP_TILE;
		const uint32_t len = min((uint32_t)RP_TILE, cnt - first);
		const uint64_t* e1 = Ein + (size_t)sb * cap_in + first;
		const uint32_t* k1 = Kin + (size_t)sb * cap_in + first;
		uint64_t ev[RP_ITEMS];
		uint32_t kv[RP_ITEMS], sl[RP_ITEMS];
#pragma unroll
		for (int u = 0; u < RP_ITEMS; ++u) {
			const uint32_t x = u * RP_THREADS + tid;
			if (x < len) { ev[u] = e1[x]; kv[u] = k1[x]; }
		}
#pragma unroll
		for (int u = 0; u < RP_ITEMS; ++u) {
			const uint32_t x = u * RP_THREADS + tid;
			sl[u] = 0xFFFFFFFFu;
			if (x < len) {
				if ((kv[u] >> shift1) >= nb1) { set_err(err, -5); continue; }      // a k-mer outside this rank's range
				sl[u] = atomicAdd(&hist[kv[u] >> shift1], 1u);
			}
		}
		__syncthreads();
		rp_reserve(hist, off, gbase, nb1, gcur1, BCNT_STRIDE, cap1, err);
#pragma unroll
		for (int u = 0; u < RP_ITEMS; ++u)
			if (sl[u] != 0xFFFFFFFFu) { const uint32_t at = off[kv[u] >> shift1] + sl[u]; SE[at] = ev[u]; SK[at] = kv[u]; }
		__syncthreads();
		const uint32_t total = off[nb1];
		for (uint32_t t = tid; t < total; t += RP_THREADS) {
			const uint32_t k = SK[t], b = k >> shift1;
			const uint32_t g = gbase[b] + (t - off[b]);
			if (g < cap1) { const size_t at = (size_t)b * cap1 + g; E1[at] = SE[t]; K1[at] = k; }
		}
	}
}

__global__ void __launch_bounds__(RP_THREADS) k_rp2(uint32_t shift1, uint32_t wshift, uint32_t nsb, uint32_t groups, uint32_t cap1, const uint32_t* __restrict__ gcur1,
		const uint32_t* __restrict__ tstart, uint32_t sb_lo, uint32_t sb_hi, const uint64_t* __restrict__ E1, const uint32_t* __restrict__ K1,
		uint32_t* __restrict__ bcnt, uint64_t* __restrict__ partE, uint16_t* __restrict__ partK, int* err)
{
	extern __shared__ __align__(16) unsigned char rsm[];
	uint64_t* SE = (uint64_t*)rsm;
	uint32_t* SK = (uint32_t*)(SE + RP_TILE);
	uint32_t* hist = SK + RP_TILE;
	uint32_t* gbase = hist + RP_NBMAX + 2;
	uint32_t* off = gbase + RP_NBMAX + 2;
	__shared__ uint32_t s_c;
	const uint32_t tid = threadIdx.x;
	const uint32_t nb2 = 1u << (shift1 - wshift), wmask = (1u << wshift) - 1u;
	const uint32_t ntiles = tstart[sb_hi];                         // this launch: the tiles of the sub-buckets [sb_lo, sb_hi)
	for (uint32_t b = tid; b <= nb2; b += RP_THREADS) hist[b] = 0;
	for (uint32_t tile = tstart[sb_lo] + blockIdx.x; tile < ntiles; tile += gridDim.x) {
		if (tid == 0) {
			uint32_t a = 0, b = nsb;                                // last sub-bucket sb with tstart[sb] <= tile (empty ones have equal starts: take the last)
			while (b - a > 1) { const uint32_t c = (a + b) >> 1; if (tstart[c] <= tile) a = c; else b = c; }
			s_c = a;
		}
		__syncthreads();
		const uint32_t sb = s_c, c = sb / groups;
		const uint32_t cnt = min(gcur1[(size_t)sb * BCNT_STRIDE], cap1), first = (tile - tstart[sb]) * RP_TILE;
		const uint32_t len = min((uint32_t)RP_TILE, cnt - first);
		const uint64_t* e1 = E1 + (size_t)sb * cap1 + first;
		const uint32_t* k1 = K1 + (size_t)sb * cap1 + first;
		const uint32_t fbase = c << (shift1 - wshift);              // first fine bucket of this coarse bucket
		uint64_t ev[RP_ITEMS];
		uint32_t kv[RP_ITEMS], sl[RP_ITEMS];
#pragma unroll
		for (int u = 0; u < RP_ITEMS; ++u) {                      // every load of the tile in flight before anything waits
			const uint32_t x = u * RP_THREADS + tid;
			if (x < len) { ev[u] = e1[x]; kv[u] = k1[x]; }
		}
#pragma unroll
		for (int u = 0; u < RP_ITEMS; ++u) {
			const uint32_t x = u * RP_THREADS + tid;
			sl[u] = 0xFFFFFFFFu;
			if (x < len) sl[u] = atomicAdd(&hist[(kv[u] >> wshift) - fbase], 1u);
		}
		__syncthreads();
		rp_reserve(hist, off, gbase, nb2, bcnt + (size_t)fbase * BCNT_STRIDE, BCNT_STRIDE, BUCKET_CAP, err);
#pragma unroll
		for (int u = 0; u < RP_ITEMS; ++u)
			if (sl[u] != 0xFFFFFFFFu) { const uint32_t at = off[(kv[u] >> wshift) - fbase] + sl[u]; SE[at] = ev[u]; SK[at] = kv[u]; }
		__syncthreads();
		for (uint32_t t = tid; t < len; t += RP_THREADS) {
			const uint32_t k = SK[t], f = k >> wshift, b = f - fbase;
			const uint32_t g = gbase[b] + (t - off[b]);
			if (g < BUCKET_CAP) { const size_t at = (size_t)f * BUCKET_CAP + g; partE[at] = SE[t]; partK[at] = (uint16_t)(k & wmask); }
		}
	}
}

// Where the buckets [f_lo, f_hi) start in A: boff[f] = boff[f_lo] + sizes before f (boff[f_lo] was left by the launch for the
// previous range; 0 for the first), boff[f_hi] = the end.  One CTA: the ranges are consumed one after the other, while the
// level-2 partition of the next range is still running.
__global__ void __launch_bounds__(1024) k_bucket_offsets(uint32_t f_lo, uint32_t f_hi, const uint32_t* __restrict__ bcnt, uint32_t* __restrict__ boff)
{
	__shared__ uint32_t s_w[33];
	const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	uint32_t carry = f_lo ? boff[f_lo] : 0;
	if (tid == 0 && !f_lo) boff[0] = 0;
	for (uint32_t base = f_lo; base < f_hi; base += 1024) {
		const uint32_t f = base + tid;
		const uint32_t x = f < f_hi ? min(bcnt[(size_t)f * BCNT_STRIDE], BUCKET_CAP) : 0;
		uint32_t v = x;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(FULL, v, o); if (lane >= (uint32_t)o) v += y; }
		if (lane == 31) s_w[wid] = v;
		__syncthreads();
		if (wid == 0) {
			uint32_t w = s_w[lane], ws = w;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(FULL, ws, o); if (lane >= (uint32_t)o) ws += y; }
			s_w[lane] = ws - w;
			if (lane == 31) s_w[32] = ws;
		}
		__syncthreads();
		if (f < f_hi) boff[f + 1] = carry + s_w[wid] + v;
		carry += s_w[32];
		__syncthreads();
	}
}

// One CTA per bucket: the bucket arrives in shared memory as two bulk async copies (TMA 1-D); counting sort by k-mer;
// every entry then finds its place among the 2..8 entries of its column by read id and goes out to Aent (A's columns:
// rows ascending) with the per-column product counts (== estimateFLOP, overlap.hpp:157-202) and A's colptr.
constexpr size_t BUCKET_SMEM = (size_t)BUCKET_CAP * (8 + 2 + 2) + ((size_t)BUCKET_WMAX + 2) * 4;
__global__ void __launch_bounds__(256, 4) k_bucket(uint32_t klo, uint32_t m, uint32_t lo, uint32_t hi, uint32_t wshift, uint32_t b_lo, uint32_t b_hi, uint32_t nb,
		const uint32_t* __restrict__ boff, const uint64_t* __restrict__ partE, const uint16_t* __restrict__ partK,
		uint32_t* __restrict__ Acolptr, uint64_t* __restrict__ Aent, uint8_t* __restrict__ Ainfo, uint32_t* __restrict__ flop32, const int* err)
{
	extern __shared__ __align__(128) unsigned char bsm[];
	uint64_t* E = (uint64_t*)bsm;                              // [BUCKET_CAP]
	uint16_t* KR = (uint16_t*)(E + BUCKET_CAP);                // [BUCKET_CAP] k-mer of the entry (arrival order, later grouped order)
	uint16_t* SL = KR + BUCKET_CAP;                            // [BUCKET_CAP] arrival slot inside its column
	uint32_t* off = (uint32_t*)(SL + BUCKET_CAP);              // [BUCKET_WMAX + 2]
	__shared__ uint32_t s_tmp[34];
	__shared__ __align__(8) uint64_t s_bar;
	const uint32_t tid = threadIdx.x, nt = blockDim.x;
	const uint32_t W = 1u << wshift;
	if (*err != 0) return;                                     // a bucket overflowed: the host retries with narrower buckets
	if (tid == 0) mbar_init(&s_bar, 1);
	__syncthreads();
	uint32_t phase = 0;
	for (uint32_t b = b_lo + blockIdx.x; b < b_hi; b += gridDim.x) {     // this launch: the buckets [b_lo, b_hi) of nb
		const uint32_t o0 = boff[b], size = boff[b + 1] - o0;
		const uint32_t kbase = b * W, kw = min(W, m - kbase);      // k-mer ids are klo + kbase + k; A's colptr is local to [klo, klo + m)
		PHASE_BEGIN();
		if (tid == 0 && size) {
			fence_proxy_async();
			const uint32_t be = (size * 8u + 15u) & ~15u, bk2 = (size * 2u + 15u) & ~15u;
			asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&s_bar)), "r"(be + bk2) : "memory");
			asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
				::"r"(smem_u32(E)), "l"(partE + (size_t)b * BUCKET_CAP), "r"(be), "r"(smem_u32(&s_bar)) : "memory");
			asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
				::"r"(smem_u32(KR)), "l"(partK + (size_t)b * BUCKET_CAP), "r"(bk2), "r"(smem_u32(&s_bar)) : "memory");
		}
		for (uint32_t k = tid; k <= kw; k += nt) off[k] = 0;
		if (size) { mbar_wait(&s_bar, phase); phase ^= 1; }
		__syncthreads();
		for (uint32_t x = tid; x < size; x += nt) SL[x] = (uint16_t)atomicAdd(&off[KR[x]], 1u);
		__syncthreads();
		PHASE(16);
		block_excl_scan<uint32_t>(off, kw, s_tmp);
		PHASE(17);
		// group by k-mer in place through registers: every thread first reads its elements, then all write
		uint64_t ev[BUCKET_CAP / 256];
		uint32_t pk[BUCKET_CAP / 256];
#pragma unroll
		for (int q = 0; q < (int)(BUCKET_CAP / 256); ++q) {
			const uint32_t x = tid + q * 256;
			if (x < size) { const uint32_t k = KR[x]; ev[q] = E[x]; pk[q] = (off[k] + SL[x]) | (k << 16); }
		}
		__syncthreads();
#pragma unroll
		for (int q = 0; q < (int)(BUCKET_CAP / 256); ++q) {
			const uint32_t x = tid + q * 256;
			if (x < size) { E[pk[q] & 0xFFFFu] = ev[q]; KR[pk[q] & 0xFFFFu] = (uint16_t)(pk[q] >> 16); }
		}
		__syncthreads();
		PHASE(18);
		// every entry: rank by read id among its column, straight to its final place (the stores of a column's entries fall
		// into the same one or two sectors, so the warp's store is as coalesced as an ordered one)
		for (uint32_t y = tid; y < size; y += nt) {
			const uint32_t k = KR[y], s = off[k], e = off[k + 1];
			const uint64_t x = E[y];
			const uint32_t r = ent_row(x);
			uint32_t rho = 0;
			for (uint32_t z = s; z < e; ++z) rho += ent_row(E[z]) < r;
			Aent[o0 + s + rho] = x;
			const uint32_t after = e - s - 1 - rho;                 // products this entry's read collects as the column read
			Ainfo[o0 + s + rho] = (uint8_t)min(after, AINFO_ESC);
			if (after && r >= lo && r < hi) atomicAdd(&flop32[r - lo], after);
		}
		for (uint32_t k = tid; k < kw; k += nt) Acolptr[kbase + k] = o0 + off[k];
		if (b == nb - 1 && tid == 0) Acolptr[m] = o0 + size;
		__syncthreads();
		PHASE(19);
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
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		meta->n_units = U;
		for (int r = 0; r <= NRANGE; ++r) { meta->range_col[r] = ncols; meta->range_unit[r] = U; }
	}
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
		unsigned long long* __restrict__ ccur, uint32_t* __restrict__ lists, uint8_t* __restrict__ refine, uint32_t round, uint32_t nrange, Meta* meta, const int* err)
{
	if (*err != 0) return;
	const uint32_t U = meta->n_units;
	const unsigned long long total = max((unsigned long long)uptr[U], 1ull);
	for (uint32_t u = blockIdx.x * blockDim.x + threadIdx.x; u < U; u += gridDim.x * blockDim.x) {
		const uint32_t f = ucount[u];
		ucur[u] = uptr[u];
		// column ranges of about equal product counts (units are ordered by column, so a range is an interval of columns);
		// a column belongs to the range of its FIRST unit with all its units: the scatter of a range covers whole columns
		const uint32_t rg = nrange > 1 ? min(nrange - 1, (uint32_t)((unsigned long long)uptr[colinfo[ucol[u]].ubase] * nrange / total)) : 0u;
		{
			const uint32_t li = ucol[u];                             // the scatter's per-column cursor: the region of a light column, or the heavy mark
			ccur[(size_t)li * CCUR_STRIDE] = colinfo[li].sh == 31 ? (unsigned long long)uptr[u] : CCUR_HEAVY;
			atomicMin(&meta->range_col[rg], li);
			atomicMin(&meta->range_unit[rg], u);
		}
		if (!f) continue;
		int c = f <= CLASS_CAP[0] ? 0 : f <= CLASS_CAP[1] ? 1 : f <= CLASS_CAP[2] ? 2 : 3;
		if (c == 3) {
			const uint32_t li = ucol[u];
			if (colinfo[li].sh != 0) { refine[li] = (uint8_t)(round + 1); atomicAdd(&meta->n_refine, 1u); continue; }
		}
		uint32_t idx = atomicAdd(&meta->class_count[rg][c], 1u);
		lists[((size_t)rg * (NCLASS + 1) + c) * ucap + idx] = u;
	}
}

// ================================ scatter ===================================================
// One thread per entry of A.  In its k-mer column (r_0 < r_1 < ...) entry a owns the run of products (col r_a, row r_b),
// b > a: `Ainfo` (written by k_bucket) says how many entries follow it in the column, so the thread needs nothing but its
// own entry, one cursor atomic on its output column and the entries behind it.  ccur: one cursor per output column in its
// own 32-byte sector (CCUR_STRIDE words of 8 bytes); bit 63 marks a heavy column, whose products go to row-range units
// one by one through `ucur`.

__device__ __forceinline__ uint32_t entries_after(uint32_t x, uint32_t info, const uint32_t* __restrict__ Acolptr, uint32_t m)
{
	if (info < AINFO_ESC) return info;
	uint32_t a = 0, b = m;                                     // last column c with Acolptr[c] <= x
	while (b - a > 1) { const uint32_t c = (a + b) >> 1; if (Acolptr[c] <= x) a = c; else b = c; }
	return Acolptr[a + 1] - 1 - x;
}

__global__ void __launch_bounds__(256) k_scatter(const uint32_t* __restrict__ nnzA_ptr, uint32_t m, uint32_t lo, uint32_t hi, const uint32_t* __restrict__ Acolptr,
		const uint64_t* __restrict__ Aent, const uint8_t* __restrict__ Ainfo, unsigned long long* __restrict__ ccur,
		const ColInfo* __restrict__ colinfo, unsigned long long* __restrict__ ucur, uint64_t* __restrict__ raw,
		const Meta* __restrict__ meta, uint32_t range, uint32_t nrange, const int* __restrict__ stop)
{
	if (stop && *stop) return;                                  // multi-GPU: the exchange plan found a buffer too small
	constexpr int ILP = 4;
	constexpr uint64_t LOW48 = 0x0000FFFFFFFFFFFFull;
	const uint32_t stride = gridDim.x * blockDim.x;
	const uint32_t base = lo;                                  // the cursors are indexed by the local column
	if (nrange > 1) {                                          // this launch: the output columns of one range of the pipeline
		uint32_t c0 = meta->range_col[range], c1 = meta->range_col[nrange];
		for (uint32_t r = range + 1; r < nrange; ++r) c1 = min(c1, meta->range_col[r]);
		if (c0 >= c1) return;
		hi = lo + c1; lo = lo + c0;
	}
	const uint32_t nnzA = *nnzA_ptr;                           // entries of A on this device (the end of the last transpose bucket)
	for (uint32_t x0 = blockIdx.x * blockDim.x + threadIdx.x; x0 < nnzA; x0 += stride * ILP) {
		uint32_t after[ILP];
		uint64_t ea[ILP];
		unsigned long long q[ILP];
#pragma unroll
		for (int u = 0; u < ILP; ++u) {
			const uint32_t x = x0 + u * stride;
			after[u] = 0;
			if (x < nnzA) { after[u] = Ainfo[x]; ea[u] = Aent[x]; }
		}
#pragma unroll
		for (int u = 0; u < ILP; ++u) {
			if (!after[u]) continue;
			const uint32_t ra = ent_row(ea[u]);
			if (ra < lo || ra >= hi) { after[u] = 0; continue; }
			if (after[u] == AINFO_ESC) after[u] = entries_after(x0 + u * stride, AINFO_ESC, Acolptr, m);
			q[u] = atomicAdd(&ccur[(size_t)(ra - base) * CCUR_STRIDE], (unsigned long long)after[u]);
		}
#pragma unroll
		for (int u = 0; u < ILP; ++u) {
			if (!after[u]) continue;
			const uint64_t* nxt = Aent + (x0 + u * stride) + 1;
			const uint64_t top = ea[u] & ~LOW48;
			if (!(q[u] & CCUR_HEAVY)) {
				uint64_t* dst = raw + q[u];
				#pragma unroll 1
				for (uint32_t b = 0; b < after[u]; ++b) dst[b] = (nxt[b] & LOW48) | top;
			} else {
				const uint32_t ra = ent_row(ea[u]);
				const ColInfo ci = colinfo[ra - base];
				#pragma unroll 1
				for (uint32_t b = 0; b < after[u]; ++b) {
					const uint64_t eb = nxt[b];
					raw[atomicAdd(&ucur[unit_of(ci, ra, ent_row(eb))], 1ull)] = (eb & LOW48) | top;
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

// One warp folds a pair of P products (hv/ov in fold order; hv + P0 is the pair's slice of a
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

template <int CAP>
struct GF {
	static constexpr size_t PROD = 0;                                  // u64[CAP]  raw -> packed (h, jr, pair, slot); later hv u32[CAP] + ov u16[CAP] in fold order + acc u32[CAP/2]
	static constexpr size_t REC = PROD + 8 * (size_t)CAP;              // u64[CAP]  bits u32[CAP] + pre u16[CAP+2], then the products sorted by pair
	static constexpr size_t CNT = REC + 8 * (size_t)CAP + 16;          // u16[CAP+2] per pair count -> offsets
	static constexpr size_t ROW = CNT + 2 * (size_t)CAP + 16;          // u32[CAP]  row id of the pair
	static constexpr size_t LST = ROW + 4 * (size_t)CAP;               // u16[CAP/2] pairs for the warp path: longer than EP_MAX, or several bins
	static constexpr size_t HEAD = LST + 2 * ((size_t)CAP / 2);        // u32[CAP/32+2] bit y set <=> a pair starts at sorted position y (+ sentinel at Fi)
	static constexpr size_t L1 = HEAD + 4 * ((size_t)CAP / 32 + 2);    // u32[l1cap+1] level-1 bitmap + u32[l1cap+2] prefix (two-level units only)
	static size_t bytes(uint32_t l1cap) { return L1 + 8 * ((size_t)l1cap + 2); }
};

// One warp: the products of a long pair (keys, arrival order) -> hv/ov in fold order (position in B's
// column).  The positions of one pair are distinct, so a bitmap over them ranks in O(len + L/32).
__device__ __forceinline__ void warp_rank_pair(const uint64_t* key, uint32_t* hv, uint16_t* ov, uint32_t len, uint32_t L, uint32_t* scr, uint32_t lane)
{
	if (L <= JR_BITMAP_MAX) {
		const uint32_t Lw = (L + 31) >> 5;
		uint16_t* jpre = (uint16_t*)(scr + 128);
		#pragma unroll 1
		for (uint32_t w = lane; w < Lw; w += 32) scr[w] = 0;
		__syncwarp();
		#pragma unroll 1
		for (uint32_t y = lane; y < len; y += 32) {
			const uint32_t jr = (uint32_t)(key[y] >> 48);
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
			const uint64_t k = key[y];
			const uint32_t jr = (uint32_t)(k >> 48);
			const uint32_t rank = jpre[jr >> 5] + __popc(scr[jr >> 5] & ((1u << (jr & 31)) - 1u));
			hv[rank] = (uint32_t)k; ov[rank] = (uint16_t)(k >> 32);
		}
	} else {
		#pragma unroll 1
		for (uint32_t y = lane; y < len; y += 32) {
			const uint64_t k = key[y];
			const uint32_t jr = (uint32_t)(k >> 48);
			uint32_t rank = 0;
			#pragma unroll 1
			for (uint32_t z = 0; z < len; ++z) rank += ((uint32_t)(key[z] >> 48) < jr);
			hv[rank] = (uint32_t)k; ov[rank] = (uint16_t)(k >> 32);
		}
	}
	__syncwarp();
}

constexpr uint32_t EP_LIMIT = 128;         // upper bound of Params::ep_max (the per-pair accumulator packs counts for pairs up to this size)

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
	uint16_t* todo = (uint16_t*)(smem + GF<CAP>::LST);
	uint32_t* head = (uint32_t*)(smem + GF<CAP>::HEAD);
	uint32_t* acc = (uint32_t*)(smem + GF<CAP>::PROD + 6 * (size_t)CAP);       // u32[CAP/2], indexed by (pair start) / 2
	uint32_t* l1 = (uint32_t*)(smem + GF<CAP>::L1);
	uint32_t* l1pre = l1 + l1cap + 1;
	const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	const uint32_t K = P.K, EP_MAX = P.ep_max;
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
		#pragma unroll 1
		for (uint32_t s = tid; s < (Fi >> 5) + 2; s += NT) head[s] = 0;
		if (tid == 0) { s_nlong = 0; s_nhuge = 0; s_next = 0; s_wide = 0; }
		// what the multiply needs of the column read: where its k-mers start in B, and its length
		const uint32_t j0 = P.B_colptr[i];
		const int lenV = (int)P.read_len[i];
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
		// --- placement by pair, fused with the multiply (overlapop): key = jrank<<48 | overlap<<32 | v<<16 | h ---
		bool wide = false;
		#pragma unroll 1
		for (uint32_t x = tid; x < Fi; x += NT) {
			const uint64_t t = prodS[x];
			const uint32_t p = (uint32_t)(t >> 34) & 0x3FFFu;
			const uint32_t h = (uint32_t)t & 0xFFFFu, jr = (uint32_t)t >> 16, sH = (uint32_t)(t >> 32) & 1u;
			const uint32_t jg = j0 + jr;
			const uint32_t v = P.B_values[jg], sV = strand_of(P.B_strand, P.B_rowids, jg);
			const uint32_t ov = overlap_estimate((int)P.read_len[rowS[p]], lenV, h, v, sH == sV, K);
			wide |= max(h, v) > 65535u - K;                          // positions this large need the 32-bit far test
			sorted[poff[p] + ((uint32_t)(t >> 48) & 0x3FFFu)] = ((uint64_t)jr << 48) | ((uint64_t)ov << 32) | (uint64_t)(h | (v << 16));
		}
		#pragma unroll 1
		for (uint32_t p = tid; p <= Z; p += NT) { const uint32_t y = poff[p]; atomicOr(&head[y >> 5], 1u << (y & 31)); }   // p == Z: sentinel at Fi
		if (!EXACT && (wide || K > 16383u)) s_wide = 1;
		__syncthreads();
		PHASE(5);

		// --- fold order = position in B's column.  One thread per product: its pair is [start, end) between two head
		//     bits; its rank is the number of smaller positions in the pair (pairs up to EP_MAX products) ---
		uint4* out = P.out + base;
		const uint32_t Lcol = P.B_colptr[i + 1] - j0;
		#pragma unroll 1
		for (uint32_t s = tid; s < ((Fi + 1) >> 1); s += NT) acc[s] = 0;
		auto pair_of = [&](uint32_t y, uint32_t& start, uint32_t& end) {
			uint32_t w = y >> 5;
			const uint32_t below = (2u << (y & 31)) - 1u;               // bits <= y
			uint32_t m = head[w] & below;
			while (!m) m = head[--w];                                  // bit 0 of word 0 is always set
			start = (w << 5) + 31u - (uint32_t)__clz(m);
			w = y >> 5;
			m = head[w] & ~below;
			while (!m) m = head[++w];                                  // the sentinel at Fi ends the search
			end = (w << 5) + (uint32_t)__ffs(m) - 1u;
		};
		#pragma unroll 1
		for (uint32_t y = tid; y < Fi; y += NT) {
			uint32_t start, end;
			pair_of(y, start, end);
			const uint32_t len = end - start;
			if (len > EP_MAX) continue;
			const uint64_t k = sorted[y];
			const uint32_t jr = (uint32_t)(k >> 48);
			uint32_t rank = 0;
			const uint32_t* hi = (const uint32_t*)sorted + 1;
			#pragma unroll 1
			for (uint32_t z = start; z < end; ++z) rank += (hi[2 * z] >> 16) < jr;
			hvL[start + rank] = (uint32_t)k; ovL[start + rank] = (uint16_t)(k >> 32);
		}
		__syncthreads();
		PHASE(6);

		// --- fold (chain.hpp:100-150), the one-bin case: product s is dropped at the first later product that is near it ---
		{
			const uint32_t lim = far_limit<EXACT>(K);
			#pragma unroll 1
			for (uint32_t sx = tid; sx < Fi; sx += NT) {
				uint32_t start, end;
				pair_of(sx, start, end);
				const uint32_t len = end - start;
				if (len < 2 || len > EP_MAX) continue;
				if (sx > start && abs((int)ovL[sx] - (int)ovL[sx - 1]) >= BIN) atomicOr(&acc[start >> 1], 0x80000000u);   // several bins
				const FarKey key = far_key<EXACT>(hvL[sx], K);
				uint32_t t = sx + 1;
				#pragma unroll 1
				for (; t < end; ++t) if (!is_far<EXACT>(hvL[t], key, lim)) break;
				atomicAdd(&acc[start >> 1], (t - sx - 1) + ((uint32_t)(t == end) << 20));
			}
		}
		__syncthreads();
		// --- one thread per pair: result of the one-bin pairs; the others go to the warp path ---
		#pragma unroll 1
		for (uint32_t p = tid; p < Z; p += NT) {
			const uint32_t start = poff[p], len = poff[p + 1] - start;
			PairResult R;
			if (len == 1) { R.count = 1; R.hv = hvL[start]; R.nbins = 1; R.sup = 1; R.ov = ovL[start]; }
			else {
				const uint32_t a = len <= EP_MAX ? acc[start >> 1] : 0x80000000u;
				if (a >> 31) { todo[atomicAdd(&s_nlong, 1u)] = (uint16_t)p; continue; }
				R.count = (len + (a & 0xFFFFFu)) & 0xFFFFu; R.hv = hvL[start + len - 1]; R.nbins = 1; R.sup = (a >> 20) & 0x7FFu; R.ov = ovL[start + len - 1];
			}
			out[p] = pack_result(rowS[p], R);
		}
		__syncthreads();
		PHASE(7);
		// --- warp path: one warp per remaining pair (longer than EP_MAX: ranked through a position bitmap; several bins: the forest) ---
		const uint32_t nlong = s_nlong;
		uint32_t* scr = wscr + wid * WSCR_WORDS;
		for (;;) {
			uint32_t q = 0;
			if (lane == 0) q = atomicAdd(&s_next, 1u);
			q = __shfl_sync(FULL, q, 0);
			if (q >= nlong) break;
			const uint32_t p = todo[q], s0 = poff[p], len = poff[p + 1] - s0;
			if (len > 1024) { if (lane == 0) hugelist[atomicAdd(&s_nhuge, 1u)] = (uint16_t)p; continue; }
			if (len > EP_MAX) warp_rank_pair(sorted + s0, hvL + s0, ovL + s0, len, Lcol, scr, lane);
			PairResult R = warp_fold_pair<EXACT>(hvL + s0, ovL + s0, sorted + s0, len, K, BIN, lane);
			if (lane == 0) out[p] = pack_result(rowS[p], R);
		}
		__syncthreads();
		PHASE(8);
		const uint32_t nhuge = s_nhuge;                             // at most CAP/1024 pairs: the whole CTA takes each
		#pragma unroll 1
		for (uint32_t q = 0; q < nhuge; ++q) {
			const uint32_t p = hugelist[q], s0 = poff[p], len = poff[p + 1] - s0;
			const uint32_t row = rowS[p];
			#pragma unroll 1
			for (uint32_t y = tid; y < len; y += NT) {
				const uint64_t k = sorted[s0 + y];
				const uint32_t jr = (uint32_t)(k >> 48);
				uint32_t rank = 0;
				#pragma unroll 1
				for (uint32_t z = s0; z < s0 + len; ++z) rank += ((uint32_t)(sorted[z] >> 48) < jr);
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

__global__ void k_mg_colinfo(uint32_t n, const uint64_t* __restrict__ sendoff, unsigned long long* __restrict__ ccur)
{
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) ccur[(size_t)i * CCUR_STRIDE] = sendoff[i];
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

// ---- multi-GPU over NVLink peer memory (mode "nvlink" of bella_b200/distributed.py) ----
// A 32-bit array of this rank to the same place on every rank (its per-column product counts: row `me` of counts_all;
// the lengths of its reads): coalesced remote stores.
struct MgPeers { void* p[MG_MAXW]; };
__global__ void k_mg_post(uint64_t count, const uint32_t* __restrict__ src, uint32_t world, uint64_t at, const MgPeers dst)
{
	for (uint64_t c = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; c < count; c += (uint64_t)gridDim.x * blockDim.x) {
		const uint32_t v = src[c];
		for (uint32_t d = 0; d < world; ++d) ((uint32_t*)dst.p[d])[at + c] = v;
	}
}

// The exchange plan, on the device, from the flat exclusive scan S of counts_all (u64 [world * n + 1]):
//   sendoff[c]        where this rank's products for column c start in its send buffer (its own row's prefix)
//   segoff[s][li]     per source rank s, the prefix of its counts over the columns this rank owns
//   recvbase[s]       where source s's block starts in this rank's receive buffer
//   push[d] = {first element of the send buffer that goes to rank d, count, offset in d's receive buffer}
// cuts[world+1]: the column ranges the ranks own.  One launch, grid-stride; no host involvement.
__global__ void k_mg_plan(uint32_t world, uint32_t me, uint32_t n, const uint32_t* __restrict__ cuts, const unsigned long long* __restrict__ S,
		unsigned long long cap_recv, unsigned long long cap_send, const int* __restrict__ errs_all, unsigned long long* __restrict__ sendoff,
		unsigned long long* __restrict__ segoff, unsigned long long* __restrict__ recvbase, unsigned long long* __restrict__ push, int* err)
{
	// Every rank sees the same counts, so every rank reaches the same verdict: an error any rank has posted (errs_all[s]), or a
	// send / receive buffer that is too small on ANY rank (needs in push[3 * world + 0 / 1]), stops the step everywhere.
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		unsigned long long need_recv = 0, need_send = 0;
		for (uint32_t d = 0; d < world; ++d) {
			unsigned long long r = 0;
			for (uint32_t s = 0; s < world; ++s) r += S[(size_t)s * n + cuts[d + 1]] - S[(size_t)s * n + cuts[d]];
			need_recv = max(need_recv, r);
			need_send = max(need_send, (unsigned long long)(S[(size_t)(d + 1) * n] - S[(size_t)d * n]));
		}
		push[3 * world + 0] = need_recv; push[3 * world + 1] = need_send;
		for (uint32_t s = 0; s < world; ++s) if (errs_all[s]) set_err(err, errs_all[s]);
		if (need_recv > cap_recv || need_send > cap_send) set_err(err, -9);
	}
	const uint32_t lo = cuts[me], hi = cuts[me + 1], ncols = hi - lo;
	const uint64_t t0 = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x, nt = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t c = t0; c <= n; c += nt) sendoff[c] = S[(size_t)me * n + c] - S[(size_t)me * n];
	for (uint64_t x = t0; x < (uint64_t)world * (ncols + 1); x += nt) {
		const uint32_t s = (uint32_t)(x / (ncols + 1)), li = (uint32_t)(x % (ncols + 1));
		segoff[x] = S[(size_t)s * n + lo + li] - S[(size_t)s * n + lo];
	}
	if (t0 < world) {
		const uint32_t d = (uint32_t)t0;                          // as a sender to d: what is in front of me in d's buffer
		unsigned long long before = 0;
		for (uint32_t s = 0; s < me; ++s) before += S[(size_t)s * n + cuts[d + 1]] - S[(size_t)s * n + cuts[d]];
		push[3 * d + 0] = S[(size_t)me * n + cuts[d]] - S[(size_t)me * n];
		push[3 * d + 1] = S[(size_t)me * n + cuts[d + 1]] - S[(size_t)me * n + cuts[d]];
		push[3 * d + 2] = before;
		unsigned long long mine = 0;                              // as the owner: source d's block
		for (uint32_t s = 0; s < d; ++s) mine += S[(size_t)s * n + hi] - S[(size_t)s * n + lo];
		recvbase[d] = mine;
	}
}

// send buffer -> the owners' receive buffers (peer memory): one contiguous block per destination, 16-byte stores
__global__ void __launch_bounds__(256) k_mg_push(uint32_t world, uint32_t me, const uint64_t* __restrict__ send, const unsigned long long* __restrict__ push, const MgPeers recv,
		const int* __restrict__ stop)
{
	if (*stop) return;
	for (uint32_t i = 0; i < world; ++i) {
		const uint32_t d = (me + 1 + i) % world;                    // staggered: at any time every rank receives from ONE sender (no incast)
		const unsigned long long first = push[3 * d], cnt = push[3 * d + 1], at = push[3 * d + 2];
		const uint64_t* src = send + first;
		uint64_t* dst = (uint64_t*)recv.p[d] + at;
		const uint64_t t0 = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x, nt = (uint64_t)gridDim.x * blockDim.x;
		// src and dst may differ in their alignment to 16 bytes: scalar 8-byte stores then (still full 32-byte sectors per warp)
		if ((((uintptr_t)src | (uintptr_t)dst) & 15) == 0) {
			const ulonglong2* s2 = (const ulonglong2*)src;
			ulonglong2* d2 = (ulonglong2*)dst;
			for (uint64_t x = t0; x < cnt / 2; x += nt) d2[x] = s2[x];
			if (t0 == 0 && (cnt & 1)) dst[cnt - 1] = src[cnt - 1];
		} else {
			for (uint64_t x = t0; x < cnt; x += nt) dst[x] = src[x];
		}
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

// Where the units [u0, u1) start in C: uoff[u] = uoff[u0] + pairs of the units before u (uoff[u0] was left by the launch for the
// previous range of units; 0 for the first), uoff[u1] = the end.  One CTA; used when the results of a column range are
// compacted and copied out while the later ranges still fold (bella_b200_set_output_buffers).
__global__ void __launch_bounds__(1024) k_uoff_range(uint32_t u0, uint32_t u1, const uint32_t* __restrict__ unnz, uint32_t* __restrict__ uoff)
{
	__shared__ uint32_t s_w[33];
	const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	uint32_t carry = u0 ? uoff[u0] : 0;
	if (tid == 0 && !u0) uoff[0] = 0;
	for (uint32_t base = u0; base < u1; base += 1024) {
		const uint32_t u = base + tid;
		const uint32_t x = u < u1 ? unnz[u] : 0;
		uint32_t v = x;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(FULL, v, o); if (lane >= (uint32_t)o) v += y; }
		if (lane == 31) s_w[wid] = v;
		__syncthreads();
		if (wid == 0) {
			uint32_t w = s_w[lane], ws = w;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(FULL, ws, o); if (lane >= (uint32_t)o) ws += y; }
			s_w[lane] = ws - w;
			if (lane == 31) s_w[32] = ws;
		}
		__syncthreads();
		if (u < u1) uoff[u + 1] = carry + s_w[wid] + v;
		carry += s_w[32];
		__syncthreads();
	}
}

// one warp per unit: per-unit pair records -> C (SoA), at the unit's final offset
__global__ void __launch_bounds__(256) k_compact(uint32_t u_lo, uint32_t U, const uint64_t* __restrict__ uptr, const uint32_t* __restrict__ uoff,
		const uint4* __restrict__ out, uint32_t* __restrict__ rowsC, uint16_t* __restrict__ countC, uint16_t* __restrict__ posH,
		uint16_t* __restrict__ posV, uint16_t* __restrict__ aux, unsigned long long* __restrict__ n_unpinned)
{
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
	for (uint32_t u = u_lo + warp; u < U; u += nwarps) {              // the units [u_lo, U)
		const uint32_t g0 = uoff[u], Z = uoff[u + 1] - g0;
		const uint4* src = out + uptr[u];
		for (uint32_t p = lane; p < Z; p += 32) {
			const uint4 r = src[p];
			// choose() (common.h:162-170) runs std::sort over the bins: its tie order is pinned for up to 16 bins only
			if ((r.y >> 16) > 16u) atomicAdd(n_unpinned, 1ull);
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
