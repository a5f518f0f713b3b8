// bella_b200.cu -- B200 (sm_100a) overlap-detection SpGEMM  C = A * A^T  with BELLA's binning
// semiring, and the C-ABI of include/bella_b200.h.  See DESIGN.md for the data layout and the
// kernel-by-kernel roofline.  There is no CPU fallback in this file.
//
// Reference behaviour reproduced (bit-exact), by reference file:line:
//   Transpose()         src/CSC.cpp:289-299           -> k_rp1 + k_rp2 + k_bucket (A derived from B, columns sorted by read)
//   estimateFLOP        include/overlap.hpp:157-202   -> k_bucket (per-column kept-product count)
//   estimateNNZ_Hash    include/overlap.hpp:205-276   -> k_group_fold (distinct rows > col, two-level bitmap)
//   LocalSpGEMM         include/overlap.hpp:281-363   -> k_scatter + k_group_fold (products grouped by pair
//                                                        in B-column order, then folded)
//   multiop/overlapop   include/chain.hpp:47-86       -> overlap_estimate()
//   chainop             include/chain.hpp:100-150     -> the fold phases of k_group_fold (thread per product / warp per pair / fold_cta)
//   choose()            include/common/common.h:162-170 -> the pair results of the same phases
#include "bella_b200.h"

#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <thrust/iterator/transform_iterator.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "kernels.cuh"

namespace {
using namespace bk;

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
	int device = 0, sms = 148;
	cudaStream_t stream = nullptr;
	bool own_stream = true;
	std::string err;
	// problem
	uint32_t n = 0, m = 0, lo = 0, hi = 0;
	uint64_t nnzB = 0;
	uint16_t K = 17, BIN = 500;
	uint32_t ep_max = 16;                      // tuning knob (BELLA_B200_EP_MAX), see Params::ep_max
	bool have_inputs = false, symbolic_done = false, numeric_done = false;
	// input device pointers (borrowed or pointing into the owned buffers below)
	const uint32_t *dB_colptr = nullptr, *dB_rowids = nullptr, *d_len = nullptr;
	const uint16_t* dB_values = nullptr;
	const uint8_t* dB_strand = nullptr;
	DevBuf oB_colptr, oB_rowids, oB_values, oB_strand, o_len;
	// chunked upload of host inputs, overlapped with the transpose (bella_b200_set_inputs)
	static constexpr int MAX_CHUNKS = 8;
	// results streamed to the caller's host buffers range by range (bella_b200_set_output_buffers)
	uint32_t* so_rows = nullptr; uint16_t *so_count = nullptr, *so_posH = nullptr, *so_posV = nullptr;
	uint64_t so_cap = 0;
	bool so_done = false;                      // the last symbolic pass delivered the whole result to those buffers
	float so_ms = 0;
	cudaStream_t out_stream = nullptr;
	cudaEvent_t so_main[4]{}, so_aux[4]{}, so_out[4]{}, so_t0 = nullptr, so_t1 = nullptr;
	uint32_t* so_z = nullptr;                  // pinned: the C offsets of the range boundaries
	cudaStream_t sc_stream = nullptr;          // the scatter passes of the scatter | group pipeline
	cudaEvent_t sc_t0 = nullptr, sc_t1 = nullptr;
	uint32_t nrange = 1;                       // ranges of that pipeline in the current pass (1: multi-GPU finish)
	cudaStream_t copy_stream = nullptr, aux_stream = nullptr;   // aux: the larger group classes run beside the 2048 class
	cudaEvent_t aux_fork = nullptr, aux_join = nullptr;
	static constexpr int PIPE = 4;             // ranges of the level-2 partition | bucket pipeline and of the scatter | group pipeline
	cudaEvent_t pipe_ev[PIPE]{};
	cudaEvent_t chunk_ev[MAX_CHUNKS]{}, copy_begin = nullptr, copy_end = nullptr;
	uint32_t chunk_lo[MAX_CHUNKS + 1]{};
	int n_chunks = 0;                          // > 0: reads [chunk_lo[c], chunk_lo[c+1]) are on the device once chunk_ev[c] has fired
	// transpose
	uint32_t W = 0, NB = 0;                    // k-mers per bucket, buckets
	uint32_t klo = 0, khi = 0;                 // k-mer range this handle transposes ([0, m) unless multi-GPU)
	// multi-GPU product exchange (borrowed device pointers, valid during bella_b200_mg_finish)
	const uint64_t *mg_recv = nullptr, *mg_segoff = nullptr, *mg_recvbase = nullptr;
	const uint32_t* mg_counts = nullptr;
	int mg_world = 0;
	DevBuf mg_colinfo, mg_ucur, mg_rcur, mg_rtiles;
	// matrix construction from tuples (bella_b200_set_inputs_tuples)
	DevBuf tp_kmer, tp_read, tp_pos, tp_strand, tp_rs, tp_re, tp_nruns, tp_cnt, tp_cp, tp_merged, tp_tmpK, tp_tmpV, tp_slab;
	float t_build_ms = 0;
	DevBuf boff, bcur, part, partK, Aent, Ainfo, Acolptr, flop32;
	DevBuf rp_cur, rp_tiles, rp_E, rp_K;       // level 1 of the two-level partition: cursors, tile starts, coarse buckets
	double cap_scale = 1.25;                   // head room of the coarse buckets over the average
	// plan
	uint32_t ucap = 0, U = 0, round = 0;
	DevBuf nunits, ubase, shv, refine, colinfo, ucol, ucount, uptr, ucur, ccur, lists, redo, unnz, uoff;
	// products and results
	DevBuf raw, out, colptrC, rowsC, countC, posH, posV, aux;
	DevBuf meta, errflag, cubtmp, unpinned;
	Meta hmeta{};
	uint64_t flops = 0, Z = 0;
	cudaEvent_t ev[12]{};
	float t_ms[8]{};
	float t_scans = 0;
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

constexpr int ERR_BUCKET = -6;     // internal: a transpose bucket overflowed -> retry with narrower buckets
constexpr int ERR_UCAP = -7;       // internal: more units than the arrays hold -> retry with larger arrays

inline int grid_for(uint64_t work, int threads, int cap = 148 * 16)
{
	uint64_t g = (work + threads - 1) / threads;
	if (g < 1) g = 1;
	if (g > (uint64_t)cap) g = cap;
	return (int)g;
}

struct Widen { __host__ __device__ unsigned long long operator()(uint32_t x) const { return x; } };
struct PadEven { __host__ __device__ unsigned long long operator()(uint32_t x) const { return ((unsigned long long)x + 1ull) & ~1ull; } };

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

int read_flags(bella_b200_handle* h, int* e)
{
	CK(cudaMemcpyAsync(&h->hmeta, h->meta.p, sizeof(Meta), cudaMemcpyDeviceToHost, h->stream));
	CK(cudaMemcpyAsync(e, h->errflag.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
	CK(cudaStreamSynchronize(h->stream));
	return 0;
}

int report_device_error(bella_b200_handle* h, int e)
{
	if (e == -9) return fail(h, BELLA_B200_ERR_CAPACITY, "a multi-GPU exchange buffer is too small on some rank (the needed sizes are behind the push plan)");
	if (e == ERR_BUCKET) return fail(h, BELLA_B200_ERR_CAPACITY, "a transpose bucket overflowed on some rank");
	if (e == BELLA_B200_ERR_RANGE)
		return fail(h, e, "a read has more than 65536 k-mers, or one read pair shares more than 65535 k-mers");
	return fail(h, BELLA_B200_ERR_INTERNAL, "device-side error %d", e);
}

// Geometry of the two-level partition of `ml` k-mer ids holding about `nnz` nonzeros: fine buckets of W = 2^wshift k-mers
// (about 0.7 * BUCKET_CAP entries on average), coarse buckets of 2^shift1 k-mers, at most RP_NBMAX of each per level.
struct RpGeom { uint32_t W, wshift, NB, l2, shift1, nb1; };
int rp_geometry(uint64_t ml, uint64_t nnz, uint32_t W_forced, RpGeom* g)
{
	uint32_t W = W_forced;
	if (!W) {
		double avg = (double)nnz / (double)(ml ? ml : 1);
		double w = 0.7 * BUCKET_CAP / (avg > 0.25 ? avg : 0.25);
		W = 1;
		while (W * 2 <= BUCKET_WMAX && (double)(W * 2) <= w) W *= 2;
	}
	g->W = W; g->wshift = 0;
	while ((1u << g->wshift) < W) ++g->wshift;
	g->NB = (uint32_t)((ml + W - 1) / W);
	uint32_t bits = 0;
	while ((1ull << bits) < g->NB) ++bits;
	uint32_t l2 = bits / 2, l1 = bits - l2;
	if (l1 > 10) { l1 = 10; l2 = bits - l1; }
	if (l2 > 10) return -1;
	g->l2 = l2; g->shift1 = g->wshift + l2;
	g->nb1 = (uint32_t)((ml + (1ull << g->shift1) - 1) >> g->shift1);
	return 0;
}

// transpose: B (read-major) -> Aent (k-mer-major, columns sorted by read id) for the k-mers [klo, khi)
// and the reads >= row_lo, + the product counts of the output columns [cnt_lo, cnt_hi) into flop_out
int run_transpose(bella_b200_handle* h, uint32_t row_lo, uint32_t cnt_lo, uint32_t cnt_hi, uint32_t* flop_out)
{
	const uint32_t n = h->n, ml = h->khi - h->klo;
	const uint64_t nnz = h->nnzB;
	const uint32_t ncols = cnt_hi - cnt_lo;
	ENSURE(h->Acolptr, sizeof(uint32_t) * ((size_t)ml + 2));
	ENSURE(h->Aent, sizeof(uint64_t) * (nnz + 2));
	ENSURE(h->Ainfo, nnz + 16);
	CK(cudaMemsetAsync(flop_out, 0, sizeof(uint32_t) * (size_t)ncols, h->stream));
	if (!nnz || !ml || !ncols) {
		CK(cudaMemsetAsync(h->Acolptr.p, 0, sizeof(uint32_t) * ((size_t)ml + 2), h->stream));
		ENSURE(h->boff, sizeof(uint32_t) * 2);
		CK(cudaMemsetAsync(h->boff.p, 0, sizeof(uint32_t) * 2, h->stream));   // boff[NB] is the number of entries of A: none
		h->NB = 0;
		return 0;
	}
	RpGeom geo;
	if (rp_geometry(ml, (uint64_t)((double)nnz * ml / (h->m ? h->m : 1)), h->W, &geo))
		return fail(h, BELLA_B200_ERR_RANGE, "%u k-mers are more than the two-level partition addresses", ml);
	h->W = geo.W;
	const uint32_t W = geo.W, wshift = geo.wshift, NB = geo.NB;
	h->NB = NB;
	ENSURE(h->bcur, sizeof(uint32_t) * ((size_t)NB + 2) * BCNT_STRIDE);
	ENSURE(h->boff, sizeof(uint32_t) * ((size_t)NB + 2));
	ENSURE(h->part, sizeof(uint64_t) * (size_t)NB * BUCKET_CAP + 64);
	ENSURE(h->partK, sizeof(uint16_t) * (size_t)NB * BUCKET_CAP + 64);
	CK(cudaMemsetAsync(h->bcur.p, 0, sizeof(uint32_t) * ((size_t)NB + 2) * BCNT_STRIDE, h->stream));
	uint32_t* bcnt = h->bcur.as<uint32_t>();
	uint64_t* partE = h->part.as<uint64_t>();
	uint16_t* partK = h->partK.as<uint16_t>();
	{
		const uint32_t l2 = geo.l2, shift1 = geo.shift1, nb1 = geo.nb1;
		const double share = (double)ml / (double)(h->m ? h->m : 1);           // multi-GPU: only the k-mers [klo, khi) are this handle's
		const uint32_t groups = 1;                                             // one writer per coarse bucket on one GPU
		const uint32_t nsb = nb1 * groups;
		const uint64_t cap64 = (uint64_t)((double)nnz * share / nsb * h->cap_scale) + 2 * RP_TILE;
		if (cap64 > 0x7FFFFFFFull) return fail(h, BELLA_B200_ERR_RANGE, "coarse transpose bucket too large");
		const uint32_t cap1 = (uint32_t)((cap64 + 15) & ~15ull);
		ENSURE(h->rp_cur, sizeof(uint32_t) * ((size_t)nsb + 2) * BCNT_STRIDE);
		ENSURE(h->rp_tiles, sizeof(uint32_t) * ((size_t)nsb + 2));
		ENSURE(h->rp_E, sizeof(uint64_t) * (size_t)nsb * cap1 + 64);
		ENSURE(h->rp_K, sizeof(uint32_t) * (size_t)nsb * cap1 + 64);
		CK(cudaMemsetAsync(h->rp_cur.p, 0, sizeof(uint32_t) * ((size_t)nsb + 2) * BCNT_STRIDE, h->stream));
		CK(cudaFuncSetAttribute(k_rp1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RP_SMEM));
		CK(cudaFuncSetAttribute(k_rp2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RP_SMEM));
		CK(cudaEventRecord(h->ev[8], h->stream));
		const uint32_t warps_per_cta = RP_THREADS / 32;
		auto rp1 = [&](uint32_t r0, uint32_t r1) {
			uint32_t g = (r1 - r0 + warps_per_cta - 1) / warps_per_cta;
			if (g > (uint32_t)h->sms * 2) g = (uint32_t)h->sms * 2;
			RpOut O{};
			O.groups = 1; O.me = 0; O.nb1_loc = nb1; O.kpr = 0;
			O.E[0] = h->rp_E.as<uint64_t>(); O.K[0] = h->rp_K.as<uint32_t>();
			k_rp1<<<g, RP_THREADS, RP_SMEM, h->stream>>>(r1, r0, h->klo, h->khi, h->dB_colptr, h->dB_rowids, h->dB_values, h->dB_strand,
				shift1, nb1, cap1, h->rp_cur.as<uint32_t>(), O, h->errflag.as<int>());
		};
		if (h->n_chunks > 0) {
			// host inputs are still arriving chunk by chunk: partition each range of reads as soon as it is on the device
			for (int c = 0; c < h->n_chunks; ++c) {
				const uint32_t c0 = h->chunk_lo[c] > row_lo ? h->chunk_lo[c] : row_lo, c1 = h->chunk_lo[c + 1];
				CK(cudaStreamWaitEvent(h->stream, h->chunk_ev[c], 0));
				if (c0 >= c1) continue;
				rp1(c0, c1);
				LAUNCHED();
			}
		} else if (n > row_lo) {
			rp1(row_lo, n);
			LAUNCHED();
		}
		k_rp_tiles<<<1, 1024, 0, h->stream>>>(nsb, cap1, h->rp_cur.as<uint32_t>(), h->rp_tiles.as<uint32_t>());
		LAUNCHED();
		// Level 2 and the bucket kernel run as a pipeline over a few ranges of coarse buckets: the partition is bound by
		// memory latency, the bucket kernel by instruction issue, so the bucket kernel of one range runs (on the second
		// stream) beside the level-2 partition of the next.
		const uint32_t G = (nb1 >= 8 && getenv("BELLA_B200_PIPE_BUCKET")) ? (uint32_t)bella_b200_handle::PIPE : 1u;   // measured: no gain (profiles/README.md), off
		CK(cudaFuncSetAttribute(k_bucket, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BUCKET_SMEM));
		for (uint32_t g = 0; g < G; ++g) {
			const uint32_t c0 = (uint32_t)((uint64_t)nb1 * g / G), c1 = (uint32_t)((uint64_t)nb1 * (g + 1) / G);
			uint64_t f0 = (uint64_t)c0 << l2, f1 = (uint64_t)c1 << l2;
			if (f0 > NB) f0 = NB;
			if (f1 > NB || g + 1 == G) f1 = NB;
			k_rp2<<<h->sms * 2, RP_THREADS, RP_SMEM, h->stream>>>(shift1, wshift, nsb, groups, cap1, h->rp_cur.as<uint32_t>(), h->rp_tiles.as<uint32_t>(),
				c0 * groups, c1 * groups, h->rp_E.as<uint64_t>(), h->rp_K.as<uint32_t>(), bcnt, partE, partK, h->errflag.as<int>());
			LAUNCHED();
			CK(cudaEventRecord(h->pipe_ev[g], h->stream));
			CK(cudaStreamWaitEvent(h->aux_stream, h->pipe_ev[g], 0));
			if (f1 > f0) {
				const uint32_t nbk = (uint32_t)(f1 - f0);
				k_bucket_offsets<<<1, 1024, 0, h->aux_stream>>>((uint32_t)f0, (uint32_t)f1, bcnt, h->boff.as<uint32_t>());
				LAUNCHED();
				k_bucket<<<nbk < (uint32_t)h->sms * 16 ? nbk : (uint32_t)h->sms * 16, 256, BUCKET_SMEM, h->aux_stream>>>(h->klo, ml, cnt_lo, cnt_hi, wshift, (uint32_t)f0, (uint32_t)f1, NB,
					h->boff.as<uint32_t>(), partE, partK, h->Acolptr.as<uint32_t>(), h->Aent.as<uint64_t>(), h->Ainfo.as<uint8_t>(), flop_out, h->errflag.as<int>());
				LAUNCHED();
			}
		}
		CK(cudaEventRecord(h->ev[9], h->stream));
		CK(cudaEventRecord(h->aux_join, h->aux_stream));
		CK(cudaStreamWaitEvent(h->stream, h->aux_join, 0));
		return 0;
	}
	return 0;
}

// plan: columns -> units, unit sizes, region offsets, classes.  Everything is enqueued without a host
// synchronisation; the caller reads meta + error flag afterwards and retries when asked to.
int run_plan(bella_b200_handle* h)
{
	const uint32_t ncols = h->hi - h->lo;
	if (h->ucap < ncols + 1) h->ucap = ncols + 1;
	const uint32_t ucap = h->ucap;
	ENSURE(h->nunits, sizeof(uint32_t) * ((size_t)ncols + 1));
	ENSURE(h->ubase, sizeof(uint32_t) * ((size_t)ncols + 1));
	ENSURE(h->shv, (size_t)ncols + 1);
	ENSURE(h->colinfo, sizeof(ColInfo) * ((size_t)ncols + 1));
	ENSURE(h->ucol, sizeof(uint32_t) * ((size_t)ucap + 1));
	ENSURE(h->ucount, sizeof(uint32_t) * ((size_t)ucap + 2));
	ENSURE(h->uptr, sizeof(uint64_t) * ((size_t)ucap + 2));
	ENSURE(h->ucur, sizeof(uint64_t) * ((size_t)ucap + 2));
	ENSURE(h->ccur, sizeof(uint64_t) * ((size_t)ncols + 1) * CCUR_STRIDE);
	ENSURE(h->unnz, sizeof(uint32_t) * ((size_t)ucap + 2));
	ENSURE(h->uoff, sizeof(uint32_t) * ((size_t)ucap + 2));
	ENSURE(h->lists, sizeof(uint32_t) * (size_t)NRANGE * (NCLASS + 1) * ((size_t)ucap + 1));
	CK(cudaMemsetAsync(h->meta.p, 0, sizeof(Meta), h->stream));
	CK(cudaMemsetAsync(h->nunits.p, 0, sizeof(uint32_t) * ((size_t)ncols + 1), h->stream));
	CK(cudaMemsetAsync(h->ucount.p, 0, sizeof(uint32_t) * ((size_t)ucap + 2), h->stream));
	CK(cudaMemsetAsync(h->unnz.p, 0, sizeof(uint32_t) * ((size_t)ucap + 2), h->stream));
	if (!ncols) return 0;
	k_plan<<<grid_for(ncols, 256), 256, 0, h->stream>>>(h->n, h->lo, ncols, h->flop32.as<uint32_t>(), h->refine.as<uint8_t>(),
		h->nunits.as<uint32_t>(), h->shv.as<uint8_t>(), h->meta.as<Meta>());
	LAUNCHED();
	if (int rc = exclusive_scan(h, h->nunits.as<uint32_t>(), h->ubase.as<uint32_t>(), ncols + 1)) return rc;
	k_units_init<<<grid_for(ncols, 256), 256, 0, h->stream>>>(ncols, ucap, h->flop32.as<uint32_t>(), h->ubase.as<uint32_t>(), h->shv.as<uint8_t>(),
		h->colinfo.as<ColInfo>(), h->ucol.as<uint32_t>(), h->ucount.as<uint32_t>(), h->meta.as<Meta>(), h->errflag.as<int>());
	LAUNCHED();
	if (h->mg_recv) {
		k_regroup<true><<<h->sms * 8, 256, 0, h->stream>>>(h->n, h->lo, ncols, (uint32_t)h->mg_world, h->mg_counts, h->mg_segoff, h->mg_recvbase,
			h->mg_recv, h->colinfo.as<ColInfo>(), h->ucount.as<uint32_t>(), nullptr, nullptr, h->meta.as<Meta>(), h->errflag.as<int>());
	} else {
		const uint32_t ml = h->khi - h->klo;
		k_count_units<<<grid_for(ml, 256), 256, 0, h->stream>>>(ml, h->lo, h->hi, h->Acolptr.as<uint32_t>(), h->Aent.as<uint64_t>(),
			h->colinfo.as<ColInfo>(), h->ucount.as<uint32_t>(), h->meta.as<Meta>(), h->errflag.as<int>());
	}
	LAUNCHED();
	{
		auto padded = thrust::make_transform_iterator((const uint32_t*)h->ucount.as<uint32_t>(), PadEven());
		if (int rc = exclusive_scan(h, padded, h->uptr.as<unsigned long long>(), ucap + 1)) return rc;
	}
	k_classify_units<<<grid_for(ucap, 256), 256, 0, h->stream>>>(ucap, h->ucol.as<uint32_t>(), h->ucount.as<uint32_t>(), h->colinfo.as<ColInfo>(),
		h->uptr.as<uint64_t>(), h->ucur.as<unsigned long long>(), h->ccur.as<unsigned long long>(), h->lists.as<uint32_t>(), h->refine.as<uint8_t>(), h->round,
		h->nrange, h->meta.as<Meta>(), h->errflag.as<int>());
	LAUNCHED();
	return 0;
}

Params make_params(bella_b200_handle* h)
{
	Params P{};
	P.n = h->n; P.m = h->m; P.lo = h->lo; P.hi = h->hi; P.K = h->K; P.BIN = h->BIN;
	P.ep_max = h->ep_max;
	P.B_colptr = h->dB_colptr; P.B_rowids = h->dB_rowids; P.B_values = h->dB_values; P.B_strand = h->dB_strand; P.read_len = h->d_len;
	P.A_colptr = h->Acolptr.as<uint32_t>(); P.Aent = h->Aent.as<uint64_t>();
	P.colinfo = h->colinfo.as<ColInfo>(); P.ucol = h->ucol.as<uint32_t>(); P.ucount = h->ucount.as<uint32_t>();
	P.uptr = h->uptr.as<uint64_t>(); P.ucur = h->ucur.as<unsigned long long>();
	P.raw = h->raw.as<uint64_t>(); P.out = h->out.as<uint4>(); P.unnz = h->unnz.as<uint32_t>();
	P.err = h->errflag.as<int>();
	return P;
}

template <int CAP, int NT>
int launch_group(bella_b200_handle* h, const Params& P, int rg, int cls, uint32_t count, uint32_t l1cap, int ctas_per_sm, cudaStream_t st)
{
	// fast instance over the class list of the range, then the exact instance over whatever the fast one handed back
	const uint32_t ucap = h->ucap;
	const uint32_t* list = h->lists.as<uint32_t>() + ((size_t)rg * (NCLASS + 1) + cls) * ucap;
	uint32_t* redo = h->redo.as<uint32_t>() + ((size_t)rg * NCLASS + cls) * (ucap + 1);
	uint32_t* redo_count = redo + ucap;
	const uint32_t* class_count = &h->meta.as<Meta>()->class_count[rg][cls];
	const size_t smem = GF<CAP>::bytes(l1cap);
	CK(cudaFuncSetAttribute(k_group_fold<CAP, NT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	CK(cudaFuncSetAttribute(k_group_fold<CAP, NT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	if (count) {
		const uint32_t resident = (uint32_t)h->sms * ctas_per_sm;
		uint32_t grid = count < resident ? count : resident;
		k_group_fold<CAP, NT, false><<<grid, NT, smem, st>>>(P, list, class_count, l1cap, redo, redo_count);
		LAUNCHED();
		k_group_fold<CAP, NT, true><<<h->sms, NT, smem, st>>>(P, redo, redo_count, l1cap, nullptr, nullptr);
		LAUNCHED();
	}
	return 0;
}

// plan until it converges (bucket overflow -> narrower buckets and a new transpose; more units than the
// arrays hold -> larger arrays; units that do not fit shared memory -> finer row ranges)
int plan_loop(bella_b200_handle* h, bool own_transpose)
{
	const uint32_t ncols = h->hi - h->lo;
	ENSURE(h->refine, (size_t)ncols + 1);
	ENSURE(h->colptrC, sizeof(uint32_t) * ((size_t)ncols + 1));
	CK(cudaMemsetAsync(h->refine.p, 0, (size_t)ncols + 1, h->stream));
	h->round = 0;
	bool need_transpose = own_transpose;
	for (int attempt = 0;; ++attempt) {
		if (attempt > 40) return fail(h, BELLA_B200_ERR_INTERNAL, "planning did not converge");
		if (need_transpose) {
			ENSURE(h->flop32, sizeof(uint32_t) * ((size_t)ncols + 1));
			if (int rc = run_transpose(h, h->lo, h->lo, h->hi, h->flop32.as<uint32_t>())) return rc;
		}
		if (int rc = run_plan(h)) return rc;
		int e = 0;
		if (int rc = read_flags(h, &e)) return rc;
		if (e == ERR_BUCKET && own_transpose) {
			if (h->W <= 1) return fail(h, BELLA_B200_ERR_RANGE, "a k-mer occurs in more than %u reads", BUCKET_CAP);
			h->W = h->W / 2;
			h->cap_scale *= 1.5;
			need_transpose = true;
			CK(cudaMemsetAsync(h->errflag.p, 0, sizeof(int), h->stream));
			continue;
		}
		need_transpose = false;
		if (e == ERR_UCAP) {
			h->ucap = h->hmeta.n_units + 1;
			CK(cudaMemsetAsync(h->errflag.p, 0, sizeof(int), h->stream));
			continue;
		}
		if (e) return report_device_error(h, e);
		if (h->hmeta.n_refine) { ++h->round; continue; }
		break;
	}
	h->flops = h->hmeta.flops;
	h->U = h->hmeta.n_units;
	ENSURE(h->raw, sizeof(uint64_t) * (h->flops + h->U + 2));
	ENSURE(h->out, sizeof(uint4) * (h->flops + h->U + 2));
	return 0;
}

// scatter (unless the products are already in the unit regions: multi-GPU) + group + fold, then C's colptr; ends with the one
// host synchronisation that returns nnz(C).  Single GPU: a pipeline over h->nrange ranges of output columns -- the scatter
// is bound by scattered memory transactions and the group + fold kernels by instruction issue, so the scatter of one range
// runs (on its own stream) beside the group + fold kernels of the range before.
int group_and_output(bella_b200_handle* h, bool do_scatter)
{
	const uint32_t ncols = h->hi - h->lo;
	const uint32_t U = h->U, ucap = h->ucap;
	const int R = (int)h->nrange;
	const bool stream_out = do_scatter && h->so_rows && h->flops && h->U;
	Params P = make_params(h);
	CK(cudaEventRecord(h->ev[3], h->stream));
	// level-1 bitmap words a unit can need: light columns span at most n rows, heavy units at most 2^MAX_SPAN_SHIFT
	uint32_t span = h->n < (1u << MAX_SPAN_SHIFT) ? h->n : (1u << MAX_SPAN_SHIFT);
	const uint32_t l1cap = (span + 1023 + 32) / 1024 + 1;
	const uint32_t* lists = h->lists.as<uint32_t>();
	ENSURE(h->redo, sizeof(uint32_t) * (size_t)NRANGE * NCLASS * ((size_t)ucap + 1));
	for (int c = 0; c < R * NCLASS; ++c)
		CK(cudaMemsetAsync(h->redo.as<uint32_t>() + (size_t)c * (ucap + 1) + ucap, 0, sizeof(uint32_t), h->stream));
	CK(cudaEventRecord(h->aux_fork, h->stream));
	if (do_scatter && h->flops) {
		CK(cudaStreamWaitEvent(h->sc_stream, h->aux_fork, 0));
		CK(cudaEventRecord(h->sc_t0, h->sc_stream));
		for (int r = 0; r < R; ++r) {
			k_scatter<<<grid_for((h->nnzB + 3) / 4, 256, 148 * 32), 256, 0, h->sc_stream>>>(h->boff.as<uint32_t>() + h->NB, h->m, h->lo, h->hi, P.A_colptr, P.Aent,
				h->Ainfo.as<uint8_t>(), h->ccur.as<unsigned long long>(), P.colinfo, P.ucur, P.raw, h->meta.as<Meta>(), (uint32_t)r, (uint32_t)R, nullptr);
			LAUNCHED();
			CK(cudaEventRecord(h->pipe_ev[r], h->sc_stream));
		}
		CK(cudaEventRecord(h->sc_t1, h->sc_stream));
	}
	for (int r = 0; r < R; ++r) {
		const uint32_t* cc = h->hmeta.class_count[r];
		if (do_scatter && h->flops) {
			CK(cudaStreamWaitEvent(h->stream, h->pipe_ev[r], 0));
			CK(cudaEventRecord(h->aux_fork, h->stream));
		}
		// the larger classes (few, long-running CTAs) go to a second stream; the 2048 class fills the rest of the GPU
		CK(cudaStreamWaitEvent(h->aux_stream, h->aux_fork, 0));
		if (int rc = launch_group<8192, 1024>(h, P, r, 2, cc[2], l1cap, 1, h->aux_stream)) return rc;
		if (int rc = launch_group<4096, 512>(h, P, r, 1, cc[1], l1cap, 2, h->aux_stream)) return rc;
		if (cc[3]) {
			k_huge_pair<<<cc[3] < (uint32_t)h->sms ? cc[3] : (uint32_t)h->sms, 1024, 0, h->aux_stream>>>(P, lists + ((size_t)r * (NCLASS + 1) + 3) * ucap, cc[3]);
			LAUNCHED();
		}
		if (int rc = launch_group<2048, 256>(h, P, r, 0, cc[0], l1cap, 4, h->stream)) return rc;
		if (stream_out) { CK(cudaEventRecord(h->so_main[r], h->stream)); CK(cudaEventRecord(h->so_aux[r], h->aux_stream)); }
	}
	h->so_done = false;
	if (stream_out) {
		// The caller's host buffers are known (bella_b200_set_output_buffers): a range of columns is compacted and copied out as
		// soon as its group + fold kernels are done, while the later ranges still fold.  C's arrays are sized by the product
		// count (an upper bound of nnz(C)); the offsets of the range boundaries come back through pinned memory.
		const uint64_t zmax = h->flops + 1;
		ENSURE(h->rowsC, sizeof(uint32_t) * zmax); ENSURE(h->countC, sizeof(uint16_t) * zmax); ENSURE(h->posH, sizeof(uint16_t) * zmax);
		ENSURE(h->posV, sizeof(uint16_t) * zmax); ENSURE(h->aux, sizeof(uint16_t) * 3 * zmax);
		ENSURE(h->unpinned, sizeof(unsigned long long));
		CK(cudaMemsetAsync(h->unpinned.p, 0, sizeof(unsigned long long), h->out_stream));
		uint32_t ub[NRANGE + 1];
		for (int r = 0; r <= R; ++r) ub[r] = r < R ? h->hmeta.range_unit[r] : U;
		for (int r = R - 1; r >= 0; --r) if (ub[r] > ub[r + 1]) ub[r] = ub[r + 1];      // an empty range starts where the next one does
		ub[0] = 0;
		for (int r = 0; r < R; ++r) {
			CK(cudaStreamWaitEvent(h->out_stream, h->so_main[r], 0));
			CK(cudaStreamWaitEvent(h->out_stream, h->so_aux[r], 0));
			if (ub[r + 1] > ub[r]) {
				k_uoff_range<<<1, 1024, 0, h->out_stream>>>(ub[r], ub[r + 1], h->unnz.as<uint32_t>(), h->uoff.as<uint32_t>());
				LAUNCHED();
				k_compact<<<grid_for((uint64_t)(ub[r + 1] - ub[r]) * 32, 256), 256, 0, h->out_stream>>>(ub[r], ub[r + 1], h->uptr.as<uint64_t>(), h->uoff.as<uint32_t>(),
					h->out.as<uint4>(), h->rowsC.as<uint32_t>(), h->countC.as<uint16_t>(), h->posH.as<uint16_t>(), h->posV.as<uint16_t>(), h->aux.as<uint16_t>(),
					h->unpinned.as<unsigned long long>());
				LAUNCHED();
			}
			CK(cudaMemcpyAsync(&h->so_z[r + 1], h->uoff.as<uint32_t>() + ub[r + 1], sizeof(uint32_t), cudaMemcpyDeviceToHost, h->out_stream));
			CK(cudaEventRecord(h->so_out[r], h->out_stream));
		}
		h->so_z[0] = 0;
		bool fits = true;
		for (int r = 0; r < R; ++r) {
			// the host waits for range r only (its compaction is done: the offsets are known); the copies go to the copy stream,
			// which is idle, so they run at once -- beside the group + fold kernels of the later ranges, which are all enqueued
			CK(cudaEventSynchronize(h->so_out[r]));
			if (r == 0) CK(cudaEventRecord(h->so_t0, h->copy_stream));
			const uint64_t z0 = ub[r] ? h->so_z[r] : 0, z1 = h->so_z[r + 1];
			if (z1 > h->so_cap) fits = false;
			if (fits && z1 > z0) {
				cudaStream_t cs = h->copy_stream;
				CK(cudaMemcpyAsync(h->so_rows + z0, h->rowsC.as<uint32_t>() + z0, sizeof(uint32_t) * (z1 - z0), cudaMemcpyDeviceToHost, cs));
				CK(cudaMemcpyAsync(h->so_count + z0, h->countC.as<uint16_t>() + z0, sizeof(uint16_t) * (z1 - z0), cudaMemcpyDeviceToHost, cs));
				CK(cudaMemcpyAsync(h->so_posH + z0, h->posH.as<uint16_t>() + z0, sizeof(uint16_t) * (z1 - z0), cudaMemcpyDeviceToHost, cs));
				CK(cudaMemcpyAsync(h->so_posV + z0, h->posV.as<uint16_t>() + z0, sizeof(uint16_t) * (z1 - z0), cudaMemcpyDeviceToHost, cs));
			}
		}
		CK(cudaEventRecord(h->so_t1, h->copy_stream));
		CK(cudaStreamWaitEvent(h->stream, h->so_out[R - 1], 0));     // the scans below rewrite uoff (same values) after the last compaction
		h->so_done = fits;
	}
	CK(cudaEventRecord(h->aux_join, h->aux_stream));
	CK(cudaStreamWaitEvent(h->stream, h->aux_join, 0));
	CK(cudaEventRecord(h->ev[4], h->stream));
	if (int rc = exclusive_scan(h, h->unnz.as<uint32_t>(), h->uoff.as<uint32_t>(), U + 1)) return rc;
	k_colptr<<<grid_for(ncols + 1, 256), 256, 0, h->stream>>>(ncols, h->ubase.as<uint32_t>(), h->uoff.as<uint32_t>(), h->colptrC.as<uint32_t>());
	LAUNCHED();
	uint32_t z32 = 0;
	int e = 0;
	CK(cudaMemcpyAsync(&z32, h->uoff.as<uint32_t>() + U, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
	CK(cudaMemcpyAsync(&e, h->errflag.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
	CK(cudaEventRecord(h->ev[5], h->stream));
	CK(cudaStreamSynchronize(h->stream));
	if (e) return report_device_error(h, e);
	h->Z = z32;
	CK(cudaEventElapsedTime(&h->t_ms[1], h->ev[3], h->ev[4]));    // scatter | group + fold pipeline (the scatter passes run beside the group kernels)
	h->t_ms[7] = 0;
	if (do_scatter && h->flops) CK(cudaEventElapsedTime(&h->t_ms[7], h->sc_t0, h->sc_t1));   // the scatter passes alone, first to last
	CK(cudaEventElapsedTime(&h->t_scans, h->ev[4], h->ev[5]));    // scans + colptr
	h->symbolic_done = true;
	h->numeric_done = false;
	if (stream_out) {
		CK(cudaStreamSynchronize(h->out_stream));
		CK(cudaStreamSynchronize(h->copy_stream));             // the caller's buffers are complete when the symbolic phase returns
		h->numeric_done = true;                                 // C's arrays are on the device already (and in the caller's buffers when so_done)
		CK(cudaEventElapsedTime(&h->so_ms, h->so_t0, h->so_t1));   // first streamed copy to the last, on the copy stream
	}
	return 0;
}

// the whole symbolic phase; on return C's colptr and the per-unit results are on the device
int run_symbolic(bella_b200_handle* h)
{
	ENSURE(h->meta, sizeof(Meta));
	ENSURE(h->errflag, 4 * sizeof(int));
	CK(cudaMemsetAsync(h->errflag.p, 0, sizeof(int), h->stream));
	h->klo = 0; h->khi = h->m;
	h->mg_recv = nullptr;
	h->nrange = h->nnzB >= (1u << 22) ? 2u : 1u;                      // two ranges measured best (profiles/README.md); small inputs: one
	if (h->so_rows && h->nnzB >= (1u << 22)) h->nrange = (uint32_t)NRANGE;      // streamed output: more, smaller ranges leave less of the copy exposed
	if (const char* e = getenv("BELLA_B200_NRANGE")) { int v = atoi(e); if (v >= 1 && v <= NRANGE) h->nrange = (uint32_t)v; }
	CK(cudaEventRecord(h->ev[0], h->stream));
	if (int rc = plan_loop(h, true)) return rc;
	Params P = make_params(h);
	CK(cudaEventRecord(h->ev[2], h->stream));
	if (int rc = group_and_output(h, true)) return rc;
	if (h->nnzB && h->m && h->hi > h->lo) {
		CK(cudaEventElapsedTime(&h->t_ms[0], h->ev[8], h->ev[9]));    // k_partition (includes waiting for the upload when it is still running)
		CK(cudaEventElapsedTime(&h->t_ms[6], h->ev[9], h->ev[2]));    // k_bucket + plan
	} else h->t_ms[0] = h->t_ms[6] = 0;
	return 0;
}

// numeric: the values were produced together with the structure; what is left is the compaction
// of the per-unit records into C's arrays
int run_numeric(bella_b200_handle* h)
{
	const uint64_t Z = h->Z;
	ENSURE(h->rowsC, sizeof(uint32_t) * (Z + 1));
	ENSURE(h->countC, sizeof(uint16_t) * (Z + 1));
	ENSURE(h->posH, sizeof(uint16_t) * (Z + 1));
	ENSURE(h->posV, sizeof(uint16_t) * (Z + 1));
	ENSURE(h->aux, sizeof(uint16_t) * 3 * (Z + 1));
	ENSURE(h->unpinned, sizeof(unsigned long long));
	CK(cudaMemsetAsync(h->unpinned.p, 0, sizeof(unsigned long long), h->stream));
	if (Z && h->U) {
		k_compact<<<grid_for((uint64_t)h->U * 32, 256), 256, 0, h->stream>>>(0u, h->U, h->uptr.as<uint64_t>(), h->uoff.as<uint32_t>(), h->out.as<uint4>(),
			h->rowsC.as<uint32_t>(), h->countC.as<uint16_t>(), h->posH.as<uint16_t>(), h->posV.as<uint16_t>(), h->aux.as<uint16_t>(),
			h->unpinned.as<unsigned long long>());
		LAUNCHED();
	}
	h->numeric_done = true;
	return 0;
}

int validate_views(bella_b200_handle* h, const bella_csc_view* A, const bella_csc_view* B, const uint32_t* read_len,
		const uint8_t* sB)
{
	if (!h) return BELLA_B200_ERR_ARG;
	if (!B || !B->colptr || !read_len) return fail(h, BELLA_B200_ERR_ARG, "B and read_len are required");
	if (!sB && B->rows > 0x80000000u) return fail(h, BELLA_B200_ERR_RANGE, "strand bits inside B.rowids need fewer than 2^31 k-mers");
	if (B->nnz && (!B->rowids || !B->values)) return fail(h, BELLA_B200_ERR_ARG, "B.rowids/B.values missing");
	if (B->cols > 0x7FFFFFFFu) return fail(h, BELLA_B200_ERR_RANGE, "more than 2^31-1 reads");
	if (A && (A->rows != B->cols || A->cols != B->rows || A->nnz != B->nnz))
		return fail(h, BELLA_B200_ERR_ARG, "A is %ux%u (nnz %u) but B is %ux%u (nnz %u): A must be the transpose of B", A->rows, A->cols,
			A->nnz, B->rows, B->cols, B->nnz);
	return 0;
}

void reset_problem(bella_b200_handle* h, const bella_csc_view* B, uint16_t K, uint16_t BIN)
{
	h->n = B->cols; h->m = B->rows; h->nnzB = B->nnz;
	h->lo = 0; h->hi = h->n; h->K = K; h->BIN = BIN;
	h->W = 0; h->cap_scale = 1.25;
	h->klo = 0; h->khi = h->m;
	h->mg_recv = nullptr;
	h->have_inputs = true; h->symbolic_done = h->numeric_done = false;
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
	h->sms = prop.multiProcessorCount;
	if (const char* e = getenv("BELLA_B200_EP_MAX")) { int v = atoi(e); if (v >= 1 && v <= (int)EP_LIMIT) h->ep_max = (uint32_t)v; }
	if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return BELLA_B200_ERR_CUDA; }
	for (auto& e : h->ev) cudaEventCreate(&e);
	cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking);
	cudaStreamCreateWithFlags(&h->aux_stream, cudaStreamNonBlocking);
	cudaStreamCreateWithFlags(&h->sc_stream, cudaStreamNonBlocking);
	cudaStreamCreateWithFlags(&h->out_stream, cudaStreamNonBlocking);
	for (int i = 0; i < 4; ++i) {
		cudaEventCreateWithFlags(&h->so_main[i], cudaEventDisableTiming); cudaEventCreateWithFlags(&h->so_aux[i], cudaEventDisableTiming);
		cudaEventCreateWithFlags(&h->so_out[i], cudaEventDisableTiming);
	}
	cudaEventCreate(&h->so_t0); cudaEventCreate(&h->so_t1);
	cudaHostAlloc((void**)&h->so_z, sizeof(uint32_t) * 16, cudaHostAllocDefault);
	cudaEventCreate(&h->sc_t0); cudaEventCreate(&h->sc_t1);
	cudaEventCreateWithFlags(&h->aux_fork, cudaEventDisableTiming); cudaEventCreateWithFlags(&h->aux_join, cudaEventDisableTiming);
	for (auto& e : h->chunk_ev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
	for (auto& e : h->pipe_ev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
	cudaEventCreate(&h->copy_begin); cudaEventCreate(&h->copy_end);
	*out = h;
	return BELLA_B200_OK;
}

int bella_b200_destroy(bella_b200_handle* h)
{
	if (!h) return BELLA_B200_ERR_ARG;
	cudaSetDevice(h->device);
	cudaStreamSynchronize(h->stream);
	DevBuf* bufs[] = {&h->oB_colptr, &h->oB_rowids, &h->oB_values, &h->oB_strand, &h->o_len, &h->boff, &h->bcur, &h->part, &h->partK, &h->Ainfo, &h->ccur, &h->rp_cur, &h->rp_tiles, &h->rp_E, &h->rp_K,
		&h->Aent, &h->Acolptr, &h->flop32, &h->nunits, &h->ubase, &h->shv, &h->refine, &h->colinfo, &h->ucol, &h->ucount,
		&h->uptr, &h->ucur, &h->lists, &h->redo, &h->unnz, &h->uoff, &h->raw, &h->out, &h->colptrC, &h->rowsC, &h->countC, &h->posH, &h->posV,
		&h->aux, &h->meta, &h->errflag, &h->cubtmp, &h->unpinned, &h->mg_colinfo, &h->mg_ucur, &h->mg_rcur, &h->mg_rtiles, &h->tp_kmer, &h->tp_read, &h->tp_pos, &h->tp_strand,
		&h->tp_rs, &h->tp_re, &h->tp_nruns, &h->tp_cnt, &h->tp_cp, &h->tp_merged, &h->tp_tmpK, &h->tp_tmpV, &h->tp_slab};
	for (DevBuf* b : bufs) b->release();
	for (auto& e : h->ev) if (e) cudaEventDestroy(e);
	for (auto& e : h->chunk_ev) if (e) cudaEventDestroy(e);
	for (auto& e : h->pipe_ev) if (e) cudaEventDestroy(e);
	if (h->copy_begin) cudaEventDestroy(h->copy_begin);
	if (h->copy_end) cudaEventDestroy(h->copy_end);
	if (h->copy_stream) { cudaStreamSynchronize(h->copy_stream); cudaStreamDestroy(h->copy_stream); }
	if (h->aux_stream) { cudaStreamSynchronize(h->aux_stream); cudaStreamDestroy(h->aux_stream); }
	if (h->sc_stream) { cudaStreamSynchronize(h->sc_stream); cudaStreamDestroy(h->sc_stream); }
	if (h->out_stream) { cudaStreamSynchronize(h->out_stream); cudaStreamDestroy(h->out_stream); }
	for (int i = 0; i < 4; ++i) { if (h->so_main[i]) cudaEventDestroy(h->so_main[i]); if (h->so_aux[i]) cudaEventDestroy(h->so_aux[i]); if (h->so_out[i]) cudaEventDestroy(h->so_out[i]); }
	if (h->so_t0) cudaEventDestroy(h->so_t0);
	if (h->so_t1) cudaEventDestroy(h->so_t1);
	if (h->so_z) cudaFreeHost(h->so_z);
	if (h->sc_t0) cudaEventDestroy(h->sc_t0);
	if (h->sc_t1) cudaEventDestroy(h->sc_t1);
	if (h->aux_fork) cudaEventDestroy(h->aux_fork);
	if (h->aux_join) cudaEventDestroy(h->aux_join);
	if (h->own_stream) cudaStreamDestroy(h->stream);
	delete h;
	return BELLA_B200_OK;
}

const char* bella_b200_last_error(const bella_b200_handle* h) { return h ? h->err.c_str() : "null handle"; }

int bella_b200_set_inputs(bella_b200_handle* h, const bella_csc_view* A, const bella_csc_view* B,
		const uint32_t* read_len, const uint8_t* strand_A, const uint8_t* strand_B, uint16_t kmer_size, uint16_t bin_size)
{
	(void)strand_A;
	if (int rc = validate_views(h, A, B, read_len, strand_B)) return rc;
	CK(cudaSetDevice(h->device));
	CK(cudaStreamSynchronize(h->copy_stream));                  // a previous upload must not be overwritten mid-flight
	reset_problem(h, B, kmer_size, bin_size);
	const size_t n = B->cols, nnz = B->nnz;
	ENSURE(h->oB_colptr, sizeof(uint32_t) * (n + 1) + 16);
	ENSURE(h->oB_rowids, sizeof(uint32_t) * nnz + 16);
	ENSURE(h->oB_values, sizeof(uint16_t) * nnz + 16);
	ENSURE(h->o_len, sizeof(uint32_t) * n + 16);
	if (strand_B) ENSURE(h->oB_strand, (nnz + 7) / 8 + 16);
	h->dB_colptr = h->oB_colptr.as<uint32_t>(); h->dB_rowids = h->oB_rowids.as<uint32_t>(); h->dB_values = h->oB_values.as<uint16_t>();
	h->dB_strand = strand_B ? h->oB_strand.as<uint8_t>() : nullptr; h->d_len = h->o_len.as<uint32_t>();
	// The upload runs on its own stream in up to MAX_CHUNKS ranges of reads of about equal nnz; the transpose
	// (bella_b200_symbolic) consumes each range as soon as it has landed.  The host arrays must therefore stay
	// valid until bella_b200_symbolic returns (page-locked memory makes the copies truly asynchronous).
	cudaStream_t cs = h->copy_stream;
	CK(cudaEventRecord(h->copy_begin, cs));
	CK(cudaMemcpyAsync(h->oB_colptr.p, B->colptr, sizeof(uint32_t) * (n + 1), cudaMemcpyHostToDevice, cs));
	CK(cudaMemcpyAsync(h->o_len.p, read_len, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, cs));
	int nc = nnz >= (1u << 22) ? bella_b200_handle::MAX_CHUNKS : 1;
	h->chunk_lo[0] = 0;
	for (int c = 1; c < nc; ++c) {
		const uint32_t target = (uint32_t)((uint64_t)nnz * c / nc);
		const uint32_t* it = std::lower_bound(B->colptr, B->colptr + n + 1, target);
		uint32_t r = (uint32_t)(it - B->colptr);
		if (r > n) r = (uint32_t)n;
		h->chunk_lo[c] = r < h->chunk_lo[c - 1] ? h->chunk_lo[c - 1] : r;
	}
	h->chunk_lo[nc] = (uint32_t)n;
	for (int c = 0; c < nc; ++c) {
		const size_t j0 = B->colptr[h->chunk_lo[c]], j1 = B->colptr[h->chunk_lo[c + 1]];
		if (j1 > j0) {
			CK(cudaMemcpyAsync(h->oB_rowids.as<uint32_t>() + j0, B->rowids + j0, sizeof(uint32_t) * (j1 - j0), cudaMemcpyHostToDevice, cs));
			CK(cudaMemcpyAsync(h->oB_values.as<uint16_t>() + j0, B->values + j0, sizeof(uint16_t) * (j1 - j0), cudaMemcpyHostToDevice, cs));
			if (strand_B) {
				const size_t b0 = j0 / 8, b1 = (j1 + 7) / 8;        // whole bytes: neighbouring chunks rewrite a shared byte with the same value
				CK(cudaMemcpyAsync(h->oB_strand.as<uint8_t>() + b0, strand_B + b0, b1 - b0, cudaMemcpyHostToDevice, cs));
			}
		}
		CK(cudaEventRecord(h->chunk_ev[c], cs));
	}
	CK(cudaEventRecord(h->copy_end, cs));
	h->n_chunks = nc;
	return BELLA_B200_OK;
}

int bella_b200_set_inputs_csr(bella_b200_handle* h, const bella_csr_view* A_csr, const uint32_t* read_len,
		const uint8_t* strand, uint16_t kmer_size, uint16_t bin_size)
{
	if (!h) return BELLA_B200_ERR_ARG;
	bella_csc_view B;
	if (!A_csr) return fail(h, BELLA_B200_ERR_ARG, "A_csr is required");
	if (bella_csr_as_transposed_csc(A_csr, &B)) return fail(h, BELLA_B200_ERR_ARG, "one-based CSR views (CSR::ConvertOneBased) are not accepted");
	return bella_b200_set_inputs(h, nullptr, &B, read_len, nullptr, strand, kmer_size, bin_size);
}

int bella_b200_set_inputs_device(bella_b200_handle* h, const bella_csc_view* A, const bella_csc_view* B,
		const uint32_t* read_len, const uint8_t* strand_A, const uint8_t* strand_B, uint16_t kmer_size, uint16_t bin_size)
{
	(void)strand_A;
	if (int rc = validate_views(h, A, B, read_len, strand_B)) return rc;
	CK(cudaSetDevice(h->device));
	reset_problem(h, B, kmer_size, bin_size);
	h->dB_colptr = B->colptr; h->dB_rowids = B->rowids; h->dB_values = B->values; h->dB_strand = strand_B; h->d_len = read_len;
	h->n_chunks = 0;
	h->t_ms[3] = 0;
	return BELLA_B200_OK;
}

int bella_b200_set_column_range(bella_b200_handle* h, uint32_t col_lo, uint32_t col_hi)
{
	if (!h || !h->have_inputs) return fail(h, BELLA_B200_ERR_ARG, "set_inputs first");
	if (col_lo > col_hi || col_hi > h->n) return fail(h, BELLA_B200_ERR_ARG, "column range [%u,%u) outside [0,%u)", col_lo, col_hi, h->n);
	h->lo = col_lo; h->hi = col_hi;
	h->symbolic_done = h->numeric_done = false;
	return BELLA_B200_OK;
}

static int do_symbolic(bella_b200_handle* h)
{
	CK(cudaSetDevice(h->device));
	h->launches = 0;
	if (h->n_chunks > 0) {
		// small arrays (colptr, read lengths) first: everything that follows reads them
		CK(cudaStreamWaitEvent(h->stream, h->copy_begin, 0));
		CK(cudaStreamWaitEvent(h->stream, h->chunk_ev[0], 0));
	}
	int rc = run_symbolic(h);
	if (h->n_chunks > 0) {
		CK(cudaStreamSynchronize(h->copy_stream));
		CK(cudaEventElapsedTime(&h->t_ms[3], h->copy_begin, h->copy_end));
	}
	return rc;
}

int bella_b200_symbolic(bella_b200_handle* h, uint64_t* flops, uint32_t* flopC, uint32_t* colptrC)
{
	if (!h || !h->have_inputs) return fail(h, BELLA_B200_ERR_ARG, "set_inputs first");
	if (int rc = do_symbolic(h)) return rc;
	const uint32_t ncols = h->hi - h->lo;
	CK(cudaEventRecord(h->ev[10], h->stream));
	if (flopC && ncols) CK(cudaMemcpyAsync(flopC, h->flop32.p, sizeof(uint32_t) * ncols, cudaMemcpyDeviceToHost, h->stream));
	if (colptrC) CK(cudaMemcpyAsync(colptrC, h->colptrC.p, sizeof(uint32_t) * ((size_t)ncols + 1), cudaMemcpyDeviceToHost, h->stream));
	CK(cudaEventRecord(h->ev[11], h->stream));
	CK(cudaStreamSynchronize(h->stream));
	CK(cudaEventElapsedTime(&h->t_ms[4], h->ev[10], h->ev[11]));
	if (h->so_done) h->t_ms[4] += h->so_ms;                    // + the result copies that ran beside the fold
	if (flops) *flops = h->flops;
	return BELLA_B200_OK;
}

int bella_b200_numeric_device(bella_b200_handle* h)
{
	if (!h || !h->symbolic_done) return fail(h, BELLA_B200_ERR_ARG, "bella_b200_symbolic first");
	if (h->numeric_done) return BELLA_B200_OK;
	CK(cudaSetDevice(h->device));
	CK(cudaEventRecord(h->ev[6], h->stream));
	if (int rc = run_numeric(h)) return rc;
	CK(cudaEventRecord(h->ev[7], h->stream));
	CK(cudaStreamSynchronize(h->stream));
	CK(cudaEventElapsedTime(&h->t_ms[2], h->ev[6], h->ev[7]));
	h->t_ms[2] += h->t_scans;                                  // output = C's colptr scans + compaction
	h->t_scans = 0;
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

int bella_b200_set_output_buffers(bella_b200_handle* h, uint32_t* rowidsC, uint16_t* count, uint16_t* posH, uint16_t* posV, uint64_t capacity)
{
	if (!h) return BELLA_B200_ERR_ARG;
	if (rowidsC && (!count || !posH || !posV || !capacity)) return fail(h, BELLA_B200_ERR_ARG, "all four output buffers and a capacity are required");
	h->so_rows = rowidsC; h->so_count = count; h->so_posH = posH; h->so_posV = posV; h->so_cap = rowidsC ? capacity : 0;
	h->so_done = false;
	return BELLA_B200_OK;
}

int bella_b200_numeric(bella_b200_handle* h, uint32_t col_begin, uint32_t col_end,
		uint32_t* rowidsC, uint16_t* count, uint16_t* posH, uint16_t* posV)
{
	// already delivered by the symbolic phase (bella_b200_set_output_buffers): nothing to copy
	if (h && h->symbolic_done && h->so_done && col_begin == h->lo && col_end == h->hi && rowidsC == h->so_rows && count == h->so_count
			&& posH == h->so_posH && posV == h->so_posV)
		return BELLA_B200_OK;
	if (int rc = bella_b200_numeric_device(h)) return rc;
	uint64_t off, cnt;
	if (int rc = copy_range(h, col_begin, col_end, &off, &cnt)) return rc;
	CK(cudaEventRecord(h->ev[10], h->stream));
	if (cnt) {
		if (rowidsC) CK(cudaMemcpyAsync(rowidsC, h->rowsC.as<uint32_t>() + off, sizeof(uint32_t) * cnt, cudaMemcpyDeviceToHost, h->stream));
		if (count) CK(cudaMemcpyAsync(count, h->countC.as<uint16_t>() + off, sizeof(uint16_t) * cnt, cudaMemcpyDeviceToHost, h->stream));
		if (posH) CK(cudaMemcpyAsync(posH, h->posH.as<uint16_t>() + off, sizeof(uint16_t) * cnt, cudaMemcpyDeviceToHost, h->stream));
		if (posV) CK(cudaMemcpyAsync(posV, h->posV.as<uint16_t>() + off, sizeof(uint16_t) * cnt, cudaMemcpyDeviceToHost, h->stream));
	}
	CK(cudaEventRecord(h->ev[11], h->stream));
	CK(cudaStreamSynchronize(h->stream));
	float t; CK(cudaEventElapsedTime(&t, h->ev[10], h->ev[11]));
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

int bella_b200_get_flops(bella_b200_handle* h, uint64_t* flops)
{
	if (!h || !h->symbolic_done || !flops) return fail(h, BELLA_B200_ERR_ARG, "the symbolic phase has not run");
	*flops = h->flops;
	return BELLA_B200_OK;
}

int bella_b200_n_unpinned(bella_b200_handle* h, uint64_t* n_unpinned)
{
	if (!h || !h->numeric_done || !n_unpinned) return fail(h, BELLA_B200_ERR_ARG, "numeric phase has not run");
	unsigned long long v = 0;
	CK(cudaMemcpyAsync(&v, h->unpinned.p, sizeof v, cudaMemcpyDeviceToHost, h->stream));
	CK(cudaStreamSynchronize(h->stream));
	*n_unpinned = v;
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
	h->symbolic_done = h->numeric_done = false;
	if (int rc = do_symbolic(h)) return rc;
	if (int rc = bella_b200_numeric_device(h)) return rc;
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

/* ---- multi-GPU: k-mer-range transposition + product exchange (bella_b200/distributed.py drives the collectives) ---- */

int bella_b200_mg_transpose(bella_b200_handle* h, uint32_t kmer_lo, uint32_t kmer_hi, uint32_t* cnt_local_dev)
{
	if (!h || !h->have_inputs) return fail(h, BELLA_B200_ERR_ARG, "set_inputs first");
	if (kmer_lo > kmer_hi || kmer_hi > h->m || !cnt_local_dev) return fail(h, BELLA_B200_ERR_ARG, "bad k-mer range [%u,%u) of %u", kmer_lo, kmer_hi, h->m);
	CK(cudaSetDevice(h->device));
	h->launches = 0;
	ENSURE(h->meta, sizeof(Meta));
	ENSURE(h->errflag, 4 * sizeof(int));
	h->klo = kmer_lo; h->khi = kmer_hi;
	h->mg_recv = nullptr;
	h->symbolic_done = h->numeric_done = false;
	CK(cudaEventRecord(h->ev[0], h->stream));
	for (;;) {
		CK(cudaMemsetAsync(h->errflag.p, 0, sizeof(int), h->stream));
		if (int rc = run_transpose(h, 0, 0, h->n, cnt_local_dev)) return rc;
		int e = 0;
		CK(cudaMemcpyAsync(&e, h->errflag.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
		CK(cudaStreamSynchronize(h->stream));
		if (e == ERR_BUCKET) {
			if (h->W <= 1) return fail(h, BELLA_B200_ERR_RANGE, "a k-mer occurs in more than %u reads", BUCKET_CAP);
			h->W = h->W / 2;
			h->cap_scale *= 1.5;
			continue;
		}
		if (e) return report_device_error(h, e);
		break;
	}
	CK(cudaEventRecord(h->ev[1], h->stream));
	return BELLA_B200_OK;
}

int bella_b200_mg_scatter(bella_b200_handle* h, const uint64_t* sendoff_dev, uint64_t* sendbuf_dev)
{
	if (!h || !h->have_inputs || !sendoff_dev) return fail(h, BELLA_B200_ERR_ARG, "bella_b200_mg_transpose first");
	CK(cudaSetDevice(h->device));
	const uint32_t n = h->n, ml = h->khi - h->klo;
	ENSURE(h->mg_ucur, sizeof(uint64_t) * ((size_t)n + 1) * CCUR_STRIDE);
	CK(cudaEventRecord(h->ev[10], h->stream));
	if (n && ml) {
		// every output column is "light" here: its cursor is its place in the send buffer
		k_mg_colinfo<<<grid_for(n, 256), 256, 0, h->stream>>>(n, sendoff_dev, h->mg_ucur.as<unsigned long long>());
		LAUNCHED();
		k_scatter<<<grid_for((h->nnzB + 3) / 4, 256, 148 * 32), 256, 0, h->stream>>>(h->boff.as<uint32_t>() + h->NB, ml, 0, n, h->Acolptr.as<uint32_t>(), h->Aent.as<uint64_t>(),
			h->Ainfo.as<uint8_t>(), h->mg_ucur.as<unsigned long long>(), nullptr, nullptr, sendbuf_dev, nullptr, 0u, 1u, h->errflag.as<int>());
		LAUNCHED();
	}
	CK(cudaEventRecord(h->ev[2], h->stream));
	return BELLA_B200_OK;
}

int bella_b200_mg_finish(bella_b200_handle* h, uint32_t col_lo, uint32_t col_hi, int world, const uint32_t* counts_all_dev,
		const uint64_t* segoff_dev, const uint64_t* recvbase_dev, const uint64_t* recv_dev)
{
	if (!h || !h->have_inputs) return fail(h, BELLA_B200_ERR_ARG, "set_inputs first");
	if (col_lo > col_hi || col_hi > h->n || world < 1 || !counts_all_dev || !segoff_dev || !recvbase_dev)
		return fail(h, BELLA_B200_ERR_ARG, "bad arguments to bella_b200_mg_finish");
	CK(cudaSetDevice(h->device));
	h->lo = col_lo; h->hi = col_hi;
	const uint32_t ncols = col_hi - col_lo;
	h->mg_recv = recv_dev ? recv_dev : (const uint64_t*)h->errflag.p;     // non-null marks the exchange mode even when nothing was received
	h->mg_counts = counts_all_dev; h->mg_segoff = segoff_dev; h->mg_recvbase = recvbase_dev; h->mg_world = world;
	CK(cudaEventRecord(h->ev[6], h->stream));
	ENSURE(h->flop32, sizeof(uint32_t) * ((size_t)ncols + 1));
	if (ncols) {
		k_mg_sum_counts<<<grid_for(ncols, 256), 256, 0, h->stream>>>(h->n, col_lo, ncols, (uint32_t)world, counts_all_dev, h->flop32.as<uint32_t>(),
			h->errflag.as<int>());
		LAUNCHED();
	}
	h->nrange = 1;
	int rc = plan_loop(h, false);
	if (!rc && h->flops) {
		k_regroup<false><<<h->sms * 8, 256, 0, h->stream>>>(h->n, col_lo, ncols, (uint32_t)world, counts_all_dev, segoff_dev, recvbase_dev, h->mg_recv,
			h->colinfo.as<ColInfo>(), nullptr, h->ucur.as<unsigned long long>(), h->raw.as<uint64_t>(), h->meta.as<Meta>(), h->errflag.as<int>());
		LAUNCHED();
	}
	if (!rc) rc = group_and_output(h, false);
	h->mg_recv = nullptr;
	if (rc) return rc;
	if (h->nnzB && h->khi > h->klo && h->n) {
		CK(cudaEventElapsedTime(&h->t_ms[0], h->ev[8], h->ev[9]));    // k_partition of this GPU's k-mer range
		CK(cudaEventElapsedTime(&h->t_ms[6], h->ev[9], h->ev[1]));    // k_bucket
	} else h->t_ms[0] = h->t_ms[6] = 0;
	CK(cudaEventElapsedTime(&h->t_ms[7], h->ev[10], h->ev[2]));    // expansion into the send buffer
	return BELLA_B200_OK;
}

int bella_b200_get_colptr(bella_b200_handle* h, uint32_t* colptrC_host)
{
	if (!h || !h->symbolic_done || !colptrC_host) return fail(h, BELLA_B200_ERR_ARG, "the symbolic phase has not run");
	CK(cudaMemcpyAsync(colptrC_host, h->colptrC.p, sizeof(uint32_t) * ((size_t)(h->hi - h->lo) + 1), cudaMemcpyDeviceToHost, h->stream));
	CK(cudaStreamSynchronize(h->stream));
	return BELLA_B200_OK;
}

/* ---- multi-GPU over NVLink peer memory: no collective on the data path (bella_b200/distributed.py, mode "nvlink") ---- */

int bella_b200_mg_geometry(uint32_t n_kmers, uint64_t nnz_total, int world, uint32_t* out8)
{
	if (!out8 || world < 1 || world > MG_MAXW) return BELLA_B200_ERR_ARG;
	RpGeom g;
	if (rp_geometry(n_kmers, nnz_total, 0, &g)) return BELLA_B200_ERR_RANGE;
	const uint32_t nb1_loc = (g.nb1 + (uint32_t)world - 1) / (uint32_t)world;
	const uint64_t kpr = (uint64_t)nb1_loc << g.shift1;
	if (kpr > 0xFFFFFFFFull) return BELLA_B200_ERR_RANGE;
	// a slot holds what ONE rank sends to ONE rank: 1/world^2 of the nonzeros on average (the caller may raise it)
	const uint64_t cap1 = (((uint64_t)((double)nnz_total / ((double)world * world) * 1.3) + 2 * RP_TILE) + 15) & ~15ull;
	if (cap1 > 0x7FFFFFFFull) return BELLA_B200_ERR_RANGE;
	out8[0] = g.wshift; out8[1] = g.shift1; out8[2] = g.nb1; out8[3] = nb1_loc; out8[4] = (uint32_t)kpr; out8[5] = (uint32_t)cap1;
	out8[6] = g.NB; out8[7] = g.l2;
	return BELLA_B200_OK;
}

int bella_b200_mg_route_push(bella_b200_handle* h, uint32_t read_lo, uint32_t read_hi, const uint32_t* colptr_global_dev, const uint32_t* rowids_dev,
		const uint16_t* values_dev, uint32_t n_kmers, const uint32_t* geom8, int world, int me, void* const* peer_E, void* const* peer_K, void* const* peer_cnt)
{
	if (!h || !geom8 || world < 1 || world > MG_MAXW || me < 0 || me >= world || !peer_E || !peer_K || !peer_cnt)
		return fail(h, BELLA_B200_ERR_ARG, "bad arguments to bella_b200_mg_route_push");
	CK(cudaSetDevice(h->device));
	h->launches = 0;
	ENSURE(h->errflag, 4 * sizeof(int));
	// one bucket per destination rank: geom8[5] = records one rank may send to one rank
	const uint32_t shift1 = geom8[1], nb1 = (uint32_t)world, nb1_loc = 1, kpr = geom8[4], cap1 = geom8[5];
	ENSURE(h->mg_rcur, sizeof(uint32_t) * ((size_t)nb1 + 2) * BCNT_STRIDE);
	CK(cudaMemsetAsync(h->mg_rcur.p, 0, sizeof(uint32_t) * ((size_t)nb1 + 2) * BCNT_STRIDE, h->stream));
	CK(cudaMemsetAsync(h->errflag.p, 0, sizeof(int), h->stream));
	CK(cudaFuncSetAttribute(k_rp1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RP_SMEM));
	CK(cudaEventRecord(h->ev[8], h->stream));
	RpOut O{};
	O.groups = (uint32_t)world; O.me = (uint32_t)me; O.nb1_loc = nb1_loc; O.kpr = kpr; O.route = 1;
	RpPost P{};
	for (int d = 0; d < world; ++d) { O.E[d] = (uint64_t*)peer_E[d]; O.K[d] = (uint32_t*)peer_K[d]; P.cnt[d] = (uint32_t*)peer_cnt[d]; }
	if (read_hi > read_lo) {
		const uint32_t wpc = RP_THREADS / 32;
		uint32_t g = (read_hi - read_lo + wpc - 1) / wpc;
		if (g > (uint32_t)h->sms * 2) g = (uint32_t)h->sms * 2;
		// the strand bit rides in bit 31 of the row ids (panel format): Bstrand == NULL
		k_rp1<<<g, RP_THREADS, RP_SMEM, h->stream>>>(read_hi, read_lo, 0u, n_kmers, colptr_global_dev, rowids_dev, values_dev, nullptr,
			shift1, nb1, cap1, h->mg_rcur.as<uint32_t>(), O, h->errflag.as<int>());
		LAUNCHED();
	}
	k_rp_post<<<grid_for(nb1, 256), 256, 0, h->stream>>>(nb1, cap1, (uint32_t)world, (uint32_t)me, nb1_loc, h->mg_rcur.as<uint32_t>(), P);
	LAUNCHED();
	CK(cudaEventRecord(h->ev[9], h->stream));
	return BELLA_B200_OK;
}

int bella_b200_mg_transpose_coarse(bella_b200_handle* h, uint32_t kmer_lo, uint32_t kmer_hi, const uint32_t* geom8, int world,
		const uint64_t* E_dev, const uint32_t* K_dev, const uint32_t* cnt_dev, uint64_t nnz_cap, uint32_t* cnt_local_dev)
{
	if (!h || !h->have_inputs) return fail(h, BELLA_B200_ERR_ARG, "set_inputs first");
	if (kmer_lo > kmer_hi || kmer_hi > h->m || !geom8 || !cnt_local_dev || world < 1 || world > MG_MAXW)
		return fail(h, BELLA_B200_ERR_ARG, "bad arguments to bella_b200_mg_transpose_coarse");
	CK(cudaSetDevice(h->device));
	ENSURE(h->meta, sizeof(Meta));
	ENSURE(h->errflag, 4 * sizeof(int));
	h->klo = kmer_lo; h->khi = kmer_hi;
	h->mg_recv = nullptr;
	h->symbolic_done = h->numeric_done = false;
	const uint32_t n = h->n, ml = kmer_hi - kmer_lo;
	const uint32_t wshift = geom8[0], shift1 = geom8[1], nb1_loc = geom8[3], cap1 = geom8[5], l2 = geom8[7];
	const uint32_t W = 1u << wshift;
	const uint32_t NB = (uint32_t)(((uint64_t)ml + W - 1) / W);
	const uint32_t nb1_mine = (uint32_t)(((uint64_t)ml + (1ull << shift1) - 1) >> shift1);
	if (nb1_mine > nb1_loc) return fail(h, BELLA_B200_ERR_ARG, "k-mer range wider than the rank's coarse buckets");
	h->W = W; h->NB = NB;
	ENSURE(h->Acolptr, sizeof(uint32_t) * ((size_t)ml + 2));
	const uint64_t rec_max = (uint64_t)world * geom8[5];           // what the slots can hold at most
	ENSURE(h->Aent, sizeof(uint64_t) * (rec_max + 2));
	ENSURE(h->Ainfo, rec_max + 16);
	ENSURE(h->boff, sizeof(uint32_t) * ((size_t)NB + 2));
	CK(cudaMemsetAsync(cnt_local_dev, 0, sizeof(uint32_t) * (size_t)n, h->stream));
	if (!ml || !n) {
		CK(cudaMemsetAsync(h->Acolptr.p, 0, sizeof(uint32_t) * ((size_t)ml + 2), h->stream));
		CK(cudaMemsetAsync(h->boff.p, 0, sizeof(uint32_t) * 2, h->stream));
		h->NB = 0;
		CK(cudaEventRecord(h->ev[1], h->stream));
		return BELLA_B200_OK;
	}
	ENSURE(h->bcur, sizeof(uint32_t) * ((size_t)NB + 2) * BCNT_STRIDE);
	ENSURE(h->part, sizeof(uint64_t) * (size_t)NB * BUCKET_CAP + 64);
	ENSURE(h->partK, sizeof(uint16_t) * (size_t)NB * BUCKET_CAP + 64);
	// the received records (one slot of cap1 records per source rank) -> this rank's coarse buckets -> fine buckets
	const uint32_t capR = cap1, nsbR = (uint32_t)world;
	const uint64_t capc64 = (((uint64_t)((double)nnz_cap / nb1_mine * 1.3) + 2 * RP_TILE) + 15) & ~15ull;     // nnz_cap: records this rank expects
	if (capc64 > 0x7FFFFFFFull) return fail(h, BELLA_B200_ERR_RANGE, "coarse transpose bucket too large");
	const uint32_t capC = (uint32_t)capc64;
	ENSURE(h->mg_rtiles, sizeof(uint32_t) * ((size_t)nsbR + 2));
	ENSURE(h->rp_tiles, sizeof(uint32_t) * ((size_t)nb1_mine + 2));
	ENSURE(h->rp_cur, sizeof(uint32_t) * ((size_t)nb1_mine + 2) * BCNT_STRIDE);
	ENSURE(h->rp_E, sizeof(uint64_t) * (size_t)nb1_mine * capC + 64);
	ENSURE(h->rp_K, sizeof(uint32_t) * (size_t)nb1_mine * capC + 64);
	CK(cudaMemsetAsync(h->rp_cur.p, 0, sizeof(uint32_t) * ((size_t)nb1_mine + 2) * BCNT_STRIDE, h->stream));
	CK(cudaMemsetAsync(h->bcur.p, 0, sizeof(uint32_t) * ((size_t)NB + 2) * BCNT_STRIDE, h->stream));
	CK(cudaFuncSetAttribute(k_rp1b, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RP_SMEM));
	CK(cudaFuncSetAttribute(k_rp2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RP_SMEM));
	CK(cudaFuncSetAttribute(k_bucket, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BUCKET_SMEM));
	(void)l2;
	k_rp_tiles<<<1, 1024, 0, h->stream>>>(nsbR, capR, cnt_dev, h->mg_rtiles.as<uint32_t>());
	LAUNCHED();
	k_rp1b<<<h->sms * 2, RP_THREADS, RP_SMEM, h->stream>>>(shift1, nb1_mine, nsbR, capR, cnt_dev, h->mg_rtiles.as<uint32_t>(), E_dev, K_dev,
		capC, h->rp_cur.as<uint32_t>(), h->rp_E.as<uint64_t>(), h->rp_K.as<uint32_t>(), h->errflag.as<int>());
	LAUNCHED();
	k_rp_tiles<<<1, 1024, 0, h->stream>>>(nb1_mine, capC, h->rp_cur.as<uint32_t>(), h->rp_tiles.as<uint32_t>());
	LAUNCHED();
	k_rp2<<<h->sms * 2, RP_THREADS, RP_SMEM, h->stream>>>(shift1, wshift, nb1_mine, 1u, capC, h->rp_cur.as<uint32_t>(), h->rp_tiles.as<uint32_t>(), 0u, nb1_mine,
		h->rp_E.as<uint64_t>(), h->rp_K.as<uint32_t>(), h->bcur.as<uint32_t>(), h->part.as<uint64_t>(), h->partK.as<uint16_t>(), h->errflag.as<int>());
	LAUNCHED();
	k_bucket_offsets<<<1, 1024, 0, h->stream>>>(0u, NB, h->bcur.as<uint32_t>(), h->boff.as<uint32_t>());
	LAUNCHED();
	k_bucket<<<NB < (uint32_t)h->sms * 16 ? NB : (uint32_t)h->sms * 16, 256, BUCKET_SMEM, h->stream>>>(kmer_lo, ml, 0u, n, wshift, 0u, NB, NB, h->boff.as<uint32_t>(),
		h->part.as<uint64_t>(), h->partK.as<uint16_t>(), h->Acolptr.as<uint32_t>(), h->Aent.as<uint64_t>(), h->Ainfo.as<uint8_t>(), cnt_local_dev, h->errflag.as<int>());
	LAUNCHED();
	CK(cudaEventRecord(h->ev[1], h->stream));
	return BELLA_B200_OK;
}

int bella_b200_mg_post(bella_b200_handle* h, uint64_t count, const uint32_t* src_dev, int world, uint64_t at, void* const* peer_dst)
{
	if (!h || world < 1 || world > MG_MAXW || !peer_dst || (count && !src_dev)) return fail(h, BELLA_B200_ERR_ARG, "bad arguments to bella_b200_mg_post");
	CK(cudaSetDevice(h->device));
	MgPeers P{};
	for (int d = 0; d < world; ++d) P.p[d] = peer_dst[d];
	if (count) { k_mg_post<<<grid_for(count, 256), 256, 0, h->stream>>>(count, src_dev, (uint32_t)world, at, P); LAUNCHED(); }
	return BELLA_B200_OK;
}

int bella_b200_mg_exchange(bella_b200_handle* h, int world, int me, const uint32_t* cuts_dev, const uint32_t* cnt_local_dev, void* const* peer_counts_all,
		int phase, const uint32_t* counts_all_dev, uint64_t* scan_dev, uint64_t cap_recv, uint64_t cap_send, uint64_t* sendoff_dev, uint64_t* segoff_dev,
		uint64_t* recvbase_dev, uint64_t* push_dev, uint64_t* sendbuf_dev, void* const* peer_recv)
{
	// phase 0: post this rank's per-column counts to every rank      (then the caller's barrier)
	// phase 1: plan on the device, expand the products into the send buffer, push the blocks to their owners   (then a barrier)
	if (!h || !h->have_inputs || world < 1 || world > MG_MAXW || me < 0 || me >= world) return fail(h, BELLA_B200_ERR_ARG, "bad arguments to bella_b200_mg_exchange");
	CK(cudaSetDevice(h->device));
	const uint32_t n = h->n;
	if (phase == 0) {
		MgPeers P{};
		for (int d = 0; d < world; ++d) P.p[d] = peer_counts_all[d];
		if (n) { k_mg_post<<<grid_for(n, 256), 256, 0, h->stream>>>((uint64_t)n, cnt_local_dev, (uint32_t)world, (uint64_t)me * n, P); LAUNCHED(); }
		// and this rank's error flag (a bucket that overflowed, ...): slot `me` behind the counts, 16 words in
		k_mg_post<<<1, 32, 0, h->stream>>>(1ull, (const uint32_t*)h->errflag.p, (uint32_t)world, (uint64_t)world * n + 16 + me, P);
		LAUNCHED();
		return BELLA_B200_OK;
	}
	// flat exclusive scan of counts_all (world rows of n) -> scan_dev u64 [world * n + 1]
	{
		auto in = thrust::make_transform_iterator(counts_all_dev, Widen());
		if (int rc = exclusive_scan(h, in, (unsigned long long*)scan_dev, (uint32_t)((size_t)world * n + 1))) return rc;
	}
	k_mg_plan<<<grid_for((uint64_t)n + 1, 256), 256, 0, h->stream>>>((uint32_t)world, (uint32_t)me, n, cuts_dev, (const unsigned long long*)scan_dev, cap_recv, cap_send,
		(const int*)(counts_all_dev + (size_t)world * n + 16), (unsigned long long*)sendoff_dev, (unsigned long long*)segoff_dev, (unsigned long long*)recvbase_dev, (unsigned long long*)push_dev, h->errflag.as<int>());
	LAUNCHED();
	if (int rc = bella_b200_mg_scatter(h, sendoff_dev, sendbuf_dev)) return rc;
	MgPeers R{};
	for (int d = 0; d < world; ++d) R.p[d] = peer_recv[d];
	k_mg_push<<<h->sms * 4, 256, 0, h->stream>>>((uint32_t)world, (uint32_t)me, sendbuf_dev, (const unsigned long long*)push_dev, R, h->errflag.as<int>());
	LAUNCHED();
	return BELLA_B200_OK;
}

/* ---- matrix construction on the device: tuples -> B (reference src/CSC.cpp:422-479 + MergeDuplicates :301-420) ---- */

int bella_b200_set_inputs_tuples(bella_b200_handle* h, uint32_t n_kmers, uint32_t n_reads, uint64_t ntuples, const uint32_t* t_kmer,
		const uint32_t* t_read, const uint16_t* t_pos, const uint8_t* t_strand, const uint32_t* read_len, uint16_t kmer_size, uint16_t bin_size)
{
	if (!h) return BELLA_B200_ERR_ARG;
	if (!read_len || (ntuples && (!t_kmer || !t_read || !t_pos || !t_strand))) return fail(h, BELLA_B200_ERR_ARG, "tuple arrays, strand bits and read_len are required");
	if (ntuples > 0xFFFFFFFFull) return fail(h, BELLA_B200_ERR_RANGE, "more than 2^32-1 tuples (CSC<uint32_t,...> cannot index them either)");
	if (n_reads > 0x7FFFFFFFu || n_kmers > 0x80000000u) return fail(h, BELLA_B200_ERR_RANGE, "more than 2^31-1 reads or 2^31 k-mers");
	CK(cudaSetDevice(h->device));
	CK(cudaStreamSynchronize(h->copy_stream));
	const size_t T = (size_t)ntuples, n = n_reads;
	ENSURE(h->meta, sizeof(Meta));
	ENSURE(h->errflag, 4 * sizeof(int));
	ENSURE(h->tp_kmer, sizeof(uint32_t) * T + 16); ENSURE(h->tp_read, sizeof(uint32_t) * T + 16);
	ENSURE(h->tp_pos, sizeof(uint16_t) * T + 16); ENSURE(h->tp_strand, (T + 7) / 8 + 16);
	ENSURE(h->tp_rs, sizeof(uint32_t) * (n + 1)); ENSURE(h->tp_re, sizeof(uint32_t) * (n + 1)); ENSURE(h->tp_nruns, sizeof(uint32_t) * (n + 1));
	ENSURE(h->tp_cnt, sizeof(uint32_t) * (n + 2)); ENSURE(h->tp_cp, sizeof(uint32_t) * (n + 2)); ENSURE(h->tp_merged, sizeof(uint32_t) * (n + 2));
	ENSURE(h->tp_tmpK, sizeof(uint32_t) * T + 16); ENSURE(h->tp_tmpV, sizeof(uint16_t) * T + 16);
	ENSURE(h->oB_colptr, sizeof(uint32_t) * (n + 2)); ENSURE(h->oB_rowids, sizeof(uint32_t) * T + 16); ENSURE(h->oB_values, sizeof(uint16_t) * T + 16);
	ENSURE(h->o_len, sizeof(uint32_t) * n + 16);
	cudaStream_t st = h->stream;
	CK(cudaEventRecord(h->ev[10], st));
	CK(cudaMemcpyAsync(h->tp_kmer.p, t_kmer, sizeof(uint32_t) * T, cudaMemcpyHostToDevice, st));
	CK(cudaMemcpyAsync(h->tp_read.p, t_read, sizeof(uint32_t) * T, cudaMemcpyHostToDevice, st));
	CK(cudaMemcpyAsync(h->tp_pos.p, t_pos, sizeof(uint16_t) * T, cudaMemcpyHostToDevice, st));
	CK(cudaMemcpyAsync(h->tp_strand.p, t_strand, (T + 7) / 8, cudaMemcpyHostToDevice, st));
	CK(cudaMemcpyAsync(h->o_len.p, read_len, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, st));
	CK(cudaEventRecord(h->ev[11], st));
	CK(cudaMemsetAsync(h->errflag.p, 0, 4 * sizeof(int), st));
	CK(cudaMemsetAsync(h->tp_nruns.p, 0, sizeof(uint32_t) * (n + 1), st));
	CK(cudaMemsetAsync(h->tp_cnt.p, 0, sizeof(uint32_t) * (n + 2), st));
	CK(cudaMemsetAsync(h->tp_merged.p, 0, sizeof(uint32_t) * (n + 2), st));
	uint32_t nnz = 0;
	if (T && n) {
		k_tuple_runs<<<grid_for(T, 256), 256, 0, st>>>(T, h->tp_read.as<uint32_t>(), n_reads, h->tp_rs.as<uint32_t>(), h->tp_re.as<uint32_t>(),
			h->tp_nruns.as<uint32_t>(), h->errflag.as<int>());
		LAUNCHED();
		k_tuple_counts<<<grid_for(n, 256), 256, 0, st>>>(n_reads, h->tp_rs.as<uint32_t>(), h->tp_re.as<uint32_t>(), h->tp_nruns.as<uint32_t>(),
			h->tp_cnt.as<uint32_t>(), h->errflag.as<int>());
		LAUNCHED();
		int maxcnt = 0;
		CK(cudaMemcpyAsync(&maxcnt, h->errflag.as<int>() + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
		if (int rc = exclusive_scan(h, h->tp_cnt.as<uint32_t>(), h->tp_cp.as<uint32_t>(), n_reads + 1)) return rc;
		CK(cudaStreamSynchronize(st));
		constexpr int HT = 2048, WARPS = 12;
		const size_t smem = (size_t)WARPS * 2 * HT * sizeof(uint32_t);
		CK(cudaFuncSetAttribute(k_merge_duplicates<HT, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		k_merge_duplicates<HT, WARPS><<<h->sms, WARPS * 32, smem, st>>>(n_reads, h->tp_rs.as<uint32_t>(), h->tp_cnt.as<uint32_t>(), h->tp_cp.as<uint32_t>(),
			h->tp_kmer.as<uint32_t>(), h->tp_pos.as<uint16_t>(), h->tp_strand.as<uint8_t>(), h->tp_tmpK.as<uint32_t>(), h->tp_tmpV.as<uint16_t>(),
			h->tp_merged.as<uint32_t>(), false, nullptr, 0);
		LAUNCHED();
		if (maxcnt > HT) {
			// reads with more than HT tuples: their tables (up to 65536 slots) live in a global slab
			uint32_t big_ht = 16;
			while (big_ht < (uint32_t)maxcnt) big_ht <<= 1;
			const uint32_t big_ctas = 74;
			ENSURE(h->tp_slab, sizeof(uint32_t) * (size_t)big_ctas * WARPS * 2 * big_ht);
			k_merge_duplicates<HT, WARPS><<<big_ctas, WARPS * 32, smem, st>>>(n_reads, h->tp_rs.as<uint32_t>(), h->tp_cnt.as<uint32_t>(),
				h->tp_cp.as<uint32_t>(), h->tp_kmer.as<uint32_t>(), h->tp_pos.as<uint16_t>(), h->tp_strand.as<uint8_t>(), h->tp_tmpK.as<uint32_t>(),
				h->tp_tmpV.as<uint16_t>(), h->tp_merged.as<uint32_t>(), true, h->tp_slab.as<uint32_t>(), big_ht);
			LAUNCHED();
		}
		if (int rc = exclusive_scan(h, h->tp_merged.as<uint32_t>(), h->oB_colptr.as<uint32_t>(), n_reads + 1)) return rc;
		k_compact_B<<<grid_for((uint64_t)n * 32, 256), 256, 0, st>>>(n_reads, h->tp_cp.as<uint32_t>(), h->oB_colptr.as<uint32_t>(),
			h->tp_tmpK.as<uint32_t>(), h->tp_tmpV.as<uint16_t>(), h->oB_rowids.as<uint32_t>(), h->oB_values.as<uint16_t>());
		LAUNCHED();
		int e = 0;
		CK(cudaMemcpyAsync(&nnz, h->oB_colptr.as<uint32_t>() + n_reads, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
		CK(cudaMemcpyAsync(&e, h->errflag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
		CK(cudaEventRecord(h->ev[9], st));
		CK(cudaStreamSynchronize(st));
		if (e == -8) return fail(h, BELLA_B200_ERR_ARG, "the tuples of a read must be contiguous (the order BELLA's tuple emission produces, src/main.cpp:393-416)");
		if (e == -1) return fail(h, BELLA_B200_ERR_ARG, "a tuple names a read id >= n_reads");
		if (e) return report_device_error(h, e);
		CK(cudaEventElapsedTime(&h->t_build_ms, h->ev[11], h->ev[9]));
		CK(cudaEventElapsedTime(&h->t_ms[3], h->ev[10], h->ev[11]));
	} else {
		CK(cudaMemsetAsync(h->oB_colptr.p, 0, sizeof(uint32_t) * (n + 2), st));
		CK(cudaStreamSynchronize(st));
	}
	bella_csc_view B{n_kmers, n_reads, nnz, h->oB_colptr.as<uint32_t>(), h->oB_rowids.as<uint32_t>(), h->oB_values.as<uint16_t>()};
	reset_problem(h, &B, kmer_size, bin_size);
	h->dB_colptr = B.colptr; h->dB_rowids = B.rowids; h->dB_values = B.values; h->dB_strand = nullptr; h->d_len = h->o_len.as<uint32_t>();
	h->n_chunks = 0;
	return BELLA_B200_OK;
}

int bella_b200_get_B(bella_b200_handle* h, uint32_t* nnz, uint32_t* colptr_host, uint32_t* rowids_host, uint16_t* values_host, uint8_t* strand_host,
		float* build_ms)
{
	if (!h || !h->have_inputs) return fail(h, BELLA_B200_ERR_ARG, "set_inputs first");
	CK(cudaSetDevice(h->device));
	CK(cudaStreamSynchronize(h->copy_stream));
	const size_t nz = h->nnzB;
	if (nnz) *nnz = (uint32_t)nz;
	if (build_ms) *build_ms = h->t_build_ms;
	if (colptr_host) CK(cudaMemcpyAsync(colptr_host, h->dB_colptr, sizeof(uint32_t) * ((size_t)h->n + 1), cudaMemcpyDeviceToHost, h->stream));
	if (values_host && nz) CK(cudaMemcpyAsync(values_host, h->dB_values, sizeof(uint16_t) * nz, cudaMemcpyDeviceToHost, h->stream));
	std::vector<uint32_t> rows;
	std::vector<uint8_t> bits;
	if ((rowids_host || strand_host) && nz) {
		rows.resize(nz);
		CK(cudaMemcpyAsync(rows.data(), h->dB_rowids, sizeof(uint32_t) * nz, cudaMemcpyDeviceToHost, h->stream));
		if (h->dB_strand && strand_host) {
			bits.resize((nz + 7) / 8);
			CK(cudaMemcpyAsync(bits.data(), h->dB_strand, (nz + 7) / 8, cudaMemcpyDeviceToHost, h->stream));
		}
	}
	CK(cudaStreamSynchronize(h->stream));
	const bool packed = h->dB_strand == nullptr;                   // strand bit in bit 31 of the row ids
	if (strand_host && nz) memset(strand_host, 0, (nz + 7) / 8);
	for (size_t j = 0; j < rows.size(); ++j) {
		if (rowids_host) rowids_host[j] = packed ? rows[j] & 0x7FFFFFFFu : rows[j];
		if (strand_host) {
			const uint32_t b = packed ? rows[j] >> 31 : (bits[j >> 3] >> (j & 7)) & 1u;
			if (b) strand_host[j >> 3] |= (uint8_t)(1u << (j & 7));
		}
	}
	return BELLA_B200_OK;
}

#ifdef BELLA_PHASE_CLOCKS
int bella_b200_debug_phases(unsigned long long* out32, int reset)
{
	if (out32) cudaMemcpyFromSymbol(out32, bk::g_phase, sizeof(unsigned long long) * 32);
	if (reset) { unsigned long long z[32] = {}; cudaMemcpyToSymbol(bk::g_phase, z, sizeof z); }
	return 0;
}
#endif

} // extern "C"
