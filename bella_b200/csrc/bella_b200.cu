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
	DevBuf Aent, Bent, tA_colptr, tcursor, flop64, flopptr, cursor, raw, bcount, nnzC, colptrC, lists, meta, errflag, cubtmp, slab;
	DevBuf prod, prow, pdesc, rowsC, countC, posH, posV, aux, fdesc, flists;
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

// layout: Aent (columns sorted by read id) + per-column product counts for the handle's column range
int run_layout(bella_b200_handle* h)
{
	const uint32_t n = h->n, m = h->m;
	const uint64_t nnz = h->nnzB;
	const uint32_t ncols = h->hi - h->lo;
	ENSURE(h->Aent, sizeof(uint64_t) * (nnz + 1));
	ENSURE(h->flop64, sizeof(uint64_t) * ((size_t)ncols + 1));
	ENSURE(h->errflag, sizeof(int));
	CK(cudaMemsetAsync(h->errflag.p, 0, sizeof(int), h->stream));
	CK(cudaMemsetAsync(h->flop64.p, 0, sizeof(uint64_t) * ((size_t)ncols + 1), h->stream));
	if (!nnz || !m) { h->layout_done = true; return 0; }
	if (h->have_A) {
		k_build_A<false><<<grid_for(m, 256), 256, 0, h->stream>>>(m, h->lo, h->hi, h->dA_colptr, h->dA_rowids, h->dA_values,
			h->dA_strand, h->Aent.as<uint64_t>(), h->flop64.as<unsigned long long>(), h->errflag.as<int>());
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
		h->dA_colptr = h->tA_colptr.as<uint32_t>();
		k_build_A<true><<<grid_for(m, 256), 256, 0, h->stream>>>(m, h->lo, h->hi, h->dA_colptr, nullptr, nullptr, nullptr,
			h->Aent.as<uint64_t>(), h->flop64.as<unsigned long long>(), h->errflag.as<int>());
		LAUNCHED();
	}
	h->layout_done = true;
	return 0;
}

Params make_params(bella_b200_handle* h)
{
	Params P{};
	P.n = h->n; P.m = h->m; P.lo = h->lo; P.hi = h->hi; P.K = h->K; P.BIN = h->BIN;
	P.B_colptr = h->dB_colptr; P.B_rowids = h->dB_rowids; P.A_colptr = h->dA_colptr; P.read_len = h->d_len;
	P.Aent = h->Aent.as<uint64_t>(); P.Bent = h->Bent.as<uint64_t>();
	P.flop64 = h->flop64.as<unsigned long long>(); P.flopptr = h->flopptr.as<uint64_t>();
	P.cursor = h->cursor.as<uint32_t>(); P.raw = h->raw.as<uint4>(); P.bcount = h->bcount.as<uint32_t>();
	P.nnzC = h->nnzC.as<uint32_t>(); P.colptrC = h->colptrC.as<uint32_t>();
	P.prod = h->prod.as<uint64_t>(); P.prod_half = h->flops; P.prow = h->prow.as<uint32_t>(); P.pdesc = h->pdesc.as<uint2>();
	P.rowsC = h->rowsC.as<uint32_t>(); P.countC = h->countC.as<uint16_t>(); P.posH = h->posH.as<uint16_t>();
	P.posV = h->posV.as<uint16_t>(); P.aux = h->aux.as<uint16_t>(); P.err = h->errflag.as<int>();
	return P;
}

// symbolic: product-count scan, outer-product scatter, per-column grouping (distinct rows = nnz(C)), scans
int run_symbolic(bella_b200_handle* h)
{
	const uint32_t ncols = h->hi - h->lo;
	const uint32_t m = h->m;
	ENSURE(h->flopptr, sizeof(uint64_t) * ((size_t)ncols + 1));
	ENSURE(h->cursor, sizeof(uint32_t) * ((size_t)ncols + 1));
	ENSURE(h->bcount, sizeof(uint32_t) * ((size_t)NBUCKETS * ncols + 1));
	ENSURE(h->nnzC, sizeof(uint32_t) * ((size_t)ncols + 1));
	ENSURE(h->colptrC, sizeof(uint32_t) * ((size_t)ncols + 1));
	ENSURE(h->lists, sizeof(uint32_t) * (size_t)N_CLASSES * (ncols + 1));
	ENSURE(h->meta, sizeof(Meta));
	if (int rc = exclusive_scan(h, h->flop64.as<unsigned long long>(), h->flopptr.as<unsigned long long>(), ncols + 1)) return rc;
	CK(cudaMemsetAsync(h->meta.p, 0, sizeof(Meta), h->stream));
	CK(cudaMemsetAsync(h->nnzC.p, 0, sizeof(uint32_t) * ((size_t)ncols + 1), h->stream));
	CK(cudaMemsetAsync(h->cursor.p, 0, sizeof(uint32_t) * ((size_t)ncols + 1), h->stream));
	CK(cudaMemsetAsync(h->bcount.p, 0, sizeof(uint32_t) * ((size_t)NBUCKETS * ncols + 1), h->stream));
	if (ncols) {
		k_classify<<<grid_for(ncols, 256), 256, 0, h->stream>>>(h->lo, ncols, h->flop64.as<unsigned long long>(), h->dB_colptr,
			h->lists.as<uint32_t>(), h->meta.as<Meta>(), h->errflag.as<int>());
		LAUNCHED();
	}
	k_set_total<<<1, 1, 0, h->stream>>>(h->meta.as<Meta>(), h->flopptr.as<uint64_t>(), ncols);
	LAUNCHED();
	CK(cudaMemcpyAsync(&h->hmeta, h->meta.p, sizeof(Meta), cudaMemcpyDeviceToHost, h->stream));
	if (int rc = check_device_error(h)) return rc;     // synchronises
	h->flops = h->hmeta.flops;
	const uint64_t F = h->flops;
	const uint32_t* cc = h->hmeta.class_count;
	const bool need_gather = cc[2] + cc[3] > 0;
	ENSURE(h->raw, sizeof(uint4) * (F + 1));
	ENSURE(h->prod, sizeof(uint64_t) * ((need_gather ? 2 : 1) * F + 1));
	ENSURE(h->prow, sizeof(uint32_t) * (F + 1));
	ENSURE(h->pdesc, sizeof(uint2) * (F + 1));
	if (need_gather) ENSURE(h->Bent, sizeof(uint64_t) * (h->nnzB + 1));
	Params P = make_params(h);
	CK(cudaEventRecord(h->ev[6], h->stream));
	if (F) {
		k_scatter<<<grid_for(m, 256), 256, 0, h->stream>>>(m, h->lo, h->hi, h->dA_colptr, P.Aent, P.flopptr, P.cursor, P.raw);
		LAUNCHED();
	}
	const uint32_t* lists = h->lists.as<uint32_t>();
	if (cc[0]) {
		using G = GroupSmem<2048, 4096>;
		CK(cudaFuncSetAttribute(k_group<2048, 4096>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::BYTES));
		k_group<2048, 4096><<<cc[0], 256, G::BYTES, h->stream>>>(P, lists, cc[0]);
		LAUNCHED();
	}
	if (cc[1]) {
		using G = GroupSmem<4096, 8192>;
		CK(cudaFuncSetAttribute(k_group<4096, 8192>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::BYTES));
		k_group<4096, 8192><<<cc[1], 256, G::BYTES, h->stream>>>(P, lists + (size_t)ncols, cc[1]);
		LAUNCHED();
	}
	for (int c = 2; c < 4; ++c) {
		if (!cc[c]) continue;
		const uint32_t* l = lists + (size_t)c * ncols;
		k_pack_B_list<<<grid_for((uint64_t)cc[c] * 32, 256), 256, 0, h->stream>>>(h->lo, l, cc[c], h->dB_colptr, h->dB_rowids,
			h->dB_values, h->dB_strand, h->dA_colptr, P.Aent, h->Bent.as<uint64_t>(), h->errflag.as<int>());
		LAUNCHED();
		if (c == 2) {
			size_t smem = (size_t)5 * (GATHER_SMEM_LIMIT + 1) * sizeof(uint32_t);
			CK(cudaFuncSetAttribute(k_expand_gather<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
			k_expand_gather<false><<<cc[c], 256, smem, h->stream>>>(P, l, cc[c], GATHER_SMEM_LIMIT, nullptr);
			LAUNCHED();
		} else {
			uint32_t htmax = 32;
			while (htmax < h->hmeta.max_flop) htmax <<= 1;
			uint32_t ctas = cc[c] < 148 ? cc[c] : 148;
			while (ctas > 1 && (size_t)ctas * 5 * ((size_t)htmax + 1) * sizeof(uint32_t) > ((size_t)8 << 30)) ctas >>= 1;
			ENSURE(h->slab, (size_t)ctas * 5 * ((size_t)htmax + 1) * sizeof(uint32_t));
			k_expand_gather<true><<<ctas, 256, 0, h->stream>>>(P, l, cc[c], htmax, h->slab.as<uint32_t>());
			LAUNCHED();
		}
	}
	CK(cudaEventRecord(h->ev[7], h->stream));
	if (int rc = exclusive_scan(h, h->nnzC.as<uint32_t>(), h->colptrC.as<uint32_t>(), ncols + 1)) return rc;
	if (int rc = exclusive_scan(h, h->bcount.as<uint32_t>(), h->bcount.as<uint32_t>(), NBUCKETS * ncols + 1)) return rc;
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
	ENSURE(h->flists, sizeof(uint32_t) * (Z + 1));
	Params P = make_params(h);
	if (!ncols || !Z) { h->numeric_done = true; return 0; }
	FDesc* fd = h->fdesc.as<FDesc>();
	uint32_t* fl = h->flists.as<uint32_t>();
	k_flatten<<<grid_for((uint64_t)ncols * 32, 256), 256, 0, h->stream>>>(P, fd, fl);
	LAUNCHED();
	const int g = 148 * 8;
	k_fold_short<1><<<g, 128, 0, h->stream>>>(P, fd, fl, 0); LAUNCHED();
	k_fold_short<4><<<g, 128, 0, h->stream>>>(P, fd, fl, 1); LAUNCHED();
	k_fold_short<8><<<g, 128, 0, h->stream>>>(P, fd, fl, 2); LAUNCHED();
	k_fold_short<16><<<g, 128, 0, h->stream>>>(P, fd, fl, 3); LAUNCHED();
	k_fold_short<32><<<g, 128, 0, h->stream>>>(P, fd, fl, 4); LAUNCHED();
	k_fold_long<<<148 * 4, WARP_FOLD_WARPS * 32, 0, h->stream>>>(P, fd, fl, 5); LAUNCHED();
	k_fold_huge<<<148, 128, 0, h->stream>>>(P, fd, fl, 6); LAUNCHED();
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
	if (B->cols > 0x7FFFFFFFu) return fail(h, BELLA_B200_ERR_RANGE, "more than 2^31-1 reads");
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
		&h->oB_strand, &h->o_len, &h->Aent, &h->Bent, &h->tA_colptr, &h->tcursor, &h->flop64, &h->flopptr, &h->cursor, &h->raw, &h->bcount, &h->nnzC, &h->colptrC,
		&h->lists, &h->meta, &h->errflag, &h->cubtmp, &h->slab, &h->prod, &h->prow, &h->pdesc, &h->rowsC, &h->countC, &h->posH,
		&h->posV, &h->aux, &h->fdesc, &h->flists};
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
	std::vector<unsigned long long> f64;
	if (flopC && ncols) {
		f64.resize(ncols);
		CK(cudaMemcpyAsync(f64.data(), h->flop64.p, sizeof(uint64_t) * ncols, cudaMemcpyDeviceToHost, h->stream));
	}
	if (colptrC) CK(cudaMemcpyAsync(colptrC, h->colptrC.p, sizeof(uint32_t) * ((size_t)ncols + 1), cudaMemcpyDeviceToHost, h->stream));
	CK(cudaEventRecord(h->ev[4], h->stream));
	CK(cudaStreamSynchronize(h->stream));
	CK(cudaEventElapsedTime(&h->t_ms[4], h->ev[3], h->ev[4]));
	for (size_t i = 0; i < f64.size(); ++i) flopC[i] = (uint32_t)f64[i];   // < 2^32, checked on the device
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
