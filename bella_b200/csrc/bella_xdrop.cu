// bella_xdrop.cu -- kernels + C-ABI (include/bella_xdrop.h) of the "next" row f1, batched gapped X-drop seed-and-extend.
// The algorithm lives in xdrop.cuh (shared with the CPU lane emulator of the tests); this file only launches it:
//
//   k_xdrop<G,T>     persistent CTAs of 8 warps; every group of G lanes pulls extensions (2 per pair) from one queue and
//                    runs them with the anti-diagonals in registers; windows that outgrow G*T slots go to a list
//   k_xdrop_thread<W> one THREAD per extension (small x: the window is too narrow for a warp), anti-diagonals in shared memory
//   k_xdrop_wide     one warp per listed extension, anti-diagonals in global scratch
//   k_xdrop_compose  joins the two halves of each pair and applies the reference's threshold test
//
// Reads stay resident on the device between batches; seeds come as (row, col, posH, posV) arrays -- host pointers
// (bella_xdrop_align) or the overlap SpGEMM's device result (bella_xdrop_align_device / _align_csc_device).
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>

#include <cstdio>
#include <cstring>
#include <string>

#include "bella_xdrop.h"
#include "xdrop.cuh"

namespace {

constexpr int WARPS = 8;                        // per CTA

template <int G, int T>
__global__ void __launch_bounds__(WARPS * 32) k_xdrop(xd::Pairs P, xd::Queue Q, xd::JobResult* res)
{
	constexpr int PER_WARP = (32 / G) * 2 * xd::Ext<G, T>::RING;
	__shared__ char rings[WARPS * PER_WARP];
	xd::warp_main<G, T>(P, Q, res, rings + (threadIdx.x >> 5) * PER_WARP);
}

__global__ void k_xdrop_encode(char* seqs, size_t n)     // raw bases -> Dna5 codes, in place
{
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) seqs[i] = xd::dna5(seqs[i]);
}

// thread-per-extension kernel: NT threads, each with 2 anti-diagonals of W ints + 2 * W bases in shared memory
template <int W>
__global__ void k_xdrop_thread(xd::Pairs P, xd::Queue Q, xd::JobResult* res)
{
	extern __shared__ int smem_thread[];
	const int nt = blockDim.x;
	xd::thread_main<W>(P, Q, res, smem_thread, (char*)(smem_thread + 2 * W * nt), nt, threadIdx.x);
}

// packed-word form (opt-in until measured): one 32-bit word per cell carrying its two bases, NT a template constant
template <int W, int NT>
__global__ void __launch_bounds__(NT) k_xdrop_thread_packed(xd::Pairs P, xd::Queue Q, xd::JobResult* res, const int* order)
{
	extern __shared__ int smem_packed[];
	xd::thread_main_packed<W, NT>(P, Q, res, smem_packed, threadIdx.x, order);
}

// both live anti-diagonals of a column in one word: W words per thread (opt-in until measured)
template <int W, int NT>
__global__ void __launch_bounds__(NT) k_xdrop_thread_two(xd::Pairs P, xd::Queue Q, xd::JobResult* res, const int* order)
{
	extern __shared__ int smem_two[];
	xd::thread_main_two<W, NT>(P, Q, res, smem_two, threadIdx.x, order);
}

// longest-first schedule: key = 65535 - min(query segment, database segment), sorted ascending (plumbing: cub radix sort)
__global__ void k_xdrop_estimate(xd::Pairs P, unsigned short* key, int* job)
{
	const int j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j < P.n_jobs) { key[j] = (unsigned short)(65535 - xd::job_estimate(P, j)); job[j] = j; }
}

// list == nullptr: every job of the batch; otherwise the *n_list jobs the register kernel gave up on
__global__ void __launch_bounds__(WARPS * 32) k_xdrop_wide(xd::Pairs P, const int* list, const int* n_list, int* next,
		xd::JobResult* res, int* scratch, int cap, int* bad)
{
	const int warp = blockIdx.x * WARPS + (threadIdx.x >> 5);
	xd::wide_main(P, list, list ? *n_list : P.n_jobs, next, res, scratch + (size_t)warp * 3 * cap, cap, bad);
}

__global__ void k_xdrop_compose(xd::Pairs P, const xd::JobResult* res, int n_pairs, double ratiophi, double delta,
		int fixed_threshold, int32_t* out)
{
	const int p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p < n_pairs) xd::compose(P, res, p, ratiophi, delta, fixed_threshold, out);
}

struct Buf {
	void* p = nullptr; size_t cap = 0;
	cudaError_t reserve(size_t bytes)
	{
		if (bytes <= cap) return cudaSuccess;
		if (p) cudaFree(p);
		p = nullptr; cap = 0;
		const cudaError_t e = cudaMalloc(&p, bytes);
		if (e == cudaSuccess) cap = bytes;
		return e;
	}
	void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

}  // namespace

struct bella_xdrop {
	int device = 0, sms = 0;
	cudaStream_t stream = nullptr;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	std::string err;
	Buf seqs, seq_off, rows, cols, posH, posV, out, res, wide, scratch, ctr, key_in, key_out, job_in, job_out, sort_tmp;
	uint32_t n_reads = 0; int max_len = 0;
	int kmer_len = 17, xdrop = 7, fixed_threshold = -1;
	double ratiophi = 0.0, delta = 0.1;
	int lanes = -1, cells = -1, used_lanes = 0, used_cells = 0;
	int launches = 0;
	bool timed = false;
};

namespace {

int fail(bella_xdrop* h, int code, const std::string& what) { h->err = what; return code; }

#define XCUDA(call)                                                                                                   \
	do {                                                                                                              \
		const cudaError_t e_ = (call);                                                                                \
		if (e_ != cudaSuccess) return fail(h, BELLA_XDROP_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
	} while (0)

template <int G, int T>
int launch_reg(bella_xdrop* h, const xd::Pairs& P, const xd::Queue& Q, xd::JobResult* res)
{
	int per_sm = 0;
	XCUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_xdrop<G, T>, WARPS * 32, 0));
	if (per_sm < 1) per_sm = 1;
	const long groups_needed = ((long)P.n_jobs + (32 / G) * WARPS - 1) / ((32 / G) * WARPS);
	long grid = (long)h->sms * per_sm;
	if (grid > groups_needed) grid = groups_needed;
	if (grid < 1) grid = 1;
	k_xdrop<G, T><<<(unsigned)grid, WARPS * 32, 0, h->stream>>>(P, Q, res);
	XCUDA(cudaGetLastError());
	++h->launches;
	return 0;
}

template <int W>
int launch_thread(bella_xdrop* h, const xd::Pairs& P, const xd::Queue& Q, xd::JobResult* res)
{
	// as many threads per SM as the shared memory holds (10 * W bytes each), in one CTA
	int nt = (int)((size_t)(227 * 1024 - 1024) / (10 * W)) / 32 * 32;
	if (nt > 1024) nt = 1024;
	const size_t smem = (size_t)nt * 10 * W;
	XCUDA(cudaFuncSetAttribute(k_xdrop_thread<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	int per_sm = 0;
	XCUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_xdrop_thread<W>, nt, smem));
	if (per_sm < 1) return fail(h, BELLA_XDROP_ECUDA, "k_xdrop_thread does not fit an SM");
	long grid = (long)h->sms * per_sm;
	const long needed = ((long)P.n_jobs + nt - 1) / nt;
	if (grid > needed) grid = needed;
	k_xdrop_thread<W><<<(unsigned)grid, nt, smem, h->stream>>>(P, Q, res);
	XCUDA(cudaGetLastError());
	++h->launches;
	return 0;
}

typedef void (*thread_kernel_t)(xd::Pairs, xd::Queue, xd::JobResult*, const int*);

int launch_thread_ordered(bella_xdrop* h, thread_kernel_t kernel, int NT, size_t smem, const xd::Pairs& P, const xd::Queue& Q,
		xd::JobResult* res, bool longest_first)
{
	const int* order = nullptr;
	if (longest_first) {
		const int n = P.n_jobs;
		XCUDA(h->key_in.reserve((size_t)n * 2)); XCUDA(h->key_out.reserve((size_t)n * 2));
		XCUDA(h->job_in.reserve((size_t)n * 4)); XCUDA(h->job_out.reserve((size_t)n * 4));
		k_xdrop_estimate<<<(n + 255) / 256, 256, 0, h->stream>>>(P, (unsigned short*)h->key_in.p, (int*)h->job_in.p);
		XCUDA(cudaGetLastError());
		++h->launches;
		size_t tmp = 0;
		XCUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp, (const unsigned short*)h->key_in.p, (unsigned short*)h->key_out.p,
				(const int*)h->job_in.p, (int*)h->job_out.p, n, 0, 16, h->stream));
		XCUDA(h->sort_tmp.reserve(tmp ? tmp : 1));
		XCUDA(cub::DeviceRadixSort::SortPairs(h->sort_tmp.p, tmp, (const unsigned short*)h->key_in.p, (unsigned short*)h->key_out.p,
				(const int*)h->job_in.p, (int*)h->job_out.p, n, 0, 16, h->stream));
		order = (const int*)h->job_out.p;
	}
	XCUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	int per_sm = 0;
	XCUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, NT, smem));
	if (per_sm < 1) return fail(h, BELLA_XDROP_ECUDA, "the thread kernel does not fit an SM");
	long grid = (long)h->sms * per_sm;
	const long needed = ((long)P.n_jobs + NT - 1) / NT;
	if (grid > needed) grid = needed;
	kernel<<<(unsigned)grid, NT, smem, h->stream>>>(P, Q, res, order);
	XCUDA(cudaGetLastError());
	++h->launches;
	return 0;
}

template <int W, int NT>
int launch_thread_packed(bella_xdrop* h, const xd::Pairs& P, const xd::Queue& Q, xd::JobResult* res, bool longest_first)
{
	return launch_thread_ordered(h, k_xdrop_thread_packed<W, NT>, NT, (size_t)2 * W * NT * sizeof(int), P, Q, res, longest_first);
}

template <int W, int NT>
int launch_thread_two(bella_xdrop* h, const xd::Pairs& P, const xd::Queue& Q, xd::JobResult* res, bool longest_first)
{
	return launch_thread_ordered(h, k_xdrop_thread_two<W, NT>, NT, (size_t)W * NT * sizeof(int), P, Q, res, longest_first);
}

void pick_shape(const bella_xdrop* h, int& G, int& T)
{
	if (h->lanes >= 0) { G = h->lanes; T = h->cells; return; }
	// the live window is about 2 * xdrop columns wide on real overlaps (xdrop + 9 near the seed), with a tail several
	// times that; whatever outgrows the chosen shape takes the wide path.  Measured on a B200 (profiles/xdrop_r01.md):
	// at x = 7 one thread per extension beats every warp shape.
	// Round 2, all shapes on the bench batch (60 k pairs, x = 7; profiles/xdrop_r02.md): (3,64) -- the packed thread kernel with
	// the jobs started longest first -- 38.3 ms, (2,64) 40.0, (1,64) 47.6, the warp shapes 60-115 ms; LOGAN on the same box 1057 ms.
	if (h->xdrop <= 12) { G = 3; T = 64; }
	else if (h->xdrop <= 40) { G = 32; T = 2; }
	else if (h->xdrop <= 100) { G = 32; T = 4; }
	else { G = 0; T = 0; }
}

int run_batch(bella_xdrop* h, uint64_t n_pairs, const uint32_t* d_rows, const uint32_t* d_cols, const uint16_t* d_posH,
		const uint16_t* d_posV, int32_t* d_out, const uint32_t* d_colptr = nullptr, int n_cols = 0)
{
	if (!h->seqs.p) return fail(h, BELLA_XDROP_EINVAL, "bella_xdrop_set_reads has not been called");
	if (n_pairs > (1u << 30) - 1) return fail(h, BELLA_XDROP_EINVAL, "more than 2^30 - 1 pairs in one batch");
	h->launches = 0;
	if (n_pairs == 0) return 0;
	const int n_jobs = (int)(2 * n_pairs);
	XCUDA(h->res.reserve((size_t)n_jobs * sizeof(xd::JobResult)));
	XCUDA(h->wide.reserve((size_t)n_jobs * sizeof(int)));
	XCUDA(h->ctr.reserve(4 * sizeof(int)));
	int* ctr = (int*)h->ctr.p;                  // [0] queue, [1] wide count, [2] wide queue, [3] bad seed
	XCUDA(cudaMemsetAsync(ctr, 0, 4 * sizeof(int), h->stream));
	xd::Pairs P{d_rows, d_cols, d_posH, d_posV, (const char*)h->seqs.p, (const uint64_t*)h->seq_off.p, h->kmer_len, h->xdrop, n_jobs, d_colptr, n_cols};
	xd::Queue Q{ctr, ctr + 1, (int*)h->wide.p, ctr + 3};
	xd::JobResult* res = (xd::JobResult*)h->res.p;
	int G, T;
	pick_shape(h, G, T);
	// shapes 4 and 5 keep the scores relative to the drop-off limit in 10-bit fields (xdrop.cuh ThreadExtQ): x + 5 must fit
	if ((G == 4 || G == 5) && h->xdrop > 1018) { G = 0; T = 0; }
	h->used_lanes = G; h->used_cells = T;
	const int cap = h->max_len + 3;
	// wide kernel: up to 4 CTAs per SM, as many as 1 GiB of anti-diagonal scratch allows (3 * cap ints per warp)
	const size_t per_cta = (size_t)WARPS * 3 * cap * sizeof(int);
	long per_sm_wide = (long)(((size_t)1 << 30) / per_cta / (size_t)h->sms);
	per_sm_wide = per_sm_wide < 1 ? 1 : per_sm_wide > 4 ? 4 : per_sm_wide;
	const int wide_grid = (int)(h->sms * per_sm_wide);
	XCUDA(h->scratch.reserve((size_t)wide_grid * WARPS * 3 * cap * sizeof(int)));
	XCUDA(cudaEventRecord(h->ev0, h->stream));
	int rc = 0;
	if (G == 0) {
		k_xdrop_wide<<<wide_grid, WARPS * 32, 0, h->stream>>>(P, nullptr, nullptr, ctr + 2, res, (int*)h->scratch.p, cap, ctr + 3);
		XCUDA(cudaGetLastError());
		++h->launches;
	} else {
		if (G == 32 && T == 1) rc = launch_reg<32, 1>(h, P, Q, res);
		else if (G == 32 && T == 2) rc = launch_reg<32, 2>(h, P, Q, res);
		else if (G == 32 && T == 4) rc = launch_reg<32, 4>(h, P, Q, res);
		else if (G == 16 && T == 1) rc = launch_reg<16, 1>(h, P, Q, res);
		else if (G == 16 && T == 2) rc = launch_reg<16, 2>(h, P, Q, res);
		else if (G == 16 && T == 4) rc = launch_reg<16, 4>(h, P, Q, res);
		else if (G == 8 && T == 4) rc = launch_reg<8, 4>(h, P, Q, res);
		else if (G == 8 && T == 8) rc = launch_reg<8, 8>(h, P, Q, res);
		else if (G == 1 && T == 64) rc = launch_thread<64>(h, P, Q, res);
		else if (G == 1 && T == 32) rc = launch_thread<32>(h, P, Q, res);
		else if ((G == 2 || G == 3) && T == 64) rc = launch_thread_packed<64, 128>(h, P, Q, res, G == 3);
		else if ((G == 2 || G == 3) && T == 32) rc = launch_thread_packed<32, 128>(h, P, Q, res, G == 3);
		else if ((G == 4 || G == 5) && T == 64) rc = launch_thread_two<64, 256>(h, P, Q, res, G == 5);
		else if ((G == 4 || G == 5) && T == 32) rc = launch_thread_two<32, 256>(h, P, Q, res, G == 5);
		else if ((G == 4 || G == 5) && T == 128) rc = launch_thread_two<128, 128>(h, P, Q, res, G == 5);
		else if ((G == 4 || G == 5) && T == 256) rc = launch_thread_two<256, 64>(h, P, Q, res, G == 5);
		else return fail(h, BELLA_XDROP_EINVAL, "unsupported shape (lanes, cells per lane)");
		if (rc) return rc;
		k_xdrop_wide<<<wide_grid, WARPS * 32, 0, h->stream>>>(P, (const int*)h->wide.p, ctr + 1, ctr + 2, res, (int*)h->scratch.p, cap, ctr + 3);
		XCUDA(cudaGetLastError());
		++h->launches;
	}
	k_xdrop_compose<<<(unsigned)((n_pairs + 255) / 256), 256, 0, h->stream>>>(P, res, (int)n_pairs, h->ratiophi, h->delta, h->fixed_threshold, d_out);
	XCUDA(cudaGetLastError());
	++h->launches;
	XCUDA(cudaEventRecord(h->ev1, h->stream));
	h->timed = true;
	return 0;
}

}  // namespace

extern "C" {

bella_xdrop* bella_xdrop_create(int device)
{
	if (cudaSetDevice(device) != cudaSuccess) return nullptr;
	bella_xdrop* h = new bella_xdrop;
	h->device = device;
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete h; return nullptr; }
	h->sms = prop.multiProcessorCount;
	if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess
			|| cudaEventCreate(&h->ev0) != cudaSuccess || cudaEventCreate(&h->ev1) != cudaSuccess) { delete h; return nullptr; }
	return h;
}

void bella_xdrop_destroy(bella_xdrop* h)
{
	if (!h) return;
	cudaSetDevice(h->device);
	cudaStreamSynchronize(h->stream);
	for (Buf* b : {&h->seqs, &h->seq_off, &h->rows, &h->cols, &h->posH, &h->posV, &h->out, &h->res, &h->wide, &h->scratch, &h->ctr,
			&h->key_in, &h->key_out, &h->job_in, &h->job_out, &h->sort_tmp}) b->release();
	cudaEventDestroy(h->ev0); cudaEventDestroy(h->ev1);
	cudaStreamDestroy(h->stream);
	delete h;
}

const char* bella_xdrop_last_error(const bella_xdrop* h) { return h ? h->err.c_str() : "null handle"; }

int bella_xdrop_set_reads(bella_xdrop* h, const char* seqs, const uint64_t* seq_off, uint32_t n_reads)
{
	if (!h) return BELLA_XDROP_EINVAL;
	if (!seqs || !seq_off) return fail(h, BELLA_XDROP_EINVAL, "null read arrays");
	XCUDA(cudaSetDevice(h->device));
	int max_len = 0;
	for (uint32_t r = 0; r < n_reads; ++r) {
		if (seq_off[r + 1] < seq_off[r]) return fail(h, BELLA_XDROP_EINVAL, "seq_off is not non-decreasing");
		const uint64_t len = seq_off[r + 1] - seq_off[r];
		if (len > 65535) return fail(h, BELLA_XDROP_EINVAL, "read longer than 65535 bases (positions are unsigned short in BELLA)");
		if ((int)len > max_len) max_len = (int)len;
	}
	const uint64_t total = seq_off[n_reads];
	XCUDA(h->seqs.reserve(total ? total : 1));
	XCUDA(h->seq_off.reserve(((size_t)n_reads + 1) * sizeof(uint64_t)));
	XCUDA(cudaMemcpyAsync(h->seqs.p, seqs, total, cudaMemcpyHostToDevice, h->stream));
	XCUDA(cudaMemcpyAsync(h->seq_off.p, seq_off, ((size_t)n_reads + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, h->stream));
	if (total) {
		k_xdrop_encode<<<h->sms * 8, 256, 0, h->stream>>>((char*)h->seqs.p, (size_t)total);
		XCUDA(cudaGetLastError());
	}
	XCUDA(cudaStreamSynchronize(h->stream));
	h->n_reads = n_reads; h->max_len = max_len;
	return 0;
}

int bella_xdrop_set_params(bella_xdrop* h, int kmer_len, int xdrop, double ratiophi, double delta_chernoff, int fixed_threshold)
{
	if (!h) return BELLA_XDROP_EINVAL;
	if (kmer_len < 1 || xdrop < 0) return fail(h, BELLA_XDROP_EINVAL, "kmer_len must be >= 1 and xdrop >= 0");
	h->kmer_len = kmer_len; h->xdrop = xdrop; h->ratiophi = ratiophi; h->delta = delta_chernoff; h->fixed_threshold = fixed_threshold;
	return 0;
}

int bella_xdrop_set_shape(bella_xdrop* h, int lanes, int cells_per_lane)
{
	if (!h) return BELLA_XDROP_EINVAL;
	const bool ok = (lanes == -1 && cells_per_lane == -1) || (lanes == 0 && cells_per_lane == 0)
		|| (lanes == 32 && (cells_per_lane == 1 || cells_per_lane == 2 || cells_per_lane == 4))
		|| (lanes == 16 && (cells_per_lane == 1 || cells_per_lane == 2 || cells_per_lane == 4))
		|| (lanes == 8 && (cells_per_lane == 4 || cells_per_lane == 8))
		|| (lanes >= 1 && lanes <= 5 && (cells_per_lane == 32 || cells_per_lane == 64))
		|| ((lanes == 4 || lanes == 5) && (cells_per_lane == 128 || cells_per_lane == 256));
	if (!ok) return fail(h, BELLA_XDROP_EINVAL, "unsupported shape (lanes, cells per lane)");
	h->lanes = lanes; h->cells = cells_per_lane;
	return 0;
}

int bella_xdrop_align_device(bella_xdrop* h, uint64_t n_pairs, const uint32_t* d_rows, const uint32_t* d_cols,
		const uint16_t* d_posH, const uint16_t* d_posV, int32_t* d_out)
{
	if (!h) return BELLA_XDROP_EINVAL;
	XCUDA(cudaSetDevice(h->device));
	return run_batch(h, n_pairs, d_rows, d_cols, d_posH, d_posV, d_out);
}

int bella_xdrop_align_csc_device(bella_xdrop* h, uint32_t n_cols, const uint32_t* d_colptrC, uint64_t n_pairs, const uint32_t* d_rowids,
		const uint16_t* d_posH, const uint16_t* d_posV, int32_t* d_out)
{
	if (!h) return BELLA_XDROP_EINVAL;
	if (n_cols != h->n_reads) return fail(h, BELLA_XDROP_EINVAL, "the overlap matrix must have one column per read");
	if (n_pairs && !d_colptrC) return fail(h, BELLA_XDROP_EINVAL, "null colptr");
	XCUDA(cudaSetDevice(h->device));
	return run_batch(h, n_pairs, d_rowids, nullptr, d_posH, d_posV, d_out, d_colptrC, (int)n_cols);
}

int bella_xdrop_align(bella_xdrop* h, uint64_t n_pairs, const uint32_t* rows, const uint32_t* cols,
		const uint16_t* posH, const uint16_t* posV, int32_t* out)
{
	if (!h) return BELLA_XDROP_EINVAL;
	if (n_pairs == 0) { h->launches = 0; return 0; }
	if (!rows || !cols || !posH || !posV || !out) return fail(h, BELLA_XDROP_EINVAL, "null pair arrays");
	if (!h->seqs.p) return fail(h, BELLA_XDROP_EINVAL, "bella_xdrop_set_reads has not been called");
	XCUDA(cudaSetDevice(h->device));
	for (uint64_t p = 0; p < n_pairs; ++p)
		if (rows[p] >= h->n_reads || cols[p] >= h->n_reads) return fail(h, BELLA_XDROP_EINVAL, "read index out of range");
	XCUDA(h->rows.reserve(n_pairs * 4)); XCUDA(h->cols.reserve(n_pairs * 4));
	XCUDA(h->posH.reserve(n_pairs * 2)); XCUDA(h->posV.reserve(n_pairs * 2));
	XCUDA(h->out.reserve(n_pairs * BELLA_XDROP_OUT_FIELDS * sizeof(int32_t)));
	XCUDA(cudaMemcpyAsync(h->rows.p, rows, n_pairs * 4, cudaMemcpyHostToDevice, h->stream));
	XCUDA(cudaMemcpyAsync(h->cols.p, cols, n_pairs * 4, cudaMemcpyHostToDevice, h->stream));
	XCUDA(cudaMemcpyAsync(h->posH.p, posH, n_pairs * 2, cudaMemcpyHostToDevice, h->stream));
	XCUDA(cudaMemcpyAsync(h->posV.p, posV, n_pairs * 2, cudaMemcpyHostToDevice, h->stream));
	const int rc = run_batch(h, n_pairs, (const uint32_t*)h->rows.p, (const uint32_t*)h->cols.p, (const uint16_t*)h->posH.p,
			(const uint16_t*)h->posV.p, (int32_t*)h->out.p);
	if (rc) return rc;
	XCUDA(cudaMemcpyAsync(out, h->out.p, n_pairs * BELLA_XDROP_OUT_FIELDS * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
	return bella_xdrop_sync(h);
}

int bella_xdrop_sync(bella_xdrop* h)
{
	if (!h) return BELLA_XDROP_EINVAL;
	XCUDA(cudaSetDevice(h->device));
	XCUDA(cudaStreamSynchronize(h->stream));
	if (h->ctr.p) {
		int bad = 0;
		XCUDA(cudaMemcpy(&bad, (int*)h->ctr.p + 3, sizeof(int), cudaMemcpyDeviceToHost));
		if (bad) return fail(h, BELLA_XDROP_ESEED, "a seed k-mer does not fit inside its read");
	}
	return 0;
}

int bella_xdrop_get_stats(bella_xdrop* h, double* s)
{
	if (!h || !s) return BELLA_XDROP_EINVAL;
	XCUDA(cudaSetDevice(h->device));
	s[0] = s[1] = 0.0;
	if (h->timed) {
		XCUDA(cudaEventSynchronize(h->ev1));
		float ms = 0.f;
		XCUDA(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
		s[0] = ms;
		int wide = 0;
		XCUDA(cudaMemcpy(&wide, (int*)h->ctr.p + 1, sizeof(int), cudaMemcpyDeviceToHost));
		s[1] = h->used_lanes == 0 ? 0.0 : (double)wide;
	}
	s[2] = h->launches; s[3] = h->used_lanes; s[4] = h->used_cells;
	return 0;
}

void* bella_xdrop_stream(bella_xdrop* h) { return h ? (void*)h->stream : nullptr; }

}  // extern "C"
