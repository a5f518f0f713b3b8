// bella_kmers.cu -- kernels + C-ABI (include/bella_kmers.h) of the "next" row f3.  The per-element logic is kmers.cuh; this file
// launches it around a radix sort and three prefix sums (cub: plumbing).
// STATUS: not yet run on a B200 (see include/bella_kmers.h).
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <thrust/iterator/transform_iterator.h>

#include <string>

#include "bella_kmers.h"
#include "kmers.cuh"

namespace {

#define GRID_STRIDE(i, n) for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < (n); i += (uint64_t)gridDim.x * blockDim.x)

__global__ void k_kmer_extract(uint64_t n, const char* seqs, const uint64_t* seq_off, uint32_t n_reads, int k, uint64_t* key, uint32_t* val)
{
	GRID_STRIDE(g, n) km::extract_one(g, seqs, seq_off, n_reads, k, key, val);
}
__global__ void k_kmer_classify(uint64_t n, const uint64_t* key, int lower, int upper, uint8_t* head, uint8_t* rel)
{
	GRID_STRIDE(i, n) km::classify_one(i, n, key, lower, upper, head, rel);
}
__global__ void k_kmer_place(uint64_t n, const uint32_t* val, const uint8_t* head, const uint8_t* rel, const uint32_t* scan, uint32_t* id_at, uint8_t* strand_at)
{
	GRID_STRIDE(i, n) { km::place_one(i, val, head, rel, scan, id_at); km::strand_one(i, val, strand_at); }
}
__global__ void k_kmer_emit(uint64_t n, const uint32_t* id_at, const uint8_t* strand_at, const uint64_t* slot, const uint64_t* seq_off, uint32_t n_reads,
		uint32_t* t_kmer, uint32_t* t_read, uint16_t* t_pos, uint8_t* t_strand)
{
	GRID_STRIDE(g, n) km::emit_one(g, id_at, strand_at, slot, seq_off, n_reads, t_kmer, t_read, t_pos, t_strand);
}
// one byte per tuple -> one bit per tuple, LSB first (each thread builds one output byte)
__global__ void k_kmer_pack_bits(uint64_t n_tuples, const uint8_t* t_strand, uint8_t* bits)
{
	GRID_STRIDE(b, (n_tuples + 7) / 8) {
		unsigned v = 0;
		for (int j = 0; j < 8; ++j) { const uint64_t t = b * 8 + j; if (t < n_tuples && t_strand[t]) v |= 1u << j; }
		bits[b] = (uint8_t)v;
	}
}

struct IsSet { __host__ __device__ uint64_t operator()(uint32_t id) const { return id != km::NONE ? 1ull : 0ull; } };
struct AsU32 { __host__ __device__ uint32_t operator()(uint8_t f) const { return f; } };

struct Buf {
	void* p = nullptr; size_t cap = 0;
	cudaError_t reserve(size_t bytes)
	{
		if (bytes <= cap) return cudaSuccess;
		if (p) cudaFree(p);
		p = nullptr; cap = 0;
		const cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
		if (e == cudaSuccess) cap = bytes;
		return e;
	}
	void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

}  // namespace

struct bella_kmers {
	int device = 0, sms = 0;
	cudaStream_t stream = nullptr;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	std::string err;
	Buf seqs, seq_off, key_a, key_b, val_a, val_b, head, rel, scan, id_at, strand_at, slot, tmp, t_kmer, t_read, t_pos, t_strand, bits;
	uint64_t n_bases = 0, n_kmers = 0, n_tuples = 0;
	int launches = 0;
	bool counted = false;
};

namespace {

int fail(bella_kmers* h, int code, const std::string& what) { h->err = what; return code; }

#define KCUDA(call)                                                                                                   \
	do {                                                                                                              \
		const cudaError_t e_ = (call);                                                                                \
		if (e_ != cudaSuccess) return fail(h, BELLA_KMERS_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
	} while (0)

}  // namespace

extern "C" {

bella_kmers* bella_kmers_create(int device)
{
	if (cudaSetDevice(device) != cudaSuccess) return nullptr;
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return nullptr;
	bella_kmers* h = new bella_kmers;
	h->device = device; h->sms = prop.multiProcessorCount;
	if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&h->ev0) != cudaSuccess
			|| cudaEventCreate(&h->ev1) != cudaSuccess) { delete h; return nullptr; }
	return h;
}

void bella_kmers_destroy(bella_kmers* h)
{
	if (!h) return;
	cudaSetDevice(h->device);
	cudaStreamSynchronize(h->stream);
	for (Buf* b : {&h->seqs, &h->seq_off, &h->key_a, &h->key_b, &h->val_a, &h->val_b, &h->head, &h->rel, &h->scan, &h->id_at, &h->strand_at,
			&h->slot, &h->tmp, &h->t_kmer, &h->t_read, &h->t_pos, &h->t_strand, &h->bits}) b->release();
	cudaEventDestroy(h->ev0); cudaEventDestroy(h->ev1);
	cudaStreamDestroy(h->stream);
	delete h;
}

const char* bella_kmers_last_error(const bella_kmers* h) { return h ? h->err.c_str() : "null handle"; }

int bella_kmers_count(bella_kmers* h, const char* seqs, const uint64_t* seq_off, uint32_t n_reads, int k, int lower, int upper,
		uint64_t* n_kmers_out, uint64_t* n_tuples_out)
{
	if (!h) return BELLA_KMERS_EINVAL;
	if (!seqs || !seq_off || k < 1 || k > 32 || lower < 1 || upper < lower) return fail(h, BELLA_KMERS_EINVAL, "bad arguments (1 <= k <= 32, 1 <= lower <= upper)");
	for (uint32_t r = 0; r < n_reads; ++r) {
		if (seq_off[r + 1] < seq_off[r]) return fail(h, BELLA_KMERS_EINVAL, "seq_off is not non-decreasing");
		if (seq_off[r + 1] - seq_off[r] > 65535) return fail(h, BELLA_KMERS_EINVAL, "read longer than 65535 bases (positions are unsigned short in BELLA)");
	}
	const uint64_t n = seq_off[n_reads] - seq_off[0];
	if (seq_off[0] != 0) return fail(h, BELLA_KMERS_EINVAL, "seq_off[0] must be 0");
	if (n >= (1ull << 31)) return fail(h, BELLA_KMERS_ERANGE, "more than 2^31 - 1 bases in one call");
	KCUDA(cudaSetDevice(h->device));
	h->counted = false; h->launches = 0; h->n_bases = n; h->n_kmers = h->n_tuples = 0;
	if (n_kmers_out) *n_kmers_out = 0;
	if (n_tuples_out) *n_tuples_out = 0;
	if (n == 0) { h->counted = true; return 0; }
	cudaStream_t st = h->stream;
	KCUDA(h->seqs.reserve(n)); KCUDA(h->seq_off.reserve(((size_t)n_reads + 1) * 8));
	KCUDA(h->key_a.reserve(n * 8)); KCUDA(h->key_b.reserve(n * 8)); KCUDA(h->val_a.reserve(n * 4)); KCUDA(h->val_b.reserve(n * 4));
	KCUDA(h->head.reserve(n)); KCUDA(h->rel.reserve(n)); KCUDA(h->scan.reserve(n * 4));
	KCUDA(h->id_at.reserve(n * 4)); KCUDA(h->strand_at.reserve(n)); KCUDA(h->slot.reserve(n * 8));
	KCUDA(cudaMemcpyAsync(h->seqs.p, seqs, n, cudaMemcpyHostToDevice, st));
	KCUDA(cudaMemcpyAsync(h->seq_off.p, seq_off, ((size_t)n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
	const int grid = h->sms * 8, block = 256;
	KCUDA(cudaEventRecord(h->ev0, st));
	uint64_t* key_a = (uint64_t*)h->key_a.p; uint64_t* key_b = (uint64_t*)h->key_b.p;
	uint32_t* val_a = (uint32_t*)h->val_a.p; uint32_t* val_b = (uint32_t*)h->val_b.p;
	k_kmer_extract<<<grid, block, 0, st>>>(n, (const char*)h->seqs.p, (const uint64_t*)h->seq_off.p, n_reads, k, key_a, val_a);
	KCUDA(cudaGetLastError()); ++h->launches;
	// sort by key; the sentinel ~0 needs all 64 bits, real keys 2k
	size_t tmp = 0;
	KCUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp, key_a, key_b, val_a, val_b, (int64_t)n, 0, 64, st));
	KCUDA(h->tmp.reserve(tmp));
	KCUDA(cub::DeviceRadixSort::SortPairs(h->tmp.p, tmp, key_a, key_b, val_a, val_b, (int64_t)n, 0, 64, st));
	k_kmer_classify<<<grid, block, 0, st>>>(n, key_b, lower, upper, (uint8_t*)h->head.p, (uint8_t*)h->rel.p);
	KCUDA(cudaGetLastError()); ++h->launches;
	{
		auto in = thrust::make_transform_iterator((const uint8_t*)h->head.p, AsU32());
		KCUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, in, (uint32_t*)h->scan.p, (int64_t)n, st));
		KCUDA(h->tmp.reserve(tmp));
		KCUDA(cub::DeviceScan::ExclusiveSum(h->tmp.p, tmp, in, (uint32_t*)h->scan.p, (int64_t)n, st));
	}
	k_kmer_place<<<grid, block, 0, st>>>(n, val_b, (const uint8_t*)h->head.p, (const uint8_t*)h->rel.p, (const uint32_t*)h->scan.p,
			(uint32_t*)h->id_at.p, (uint8_t*)h->strand_at.p);
	KCUDA(cudaGetLastError()); ++h->launches;
	{
		auto in = thrust::make_transform_iterator((const uint32_t*)h->id_at.p, IsSet());
		KCUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, in, (uint64_t*)h->slot.p, (int64_t)n, st));
		KCUDA(h->tmp.reserve(tmp));
		KCUDA(cub::DeviceScan::ExclusiveSum(h->tmp.p, tmp, in, (uint64_t*)h->slot.p, (int64_t)n, st));
	}
	// totals: last element of each scan + its own flag
	uint32_t last_scan = 0, last_id = 0; uint8_t last_head = 0; uint64_t last_slot = 0;
	KCUDA(cudaMemcpyAsync(&last_scan, (uint32_t*)h->scan.p + (n - 1), 4, cudaMemcpyDeviceToHost, st));
	KCUDA(cudaMemcpyAsync(&last_head, (uint8_t*)h->head.p + (n - 1), 1, cudaMemcpyDeviceToHost, st));
	KCUDA(cudaMemcpyAsync(&last_slot, (uint64_t*)h->slot.p + (n - 1), 8, cudaMemcpyDeviceToHost, st));
	KCUDA(cudaMemcpyAsync(&last_id, (uint32_t*)h->id_at.p + (n - 1), 4, cudaMemcpyDeviceToHost, st));
	KCUDA(cudaStreamSynchronize(st));
	h->n_kmers = (uint64_t)last_scan + last_head;
	h->n_tuples = last_slot + (last_id != km::NONE ? 1 : 0);
	const uint64_t nt = h->n_tuples;
	KCUDA(h->t_kmer.reserve(nt * 4)); KCUDA(h->t_read.reserve(nt * 4)); KCUDA(h->t_pos.reserve(nt * 2)); KCUDA(h->t_strand.reserve(nt));
	k_kmer_emit<<<grid, block, 0, st>>>(n, (const uint32_t*)h->id_at.p, (const uint8_t*)h->strand_at.p, (const uint64_t*)h->slot.p,
			(const uint64_t*)h->seq_off.p, n_reads, (uint32_t*)h->t_kmer.p, (uint32_t*)h->t_read.p, (uint16_t*)h->t_pos.p, (uint8_t*)h->t_strand.p);
	KCUDA(cudaGetLastError()); ++h->launches;
	KCUDA(cudaEventRecord(h->ev1, st));
	KCUDA(cudaStreamSynchronize(st));
	h->counted = true;
	if (n_kmers_out) *n_kmers_out = h->n_kmers;
	if (n_tuples_out) *n_tuples_out = h->n_tuples;
	return 0;
}

int bella_kmers_get_tuples(bella_kmers* h, uint32_t* t_kmer, uint32_t* t_read, uint16_t* t_pos, uint8_t* t_strand_bits)
{
	if (!h) return BELLA_KMERS_EINVAL;
	if (!h->counted) return fail(h, BELLA_KMERS_EINVAL, "bella_kmers_count has not run");
	const uint64_t nt = h->n_tuples;
	if (nt == 0) return 0;
	if (!t_kmer || !t_read || !t_pos) return fail(h, BELLA_KMERS_EINVAL, "null tuple arrays");
	KCUDA(cudaSetDevice(h->device));
	cudaStream_t st = h->stream;
	KCUDA(cudaMemcpyAsync(t_kmer, h->t_kmer.p, nt * 4, cudaMemcpyDeviceToHost, st));
	KCUDA(cudaMemcpyAsync(t_read, h->t_read.p, nt * 4, cudaMemcpyDeviceToHost, st));
	KCUDA(cudaMemcpyAsync(t_pos, h->t_pos.p, nt * 2, cudaMemcpyDeviceToHost, st));
	if (t_strand_bits) {
		KCUDA(h->bits.reserve((nt + 7) / 8));
		k_kmer_pack_bits<<<h->sms * 4, 256, 0, st>>>(nt, (const uint8_t*)h->t_strand.p, (uint8_t*)h->bits.p);
		KCUDA(cudaGetLastError());
		KCUDA(cudaMemcpyAsync(t_strand_bits, h->bits.p, (nt + 7) / 8, cudaMemcpyDeviceToHost, st));
	}
	KCUDA(cudaStreamSynchronize(st));
	return 0;
}

int bella_kmers_get_stats(bella_kmers* h, double* s)
{
	if (!h || !s) return BELLA_KMERS_EINVAL;
	s[0] = 0.0; s[1] = (double)h->n_bases; s[2] = h->launches;
	if (h->counted && h->n_bases) {
		KCUDA(cudaSetDevice(h->device));
		float ms = 0.f;
		KCUDA(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
		s[0] = ms;
	}
	return 0;
}

}  // extern "C"
