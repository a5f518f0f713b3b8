// microbenchmark: random global atomics on B200
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t hash(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
template <int MODE, int ILP, int STRIDE = 1>
__global__ void k(uint32_t* cnt, unsigned long long* cnt64, uint32_t naddr, uint64_t n, uint32_t* sink, uint4* out)
{
	uint32_t acc = 0;
	for (uint64_t j = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * ILP; j < n; j += (uint64_t)gridDim.x * blockDim.x * ILP) {
		uint32_t q[ILP];
#pragma unroll
		for (int u = 0; u < ILP; ++u) {
			uint32_t a = (hash((uint32_t)(j + u)) % naddr) * STRIDE;
			if (MODE == 0) q[u] = atomicAdd(&cnt[a], 1u);                 // u32 with return
			else if (MODE == 1) { atomicAdd(&cnt[a], 1u); q[u] = a; }      // u32 no return (RED)
			else if (MODE == 2) q[u] = (uint32_t)atomicAdd(&cnt64[a], 1ull); // u64 with return
			else if (MODE == 3) { q[u] = a; }                               // no atomic, store only
		}
#pragma unroll
		for (int u = 0; u < ILP; ++u) {
			if (MODE == 3 || MODE == 4) out[(size_t)(hash((uint32_t)(j + u)) % naddr) * 64 + ((j + u) & 63)] = make_uint4(q[u], 0, 0, 0);
			acc += q[u];
		}
	}
	if (acc == 0xdeadbeef) *sink = acc;
}
template <int MODE, int ILP, int STRIDE = 1> void run(const char* name, uint32_t naddr, uint64_t n)
{
	uint32_t* cnt; unsigned long long* c64; uint32_t* sink; uint4* out;
	cudaMalloc(&cnt, (size_t)naddr * 4 * STRIDE); cudaMalloc(&c64, (size_t)naddr * 8 * STRIDE); cudaMalloc(&sink, 4); cudaMalloc(&out, (size_t)naddr * 64 * 16);
	cudaMemset(cnt, 0, (size_t)naddr * 4 * STRIDE); cudaMemset(c64, 0, (size_t)naddr * 8 * STRIDE);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	for (int it = 0; it < 3; ++it) {
		cudaEventRecord(e0);
		k<MODE, ILP, STRIDE><<<148 * 16, 256>>>(cnt, c64, naddr, n, sink, out);
		cudaEventRecord(e1); cudaEventSynchronize(e1);
	}
	float ms; cudaEventElapsedTime(&ms, e0, e1);
	printf("%-28s naddr=%8u ilp=%d  %.3f ms  %.1f G/s  %.1f per clk@1.9GHz\n", name, naddr, ILP, ms, n / ms / 1e6, n / ms / 1e6 / 1.9);
	cudaFree(cnt); cudaFree(c64); cudaFree(sink); cudaFree(out);
}
int main()
{
	const uint64_t n = 71000000;
	for (uint32_t naddr : {29000u, 50000u, 1000000u}) {
		run<0, 4>("atom u32 return", naddr, n);
		run<0, 8>("atom u32 return", naddr, n);
		run<1, 4>("red  u32 noreturn", naddr, n);
		run<2, 4>("atom u64 return", naddr, n);
	}
	run<3, 4>("scattered 16B store only", 29000u, n);
	run<0, 4, 8>("atom u32 ret stride 32B", 29000u, n);
	run<0, 4, 32>("atom u32 ret stride 128B", 29000u, n);
	run<0, 4, 64>("atom u32 ret stride 256B", 29000u, n);
	run<2, 4, 16>("atom u64 ret stride 128B", 50000u, n);
	run<1, 4, 32>("red u32 stride 128B", 50000u, n);
	return 0;
}
