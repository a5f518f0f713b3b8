// microbenchmark: shared-memory primitives the group / bucket kernels are built from (B200)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/smem_bench tools/smem_bench.cu && gpurun_out/smem_bench
// Every kernel: 148*4 CTAs x 256 threads, each thread does N operations on a table of TBL words in shared memory at
// pseudo-random addresses; reported: lane-operations per clock per SM (at the measured kernel time and 1.9 GHz nominal).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t hash(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
constexpr int TBL = 2048;
constexpr int N = 2048;

template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* sink, uint32_t seed)
{
	__shared__ uint32_t t[TBL];
	for (int i = threadIdx.x; i < TBL; i += 256) t[i] = 0;
	__syncthreads();
	uint32_t acc = 0, x = hash(seed + blockIdx.x * 256 + threadIdx.x);
	const uint32_t lane = threadIdx.x & 31;
#pragma unroll 4
	for (int i = 0; i < N; ++i) {
		x = x * 1664525u + 1013904223u;
		const uint32_t a = (x >> 10) & (TBL - 1);
		if (MODE == 0) acc += atomicAdd(&t[a], 1u);                       // ATOMS.ADD with return
		else if (MODE == 1) atomicAdd(&t[a], 1u);                        // no return
		else if (MODE == 2) atomicOr(&t[a], 1u << (x & 31));             // OR, no return
		else if (MODE == 3) acc += t[a];                                 // plain LDS (random banks)
		else if (MODE == 4) t[a] = x;                                    // plain STS (random banks)
		else if (MODE == 5) acc += __match_any_sync(0xFFFFFFFFu, a & 63); // MATCH.ANY
		else if (MODE == 6) { uint32_t v = t[a]; t[a] = v + 1; acc += v; } // LDS + STS (racy, cost only)
		else if (MODE == 7) acc += atomicAdd(&t[(a & ~31u) | lane], 1u);  // with return, conflict-free banks
		else if (MODE == 8) acc += atomicAdd(&t[a & 7], 1u);              // with return, 8 hot addresses
		else if (MODE == 9) acc += __popc(__ballot_sync(0xFFFFFFFFu, a & 1)); // ballot
		else if (MODE == 10) acc += __shfl_xor_sync(0xFFFFFFFFu, a, 5);   // shuffle
		else if (MODE == 11) acc += __reduce_add_sync(0xFFFFFFFFu, a);    // redux
		else if (MODE == 12) acc += atomicMax(&t[a], x);                  // ATOMS.MAX with return
		else if (MODE == 13) acc += atomicExch(&t[a], x);                 // ATOMS.EXCH
		else if (MODE == 14) acc += atomicCAS(&t[a], 0u, x);              // ATOMS.CAS
	}
	__syncthreads();
	if (threadIdx.x < 32) acc += t[threadIdx.x];
	if (acc == 0xdeadbeef) *sink = acc;
}
template <int MODE> void run(const char* name, uint32_t* sink)
{
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	float ms = 0;
	for (int it = 0; it < 3; ++it) {
		cudaEventRecord(e0);
		k<MODE><<<148 * 4, 256>>>(sink, it);
		cudaEventRecord(e1); cudaEventSynchronize(e1);
		cudaEventElapsedTime(&ms, e0, e1);
	}
	const double ops = 148.0 * 4 * 256 * N;
	printf("%-44s %.3f ms  %.1f Gop/s  %.2f lane-ops/clk/SM  (%.1f clk per warp-op per SM)\n", name, ms, ops / ms / 1e6,
		ops / ms / 1e6 / 1.9 / 148, 32.0 / (ops / ms / 1e6 / 1.9 / 148));
}
int main()
{
	uint32_t* sink; cudaMalloc(&sink, 4);
	run<0>("ATOMS.ADD return, random addr", sink);
	run<1>("ATOMS.ADD no return, random addr", sink);
	run<2>("ATOMS.OR no return, random addr", sink);
	run<7>("ATOMS.ADD return, conflict-free banks", sink);
	run<8>("ATOMS.ADD return, 8 hot addresses", sink);
	run<12>("ATOMS.MAX return, random", sink);
	run<13>("ATOMS.EXCH return, random", sink);
	run<14>("ATOMS.CAS return, random", sink);
	run<3>("LDS random", sink);
	run<4>("STS random", sink);
	run<6>("LDS+STS random", sink);
	run<5>("MATCH.ANY (6-bit keys)", sink);
	run<9>("BALLOT+POPC", sink);
	run<10>("SHFL", sink);
	run<11>("REDUX.ADD", sink);
	return 0;
}
