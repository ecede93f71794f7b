// Micro-benchmarks behind DESIGN.md's pipe-throughput numbers (B200, sm_100a):
//   1. VIMNMX.U16x2 compare-exchange throughput with 64 packed registers (the sorting network alone)
//   2. the same with 32-bit VIMNMX (one pixel per register)
//   3. shared-memory atomic histogram update rate (K4's inner loop)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sortbench sortbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../siril-0.9_b200/csrc/ssc_sortnet.h"

template <int MODE, int TPB, int MINB>
__global__ void __launch_bounds__(TPB, MINB) k_sort(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, int iters) {
	uint32_t v[64];
	size_t g = (size_t) blockIdx.x * TPB + threadIdx.x;
#pragma unroll
	for (int f = 0; f < 64; f++) v[f] = in[(g * 64 + f) & 0xFFFFF] * (f * 2654435761u + 1u);
	for (int it = 0; it < iters; it++) {
#define SSC_CE(a, b) { uint32_t lo_ = MODE == 0 ? __vminu2(v[a], v[b]) : min(v[a], v[b]); \
	v[b] = MODE == 0 ? __vmaxu2(v[a], v[b]) : max(v[a], v[b]); v[a] = lo_; }
		SSC_SORTNET_64
#undef SSC_CE
#pragma unroll
		for (int f = 0; f < 64; f++) v[f] = v[f] * 1664525u + 1013904223u + it;   // scramble (IMAD: fma pipe)
	}
	uint32_t x = 0;
#pragma unroll
	for (int f = 0; f < 64; f++) x ^= v[f];
	out[g] = x;
}

__global__ void k_hist(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, int n_per_thread, int spread) {
	extern __shared__ uint32_t sh[];
	for (int i = threadIdx.x; i < 32768; i += blockDim.x) sh[i] = 0;
	__syncthreads();
	uint32_t x = in[(blockIdx.x * blockDim.x + threadIdx.x) & 0xFFFFF];
	for (int i = 0; i < n_per_thread; i++) {
		x = x * 1664525u + 1013904223u;
		uint32_t a = (x >> 12) % spread;               // value in [0, spread)
		atomicAdd(&sh[a >> 1], 1u << ((a & 1) * 16));
	}
	__syncthreads();
	uint32_t s = 0;
	for (int i = threadIdx.x; i < 32768; i += blockDim.x) s += sh[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F> float time_ms(F f) {
	cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
	f(); cudaDeviceSynchronize();
	cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
	float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main() {
	uint32_t *in, *out;
	cudaMalloc(&in, 4u << 20); cudaMalloc(&out, 64u << 20);
	cudaMemset(in, 0x5a, 4u << 20);
	const int iters = 20, blocks = 148 * 64;
	const double ce = 543.0 * iters;
	{
		float ms = time_ms([&] { k_sort<0, 128, 4><<<blocks, 128>>>(in, out, iters); });
		double warps = blocks * 4.0;
		printf("u16x2 network, 128 thr x4/SM: %.3f ms  -> %.2f cycles per CE-pair(2 VIMNMX) per SMSP-warp @1.92GHz (4 warps/SMSP)\n", ms,
			ms * 1e-3 * 1.92e9 / (warps / (148 * 4) * ce));
	}
	{
		float ms = time_ms([&] { k_sort<0, 256, 2><<<blocks / 2, 256>>>(in, out, iters); });
		printf("u16x2 network, 256 thr x2/SM: %.3f ms\n", ms);
	}
	{
		float ms = time_ms([&] { k_sort<0, 128, 6><<<blocks, 128>>>(in, out, iters); });
		printf("u16x2 network, 128 thr x6/SM (80 regs): %.3f ms\n", ms);
	}
	{
		float ms = time_ms([&] { k_sort<1, 128, 4><<<blocks, 128>>>(in, out, iters); });
		printf("u32 network, 128 thr x4/SM: %.3f ms\n", ms);
	}
	cudaFuncSetAttribute(k_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
	for (int spread : {65536, 4096, 256, 64, 1}) {
		const int npt = 256;
		float ms = time_ms([&] { k_hist<<<148, 1024, 131072>>>(in, out, npt, spread); });
		printf("smem atomic histogram, values spread over %5d bins: %.3f ms -> %.2f updates/clk/SM\n", spread, ms,
			148.0 * 1024 * npt / (ms * 1e-3 * 1.92e9) / 148);
	}
	printf("%s\n", cudaGetErrorString(cudaGetLastError()));
	return 0;
}
