// ifetch.cu -- how fast can an SM issue straight-line code that does not fit its instruction caches?
// Body = GROUPS groups of 16 independent instructions (ALU-pipe packed min/max, FMA-pipe FFMA, or a
// 50/50 mix), unrolled; the outer loop keeps the total instruction count constant.  16 warps per SM
// (128 threads x 4 CTAs), every warp at its own position after the first few iterations.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ifetch ifetch.cu ; run: ./ifetch
#include <cstdio>
#include <cuda_runtime.h>

template <int GROUPS, int MIX>	// MIX 0: ALU only, 1: FMA only, 2: alternate
__global__ void __launch_bounds__(128, 4) k(unsigned int *out, int iters, unsigned int seed) {
	unsigned int a[8];
	float f[8];
#pragma unroll
	for (int i = 0; i < 8; i++) { a[i] = seed * (threadIdx.x + i + 1); f[i] = (float) (threadIdx.x + i); }
	const float c = 1.0001f, d = 0.5f;
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int g = 0; g < GROUPS; g++) {
#pragma unroll
			for (int i = 0; i < 8; i++) {
				if (MIX == 0) { a[i] = __vmaxu2(a[i], a[(i + 1) & 7] + g); a[(i + 3) & 7] = __vminu2(a[(i + 3) & 7], a[i]); }
				if (MIX == 1) { f[i] = fmaf(f[i], c, d + g); f[(i + 3) & 7] = fmaf(f[(i + 3) & 7], c, f[i]); }
				if (MIX == 2) { a[i] = __vmaxu2(a[i], a[(i + 1) & 7] + g); f[i] = fmaf(f[i], c, d + g); }
			}
		}
	}
	unsigned int r = 0;
#pragma unroll
	for (int i = 0; i < 8; i++) r += a[i] + __float_as_uint(f[i]);
	out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int GROUPS, int MIX>
void run(const char *name, unsigned int *out) {
	const long long total_groups = 1 << 16;
	const int iters = (int) (total_groups / GROUPS);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	int sms = 148;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
	k<GROUPS, MIX><<<sms * 4, 128>>>(out, iters, 12345u);
	cudaEventRecord(e0);
	k<GROUPS, MIX><<<sms * 4, 128>>>(out, iters, 12345u);
	cudaEventRecord(e1);
	cudaEventSynchronize(e1);
	float ms = 0;
	cudaEventElapsedTime(&ms, e0, e1);
	int khz = 0;
	cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
	// per SMSP: 4 warps x total_groups x 16 instructions
	const double instr = 4.0 * total_groups * 16.0;
	printf("%-10s body %7d instr (%6.0f KB): %8.3f ms  -> %.3f warp-instr / clk / SMSP (at %d MHz nominal)\n", name, GROUPS * 16,
			GROUPS * 16 * 16 / 1024.0, ms, instr / (ms * 1e-3 * khz * 1e3), khz / 1000);
}

int main() {
	unsigned int *out;
	cudaMalloc(&out, 148 * 4 * 128 * 4 * 2);
#define ALL(MIX, name) \
	run<4, MIX>(name, out); run<64, MIX>(name, out); run<128, MIX>(name, out); run<256, MIX>(name, out); \
	run<512, MIX>(name, out); run<1024, MIX>(name, out); run<4096, MIX>(name, out);
	ALL(0, "alu")
	ALL(1, "fma")
	ALL(2, "alu+fma")
	return 0;
}
