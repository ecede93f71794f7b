// pipes.cu -- which instruction classes of the stacking kernel compete for the same issue resources on a B200 SM?
// Every body is 256 independent-chain instructions of class A, of class B, or of A and B alternating (512), looped;
// 16 warps per SM (128 threads x 4 CTAs per SM), as the register-network kernel runs.  Printed: warp-instructions
// per clock and scheduler.  If A+B alternating reaches the SUM of the two rates the classes issue independently; if it
// stays at the rate of one of them they share a dispatch resource.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run: ./pipes
#include <cstdio>
#include <cuda_runtime.h>

enum { VMNMX = 0, IMADW, IDP2A, DFMA_, I2F64, LDS32, IMAD32, DADD_, IMADW2, NONE };
static const char *names[] = {"VIMNMX.U16x2", "IMAD.WIDE", "IDP.2A", "DFMA", "I2F.F64.U16", "LDS", "IMAD", "DADD", "IMAD.WIDE(x,x,q)", "-"};

template <int OP>
__device__ __forceinline__ void op(unsigned int (&a)[8], unsigned long long (&q)[8], double (&d)[8], unsigned int (&p)[8], int i) {
	if (OP == VMNMX) a[i] = __vminu2(a[i], a[(i + 1) & 7]);
	if (OP == IMADW) asm volatile("{ .reg .b32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, lo, %1, %0; }" : "+l"(q[i]) : "r"(a[i]));
	if (OP == IDP2A) a[i] = __dp2a_lo(a[(i + 1) & 7], 0x0001u, a[i]);
	if (OP == DFMA_) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(d[(i + 1) & 7]), "d"(d[(i + 2) & 7]));
	if (OP == DADD_) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(d[(i + 1) & 7]));
	if (OP == I2F64) asm volatile("{ .reg .b32 lo, hi; .reg .b16 h; mov.b64 {lo, hi}, %0; cvt.u16.u32 h, lo; cvt.rn.f64.u16 %0, h; }" : "+d"(d[i]));
	if (OP == LDS32) asm volatile("ld.shared.u32 %0, [%0];" : "+r"(p[i]));	// pointer chase (every word holds its own address)
	if (OP == IMAD32) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(a[(i + 1) & 7]), "r"(a[(i + 2) & 7]));
	if (OP == IMADW2) asm volatile("mad.wide.u32 %0, %1, %1, %0;" : "+l"(q[i]) : "r"(a[i]));	// (only meaningful next to an op that changes a[])
}

template <int A, int B>
__global__ void __launch_bounds__(128, 4) k(unsigned int *out, int iters, unsigned int seed) {
	__shared__ unsigned int sh[1024];
	for (int i = threadIdx.x; i < 1024; i += 128) sh[i] = (unsigned int) __cvta_generic_to_shared(sh + i);
	__syncthreads();
	unsigned int a[8], p[8];
	unsigned long long q[8];
	double d[8];
#pragma unroll
	for (int i = 0; i < 8; i++) { a[i] = seed * (threadIdx.x + i + 1); q[i] = a[i]; d[i] = 1.0 + 1e-9 * (threadIdx.x + i); p[i] = (unsigned int) __cvta_generic_to_shared(sh + threadIdx.x + i * 128); }
#pragma unroll 1
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int g = 0; g < 32; g++) {
#pragma unroll
			for (int i = 0; i < 8; i++) {
				op<A>(a, q, d, p, i);
				if (B != NONE) op<B>(a, q, d, p, i);
			}
		}
	}
	unsigned int r = 0;
#pragma unroll
	for (int i = 0; i < 8; i++) r += a[i] + p[i] + (unsigned int) q[i] + (unsigned int) (q[i] >> 32) + (unsigned int) __double2loint(d[i]);
	out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int A, int B>
void run(unsigned int *out) {
	const int iters = 2000;
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	int sms = 148, khz = 0;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
	cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
	k<A, B><<<sms * 4, 128>>>(out, iters, 12345u);
	cudaEventRecord(e0);
	k<A, B><<<sms * 4, 128>>>(out, iters, 12345u);
	cudaEventRecord(e1);
	cudaEventSynchronize(e1);
	float ms = 0;
	cudaEventElapsedTime(&ms, e0, e1);
	const double instr = 4.0 * iters * 256.0 * (B == NONE ? 1 : 2);	// per scheduler: 4 warps
	printf("%-14s + %-14s %8.3f ms  -> %.3f warp-instr / clk / SMSP (at %d MHz nominal)\n", names[A], names[B], ms,
			instr / (ms * 1e-3 * khz * 1e3), khz / 1000);
}

int main() {
	unsigned int *out;
	cudaMalloc(&out, 148 * 4 * 128 * 4 * 2);
	run<VMNMX, NONE>(out); run<IMADW, NONE>(out); run<IDP2A, NONE>(out); run<IMAD32, NONE>(out); 
	run<DFMA_, NONE>(out); run<DADD_, NONE>(out); run<I2F64, NONE>(out); run<LDS32, NONE>(out);
	run<VMNMX, IMADW>(out); run<VMNMX, IDP2A>(out); run<VMNMX, IMAD32>(out); run<VMNMX, IMADW2>(out); run<IDP2A, IMADW2>(out);
	run<VMNMX, DFMA_>(out); run<VMNMX, DADD_>(out); run<VMNMX, I2F64>(out); run<VMNMX, LDS32>(out);
	run<IMADW, DFMA_>(out); run<IMADW, IDP2A>(out); run<DFMA_, I2F64>(out); run<DFMA_, LDS32>(out);
	return 0;
}
