// ssc_sum.cu -- K1 / K1b: stack_summing (src/stacking/stacking.c:196-355); K1c: stack_addmax / stack_addmin.
//
// K1 accumulates the per-pixel sums of all frames in registers (uint32 is
// exact for up to 65537 frames of WORDs; the reference uses unsigned long),
// writes them once and folds the global maximum (stacking.c:311-313) through a
// warp/block reduction into one atomicMax per block.  K1b is the write pass
// (:328-341): round_to_WORD(sum * 65535/max) when the max exceeds 16 bits.
// Between the two sits the only exchange of this operator when the image is
// split over GPUs: an all-reduce(MAX) of one uint64.
#include "ssc_device.cuh"
#include "ssc_kernels.h"

namespace ssc {

// Fast path: no registration data.  Each thread owns 8 consecutive pixels of a
// row (one 16-byte load per frame); the source index 0 of every layer is never
// read (the reference's "ii > 0" test, stacking.c:307), so pixel 0 of each
// layer stays 0.
__global__ void k_sum_rows_vec8(ssc_band b, uint32_t *__restrict__ acc, unsigned long long *maxim) {
	const int W = b.width, N = b.n_frames;
	const int vec_per_row = W >> 3;
	unsigned long long total = (unsigned long long) b.layers * b.out_rows * vec_per_row;
	unsigned int local_max = 0;
	for (unsigned long long g = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; g < total;
			g += (unsigned long long) gridDim.x * blockDim.x) {
		int vx = (int) (g % vec_per_row);
		int rr = (int) ((g / vec_per_row) % b.out_rows);
		int layer = (int) (g / ((unsigned long long) vec_per_row * b.out_rows));
		int R = b.out_row0 + rr;	// sum stacking works in fits memory order: for SER the frames
		int srow = b.row_order == SSC_ROWS_SER ? b.height - 1 - R : R;	// were flipped by ser_read_frame
		const WORD *src = b.frames + (size_t) layer * b.layer_stride + (size_t) (srow - b.buf_row0) * W + vx * 8;
		unsigned int s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 4
		for (int f = 0; f < N; f++) {
			uint4 v = __ldg(reinterpret_cast<const uint4 *>(src + (size_t) f * b.frame_stride));
			s[0] += v.x & 0xFFFFu; s[1] += v.x >> 16;
			s[2] += v.y & 0xFFFFu; s[3] += v.y >> 16;
			s[4] += v.z & 0xFFFFu; s[5] += v.z >> 16;
			s[6] += v.w & 0xFFFFu; s[7] += v.w >> 16;
		}
		if (R == 0 && vx == 0) s[0] = 0;	// source index 0 (row 0, column 0 of the fits buffer)
		uint32_t *dst = acc + ((size_t) layer * b.out_rows + rr) * W + vx * 8;
		reinterpret_cast<uint4 *>(dst)[0] = make_uint4(s[0], s[1], s[2], s[3]);
		reinterpret_cast<uint4 *>(dst)[1] = make_uint4(s[4], s[5], s[6], s[7]);
#pragma unroll
		for (int i = 0; i < 8; i++) local_max = max(local_max, s[i]);
	}
	for (int o = 16; o > 0; o >>= 1) local_max = max(local_max, __shfl_xor_sync(0xffffffffu, local_max, o));
	__shared__ unsigned int wmax[32];
	if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = local_max;
	__syncthreads();
	if (threadIdx.x < 32) {
		unsigned int m = threadIdx.x < (blockDim.x >> 5) ? wmax[threadIdx.x] : 0;
		for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
		if (threadIdx.x == 0 && m) atomicMax(maxim, (unsigned long long) m);
	}
}

// General path: registration shifts (stacking.c:283-289,299-319), any width.
__global__ void k_sum_generic(ssc_band b, uint32_t *__restrict__ acc, unsigned long long *maxim) {
	const int W = b.width, H = b.height, N = b.n_frames;
	unsigned long long total = (unsigned long long) b.layers * b.out_rows * W;
	unsigned int local_max = 0;
	for (unsigned long long g = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; g < total;
			g += (unsigned long long) gridDim.x * blockDim.x) {
		int x = (int) (g % W);
		int rr = (int) ((g / W) % b.out_rows);
		int layer = (int) (g / ((unsigned long long) W * b.out_rows));
		int y = b.out_row0 + rr;	// row in the fits buffers the reference loops over
		const WORD *lbase = b.frames + (size_t) layer * b.layer_stride;
		unsigned int s = 0;
		for (int f = 0; f < N; f++) {
			int nx = x - (b.shiftx ? b.shiftx[f] : 0);
			int ny = y - (b.shifty ? b.shifty[f] : 0);
			if (nx < 0 || nx >= W || ny < 0 || ny >= H) continue;
			if (ny == 0 && nx == 0) continue;	// ii > 0
			int srow = b.row_order == SSC_ROWS_SER ? H - 1 - ny : ny;
			s += lbase[(size_t) f * b.frame_stride + (size_t) (srow - b.buf_row0) * W + nx];
		}
		acc[((size_t) layer * b.out_rows + rr) * W + x] = s;
		local_max = max(local_max, s);
	}
	for (int o = 16; o > 0; o >>= 1) local_max = max(local_max, __shfl_xor_sync(0xffffffffu, local_max, o));
	if ((threadIdx.x & 31) == 0 && local_max) atomicMax(maxim, (unsigned long long) local_max);
}

// K1b, stacking.c:328-341
__global__ void k_sum_finish(const uint32_t *__restrict__ acc, size_t count,
		const unsigned long long *__restrict__ maxim, WORD *__restrict__ out) {
	const unsigned long long mx = *maxim;
	const bool rescale = mx > 65535ull;
	const double ratio = rescale ? ddiv(65535.0, __ull2double_rn(mx)) : 1.0;
	for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t) gridDim.x * blockDim.x) {
		double v = (double) acc[i];
		out[i] = (WORD) round_to_word(rescale ? dmul(v, ratio) : v);
	}
}

// ---------------------------------------------------------------------------
// K1c: stack_addmax / stack_addmin (src/stacking/stacking.c:824-972, 979-1128): the brightest or
// darkest sample of each pixel, one pass, the result written as WORDs directly.  Same gather as
// the sum (shifts, "ii > 0": source index 0 is never taken, so without shifts pixel 0 of each
// layer keeps the initial 0 / 0xFFFF).  `extreme` folds the max (min) of the result, which is what
// the reference leaves in gfit.hi (:947, :1103); untouched pixels hold the neutral element.
// ---------------------------------------------------------------------------
template<bool TAKE_MAX>
__device__ __forceinline__ unsigned int pick2(unsigned int a, unsigned int b) {	// two packed WORDs
	return TAKE_MAX ? __vmaxu2(a, b) : __vminu2(a, b);
}

template<bool TAKE_MAX>
__device__ __forceinline__ void fold_extreme(unsigned int local, unsigned int *extreme) {
	for (int o = 16; o > 0; o >>= 1) {
		const unsigned int other = __shfl_xor_sync(0xffffffffu, local, o);
		local = TAKE_MAX ? max(local, other) : min(local, other);
	}
	if ((threadIdx.x & 31) == 0) {
		if (TAKE_MAX) atomicMax(extreme, local);
		else atomicMin(extreme, local);
	}
}

template<bool TAKE_MAX>
__global__ void k_extreme_rows_vec8(ssc_band b, WORD *__restrict__ out, unsigned int *extreme) {
	const int W = b.width, N = b.n_frames;
	const int vec_per_row = W >> 3;
	const unsigned int init = TAKE_MAX ? 0u : 0xFFFFFFFFu;
	unsigned long long total = (unsigned long long) b.layers * b.out_rows * vec_per_row;
	unsigned int local = TAKE_MAX ? 0u : 0xFFFFu;
	for (unsigned long long g = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; g < total;
			g += (unsigned long long) gridDim.x * blockDim.x) {
		int vx = (int) (g % vec_per_row);
		int rr = (int) ((g / vec_per_row) % b.out_rows);
		int layer = (int) (g / ((unsigned long long) vec_per_row * b.out_rows));
		int R = b.out_row0 + rr;
		int srow = b.row_order == SSC_ROWS_SER ? b.height - 1 - R : R;
		const WORD *src = b.frames + (size_t) layer * b.layer_stride + (size_t) (srow - b.buf_row0) * W + vx * 8;
		uint4 m = make_uint4(init, init, init, init);
#pragma unroll 4
		for (int f = 0; f < N; f++) {
			const uint4 v = __ldg(reinterpret_cast<const uint4 *>(src + (size_t) f * b.frame_stride));
			m.x = pick2<TAKE_MAX>(m.x, v.x); m.y = pick2<TAKE_MAX>(m.y, v.y);
			m.z = pick2<TAKE_MAX>(m.z, v.z); m.w = pick2<TAKE_MAX>(m.w, v.w);
		}
		if (R == 0 && vx == 0) m.x = (m.x & 0xFFFF0000u) | (init & 0xFFFFu);	// source index 0 is never taken
		*reinterpret_cast<uint4 *>(out + ((size_t) layer * b.out_rows + rr) * W + vx * 8) = m;
		const unsigned int e = pick2<TAKE_MAX>(pick2<TAKE_MAX>(m.x, m.y), pick2<TAKE_MAX>(m.z, m.w));
		const unsigned int e1 = TAKE_MAX ? max(e & 0xFFFFu, e >> 16) : min(e & 0xFFFFu, e >> 16);
		local = TAKE_MAX ? max(local, e1) : min(local, e1);
	}
	fold_extreme<TAKE_MAX>(local, extreme);
}

template<bool TAKE_MAX>
__global__ void k_extreme_generic(ssc_band b, WORD *__restrict__ out, unsigned int *extreme) {
	const int W = b.width, H = b.height, N = b.n_frames;
	unsigned long long total = (unsigned long long) b.layers * b.out_rows * W;
	unsigned int local = TAKE_MAX ? 0u : 0xFFFFu;
	for (unsigned long long g = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; g < total;
			g += (unsigned long long) gridDim.x * blockDim.x) {
		int x = (int) (g % W);
		int rr = (int) ((g / W) % b.out_rows);
		int layer = (int) (g / ((unsigned long long) W * b.out_rows));
		int y = b.out_row0 + rr;
		const WORD *lbase = b.frames + (size_t) layer * b.layer_stride;
		unsigned int m = TAKE_MAX ? 0u : 0xFFFFu;
		for (int f = 0; f < N; f++) {
			int nx = x - (b.shiftx ? b.shiftx[f] : 0);
			int ny = y - (b.shifty ? b.shifty[f] : 0);
			if (nx < 0 || nx >= W || ny < 0 || ny >= H) continue;
			if (ny == 0 && nx == 0) continue;	// ii > 0
			int srow = b.row_order == SSC_ROWS_SER ? H - 1 - ny : ny;
			const unsigned int v = lbase[(size_t) f * b.frame_stride + (size_t) (srow - b.buf_row0) * W + nx];
			m = TAKE_MAX ? max(m, v) : min(m, v);
		}
		out[((size_t) layer * b.out_rows + rr) * W + x] = (WORD) m;
		local = TAKE_MAX ? max(local, m) : min(local, m);
	}
	fold_extreme<TAKE_MAX>(local, extreme);
}

static int sm_count() {
	int dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	return sms;
}

int launch_sum_accumulate(const ssc_band &b, uint32_t *acc, unsigned long long *maxim, cudaStream_t stream) {
	if (b.n_frames < 2) return SSC_ERR_ARG;
	const bool aligned = (b.width % 8 == 0) && (b.frame_stride % 8 == 0) && (b.layer_stride % 8 == 0)
			&& ((uintptr_t) b.frames % 16 == 0) && ((uintptr_t) acc % 16 == 0);
	const int block = 256;
	if (!b.shiftx && !b.shifty && aligned) {
		unsigned long long total = (unsigned long long) b.layers * b.out_rows * (b.width / 8);
		unsigned long long want = (total + block - 1) / block;
		unsigned int grid = (unsigned int) (want < (unsigned long long) sm_count() * 16 ? want : sm_count() * 16);
		if (grid == 0) return SSC_OK;
		k_sum_rows_vec8<<<grid, block, 0, stream>>>(b, acc, maxim);
	} else {
		unsigned long long total = (unsigned long long) b.layers * b.out_rows * b.width;
		unsigned long long want = (total + block - 1) / block;
		unsigned int grid = (unsigned int) (want < (unsigned long long) sm_count() * 16 ? want : sm_count() * 16);
		if (grid == 0) return SSC_OK;
		k_sum_generic<<<grid, block, 0, stream>>>(b, acc, maxim);
	}
	return cudaGetLastError() == cudaSuccess ? SSC_OK : SSC_ERR_CUDA;
}

// *extreme must hold 0 (max) or 0xFFFF (min) before the first band of an image is launched
int launch_extreme(const ssc_band &b, bool take_max, WORD *out, unsigned int *extreme, cudaStream_t stream) {
	if (b.n_frames < 2) return SSC_ERR_ARG;
	const bool aligned = (b.width % 8 == 0) && (b.frame_stride % 8 == 0) && (b.layer_stride % 8 == 0)
			&& ((uintptr_t) b.frames % 16 == 0) && ((uintptr_t) out % 16 == 0);
	const int block = 256;
	const bool fast = !b.shiftx && !b.shifty && aligned;
	unsigned long long total = (unsigned long long) b.layers * b.out_rows * (fast ? b.width / 8 : b.width);
	unsigned long long want = (total + block - 1) / block;
	unsigned int grid = (unsigned int) (want < (unsigned long long) sm_count() * 16 ? want : sm_count() * 16);
	if (grid == 0) return SSC_OK;
	if (fast) {
		if (take_max) k_extreme_rows_vec8<true><<<grid, block, 0, stream>>>(b, out, extreme);
		else k_extreme_rows_vec8<false><<<grid, block, 0, stream>>>(b, out, extreme);
	} else {
		if (take_max) k_extreme_generic<true><<<grid, block, 0, stream>>>(b, out, extreme);
		else k_extreme_generic<false><<<grid, block, 0, stream>>>(b, out, extreme);
	}
	return cudaGetLastError() == cudaSuccess ? SSC_OK : SSC_ERR_CUDA;
}

int launch_sum_finish(const uint32_t *acc, size_t count, const unsigned long long *maxim, WORD *out,
		cudaStream_t stream) {
	if (count == 0) return SSC_OK;
	const int block = 256;
	size_t want = (count + block - 1) / block;
	unsigned int grid = (unsigned int) (want < (size_t) sm_count() * 16 ? want : (size_t) sm_count() * 16);
	k_sum_finish<<<grid, block, 0, stream>>>(acc, count, maxim, out);
	return cudaGetLastError() == cudaSuccess ? SSC_OK : SSC_ERR_CUDA;
}

}  // namespace ssc
