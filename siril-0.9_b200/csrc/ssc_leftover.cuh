// ssc_leftover.cuh -- the leftover list shared by the register-network kernels
// (ssc_stack_fast.cu, ssc_stack_pipe.cu) and the kernel that finishes it.
#pragma once
#include "ssc_column.cuh"
#include "ssc_kernels.h"

namespace ssc {

// ---------------------------------------------------------------------------------------------
// Leftover list: pixels the register-resident rejection could not finish (more rejections than
// its window holds, or more passes than it unrolls) are appended -- sorted column, result
// address, layer -- to a device-global list and finished by k_stack_leftover, one pixel per
// thread with full warps.  A typical deep-sky stack parks well under 1 % of its pixels here.
// ---------------------------------------------------------------------------------------------
struct LeftoverList {
	unsigned int *count;		// items appended (may exceed capacity: the excess was handled in place)
	unsigned int capacity;
	uint4 *cols;			// [capacity][NP * 2 bytes]: the sorted column of each item, contiguous
	uint2 *meta;			// [capacity]: result element index, layer
	unsigned int *ids;		// or (ssc_stack_fast.cu): [capacity] indices of pixels for the general kernel's list mode
};

// append one pixel index per flagged lane: one atomic per warp
__device__ __forceinline__ void leftover_push(const LeftoverList &list, bool flagged, unsigned int pixel, int lane) {
	const unsigned int mask = __ballot_sync(0xffffffffu, flagged);
	if (mask == 0) return;
	const int leader = __ffs(mask) - 1;
	unsigned int base = 0;
	if (lane == leader) base = atomicAdd(list.count, (unsigned int) __popc(mask));
	base = __shfl_sync(0xffffffffu, base, leader);
	if (flagged) {
		const unsigned int slot = base + __popc(mask & ((1u << lane) - 1));
		if (slot < list.capacity) list.ids[slot] = pixel;	// capacity = every pixel of the launch: always true
	}
}

// The leftover list, one pixel per thread: the column is already sorted, the sums are rebuilt.
template <int NP>
__global__ void __launch_bounds__(128)
k_stack_leftover(StackParams p, LeftoverList list) {
	__shared__ WORD col[NP * 128];
	__shared__ double s_unbias[65], s_inv_n[65], s_inv_n2[65];
	__shared__ unsigned long long blk_counters[SSC_COUNTERS];
	const int tid = threadIdx.x;
	if (tid < SSC_COUNTERS) blk_counters[tid] = 0;
	if (tid >= 1 && tid <= 64) {
		s_unbias[tid] = tid > 1 ? ddiv((double) tid, (double) (tid - 1)) : 0.0;
		s_inv_n[tid] = ddiv(1.0, (double) tid);
		s_inv_n2[tid] = ddiv(1.0, (double) (tid * tid));
	}
	const SmallN tables{s_unbias, s_inv_n, s_inv_n2};
	__syncthreads();
	unsigned int n_items = *list.count;
	if (n_items > list.capacity) n_items = list.capacity;
	for (unsigned int base = blockIdx.x * 128; base < n_items; base += gridDim.x * 128) {
		const unsigned int i = base + tid;
		if (i < n_items) {
			Column c{col + tid, 128};
			const uint4 *src = list.cols + (size_t) i * (NP / 8);
			unsigned long long S = 0, Q = 0;
#pragma unroll
			for (int q = 0; q < NP / 8; q++) {
				const uint4 w = src[q];
				const unsigned int ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
				for (int k = 0; k < 4; k++) {
					const unsigned int a = ww[k] & 0xFFFFu, bq = ww[k] >> 16;
					c.set(8 * q + 2 * k, a);
					c.set(8 * q + 2 * k + 1, bq);
					S += a + bq;
					Q += sq(a) + sq(bq);
				}
			}
			RejectState ist{0, 0, 0, 0};
			const unsigned int r = reject_sorted_and_mean<true>(c, NP, p.rejection, p.sig_low, p.sig_high, S, Q, ist, &tables);
			const uint2 m = list.meta[i];
			p.out[m.x] = (WORD) r;
			if (ist.rej_lo) atomicAdd(&blk_counters[m.y * 2 + 0], (unsigned long long) ist.rej_lo);
			if (ist.rej_hi) atomicAdd(&blk_counters[m.y * 2 + 1], (unsigned long long) ist.rej_hi);
			if (ist.guard) atomicAdd(&blk_counters[6], 1ull);
			if (ist.cap) atomicAdd(&blk_counters[7], 1ull);
		}
	}
	__syncthreads();
	if (p.counters && tid < SSC_COUNTERS && blk_counters[tid])
		atomicAdd(&p.counters[tid], blk_counters[tid]);
}

// Leftover-list storage, kept per device between launches (grown on demand; the first launch of
// a larger shape allocates, i.e. synchronises once).
struct LeftoverStore {
	unsigned int *count = nullptr;
	void *cols = nullptr;
	void *meta = nullptr;
	size_t capacity = 0, col_bytes = 0;
};
inline LeftoverStore g_leftover_store[64];

inline int leftover_list(int np, unsigned long long pixels, LeftoverList &list, cudaStream_t stream) {
	int dev = 0;
	cudaGetDevice(&dev);
	if (dev < 0 || dev >= 64) return SSC_ERR_ARG;
	LeftoverStore &st = g_leftover_store[dev];
	size_t want = (size_t) (pixels / 4) + 65536;	// a quarter of the pixels; beyond that pixels finish in place
	if (want > 0x7FFFFFF0ull) want = 0x7FFFFFF0ull;
	if (!st.count && cudaMalloc(&st.count, sizeof(unsigned int)) != cudaSuccess) return SSC_ERR_CUDA;
	if (st.capacity < want || st.col_bytes < (size_t) np * 2) {
		if (st.cols) cudaFree(st.cols);
		if (st.meta) cudaFree(st.meta);
		st.cols = st.meta = nullptr;
		st.capacity = 0;
		if (cudaMalloc(&st.cols, want * 128) != cudaSuccess) return SSC_ERR_CUDA;	// room for 64-frame columns
		if (cudaMalloc(&st.meta, want * sizeof(uint2)) != cudaSuccess) return SSC_ERR_CUDA;
		st.capacity = want;
		st.col_bytes = 128;
	}
	if (cudaMemsetAsync(st.count, 0, sizeof(unsigned int), stream) != cudaSuccess) return SSC_ERR_CUDA;
	list.count = st.count;
	list.capacity = (unsigned int) st.capacity;
	list.cols = (uint4 *) st.cols;
	list.meta = (uint2 *) st.meta;
	list.ids = nullptr;
	return SSC_OK;
}

// Index-only list for the register kernel (ssc_stack_fast.cu): room for every pixel of the launch,
// 4 bytes each, so nothing ever has to be finished in place.
struct LeftoverIdStore {
	unsigned int *count = nullptr, *ids = nullptr;
	size_t capacity = 0;
};
inline LeftoverIdStore g_leftover_id_store[64];

inline int leftover_ids(unsigned long long pixels, LeftoverList &list, cudaStream_t stream) {
	int dev = 0;
	cudaGetDevice(&dev);
	if (dev < 0 || dev >= 64 || pixels > 0xFFFFFFF0ull) return SSC_ERR_ARG;
	LeftoverIdStore &st = g_leftover_id_store[dev];
	if (!st.count && cudaMalloc(&st.count, sizeof(unsigned int)) != cudaSuccess) return SSC_ERR_CUDA;
	if (st.capacity < pixels) {
		if (st.ids) cudaFree(st.ids);
		st.ids = nullptr;
		st.capacity = 0;
		if (cudaMalloc(&st.ids, (size_t) pixels * sizeof(unsigned int)) != cudaSuccess) return SSC_ERR_CUDA;
		st.capacity = (size_t) pixels;
	}
	if (cudaMemsetAsync(st.count, 0, sizeof(unsigned int), stream) != cudaSuccess) return SSC_ERR_CUDA;
	list.count = st.count;
	list.capacity = (unsigned int) pixels;
	list.cols = nullptr;
	list.meta = nullptr;
	list.ids = st.ids;
	return SSC_OK;
}

template <int NP>
inline int launch_leftover(const StackParams &p, const LeftoverList &list, cudaStream_t stream) {
	int dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	k_stack_leftover<NP><<<sms * 8, 128, 0, stream>>>(p, list);
	return cudaGetLastError() == cudaSuccess ? SSC_OK : SSC_ERR_CUDA;
}

}  // namespace ssc
