// ssc_leftover.cuh -- the hand-back list of the register-network sigma kernel (ssc_stack_fast.cu).
#pragma once
#include "ssc_column.cuh"
#include "ssc_kernels.h"

namespace ssc {

// ---------------------------------------------------------------------------------------------
// Hand-back list: pixels the register-resident sigma clipping could not finish (more rejections
// than its candidate window holds) are appended by index to a device-global list and finished
// by k_sigma_handback, one pixel per thread.  A typical deep-sky stack parks well under 1 % of
// its pixels here.
// ---------------------------------------------------------------------------------------------
struct LeftoverList {
	unsigned int *count;		// items appended
	unsigned int capacity;		// = every pixel of the launch: the list cannot overflow
	unsigned int *ids;		// [capacity] indices of the pixels in the band ((layer * out_rows + row) * width + x)
	unsigned int *next_tile;	// work counter of the launch (ssc_stack_fast.cu: warp tiles are handed out dynamically)
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

// Index-only list for the register kernel (ssc_stack_fast.cu): room for every pixel of the launch,
// 4 bytes each, so nothing ever has to be finished in place.
struct LeftoverIdStore {
	unsigned int *count = nullptr, *ids = nullptr;
	size_t capacity = 0;
};
inline LeftoverIdStore g_leftover_id_store[64];

int launch_fit_handback(const StackParams &p, const LeftoverList &list, cudaStream_t stream);	// ssc_stack_warp.cu

inline int leftover_ids(unsigned long long pixels, LeftoverList &list, cudaStream_t stream, bool keep_items = false) {
	int dev = 0;
	cudaGetDevice(&dev);
	if (dev < 0 || dev >= 64 || pixels > 0xFFFFFFF0ull) return SSC_ERR_ARG;
	LeftoverIdStore &st = g_leftover_id_store[dev];
	if (!st.count && cudaMalloc(&st.count, 2 * sizeof(unsigned int)) != cudaSuccess) return SSC_ERR_CUDA;
	if (st.capacity < pixels) {
		if (keep_items) return SSC_ERR_ARG;	// (the deferred list is sized by its first launch)
		if (st.ids) cudaFree(st.ids);
		st.ids = nullptr;
		st.capacity = 0;
		if (cudaMalloc(&st.ids, (size_t) pixels * sizeof(unsigned int)) != cudaSuccess) return SSC_ERR_CUDA;
		st.capacity = (size_t) pixels;
	}
	if (cudaMemsetAsync(keep_items ? st.count + 1 : st.count, 0, (keep_items ? 1 : 2) * sizeof(unsigned int), stream) != cudaSuccess) return SSC_ERR_CUDA;
	list.count = st.count;
	list.capacity = (unsigned int) pixels;
	list.ids = st.ids;
	list.next_tile = st.count + 1;
	return SSC_OK;
}

}  // namespace ssc
