// ssc_peer.cu -- the multi-GPU exchanges of the stacking path as kernels over NVLink peer memory.
//
// The image is split in row bands, one per GPU (stacking.c:570-646 does the same with row blocks per
// thread); what the bands have to exchange is small and exact:
//   * the per-frame layer-0 histograms add up over the bands before the IKSS statistics,
//   * stack_summing needs the maximum over the bands between its two passes,
//   * the six rejection counters add up.
// Round 1 left these to the host (torch.distributed collectives in bench.py, host copies in the
// operator layer).  Here every GPU ("rank") owns a block of device memory that all the others map --
// cudaIpc handles between processes, cudaDeviceEnablePeerAccess inside one -- and the exchanges are
// part of the kernels:
//   K4  every rank counts its band into its own block and raises its flag at every rank;
//   K5  frame f is analysed by rank f % world: that rank waits for the flags of all ranks, adds the
//       counts of its frames straight out of the other ranks' memory (16-byte loads over NVLink: the
//       reduce-scatter, fused into the head of the statistics pass), runs the IKSS and stores the four
//       statistics of each frame into every rank's table (the all-gather), then raises that rank's flag;
//       (a first version added the counts from K4's flush with remote atomics instead: 336 k four-byte
//       atomics per rank and step ran at ~2 G/s, 0.17 ms -- slower than NCCL's reduce-scatter)
//   K6  waits for the flags of all owners and computes the coefficients locally, identical everywhere.
// Integer sums: the statistics stay exact under any sharding (tests/test_multirank_cpu.py for the
// arithmetic, tests/test_gpu_operator_paths.py on two GPUs).  Flags count up monotonically (one epoch per
// call), waits are bounded (a peer that never arrives raises SSC_PEER_TIMEOUT in the status word instead
// of hanging the GPU).
#include <cstdio>
#include <cstring>
#include <vector>
#include "ssc_device.cuh"
#include "ssc_kernels.h"

struct ssc_peer_group {
	int world = 1, rank = 0, n_frames_max = 0, owned_max = 0, device = 0;
	size_t bytes = 0;
	char *local = nullptr;			// this rank's block
	char *peer[SSC_PEER_MAX] = {};		// everybody's block as mapped here (peer[rank] == local)
	bool ipc_opened[SSC_PEER_MAX] = {};
	char **d_peer = nullptr;		// device: [world] block pointers
	bool accumulating = false;		// histogram pieces of the current exchange have been counted
	unsigned int epoch = 0;			// calls of the statistics exchange so far
	unsigned int epoch_cnt = 0, epoch_max = 0;
	bool connected = false;
};

namespace ssc {

// layout of a rank's block
struct PeerLayout {
	size_t hist_all, hist, stats, stats_own, counters, maxim, flags, status, total;
};
enum { FLAG_HIST = 0, FLAG_STATS = 1, FLAG_CNT = 2, FLAG_MAX = 3, N_FLAGS = 4 };

static PeerLayout peer_layout(int n_frames_max, int owned_max) {
	PeerLayout l;
	size_t o = 0;
	auto take = [&](size_t bytes) { const size_t at = o; o += (bytes + 255) & ~(size_t) 255; return at; };
	l.hist_all = take((size_t) n_frames_max * SSC_HIST_STRIDE * sizeof(uint32_t));	// this rank's band, every frame
	l.hist = take((size_t) owned_max * SSC_HIST_STRIDE * sizeof(uint32_t));		// all bands, the frames it owns
	l.stats = take((size_t) 4 * n_frames_max * sizeof(double));
	l.stats_own = take((size_t) 4 * owned_max * sizeof(double));
	l.counters = take((size_t) 2 * SSC_COUNTERS * sizeof(unsigned long long));	// two slots, by epoch parity
	l.maxim = take(2 * sizeof(unsigned long long));
	l.flags = take(N_FLAGS * sizeof(unsigned int));
	l.status = take(sizeof(int));
	l.total = o;
	return l;
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int load_acquire_sys(const unsigned int *p) {
	unsigned int v;
	asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}

// wait until *flag >= target (counts that wrapped around compare by difference); about ten seconds at most
__device__ __forceinline__ bool wait_flag(const unsigned int *flag, unsigned int target) {
	const long long t0 = clock64();
	while ((int) (load_acquire_sys(flag) - target) < 0) {
		__nanosleep(200);
		if (clock64() - t0 > 20000000000ll) return false;
	}
	return true;
}

__global__ void k_peer_wait(const unsigned int *flag, unsigned int target, int *status) {
	if (!wait_flag(flag, target)) atomicExch(status, SSC_PEER_TIMEOUT);
}

// everything this rank wrote before (earlier kernels of the stream) is visible system-wide before the flags move
__global__ void k_peer_signal(char *const *peer, int world, size_t flag_offset) {
	const int r = threadIdx.x;
	__threadfence_system();
	if (r < world) atomicAdd_system(reinterpret_cast<unsigned int *>(peer[r] + flag_offset), 1u);
}

// the reduce-scatter of the histograms, by the owner: slot s (frame s * world + rank) = sum over the ranks of
// their counts of that frame, read from their blocks with 16-byte loads; every CTA first waits for all ranks
constexpr int REDUCE_PARTS = 16;
__global__ void __launch_bounds__(256)
k_hist_reduce(char *const *peer, int world, int rank, int n_frames, size_t hist_all_offset, uint32_t *__restrict__ sum,
		const unsigned int *flag, unsigned int target, int *status) {
	__shared__ int ok;
	if (threadIdx.x == 0) {
		ok = wait_flag(flag, target);
		if (!ok) atomicExch(status, SSC_PEER_TIMEOUT);
	}
	__syncthreads();
	if (!ok) return;
	const int s = blockIdx.y, f = s * world + rank;
	if (f >= n_frames) return;
	constexpr int VEC = SSC_HIST_STRIDE / 4;	// 65540 words = 16385 uint4
	uint4 *dst = reinterpret_cast<uint4 *>(sum + (size_t) s * SSC_HIST_STRIDE);
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < VEC; i += gridDim.x * blockDim.x) {
		uint4 a = make_uint4(0, 0, 0, 0);
		for (int r = 0; r < world; r++) {
			const uint4 v = reinterpret_cast<const uint4 *>(peer[r] + hist_all_offset + (size_t) f * SSC_HIST_STRIDE * sizeof(uint32_t))[i];
			a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
		}
		dst[i] = a;
	}
}

// the all-gather of the statistics: slot s of this rank is frame s * world + rank
__global__ void k_stats_scatter(char *const *peer, int world, int rank, int n_frames, int owned, const double *own,
		size_t stats_offset, size_t flag_offset) {
	const int r = blockIdx.x;			// destination rank
	double *dst = reinterpret_cast<double *>(peer[r] + stats_offset);
	for (int i = threadIdx.x; i < 4 * owned; i += blockDim.x) {
		const int k = i / owned, s = i - k * owned, f = s * world + rank;
		if (f < n_frames) dst[(size_t) k * n_frames + f] = own[(size_t) k * owned + s];
	}
	__threadfence_system();
	__syncthreads();
	if (threadIdx.x == 0) atomicAdd_system(reinterpret_cast<unsigned int *>(peer[r] + flag_offset), 1u);
}

__global__ void k_counters_scatter(char *const *peer, int world, const unsigned long long *local, size_t slot_offset,
		size_t flag_offset) {
	const int r = blockIdx.x, k = threadIdx.x;
	if (k < SSC_COUNTERS && local[k])
		atomicAdd_system(reinterpret_cast<unsigned long long *>(peer[r] + slot_offset) + k, local[k]);
	__threadfence_system();
	__syncthreads();
	if (k == 0) atomicAdd_system(reinterpret_cast<unsigned int *>(peer[r] + flag_offset), 1u);
}

__global__ void k_max_scatter(char *const *peer, int world, const unsigned long long *local, size_t slot_offset,
		size_t flag_offset) {
	const int r = threadIdx.x;
	if (r < world) {
		atomicMax_system(reinterpret_cast<unsigned long long *>(peer[r] + slot_offset), *local);
		__threadfence_system();
		atomicAdd_system(reinterpret_cast<unsigned int *>(peer[r] + flag_offset), 1u);
	}
}

__global__ void k_copy_words(unsigned long long *dst, const unsigned long long *src, int n) {
	if ((int) threadIdx.x < n) dst[threadIdx.x] = src[threadIdx.x];
}

}  // namespace ssc

using namespace ssc;

namespace {
thread_local char g_peer_err[256];
int peer_fail(int code, const char *what, cudaError_t e = cudaSuccess) {
	snprintf(g_peer_err, sizeof(g_peer_err), "%s%s%s", what, e != cudaSuccess ? ": " : "", e != cudaSuccess ? cudaGetErrorString(e) : "");
	return code;
}
#define PCU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return peer_fail(SSC_ERR_CUDA, #call, e_); } while (0)

PeerLayout layout_of(const ssc_peer_group *g) { return peer_layout(g->n_frames_max, g->owned_max); }

int upload_tables(ssc_peer_group *g) {
	PCU(cudaMemcpy(g->d_peer, g->peer, (size_t) g->world * sizeof(char *), cudaMemcpyHostToDevice));
	g->connected = true;
	return SSC_OK;
}
}  // namespace

namespace ssc {
// this rank's own histograms of the running exchange (cleared when its first piece is counted: the owners
// finished reading the previous exchange's before their statistics, which this rank has waited for, went out)
int peer_local_histograms(ssc_peer_group *g, int n_frames, uint32_t **hist, cudaStream_t stream) {
	const PeerLayout l = layout_of(g);
	*hist = reinterpret_cast<uint32_t *>(g->local + l.hist_all);
	if (!g->accumulating) {
		if (cudaMemsetAsync(*hist, 0, (size_t) n_frames * SSC_HIST_STRIDE * sizeof(uint32_t), stream) != cudaSuccess) return SSC_ERR_CUDA;
		g->accumulating = true;
	}
	return SSC_OK;
}
int peer_signal_histograms(ssc_peer_group *g, cudaStream_t stream) {
	const PeerLayout l = layout_of(g);
	k_peer_signal<<<1, 64, 0, stream>>>(g->d_peer, g->world, l.flags + FLAG_HIST * sizeof(unsigned int));
	return cudaGetLastError() == cudaSuccess ? SSC_OK : SSC_ERR_CUDA;
}
}  // namespace ssc

extern "C" {

const char *ssc_peer_last_error(void) { return g_peer_err; }

int ssc_peer_create(int world, int rank, int n_frames_max, ssc_peer_group **out) {
	if (!out || world < 1 || world > SSC_PEER_MAX || rank < 0 || rank >= world || n_frames_max < 1)
		return peer_fail(SSC_ERR_ARG, "bad peer group geometry");
	ssc_peer_group *g = new ssc_peer_group;
	g->world = world;
	g->rank = rank;
	g->n_frames_max = n_frames_max;
	g->owned_max = (n_frames_max + world - 1) / world;
	PCU(cudaGetDevice(&g->device));
	const PeerLayout l = layout_of(g);
	g->bytes = l.total;
	PCU(cudaMalloc(&g->local, l.total));		// (cudaMalloc memory: exportable through cudaIpc)
	PCU(cudaMemset(g->local, 0, l.total));
	PCU(cudaMalloc(&g->d_peer, (size_t) world * sizeof(char *)));
	g->peer[rank] = g->local;
	if (world == 1) {
		int rc = upload_tables(g);
		if (rc) return rc;
	}
	PCU(cudaDeviceSynchronize());
	*out = g;
	return SSC_OK;
}

int ssc_peer_export(ssc_peer_group *g, void *handle64) {
	if (!g || !handle64) return peer_fail(SSC_ERR_ARG, "null argument");
	static_assert(sizeof(cudaIpcMemHandle_t) == SSC_PEER_HANDLE_BYTES, "handle size");
	cudaIpcMemHandle_t h;
	PCU(cudaSetDevice(g->device));
	PCU(cudaIpcGetMemHandle(&h, g->local));
	memcpy(handle64, &h, sizeof(h));
	return SSC_OK;
}

int ssc_peer_import(ssc_peer_group *g, int peer_rank, const void *handle64) {
	if (!g || !handle64 || peer_rank < 0 || peer_rank >= g->world) return peer_fail(SSC_ERR_ARG, "bad peer rank");
	if (peer_rank == g->rank) return SSC_OK;
	cudaIpcMemHandle_t h;
	memcpy(&h, handle64, sizeof(h));
	void *p = nullptr;
	PCU(cudaSetDevice(g->device));
	PCU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
	g->peer[peer_rank] = (char *) p;
	g->ipc_opened[peer_rank] = true;
	return SSC_OK;
}

int ssc_peer_attach_local(ssc_peer_group *g, int peer_rank, ssc_peer_group *other) {
	if (!g || !other || peer_rank < 0 || peer_rank >= g->world || other->rank != peer_rank
			|| other->n_frames_max != g->n_frames_max || other->world != g->world)
		return peer_fail(SSC_ERR_ARG, "peer groups do not match");
	if (peer_rank == g->rank) return SSC_OK;
	PCU(cudaSetDevice(g->device));
	if (other->device != g->device) {
		int can = 0;
		PCU(cudaDeviceCanAccessPeer(&can, g->device, other->device));
		if (!can) return peer_fail(SSC_ERR_CUDA, "no peer access between the devices");
		const cudaError_t e = cudaDeviceEnablePeerAccess(other->device, 0);
		if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return peer_fail(SSC_ERR_CUDA, "cudaDeviceEnablePeerAccess", e);
		cudaGetLastError();
	}
	g->peer[peer_rank] = other->local;
	return SSC_OK;
}

int ssc_peer_connect(ssc_peer_group *g) {
	if (!g) return peer_fail(SSC_ERR_ARG, "null group");
	for (int r = 0; r < g->world; r++)
		if (!g->peer[r]) return peer_fail(SSC_ERR_ARG, "a peer's block is missing (ssc_peer_import / ssc_peer_attach_local)");
	PCU(cudaSetDevice(g->device));
	return upload_tables(g);
}

int ssc_peer_destroy(ssc_peer_group *g) {
	if (!g) return SSC_OK;
	cudaSetDevice(g->device);
	cudaDeviceSynchronize();
	for (int r = 0; r < g->world; r++)
		if (g->ipc_opened[r]) cudaIpcCloseMemHandle(g->peer[r]);
	cudaFree(g->d_peer);
	cudaFree(g->local);
	delete g;
	return SSC_OK;
}

int ssc_peer_world(const ssc_peer_group *g) { return g ? g->world : 0; }
int ssc_peer_rank(const ssc_peer_group *g) { return g ? g->rank : -1; }

// K4 + reduce-scatter: every frame's counts of this band are added into its owner's histogram.
int ssc_dev_histogram_scatter(const ssc_band *band, ssc_peer_group *g, int last_piece, void *stream) {
	if (!band || !g || !g->connected || band->n_frames > g->n_frames_max) return peer_fail(SSC_ERR_ARG, "histogram_scatter: bad arguments");
	const PeerLayout l = layout_of(g);
	cudaStream_t st = (cudaStream_t) stream;
	uint32_t *hist = nullptr;
	int rc = peer_local_histograms(g, band->n_frames, &hist, st);
	if (!rc) rc = launch_histogram(*band, hist, st);
	if (rc) return peer_fail(rc, "histogram launch failed");
	if (!last_piece) return SSC_OK;	// more row pieces of this rank's band follow: one arrival flag per call of the exchange
	k_peer_signal<<<1, 64, 0, st>>>(g->d_peer, g->world, l.flags + FLAG_HIST * sizeof(unsigned int));
	if (cudaGetLastError() != cudaSuccess) return peer_fail(SSC_ERR_CUDA, "signal launch failed");
	return SSC_OK;
}

// K5 on the frames this rank owns once every rank's counts have arrived, then the all-gather of the statistics.
// stats (device double[4][n_frames], may be NULL) receives a copy of the gathered table after the wait in
// ssc_dev_norm_coeffs_gathered; scratch: device uint32_t [owned][65536].
int ssc_dev_ikss_gather(ssc_peer_group *g, int n_frames, uint32_t *scratch, void *stream) {
	if (!g || !g->connected || n_frames < 1 || n_frames > g->n_frames_max) return peer_fail(SSC_ERR_ARG, "ikss_gather: bad arguments");
	const PeerLayout l = layout_of(g);
	cudaStream_t st = (cudaStream_t) stream;
	g->epoch++;
	const int owned = (n_frames - g->rank + g->world - 1) / g->world;	// frames rank, rank + world, ...
	uint32_t *hist = reinterpret_cast<uint32_t *>(g->local + l.hist);
	double *own = reinterpret_cast<double *>(g->local + l.stats_own);
	int *status = reinterpret_cast<int *>(g->local + l.status);
	g->accumulating = false;
	if (owned > 0) {
		dim3 grid(REDUCE_PARTS, owned);
		k_hist_reduce<<<grid, 256, 0, st>>>(g->d_peer, g->world, g->rank, n_frames, l.hist_all, hist,
				reinterpret_cast<unsigned int *>(g->local + l.flags) + FLAG_HIST, g->epoch * (unsigned int) g->world, status);
		int rc = launch_ikss(hist, owned, own, scratch, st);
		if (rc) return peer_fail(rc, "ikss launch failed");
	} else {
		k_peer_wait<<<1, 1, 0, st>>>(reinterpret_cast<unsigned int *>(g->local + l.flags) + FLAG_HIST, g->epoch * (unsigned int) g->world, status);
	}
	k_stats_scatter<<<g->world, 128, 0, st>>>(g->d_peer, g->world, g->rank, n_frames, owned > 0 ? owned : 1, own, l.stats,
			l.flags + FLAG_STATS * sizeof(unsigned int));
	if (cudaGetLastError() != cudaSuccess) return peer_fail(SSC_ERR_CUDA, "stats scatter launch failed");
	return SSC_OK;
}

// K6 once the statistics of all owners are here; stats_out (device double[4][n_frames], may be NULL) gets a copy.
int ssc_dev_norm_coeffs_gathered(ssc_peer_group *g, int n_frames, int ref_image, int mode, double *coeff, int *status_out,
		double *stats_out, void *stream) {
	if (!g || !g->connected || n_frames < 1 || n_frames > g->n_frames_max || !coeff) return peer_fail(SSC_ERR_ARG, "norm_coeffs_gathered: bad arguments");
	const PeerLayout l = layout_of(g);
	cudaStream_t st = (cudaStream_t) stream;
	int *status = reinterpret_cast<int *>(g->local + l.status);
	const double *stats = reinterpret_cast<const double *>(g->local + l.stats);
	k_peer_wait<<<1, 1, 0, st>>>(reinterpret_cast<unsigned int *>(g->local + l.flags) + FLAG_STATS, g->epoch * (unsigned int) g->world, status);
	int rc = launch_norm_coeffs(stats, n_frames, ref_image, mode, coeff, status_out, st);
	if (rc) return peer_fail(rc, "norm_coeffs launch failed");
	if (stats_out) PCU(cudaMemcpyAsync(stats_out, stats, (size_t) 4 * n_frames * sizeof(double), cudaMemcpyDeviceToDevice, st));
	return SSC_OK;
}

// all-reduce(SUM) of the rejection counters: total (device uint64[SSC_COUNTERS]) = sum over the ranks of `local`
int ssc_dev_counters_combine(ssc_peer_group *g, const uint64_t *local, uint64_t *total, void *stream) {
	if (!g || !g->connected || !local || !total) return peer_fail(SSC_ERR_ARG, "counters_combine: bad arguments");
	const PeerLayout l = layout_of(g);
	cudaStream_t st = (cudaStream_t) stream;
	g->epoch_cnt++;
	const size_t slot = l.counters + (g->epoch_cnt & 1) * SSC_COUNTERS * sizeof(unsigned long long);
	const size_t next = l.counters + ((g->epoch_cnt + 1) & 1) * SSC_COUNTERS * sizeof(unsigned long long);
	// the other slot is cleared BEFORE this call's flag moves: a faster rank can only reach its next call after it
	PCU(cudaMemsetAsync(g->local + next, 0, SSC_COUNTERS * sizeof(unsigned long long), st));
	k_counters_scatter<<<g->world, 32, 0, st>>>(g->d_peer, g->world, reinterpret_cast<const unsigned long long *>(local), slot,
			l.flags + FLAG_CNT * sizeof(unsigned int));
	k_peer_wait<<<1, 1, 0, st>>>(reinterpret_cast<unsigned int *>(g->local + l.flags) + FLAG_CNT, g->epoch_cnt * (unsigned int) g->world,
			reinterpret_cast<int *>(g->local + l.status));
	k_copy_words<<<1, 32, 0, st>>>(reinterpret_cast<unsigned long long *>(total), reinterpret_cast<const unsigned long long *>(g->local + slot), SSC_COUNTERS);
	if (cudaGetLastError() != cudaSuccess) return peer_fail(SSC_ERR_CUDA, "counters exchange launch failed");
	return SSC_OK;
}

// all-reduce(MAX) of stack_summing's maximum (stacking.c:311-313): *maxim (device uint64) becomes the global one
int ssc_dev_max_combine(ssc_peer_group *g, unsigned long long *maxim, void *stream) {
	if (!g || !g->connected || !maxim) return peer_fail(SSC_ERR_ARG, "max_combine: bad arguments");
	const PeerLayout l = layout_of(g);
	cudaStream_t st = (cudaStream_t) stream;
	g->epoch_max++;
	const size_t slot = l.maxim + (g->epoch_max & 1) * sizeof(unsigned long long);
	const size_t next = l.maxim + ((g->epoch_max + 1) & 1) * sizeof(unsigned long long);
	PCU(cudaMemsetAsync(g->local + next, 0, sizeof(unsigned long long), st));
	k_max_scatter<<<1, 64, 0, st>>>(g->d_peer, g->world, maxim, slot, l.flags + FLAG_MAX * sizeof(unsigned int));
	k_peer_wait<<<1, 1, 0, st>>>(reinterpret_cast<unsigned int *>(g->local + l.flags) + FLAG_MAX, g->epoch_max * (unsigned int) g->world,
			reinterpret_cast<int *>(g->local + l.status));
	k_copy_words<<<1, 32, 0, st>>>(maxim, reinterpret_cast<const unsigned long long *>(g->local + slot), 1);
	if (cudaGetLastError() != cudaSuccess) return peer_fail(SSC_ERR_CUDA, "max exchange launch failed");
	return SSC_OK;
}

// 0, or SSC_PEER_TIMEOUT when a wait gave up (synchronises the stream)
int ssc_peer_status(ssc_peer_group *g, void *stream) {
	if (!g) return SSC_ERR_ARG;
	const PeerLayout l = layout_of(g);
	int h = 0;
	PCU(cudaMemcpyAsync(&h, g->local + l.status, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t) stream));
	PCU(cudaStreamSynchronize((cudaStream_t) stream));
	return h;
}

}  // extern "C"
