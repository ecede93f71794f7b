// ssc_stats.cu -- K4 / K5 / K6: the normalization statistics.
//
// The reference computes, per frame, IKSS location and scale of the non-zero
// samples of layer 0 by sorting W*H doubles (src/algos/statistics.c:152-187,
// 277-290, reached through sequence.c:1107-1118 from stacking.c:79-123).
// Here the same estimator runs on an exact 65536-bin histogram: every clipping
// window of the sorted-sample form is a run of whole bins, so its medians and
// MADs are the same order statistics, and histograms of row bands (or GPUs)
// simply add -- the statistics stay exact under any sharding.
#include "ssc_device.cuh"
#include "ssc_kernels.h"

namespace ssc {

// ---------------------------------------------------------------------------
// K4: histogram of layer 0 of every frame of a band.
//
// A CTA owns two shared-memory tables:
//   * full: all 65536 bins as 16-bit counters packed two per word (128 KB).  A 16-bit counter
//           cannot overflow before the CTA has put 65536 samples into the table, so the table is
//           flushed to the global 32-bit histogram whenever that could happen next;
//   * win:  32-bit counters for a 2048-bin window around the block's typical sample value (8 KB),
//           never flushed before the end.
// The background of an astronomical frame is narrow, so almost every sample lands in the window
// and the flush of the full table (8192 16-byte reads and two barriers) becomes rare; with
// wide-range data everything goes through the full table -- same result either way.
// Global atomics scale with the number of DISTINCT values, not with the samples.
// Measured (profiles/r01_ncu_histogram.txt): the kernel is bound by the shared-memory atomic unit
// (a warp's ATOMS costs 4-5 wavefronts whatever the bank pattern: a lane-private, conflict-free
// window table measured no faster, and neither did flushing less often), not by HBM.
// ---------------------------------------------------------------------------
constexpr int HIST_THREADS = 1024;
constexpr int HIST_DEPTH = 3;
constexpr int HIST_WIN = 2048;				// bins in the 32-bit window
constexpr unsigned int HIST_GROUP_SAMPLES = HIST_DEPTH * HIST_THREADS * 8;	// 24576
constexpr size_t HIST_SMEM = (32768 + HIST_WIN) * sizeof(uint32_t);

__global__ void __launch_bounds__(HIST_THREADS, 1)
k_histogram(ssc_band b, uint32_t *__restrict__ hist, uint32_t *const *__restrict__ frame_hist,
		uint32_t *__restrict__ local_gt255, int ctas_per_frame) {
	extern __shared__ __align__(16) uint32_t sh[];	// full[32768], then win[HIST_WIN]
	__shared__ unsigned int any_gt255, s_base, s_outside[2], s_acc[32];
	const int frame = blockIdx.x / ctas_per_frame;
	const int part = blockIdx.x % ctas_per_frame;
	const int tid = threadIdx.x, lane = tid & 31;
	const size_t band_px = (size_t) b.out_rows * b.width;	// layer-0 samples of this band
	const WORD *src = b.frames + (size_t) frame * b.frame_stride + (size_t) (b.out_row0 - b.buf_row0) * b.width;
	// frame_hist: one histogram pointer per frame instead of a dense array -- with several GPUs every frame's
	// histogram lives on the GPU that analyses it and the flushes below add into it over NVLink (ssc_peer.cu)
	uint32_t *gh = frame_hist ? frame_hist[frame] : hist + (size_t) frame * SSC_HIST_STRIDE;

	for (int i = tid; i < 32768 + HIST_WIN; i += HIST_THREADS) sh[i] = 0;
	if (tid == 0) { any_gt255 = 0; s_outside[0] = s_outside[1] = 0; }

	// this block's contiguous share of the frame's band, in 16-byte vectors (head and tail samples
	// that do not fall on a vector are left to part 0 below)
	const size_t head = ((16 - ((uintptr_t) src & 15)) & 15) >> 1;		// samples before the first aligned vector
	const size_t lead = head < band_px ? head : band_px;
	const size_t nvec_all = (band_px - lead) >> 3;
	const size_t v_beg = nvec_all * part / ctas_per_frame, v_end = nvec_all * (part + 1) / ctas_per_frame;
	const uint4 *v4 = reinterpret_cast<const uint4 *>(src + lead);

	uint4 q[HIST_DEPTH];
#pragma unroll
	for (int k = 0; k < HIST_DEPTH; k++) {
		const size_t idx = v_beg + tid + (size_t) k * HIST_THREADS;
		q[k] = idx < v_end ? __ldg(v4 + idx) : make_uint4(0, 0, 0, 0);
	}

	// window position: the mean of one sample per thread (any position gives the right histogram)
	{
		const bool have = v_beg + tid < v_end;
		const unsigned int probe = __reduce_add_sync(0xffffffffu, have ? (q[0].x & 0xFFFFu) : 0u);
		const unsigned int cnt = __reduce_add_sync(0xffffffffu, have ? 1u : 0u);
		if (lane == 0) s_acc[tid >> 5] = cnt ? probe / cnt : 0xFFFFFFFFu;
	}
	__syncthreads();
	if (tid == 0) {
		unsigned int m = 0, k = 0;
		for (int i = 0; i < HIST_THREADS / 32; i++)
			if (s_acc[i] != 0xFFFFFFFFu) { m += s_acc[i]; k++; }
		int base = (int) (k ? m / k : 0u) - HIST_WIN / 2;
		if (base < 0) base = 0;
		if (base > 65536 - HIST_WIN) base = 65536 - HIST_WIN;
		s_base = (unsigned int) base;
	}
	__syncthreads();
	const unsigned int base = s_base;

	unsigned int seen = 0, outside = 0;
	auto count1 = [&](unsigned int a) {	// one atomic, no divergence
		const unsigned int d = a - base;
		const bool in = d < (unsigned int) HIST_WIN;
		atomicAdd(&sh[in ? 32768u + d : a >> 1], in ? 1u : 1u << ((a & 1) * 16));
		outside += in ? 0u : 1u;
		seen |= a;
	};
	// Two samples at a time: with base2 = base | base << 16, t = w - base2 has both 11-bit window
	// offsets in place exactly when both samples are inside the window (a low half below the window
	// borrows and leaves bits >= 11 set in it, because base <= 65536 - HIST_WIN), so the common case of
	// an astronomical background costs one subtraction and one test per pair on top of the two atomics.
	const unsigned int base2 = base | (base << 16);
	unsigned int *const win = sh + 32768;
	auto count2 = [&](unsigned int w) {
		const unsigned int t = w - base2;
		if ((t & ((0xFFFFu ^ (HIST_WIN - 1)) * 0x10001u)) == 0) {
			atomicAdd(win + (t & 0xFFFFu), 1u);
			atomicAdd(win + (t >> 16), 1u);
			seen |= w;
		} else {
			count1(w & 0xFFFFu);
			count1(w >> 16);
		}
	};
	auto count8 = [&](const uint4 &v) {
		count2(v.x);
		count2(v.y);
		count2(v.z);
		count2(v.w);
	};
	auto flush_full = [&]() {	// non-zero 16-bit counters -> the global 32-bit histogram
		__syncthreads();
		uint4 *sh4 = reinterpret_cast<uint4 *>(sh);
		for (int i = tid; i < 8192; i += HIST_THREADS) {
			const uint4 x = sh4[i];
			if (x.x | x.y | x.z | x.w) {
				sh4[i] = make_uint4(0, 0, 0, 0);
				const unsigned int w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
				for (int k = 0; k < 4; k++) {
					if (w[k] & 0xFFFFu) atomicAdd(&gh[8 * i + 2 * k], w[k] & 0xFFFFu);
					if (w[k] >> 16) atomicAdd(&gh[8 * i + 2 * k + 1], w[k] >> 16);
				}
			}
		}
		__syncthreads();
	};

	// Software pipeline, HIST_DEPTH 16-byte loads in flight per thread (48 KB per SM): the loads of
	// the next group are issued before the current group's shared-memory atomics and stay in flight
	// across a flush.
	const size_t per_group = (size_t) HIST_DEPTH * HIST_THREADS;
	const size_t n_groups = (v_end - v_beg + per_group - 1) / per_group;	// block-uniform
	unsigned int in_full = 0, prev[2] = {0, 0};
	for (size_t g = 0; g < n_groups; g++) {
#pragma unroll
		for (int k = 0; k < HIST_DEPTH; k++) {
			const size_t idx = v_beg + tid + (g * HIST_DEPTH + k) * HIST_THREADS;
			const uint4 v = q[k];
			const size_t nxt = idx + per_group;
			q[k] = nxt < v_end ? __ldg(v4 + nxt) : make_uint4(0, 0, 0, 0);
			if (idx < v_end) count8(v);
		}
		// Samples this CTA has put into the full table since its last flush, checked every second group
		// (a barrier each time).  Two monotonic counters used alternately: a warp already in the next
		// round adds to the other one, so every thread reads the same value here and the flush decision
		// is block-uniform.
		if ((g & 1) == 1 || g + 1 == n_groups) {
			const unsigned int round = (unsigned int) (g >> 1) & 1u;
			const unsigned int w_out = __reduce_add_sync(0xffffffffu, outside);
			outside = 0;
			if (lane == 0 && w_out) atomicAdd(&s_outside[round], w_out);
			__syncthreads();
			const unsigned int t = s_outside[round];
			in_full += t - prev[round];
			prev[round] = t;
			if (in_full + 2 * HIST_GROUP_SAMPLES + 16 > 65535u) {	// the next two groups could overflow a counter
				flush_full();
				in_full = 0;
			}
		}
	}
	if (part == 0) {	// the few (< 16) samples outside the vectors
		for (size_t i = tid; i < lead; i += HIST_THREADS) count1(src[i]);
		for (size_t i = lead + (nvec_all << 3) + tid; i < band_px; i += HIST_THREADS) count1(src[i]);
	}
	flush_full();
	for (int i = tid; i < HIST_WIN; i += HIST_THREADS) {
		const unsigned int c = sh[32768 + i];
		if (c) atomicAdd(&gh[base + i], c);
	}
	// word 65537: "layer 0 of this band has a sample > 255" (consumed by k_other_layers_gt255)
	if ((seen & 0xFFFFu) > 255u || (seen >> 16) > 255u) atomicOr(&any_gt255, 1u);	// (seen collects packed pairs and single samples)
	__syncthreads();
	if (tid == 0 && any_gt255) {
		atomicOr(&gh[65537], 1u);
		if (local_gt255) local_gt255[frame] = 1u;	// the same fact kept on this GPU (gh may be a peer's memory)
	}
}

// get_normalized_value (src/core/utils.c:454-459) looks at the max over ALL
// layers.  Only when layer 0 of this band stays <= 255 can the other layers
// matter; then count their samples > 255 into word 65536.
__global__ void k_other_layers_gt255(ssc_band b, uint32_t *__restrict__ hist, uint32_t *const *__restrict__ frame_hist,
		const uint32_t *__restrict__ local_gt255) {
	const int frame = blockIdx.y;
	uint32_t *gh = frame_hist ? frame_hist[frame] : hist + (size_t) frame * SSC_HIST_STRIDE;
	if (local_gt255 ? local_gt255[frame] != 0 : gh[65537] != 0) return;
	const size_t band_px = (size_t) b.out_rows * b.width;
	unsigned int cnt = 0;
	for (int layer = 1; layer < b.layers; layer++) {
		const WORD *src = b.frames + (size_t) frame * b.frame_stride + (size_t) layer * b.layer_stride
				+ (size_t) (b.out_row0 - b.buf_row0) * b.width;
		for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < band_px;
				i += (size_t) gridDim.x * blockDim.x)
			cnt += src[i] > 255;
	}
	for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
	if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&gh[65536], cnt);
}

int launch_histogram(const ssc_band &b, uint32_t *hist, cudaStream_t stream) {
	return launch_histogram_to(b, hist, nullptr, stream);
}

int launch_histogram_to(const ssc_band &b, uint32_t *hist, uint32_t *const *frame_hist, cudaStream_t stream) {
	return launch_histogram_ex(b, hist, frame_hist, nullptr, false, stream);
}

// the second half of K4 on its own: samples > 255 in layers 1.. of the frames whose layer 0 stayed <= 255
int launch_other_layers(const ssc_band &b, uint32_t *hist, uint32_t *const *frame_hist, const uint32_t *local_gt255,
		cudaStream_t stream) {
	if (b.layers < 2) return SSC_OK;
	dim3 grid(8, b.n_frames);
	k_other_layers_gt255<<<grid, 256, 0, stream>>>(b, hist, frame_hist, local_gt255);
	return cudaGetLastError() == cudaSuccess ? SSC_OK : SSC_ERR_CUDA;
}

// local_gt255 (device uint32_t [n_frames], may be NULL): set to 1 for frames whose layer 0 has a sample > 255 in
// this band; layer0_only: leave the other layers to a later launch_other_layers (they may not be uploaded yet)
int launch_histogram_ex(const ssc_band &b, uint32_t *hist, uint32_t *const *frame_hist, uint32_t *local_gt255,
		bool layer0_only, cudaStream_t stream) {
	if (b.n_frames < 1 || (!hist && !frame_hist)) return SSC_ERR_ARG;
	int dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	if (cudaFuncSetAttribute(k_histogram, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) HIST_SMEM) != cudaSuccess)
		return SSC_ERR_CUDA;
	const size_t band_px = (size_t) b.out_rows * b.width;
	// One CTA per SM is resident (136 KB of shared memory) and all CTAs do the same amount of work, so
	// the launch runs in waves of `sms` CTAs: spread every frame over enough CTAs to fill the machine
	// at least twice and pick the count whose LAST wave is nearly full (64 frames x 5 parts = 320 CTAs
	// on 148 SMs ran 3 waves for the work of 2.16), but keep at least ~32K samples per CTA.
	const size_t max_parts = band_px / 32768 + 1;
	int per_frame = (2 * sms + b.n_frames - 1) / b.n_frames;
	{
		double best = 0.0;
		int best_p = per_frame;
		for (int pf = per_frame; pf < per_frame + 32; pf++) {
			const double waves = (double) b.n_frames * pf / sms;
			const double eff = waves / (double) (long long) (waves + 0.999999);
			if (eff > best + 1e-9) { best = eff; best_p = pf; }
			if (eff >= 0.97) break;
		}
		per_frame = best_p;
	}
	if ((size_t) per_frame > max_parts) per_frame = (int) max_parts;
	if (per_frame < 1) per_frame = 1;
	k_histogram<<<b.n_frames * per_frame, HIST_THREADS, HIST_SMEM, stream>>>(b, hist, frame_hist, local_gt255, per_frame);
	if (cudaGetLastError() != cudaSuccess) return SSC_ERR_CUDA;
	if (b.layers > 1 && !layer0_only) return launch_other_layers(b, hist, frame_hist, local_gt255, stream);
	return SSC_OK;
}

// ---------------------------------------------------------------------------
// K5: IKSS per frame from the histogram.  One CTA of 1024 threads per frame.
// Restates statistics.c:152-187 exactly as oracle/siril_oracle.c so_ikss_hist
// does (same summation order for the biweight sums -> bit-identical doubles).
// ---------------------------------------------------------------------------
constexpr int IKSS_THREADS = 1024;

// terms of siril_stats_double_bwmv, statistics.c:136-145
// (the two divisions per bin -- by the histogram norm and by 9 MAD -- go through Reciprocal: the same
// correctly rounded quotients from one real division per thread and IKSS iteration)
__device__ __forceinline__ void bwmv_term(double d, double m, const Reciprocal &inv_mad9, double &up, double &down) {
	double dm = dsub(d, m);
	double yi = inv_mad9.divide(dm);
	double yi2 = dmul(yi, yi);
	double ai = (fabs(yi) < 1.0) ? 1.0 : 0.0;
	double omy = dsub(1.0, yi2);
	double omy2 = dmul(omy, omy);
	up = dmul(dmul(ai, dmul(dm, dm)), dmul(omy2, omy2));
	down = dmul(dmul(ai, omy), dsub(1.0, dmul(5.0, yi2)));
}

__global__ void __launch_bounds__(IKSS_THREADS, 1)
k_ikss(const uint32_t *__restrict__ hist, double *__restrict__ stats, uint32_t *__restrict__ scratch, int n_frames) {
	const int frame = blockIdx.x;
	const int tid = threadIdx.x;
	const uint32_t *h = hist + (size_t) frame * SSC_HIST_STRIDE;
	uint32_t *cum = scratch + (size_t) frame * 65536;	// inclusive prefix counts, cum[0] = 0 (zeros excluded)

	__shared__ double pu[IKSS_THREADS], pd[IKSS_THREADS];
	__shared__ uint32_t wsum[32];
	__shared__ int s_bi, s_bj, s_done, s_top;
	__shared__ double s_m, s_mad, s_norm, s_loc, s_scale, s_s0;
	__shared__ unsigned long long s_n;

	// --- prefix counts: warp w owns bins [2048w, 2048w + 2048), read and written 32 consecutive bins at a
	// time (the first version gave every thread 64 consecutive bins: 32 sectors per load, ~100 us) ---
	const int lane = tid & 31, warp = tid >> 5;
	{
		uint32_t total = 0;
		int top = 0;
		for (int k = 0; k < 64; k++) {
			const int bin = warp * 2048 + k * 32 + lane;
			const uint32_t c = bin == 0 ? 0u : h[bin];
			total += c;
			if (c) top = bin;
		}
		for (int o = 16; o > 0; o >>= 1) {
			total += __shfl_xor_sync(0xffffffffu, total, o);
			top = max(top, __shfl_xor_sync(0xffffffffu, top, o));
		}
		if (lane == 0) wsum[warp] = total;
		if (tid == 0) s_top = 0;
		__syncthreads();
		if (lane == 0) atomicMax(&s_top, top);
		if (tid < 32) {
			uint32_t v = wsum[tid], inc = v;
			for (int o = 1; o < 32; o <<= 1) {
				uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
				if (tid >= o) inc += u;
			}
			wsum[tid] = inc - v;	// exclusive warp offsets
		}
		__syncthreads();
		uint32_t run = wsum[warp];
		for (int k = 0; k < 64; k += 4) {	// four independent warp scans at a time (the chain of shuffles was 64 scans long)
			uint32_t incl[4];
#pragma unroll
			for (int j = 0; j < 4; j++) {
				const int bin = warp * 2048 + (k + j) * 32 + lane;
				incl[j] = bin == 0 ? 0u : h[bin];
			}
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
				for (int j = 0; j < 4; j++) {
					const uint32_t u = __shfl_up_sync(0xffffffffu, incl[j], o);
					if (lane >= o) incl[j] += u;
				}
			}
#pragma unroll
			for (int j = 0; j < 4; j++) {
				cum[warp * 2048 + (k + j) * 32 + lane] = run + incl[j];
				run += __shfl_sync(0xffffffffu, incl[j], 31);
			}
		}
	}
	__syncthreads();

	if (tid == 0) {
		// hist_size - 1: 255 when the frame's max over all layers is <= 255 (gui/histogram.c:116)
		s_norm = (s_top <= 255 && h[65536] == 0) ? 255.0 : 65535.0;
		s_bi = 1;
		s_bj = 65536;
		s_done = 0;
		s_s0 = 1.0;
		s_loc = s_scale = 0.0;
	}
	__syncthreads();
	const double norm = s_norm;
	const Reciprocal inv_norm(norm);	// x / norm below: the same correctly rounded quotients from one real division (ssc_device.cuh)
	const unsigned long long n_total = cum[65535];

	for (int it = 0; it < SSC_ITER_CAP; it++) {
		if (tid < 32) {
			// Warp 0 finds the order statistics together: every search is a 33-ary search, 32 probes
			// of the (global-memory) prefix counts per round instead of one -- the dependent loads
			// were this kernel's critical path.  All lanes carry the same scalars; lane 0 publishes.
			const int lane = tid;
			const int bi = s_bi, bj = s_bj;
			const unsigned long long base = cum[bi - 1];
			const unsigned long long n = bj > bi ? (unsigned long long) cum[bj - 1] - base : 0ull;
			if (n < 1) {		// statistics.c:161-164
				if (lane == 0) { s_n = n; s_loc = s_scale = 0.0; s_done = 1; }
			} else {
				// smallest k in [lo, hi] with pred(k), pred monotone and pred(hi) true
				auto search = [&](int lo, int hi, auto pred) {
					while (lo < hi) {
						const int span = hi - lo, step = span / 32 + 1;
						int pos = lo + (lane + 1) * step - 1;
						if (pos > hi) pos = hi;
						const unsigned int hit = __ballot_sync(0xffffffffu, pred(pos));
						const int first = __ffs(hit) - 1;		// some lane sits on hi: hit != 0
						const int p_first = __shfl_sync(0xffffffffu, pos, first);
						const int p_prev = __shfl_sync(0xffffffffu, pos, first > 0 ? first - 1 : 0);
						hi = p_first;
						if (first > 0) lo = p_prev + 1;
					}
					return lo;
				};
				// median: bins holding ranks (n-1)/2 and n/2 (gsl_stats_median_from_sorted_data)
				const unsigned long long rk[2] = {(n - 1) >> 1, n >> 1};
				int kb[2];
				kb[0] = search(bi, bj - 1, [&](int k) { return (unsigned long long) cum[k] - base > rk[0]; });
				// (rank n/2 sits in the same bin as rank (n-1)/2 unless that bin ends exactly there: one probe, then the search)
				kb[1] = (unsigned long long) cum[kb[0]] - base > rk[1] ? kb[0]
						: search(kb[0] + 1, bj - 1, [&](int k) { return (unsigned long long) cum[k] - base > rk[1]; });
				double dl = inv_norm.divide((double) kb[0]), dr = inv_norm.divide((double) kb[1]);
				double m = rk[0] == rk[1] ? dl : dmul(dadd(dl, dr), 0.5);	// (x / 2.0 == x * 0.5 exactly)
				// MAD (statistics.c:83-100): order statistics of |d_k - m|.  With M2 = kb0 + kb1
				// (twice the median in bin units) bins order by the integer distance
				// u_k = |2k - M2|; only the mirror pair (k, M2 - k) ties, and there the FP
				// deltas decide.  C(u) = samples with u_k <= u, from the prefix counts.
				const int M2 = kb[0] + kb[1];
				auto count_within = [&](int u) -> unsigned long long {
					int khi = (M2 + u) >> 1;			// floor((M2+u)/2)
					int klo = M2 - u >= 0 ? (M2 - u + 1) >> 1 : 0;	// ceil((M2-u)/2)
					if (khi > bj - 1) khi = bj - 1;
					if (klo < bi) klo = bi;
					return khi >= klo ? (unsigned long long) cum[khi] - cum[klo - 1] : 0ull;
				};
				double dv[2];
				// the largest distance any bin of the window has: count_within(u_max) = n > rank
				const int u_max = max(M2 - 2 * bi, 2 * (bj - 1) - M2);
				int u_prev = 0;
				for (int q = 0; q < 2; q++) {
					const unsigned long long rank = rk[q];
					// (the second rank: the same distance unless the count ends exactly there -- one probe, then the search)
					const int u = q == 1 && count_within(u_prev) > rank ? u_prev
							: search(q == 1 ? u_prev + 1 : 0, u_max, [&](int uu) { return count_within(uu) > rank; });
					u_prev = u;
					const unsigned long long below = u > 0 ? count_within(u - 1) : 0ull;	// C(u - 1)
					const int k1 = (M2 - u) >> 1, k2 = (M2 + u) >> 1;	// exact: u has M2's parity here
					const bool in1 = (M2 - u) >= 0 && k1 >= bi && k1 < bj;
					const bool in2 = k2 != k1 && k2 >= bi && k2 < bj;
					const unsigned long long c1 = in1 ? (unsigned long long) cum[k1] - cum[k1 - 1] : 0ull;
					const double d1 = fabs(dsub(inv_norm.divide((double) k1), m));
					const double d2 = fabs(dsub(inv_norm.divide((double) k2), m));
					const unsigned long long pos = rk[q] - below;
					if (!in2) dv[q] = d1;
					else if (!in1) dv[q] = d2;
					else if (d1 <= d2) dv[q] = pos < c1 ? d1 : d2;
					else {
						const unsigned long long c2 = (unsigned long long) cum[k2] - cum[k2 - 1];
						dv[q] = pos < c2 ? d2 : d1;
					}
				}
				if (lane == 0) {
					s_n = n;
					s_m = m;
					s_mad = rk[0] == rk[1] ? dv[0] : dmul(dadd(dv[0], dv[1]), 0.5);
				}
			}
		}
		__syncthreads();
		if (s_done) break;

		// --- biweight midvariance sums (statistics.c:128-150), count * term per bin ---
		const double m = s_m, mad = s_mad;
		const int bi = s_bi, bj = s_bj;
		double u = 0.0, w = 0.0;
		if (mad > 0.0) {
			const Reciprocal inv_mad9(dmul(9.0, mad));
			// thread t takes bins bi + t, bi + t + 1024, ... of the clipping window: the window (a few
			// hundred bins after the first iteration) is spread over all threads, loads coalesce
			for (int bin = bi + tid; bin < bj; bin += IKSS_THREADS) {
				uint32_t c = h[bin];
				if (c == 0) continue;
				double tu, tw;
				bwmv_term(inv_norm.divide((double) bin), m, inv_mad9, tu, tw);
				u = dadd(u, dmul((double) c, tu));
				w = dadd(w, dmul((double) c, tw));
			}
		}
		pu[tid] = u;
		pd[tid] = w;
		__syncthreads();
		for (int stride = IKSS_THREADS / 2; stride >= 1; stride >>= 1) {
			if (tid < stride) {
				pu[tid] = dadd(pu[tid], pu[tid + stride]);
				pd[tid] = dadd(pd[tid], pd[tid + stride]);
			}
			__syncthreads();
		}

		if (tid == 0) {
			double bw = 0.0;
			if (mad > 0.0) {
				double down = pd[0];
				bw = dmul((double) (long long) s_n, ddiv(pu[0], dmul(down, down)));
			}
			const double s = dsqrt(bw);
			if (s < 2E-23) {			// statistics.c:168-172
				s_loc = m;
				s_scale = 0.0;
				s_done = 1;
			} else if (ddiv(dsub(s_s0, s), s) < 10E-6) {	// :173-177
				s_loc = m;
				s_scale = dmul(0.991, s);
				s_done = 1;
			} else {				// :178-184
				s_s0 = s;
				const double xlow = dsub(m, dmul(4.0, s)), xhigh = dadd(m, dmul(4.0, s));
				int nbi = bi, nbj = bj;
				{	// first bin >= bi with d_k >= xlow
					double g = dmul(xlow, norm);
					int k = g > 3.0 ? (int) g - 2 : 0;
					if (k < bi) k = bi;
					while (k < bj && inv_norm.divide((double) k) < xlow) k++;
					nbi = k;
				}
				{	// last bin < bj with d_k <= xhigh
					double g = dmul(xhigh, norm);
					int k = g < 65533.0 ? (int) g + 2 : 65535;
					if (k > bj - 1) k = bj - 1;
					while (k >= nbi && inv_norm.divide((double) k) > xhigh) k--;
					nbj = k + 1;
				}
				s_bi = nbi;
				s_bj = nbj;
			}
		}
		__syncthreads();
		if (s_done) break;
	}

	if (tid == 0) {
		stats[0 * n_frames + frame] = dmul(s_loc, norm);	// statistics.c:287-288
		stats[1 * n_frames + frame] = dmul(s_scale, norm);
		stats[2 * n_frames + frame] = norm;
		stats[3 * n_frames + frame] = (double) (long long) n_total;
	}
}

int launch_ikss(const uint32_t *hist, int n_frames, double *stats, uint32_t *scratch, cudaStream_t stream) {
	if (n_frames < 1) return SSC_ERR_ARG;
	k_ikss<<<n_frames, IKSS_THREADS, 0, stream>>>(hist, stats, scratch, n_frames);
	return cudaGetLastError() == cudaSuccess ? SSC_OK : SSC_ERR_CUDA;
}

// ---------------------------------------------------------------------------
// K6: _compute_normalization_for_image for all frames (stacking.c:95-121,
// defaults :130-134, reference index :143-145).
// ---------------------------------------------------------------------------
__global__ void k_norm_coeffs(const double *__restrict__ stats, int n, int ref, int mode,
		double *__restrict__ coeff, int *status) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const double *location = stats, *scale_in = stats + n, *ngood = stats + 3 * n;
	double offset = 0.0, mul = 1.0, scale = 1.0;
	if (ref < 0 || ref >= n) ref = 0;
	if (mode != NO_NORM) {
		if (ngood[i] < 1.0 && status) atomicExch(status, 1);	// statistics() would return NULL (:247-250)
		const double scale0 = scale_in[ref], loc0 = location[ref];
		switch (mode) {
		default:
		case ADDITIVE_SCALING:
			scale = ddiv(scale0, scale_in[i]);
			offset = dsub(dmul(scale, location[i]), loc0);
			break;
		case ADDITIVE:
			offset = dsub(dmul(scale, location[i]), loc0);
			break;
		case MULTIPLICATIVE_SCALING:
			scale = ddiv(scale0, scale_in[i]);
			mul = ddiv(loc0, location[i]);
			break;
		case MULTIPLICATIVE:
			mul = ddiv(loc0, location[i]);
			break;
		}
	}
	coeff[i] = offset;
	coeff[n + i] = mul;
	coeff[2 * n + i] = scale;
}

int launch_norm_coeffs(const double *stats, int n_frames, int ref_image, int mode, double *coeff,
		int *status, cudaStream_t stream) {
	if (n_frames < 1) return SSC_ERR_ARG;
	k_norm_coeffs<<<(n_frames + 127) / 128, 128, 0, stream>>>(stats, n_frames, ref_image, mode, coeff, status);
	return cudaGetLastError() == cudaSuccess ? SSC_OK : SSC_ERR_CUDA;
}

}  // namespace ssc
