// ssc_stack_mid.cu -- K2/K3 for stacks of 65..1024 frames without registration shifts (the shapes of
// BASELINE.json's cfg4: 200 frames, winsorized rejection, and cfg5: 1000 frames, linear fit).
//
// The same two-pixels-per-thread packed u16x2 columns as ssc_stack_fast.cu, but a column no longer fits
// the register file, so it lives in shared memory (col[i][thread], conflict-free) and is sorted in pieces:
//   1. one asynchronous 4-byte copy per frame lands the raw samples in the column, all in flight at once;
//   2. a rolled loop normalizes them in place (exact FP64, as in the 64-frame kernel) and takes the sums;
//   3. every run of 64 samples goes through the REGISTER sorting network (64 static loads, 543 packed
//      compare-exchanges, 64 static stores) -- the network's code is instantiated once and looped over;
//   4. the sorted runs are merged pairwise with bitonic merges: the long-distance stages are rolled loops
//      over shared memory (the partner of element k of one run is element L-1-k of the other), the last
//      six stages of every 64-element block run in registers again (192 compare-exchanges);
//   5. the sorted column goes through the same per-pixel rejection code as the general kernel
//      (ssc_column.cuh).
// The block has twice as many threads as pixel pairs: in 1-4 each of a pair's two threads works on one
// half of the column (only the last merge level is shared), in 5 each takes one pixel of the pair, so
// the serial per-pixel loops (winsorization, linear fit) run one pixel per thread.
// Per pixel this is ~9 k instructions against ~50 k for the general kernel's shared-memory bitonic sort.
#include <algorithm>
#include <cstdlib>
#include "ssc_column.cuh"
#include "ssc_kernels.h"
#include "ssc_network.cuh"

namespace ssc {


// the last six stages of a bitonic merge: v[0..64) bitonic -> ascending
__device__ __forceinline__ void bitonic_merge_64(unsigned int (&v)[64]) {
#pragma unroll
	for (int d = 32; d >= 1; d >>= 1) {
#pragma unroll
		for (int i = 0; i < 64; i++) {
			if ((i & d) == 0) {
				const unsigned int lo = __vminu2(v[i], v[i + d]);
				v[i + d] = __vmaxu2(v[i], v[i + d]);
				v[i] = lo;
			}
		}
	}
}

// s_coef: offset [f], mul [NPAD + f], scale [2 NPAD + f] -- or the constants of the cheap form (ssc_network.cuh):
// multiplicative kinds a [2f], c [2f + 1]; additive kinds offset [f], c [NPAD + f], a [2 NPAD + f]
template <int KIND, int MID_THREADS, int NPAD, bool CHEAP = false>
__device__ __forceinline__ void normalize_column(unsigned int *col_t, const double *s_coef, int f_begin, int N,
		unsigned int &Sa, unsigned int &Sb, unsigned long long &Qa, unsigned long long &Qb) {
	if constexpr (CHEAP) {
		constexpr bool MULT = KIND == MULTIPLICATIVE || KIND == MULTIPLICATIVE_SCALING;
		unsigned int hi_a[4], hi_b[4];	// the high words of the 8 doubles in flight (cheap_sample)
#pragma unroll
		for (int j = 0; j < 4; j++) hi_a[j] = hi_b[j] = SSC_HI52;
#pragma unroll 1
		for (int f0 = f_begin; f0 < N; f0 += 4) {
			unsigned int pa[4], pb[4];
			double a[4], c[4], off[4];
#pragma unroll
			for (int j = 0; j < 4; j++) {
				const unsigned short *h = reinterpret_cast<const unsigned short *>(col_t + (f0 + j) * MID_THREADS);
				pa[j] = lds_u16(h);
				pb[j] = lds_u16(h + 1);
				a[j] = MULT ? s_coef[2 * (f0 + j)] : s_coef[2 * NPAD + f0 + j];
				c[j] = MULT ? s_coef[2 * (f0 + j) + 1] : s_coef[NPAD + f0 + j];
				off[j] = MULT ? 0.0 : s_coef[f0 + j];
			}
#pragma unroll
			for (int j = 0; j < 4; j++) {
				if (f0 + j >= N) break;
				const unsigned int w = cheap_pair<KIND>(pa[j], pb[j], a[j], c[j], off[j], Qa, Qb, hi_a[j], hi_b[j]);
				col_t[(f0 + j) * MID_THREADS] = w;
				Sa = __dp2a_lo(w, 0x0001u, Sa);
				Sb = __dp2a_lo(w, 0x0100u, Sb);
			}
		}
	} else {
#pragma unroll 1
	for (int f0 = f_begin; f0 < N; f0 += 4) {	// frames [f_begin, N), f_begin a multiple of 4
		unsigned int w[4];
		double off[4], mul[4], sc[4];
#pragma unroll
		for (int j = 0; j < 4; j++) {	// (the column and the coefficient arrays are padded to a multiple of 4)
			w[j] = col_t[(f0 + j) * MID_THREADS];
			off[j] = 0.0; mul[j] = 1.0; sc[j] = 1.0;
			if constexpr (KIND == ADDITIVE || KIND == ADDITIVE_SCALING) off[j] = s_coef[f0 + j];
			if constexpr (KIND == MULTIPLICATIVE || KIND == MULTIPLICATIVE_SCALING) mul[j] = s_coef[NPAD + f0 + j];
			if constexpr (KIND == ADDITIVE_SCALING || KIND == MULTIPLICATIVE_SCALING) sc[j] = s_coef[2 * NPAD + f0 + j];
		}
#pragma unroll
		for (int j = 0; j < 4; j++) {
			if (f0 + j >= N) break;
			if constexpr (KIND != NO_NORM) {
				w[j] = normalize_pair_fast<KIND>(w[j], off[j], mul[j], sc[j], Qa, Qb);
				col_t[(f0 + j) * MID_THREADS] = w[j];
			} else {
				const unsigned int pa = w[j] & 0xFFFFu, pb = w[j] >> 16;
				Qa += (unsigned long long) pa * pa;
				Qb += (unsigned long long) pb * pb;
			}
			Sa = __dp2a_lo(w[j], 0x0001u, Sa);
			Sb = __dp2a_lo(w[j], 0x0100u, Sb);
		}
	}
	}
}

// NPAD: 128 .. 1024, the column length the merges work on (sentinels 0xFFFF behind the N samples);
// MID_THREADS: pixel pairs per block (the block has 2 * MID_THREADS threads, see 5. above)
template <int NPAD, int MID_THREADS>
__global__ void __launch_bounds__(2 * MID_THREADS)
k_stack_mid(StackParams p, unsigned int n_tiles, int is_median) {
	extern __shared__ unsigned int col[];			// [NPAD][MID_THREADS] packed columns
	__shared__ double s_coef[3 * NPAD];			// offset, mul, scale
	__shared__ int s_fast_norm, s_cheap;
	__shared__ unsigned long long blk_counters[SSC_COUNTERS];
	__shared__ unsigned long long s_S[4 * MID_THREADS], s_Q[4 * MID_THREADS];	// [half of the column][pixel * MID_THREADS + pair]
	__shared__ unsigned int s_out[MID_THREADS];		// per pair: result index, bit 31 = inactive
	__shared__ unsigned char s_layer[MID_THREADS];

	const ssc_band &b = p.band;
	const int W = b.width, W2 = W >> 1, N = b.n_frames, tid = threadIdx.x;
	const bool norm = p.coeff != nullptr && p.normalize != NO_NORM;
	const bool need_sort = is_median || p.rejection != NO_REJEC;

	if (tid < SSC_COUNTERS) blk_counters[tid] = 0;
	if (tid == 0) { s_fast_norm = 1; s_cheap = p.norm_bad != nullptr; }
	__syncthreads();
	for (int f = tid; f < N; f += 2 * MID_THREADS) {
		if (!norm) break;
		// domain of the 2^52 tricks (ssc_network.cuh): every intermediate stays below 2^31
		if (!norm_domain_ok(p.normalize, p.coeff[f], p.coeff[N + f], p.coeff[2 * N + f])) s_fast_norm = 0;
		if (p.norm_bad && p.norm_bad[f]) s_cheap = 0;	// the cheap form needs k_verify_cheap_norm's all-clear for every frame
	}
	__syncthreads();
	const bool cheap = norm && s_fast_norm && s_cheap;
	const bool cheap_mult = cheap && p.normalize != ADDITIVE && p.normalize != ADDITIVE_SCALING;
	for (int f = tid; f < NPAD; f += 2 * MID_THREADS) {
		const bool have = norm && f < N;
		const double off = have ? p.coeff[f] : 0.0, mul = have ? p.coeff[N + f] : 1.0, sc = have ? p.coeff[2 * N + f] : 1.0;
		if (!cheap) {
			s_coef[f] = off;
			s_coef[NPAD + f] = mul;
			s_coef[2 * NPAD + f] = sc;
		} else {	// the cheap form's constants (layout: normalize_column)
			double a = 1.0, c = 0.0;
			switch (p.normalize) {
			case ADDITIVE: cheap_constants<ADDITIVE>(off, mul, sc, a, c); break;
			case ADDITIVE_SCALING: cheap_constants<ADDITIVE_SCALING>(off, mul, sc, a, c); break;
			case MULTIPLICATIVE: cheap_constants<MULTIPLICATIVE>(off, mul, sc, a, c); break;
			default: cheap_constants<MULTIPLICATIVE_SCALING>(off, mul, sc, a, c); break;
			}
			if (cheap_mult) {
				s_coef[2 * f] = a;
				s_coef[2 * f + 1] = c;
			} else {
				s_coef[f] = off;
				s_coef[NPAD + f] = c;
				s_coef[2 * NPAD + f] = a;
			}
		}
	}
	__syncthreads();

	const unsigned int total = (unsigned int) ((unsigned long long) b.layers * b.out_rows * W2);	// < 2^32: mid_path_eligible
	// Two threads per pixel pair: thread (pair, half) gathers, normalizes, sorts and merges the half
	// [half * NPAD/2, (half + 1) * NPAD/2) of the pair's column on its own; the last merge level is
	// shared (two barriers); then each thread takes one pixel of the pair through the rejection.
	constexpr int HALF = NPAD / 2;
	const int pair = tid & (MID_THREADS - 1), half = tid / MID_THREADS;
	const int h0 = half * HALF;					// my half of the column
	const int n_end = N < h0 + HALF ? (N > h0 ? N : h0) : h0 + HALF;	// my samples: [h0, n_end)
	unsigned int *col_t = col + pair;
	const unsigned int sbase = (unsigned int) __cvta_generic_to_shared(col_t);
	const unsigned int fstride32 = (unsigned int) (b.frame_stride * sizeof(WORD));

	for (unsigned int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
		{
		unsigned int g = tile * MID_THREADS + pair;
		const bool active = g < total;
		if (!active) g = total - 1;
		const unsigned int row = g / (unsigned int) W2;
		const int x2 = (int) (g - row * (unsigned int) W2);
		const int layer = (int) (row / (unsigned int) b.out_rows);
		const int rr = (int) (row - (unsigned int) layer * (unsigned int) b.out_rows);
		const int R = b.out_row0 + rr;				// result storage row (stacking.c:731,1600)
		const int srow = b.row_order == SSC_ROWS_SER ? b.height - 1 - R : R;
		const unsigned int toff = (unsigned int) ((((size_t) layer * b.layer_stride + (size_t) (srow - b.buf_row0) * W) >> 1) + x2);
		const unsigned int out_idx = (unsigned int) ((size_t) layer * p.out_layer_stride + (size_t) (R - p.out_buf_row0) * W + 2 * x2);
		const char *src0 = reinterpret_cast<const char *>(reinterpret_cast<const unsigned int *>(b.frames) + toff);

		// ---- 1. gather (my half of the frames) ----
#pragma unroll 4
		for (int f = h0; f < n_end; f++) {
			const char *src = src0 + (unsigned long long) ((unsigned int) f) * fstride32;
			asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(sbase + f * MID_THREADS * (int) sizeof(unsigned int)), "l"(src) : "memory");
		}
		asm volatile("cp.async.commit_group;" ::: "memory");
		for (int f = n_end; f < h0 + HALF; f++) col_t[f * MID_THREADS] = 0xFFFFFFFFu;	// sentinels sort behind the samples
		asm volatile("cp.async.wait_group 0;" ::: "memory");

		// ---- 2. normalize in place (stacking.c:1635-1652; median path :750-764), exact partial sums ----
		unsigned int Sa = 0, Sb = 0;
		unsigned long long Qa = 0, Qb = 0;
		if (norm && !s_fast_norm) {
			for (int f = h0; f < n_end; f++) {
				const unsigned int w = col_t[f * MID_THREADS];
				const unsigned int pa = normalize_sample(w & 0xFFFFu, p.normalize, s_coef[f], s_coef[NPAD + f], s_coef[2 * NPAD + f]);
				const unsigned int pb = normalize_sample(w >> 16, p.normalize, s_coef[f], s_coef[NPAD + f], s_coef[2 * NPAD + f]);
				col_t[f * MID_THREADS] = pa | (pb << 16);
				Sa += pa; Sb += pb;
				Qa += (unsigned long long) pa * pa;
				Qb += (unsigned long long) pb * pb;
			}
		} else if (cheap) {
			switch (p.normalize) {
			case ADDITIVE: normalize_column<ADDITIVE, MID_THREADS, NPAD, true>(col_t, s_coef, h0, n_end, Sa, Sb, Qa, Qb); break;
			case MULTIPLICATIVE: normalize_column<MULTIPLICATIVE, MID_THREADS, NPAD, true>(col_t, s_coef, h0, n_end, Sa, Sb, Qa, Qb); break;
			case ADDITIVE_SCALING: normalize_column<ADDITIVE_SCALING, MID_THREADS, NPAD, true>(col_t, s_coef, h0, n_end, Sa, Sb, Qa, Qb); break;
			default: normalize_column<MULTIPLICATIVE_SCALING, MID_THREADS, NPAD, true>(col_t, s_coef, h0, n_end, Sa, Sb, Qa, Qb); break;
			}
		} else {
			switch (norm ? p.normalize : NO_NORM) {
			case NO_NORM: normalize_column<NO_NORM, MID_THREADS, NPAD>(col_t, s_coef, h0, n_end, Sa, Sb, Qa, Qb); break;
			case ADDITIVE: normalize_column<ADDITIVE, MID_THREADS, NPAD>(col_t, s_coef, h0, n_end, Sa, Sb, Qa, Qb); break;
			case MULTIPLICATIVE: normalize_column<MULTIPLICATIVE, MID_THREADS, NPAD>(col_t, s_coef, h0, n_end, Sa, Sb, Qa, Qb); break;
			case ADDITIVE_SCALING: normalize_column<ADDITIVE_SCALING, MID_THREADS, NPAD>(col_t, s_coef, h0, n_end, Sa, Sb, Qa, Qb); break;
			default: normalize_column<MULTIPLICATIVE_SCALING, MID_THREADS, NPAD>(col_t, s_coef, h0, n_end, Sa, Sb, Qa, Qb); break;
			}
		}
		// partial sums of my half, per pixel: [half][pixel * MID_THREADS + pair]
		s_S[half * 2 * MID_THREADS + pair] = Sa; s_S[half * 2 * MID_THREADS + MID_THREADS + pair] = Sb;
		s_Q[half * 2 * MID_THREADS + pair] = Qa; s_Q[half * 2 * MID_THREADS + MID_THREADS + pair] = Qb;
		if (half == 0) {
			s_out[pair] = out_idx | (active ? 0u : 0x80000000u);
			s_layer[pair] = (unsigned char) layer;
		}
		}

		// the last six stages of a bitonic merge for the 64-element blocks of [first, first + len)
		auto block_merges = [&](int first, int len) {
#pragma unroll 1
			for (int c0 = first; c0 < first + len; c0 += 64) {
				unsigned int v[64];
				unsigned int *blk = col_t + c0 * MID_THREADS;
#pragma unroll
				for (int k = 0; k < 64; k++) v[k] = blk[k * MID_THREADS];
				bitonic_merge_64(v);
#pragma unroll
				for (int k = 0; k < 64; k++) blk[k * MID_THREADS] = v[k];
			}
		};
		// half-cleaners at distances d_first, d_first / 2, ... 64 over [first, first + len)
		auto half_cleaners = [&](int first, int len, int d_first, int log_d_first) {
			int log_d = log_d_first;
#pragma unroll 1
			for (int d = d_first; d >= 64; d >>= 1, log_d--) {
#pragma unroll 2
				for (int j = 0; j < len / 2; j++) {
					const int blk = j >> log_d, k = j & (d - 1);
					unsigned int *lo = col_t + (first + blk * 2 * d + k) * MID_THREADS;
					unsigned int *hi = lo + d * MID_THREADS;
					const unsigned int x = *lo, y = *hi;
					*lo = __vminu2(x, y);
					*hi = __vmaxu2(x, y);
				}
			}
		};

		if (need_sort) {
			// ---- 3. my runs of 64 through the register network ----
#pragma unroll 1
			for (int c0 = h0; c0 < n_end; c0 += 64) {
				unsigned int v[64];
				unsigned int *blk = col_t + c0 * MID_THREADS;
#pragma unroll
				for (int k = 0; k < 64; k++) v[k] = blk[k * MID_THREADS];
				sort_network<64>(v);
#pragma unroll
				for (int k = 0; k < 64; k++) blk[k * MID_THREADS] = v[k];
			}
			// ---- 4a. bitonic merges inside my half: runs of L = 64, 128, ... HALF / 2 ----
			int log_l = 6;
#pragma unroll 1
			for (int L = 64; L < HALF; L <<= 1, log_l++) {
				// first stage: element k of the lower run against element L-1-k of the upper run
#pragma unroll 2
				for (int j = 0; j < HALF / 2; j++) {
					const int blk = j >> log_l, k = j & (L - 1);
					unsigned int *lo = col_t + (h0 + blk * 2 * L + k) * MID_THREADS;
					unsigned int *hi = col_t + (h0 + blk * 2 * L + 2 * L - 1 - k) * MID_THREADS;
					const unsigned int x = *lo, y = *hi;
					*lo = __vminu2(x, y);
					*hi = __vmaxu2(x, y);
				}
				half_cleaners(h0, HALF, L >> 1, log_l - 1);
				block_merges(h0, HALF);
			}
			__syncthreads();
			// ---- 4b. the last level, L = HALF: its first stage is shared between the two threads ----
#pragma unroll 2
			for (int j = half * (HALF / 2); j < (half + 1) * (HALF / 2); j++) {
				unsigned int *lo = col_t + j * MID_THREADS;
				unsigned int *hi = col_t + (NPAD - 1 - j) * MID_THREADS;
				const unsigned int x = *lo, y = *hi;
				*lo = __vminu2(x, y);
				*hi = __vmaxu2(x, y);
			}
			__syncthreads();
			{
				int log_h = 6;
				for (int t = 64; t < HALF; t <<= 1) log_h++;	// log2(HALF)
				half_cleaners(h0, HALF, HALF >> 1, log_h - 1);	// my half is bitonic now
				block_merges(h0, HALF);
			}
		}
		__syncthreads();

		// ---- 5. per-pixel tail on the sorted column: thread t takes pixel t / MID_THREADS of pair t % MID_THREADS ----
		{
			const int pair = tid & (MID_THREADS - 1), h = tid / MID_THREADS;
			const unsigned int oi = s_out[pair];
			const bool active = !(oi & 0x80000000u);
			const int layer = s_layer[pair];
			Column c{reinterpret_cast<WORD *>(col) + 2 * pair + h, 2 * MID_THREADS};
			unsigned int res;
			if (is_median) {			// stacking.c:765-767, double -> WORD truncation
				res = (c.get((N - 1) >> 1) + c.get(N >> 1)) >> 1;
			} else {
				RejectState st{0, 0, 0, 0};
				const unsigned long long S = s_S[tid] + s_S[2 * MID_THREADS + tid], Q = s_Q[tid] + s_Q[2 * MID_THREADS + tid];
				// (n <= 1024: n Q - S^2 < 2^52, the domain of the table-driven sigma / mean of ssc_device.cuh)
				const SmallN tables{p.fit.unbias, p.fit.inv, p.fit.inv_n2};
				res = p.fit.unbias ? reject_sorted_and_mean<true>(c, N, p.rejection, p.sig_low, p.sig_high, S, Q, st, &tables, &p.fit)
						: reject_sorted_and_mean<false>(c, N, p.rejection, p.sig_low, p.sig_high, S, Q, st, nullptr, &p.fit);
				if (active) {
					if (st.rej_lo) atomicAdd(&blk_counters[layer * 2 + 0], (unsigned long long) st.rej_lo);
					if (st.rej_hi) atomicAdd(&blk_counters[layer * 2 + 1], (unsigned long long) st.rej_hi);
					if (st.guard) atomicAdd(&blk_counters[6], 1ull);
					if (st.cap) atomicAdd(&blk_counters[7], 1ull);
				}
			}
			if (active) p.out[(oi & 0x7FFFFFFFu) + h] = (WORD) res;
		}
		__syncthreads();	// the columns are overwritten by the next tile
	}
	if (p.counters && tid < SSC_COUNTERS && blk_counters[tid])
		atomicAdd(&p.counters[tid], blk_counters[tid]);
}

// Same conditions as the 64-frame register kernel (ssc_stack_fast.cu), for 65..1024 frames.
bool mid_path_eligible(const StackParams &p, int is_median) {
	const ssc_band &b = p.band;
	if (b.n_frames > 1024) return false;
	if (b.n_frames <= 64) return false;	// the register kernel (ssc_stack_fast.cu) is the better sorter there
	if (!is_median && (b.shiftx || b.shifty)) return false;
	if ((b.width & 1) || (b.frame_stride & 1) || (b.layer_stride & 1)) return false;
	if (((uintptr_t) b.frames & 3) || ((uintptr_t) p.out & 3) || (p.out_layer_stride & 1)) return false;
	if ((size_t) b.layers * p.out_layer_stride >= 0x7FFFFFF0ull) return false;	// 31-bit result indices
	if ((unsigned long long) b.layers * b.out_rows * b.width >= 0xFFFFFFF0ull) return false;
	if ((unsigned long long) b.frame_stride * sizeof(WORD) >= 0xFFFFFFF0ull) return false;	// 32-bit frame stride in bytes
	return true;
}

template <int NPAD, int MID_THREADS>
static int launch_mid(const StackParams &p, int is_median, cudaStream_t stream) {
	const ssc_band &b = p.band;
	const unsigned long long total = (unsigned long long) b.layers * b.out_rows * (b.width >> 1);
	if (total == 0) return SSC_OK;
	const size_t smem = (size_t) NPAD * MID_THREADS * sizeof(unsigned int);
	if (cudaFuncSetAttribute(k_stack_mid<NPAD, MID_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem) != cudaSuccess)
		return SSC_ERR_CUDA;
	const unsigned long long tiles = (total + MID_THREADS - 1) / MID_THREADS;
	int dev = 0, sms = 148, per_sm = 1;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_stack_mid<NPAD, MID_THREADS>, 2 * MID_THREADS, smem) != cudaSuccess || per_sm < 1)
		per_sm = 1;
	unsigned long long blocks = (unsigned long long) sms * per_sm * 4;	// looping blocks, a few per resident slot
	if (blocks > tiles) blocks = tiles;
	k_stack_mid<NPAD, MID_THREADS><<<(unsigned int) blocks, 2 * MID_THREADS, smem, stream>>>(p, (unsigned int) tiles, is_median);
	return cudaGetLastError() == cudaSuccess ? SSC_OK : SSC_ERR_CUDA;
}

// columns of 64 KB per block: 64 pixel pairs up to 256 frames, 32 up to 512, 16 up to 1024
int launch_stack_mid(const StackParams &p, int is_median, cudaStream_t stream) {
	const int n = p.band.n_frames;
	if (n <= 128) return launch_mid<128, 64>(p, is_median, stream);
	if (n <= 256) return launch_mid<256, 64>(p, is_median, stream);
	if (n <= 512) return launch_mid<512, 32>(p, is_median, stream);
	return launch_mid<1024, 16>(p, is_median, stream);
}

}  // namespace ssc
