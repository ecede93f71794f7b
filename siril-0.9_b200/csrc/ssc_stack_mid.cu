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
// Only n_alloc = N rounded up to 8 rows of the column are stored: the sentinels that pad it to NPAD are VIRTUAL.  A
// compare-exchange with a +infinity partner at the higher index is a no-op and merges never move a sentinel down, so
// every step simply skips the pairs whose upper element lies at or beyond n_alloc (200 frames: 51 KB per block instead
// of 64 KB -- four resident blocks per SM instead of three -- and a fifth of the merge work gone).
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

// Two coefficients per frame (k0[f], k1[f]; the arrays are padded to a multiple of 4 frames):
//   exact form   additive kinds: k0 = offset, k1 = scale;   multiplicative kinds: k0 = mul, k1 = scale
//   cheap form   additive kinds: k0 = offset, k1 = a (c = -2^52 a on the fly);   multiplicative kinds: k0 = a, k1 = c
// (ssc_network.cuh; scale is 1.0 in ADDITIVE / MULTIPLICATIVE, stacking.c:130-134, and is not read there)
template <int KIND, int MID_THREADS, bool CHEAP = false>
__device__ __forceinline__ void normalize_column(unsigned int *col_t, const double *k0, const double *k1, int f_begin, int N,
		unsigned int &Sa, unsigned int &Sb, unsigned long long &Qa, unsigned long long &Qb) {
	constexpr bool MULT = KIND == MULTIPLICATIVE || KIND == MULTIPLICATIVE_SCALING;
	constexpr bool SCALING = KIND == ADDITIVE_SCALING || KIND == MULTIPLICATIVE_SCALING;
	if constexpr (CHEAP) {
		unsigned int hi_a[4], hi_b[4];	// the high words of the 8 doubles in flight (cheap_sample)
#pragma unroll
		for (int j = 0; j < 4; j++) hi_a[j] = hi_b[j] = SSC_HI52;
#pragma unroll 1
		for (int f0 = f_begin; f0 < N; f0 += 4) {	// frames [f_begin, N), f_begin a multiple of 4
			unsigned int pa[4], pb[4];
			double a[4], c[4], off[4];
#pragma unroll
			for (int j = 0; j < 4; j++) {
				if (f0 + j < N) {
					const unsigned short *h = reinterpret_cast<const unsigned short *>(col_t + (f0 + j) * MID_THREADS);
					pa[j] = lds_u16(h);
					pb[j] = lds_u16(h + 1);
				} else {
					pa[j] = pb[j] = 0;
				}
				if constexpr (MULT) {
					a[j] = k0[f0 + j]; c[j] = k1[f0 + j]; off[j] = 0.0;
				} else {
					off[j] = k0[f0 + j];
					a[j] = SCALING ? k1[f0 + j] : 1.0;
					c[j] = dmul(-SSC_TWO52, a[j]);
				}
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
		for (int f0 = f_begin; f0 < N; f0 += 4) {
			unsigned int w[4];
			double off[4], mul[4], sc[4];
#pragma unroll
			for (int j = 0; j < 4; j++) {
				w[j] = f0 + j < N ? col_t[(f0 + j) * MID_THREADS] : 0u;	// (rows at or beyond N may not exist)
				off[j] = 0.0; mul[j] = 1.0; sc[j] = 1.0;
				if constexpr (KIND == ADDITIVE || KIND == ADDITIVE_SCALING) off[j] = k0[f0 + j];
				if constexpr (MULT) mul[j] = k0[f0 + j];
				if constexpr (SCALING) sc[j] = k1[f0 + j];
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
k_stack_mid(StackParams p, unsigned int n_tiles, int is_median, int n_alloc) {
	// dynamic shared memory (sized for the launch's frame count: a fourth block per SM fits at 200 frames):
	//   col   [n_alloc][MID_THREADS] packed columns (rows n_alloc .. NPAD-1: virtual sentinels)
	//   k0,k1 [n_c] doubles, n_c = N rounded up to 4: the two coefficients per frame (normalize_column)
	//   x_q   [2 MID_THREADS] u64, x_s [2 MID_THREADS] u32: the partial sums one thread of a pair hands the other
	extern __shared__ __align__(16) unsigned int col[];
	__shared__ int s_fast_norm, s_cheap;
	__shared__ unsigned long long blk_counters[SSC_COUNTERS];
	__shared__ unsigned int s_out[MID_THREADS];		// per pair: result index, bit 31 = inactive
	__shared__ unsigned char s_layer[MID_THREADS];

	const ssc_band &b = p.band;
	const int W = b.width, W2 = W >> 1, N = b.n_frames, tid = threadIdx.x;
	const bool norm = p.coeff != nullptr && p.normalize != NO_NORM;
	const bool need_sort = is_median || p.rejection != NO_REJEC;
	const int n_c = (N + 3) & ~3;
	double *const k0 = reinterpret_cast<double *>(col + (size_t) n_alloc * MID_THREADS), *const k1 = k0 + n_c;
	unsigned long long *const x_q = reinterpret_cast<unsigned long long *>(k1 + n_c);
	unsigned int *const x_s = reinterpret_cast<unsigned int *>(x_q + 2 * MID_THREADS);

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
	const bool additive = p.normalize == ADDITIVE || p.normalize == ADDITIVE_SCALING;
	for (int f = tid; f < n_c; f += 2 * MID_THREADS) {
		const bool have = norm && f < N;
		const double off = have ? p.coeff[f] : 0.0, mul = have ? p.coeff[N + f] : 1.0, sc = have ? p.coeff[2 * N + f] : 1.0;
		double v0 = additive ? off : mul, v1 = sc;	// the exact form's pair (normalize_column)
		if (cheap) {
			double a = 1.0, c = 0.0;
			switch (p.normalize) {
			case ADDITIVE: cheap_constants<ADDITIVE>(off, mul, sc, a, c); break;
			case ADDITIVE_SCALING: cheap_constants<ADDITIVE_SCALING>(off, mul, sc, a, c); break;
			case MULTIPLICATIVE: cheap_constants<MULTIPLICATIVE>(off, mul, sc, a, c); break;
			default: cheap_constants<MULTIPLICATIVE_SCALING>(off, mul, sc, a, c); break;
			}
			if (additive) { v0 = off; v1 = a; } else { v0 = a; v1 = c; }
		}
		k0[f] = v0;
		k1[f] = v1;
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
		unsigned int keep_s = 0;
		unsigned long long keep_q = 0;
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
		for (int f = n_end; f < h0 + HALF && f < n_alloc; f++) col_t[f * MID_THREADS] = 0xFFFFFFFFu;	// sentinels sort behind the samples
		asm volatile("cp.async.wait_group 0;" ::: "memory");

		// ---- 2. normalize in place (stacking.c:1635-1652; median path :750-764), exact partial sums ----
		unsigned int Sa = 0, Sb = 0;
		unsigned long long Qa = 0, Qb = 0;
		if (norm && !s_fast_norm) {
			for (int f = h0; f < n_end; f++) {	// (s_fast_norm == 0: k0 / k1 hold the exact form's pair)
				const unsigned int w = col_t[f * MID_THREADS];
				const double off = additive ? k0[f] : 0.0, mul = additive ? 1.0 : k0[f], sc = k1[f];
				const unsigned int pa = normalize_sample(w & 0xFFFFu, p.normalize, off, mul, sc);
				const unsigned int pb = normalize_sample(w >> 16, p.normalize, off, mul, sc);
				col_t[f * MID_THREADS] = pa | (pb << 16);
				Sa += pa; Sb += pb;
				Qa += (unsigned long long) pa * pa;
				Qb += (unsigned long long) pb * pb;
			}
		} else if (cheap) {
			switch (p.normalize) {
			case ADDITIVE: normalize_column<ADDITIVE, MID_THREADS, true>(col_t, k0, k1, h0, n_end, Sa, Sb, Qa, Qb); break;
			case MULTIPLICATIVE: normalize_column<MULTIPLICATIVE, MID_THREADS, true>(col_t, k0, k1, h0, n_end, Sa, Sb, Qa, Qb); break;
			case ADDITIVE_SCALING: normalize_column<ADDITIVE_SCALING, MID_THREADS, true>(col_t, k0, k1, h0, n_end, Sa, Sb, Qa, Qb); break;
			default: normalize_column<MULTIPLICATIVE_SCALING, MID_THREADS, true>(col_t, k0, k1, h0, n_end, Sa, Sb, Qa, Qb); break;
			}
		} else {
			switch (norm ? p.normalize : NO_NORM) {
			case NO_NORM: normalize_column<NO_NORM, MID_THREADS>(col_t, k0, k1, h0, n_end, Sa, Sb, Qa, Qb); break;
			case ADDITIVE: normalize_column<ADDITIVE, MID_THREADS>(col_t, k0, k1, h0, n_end, Sa, Sb, Qa, Qb); break;
			case MULTIPLICATIVE: normalize_column<MULTIPLICATIVE, MID_THREADS>(col_t, k0, k1, h0, n_end, Sa, Sb, Qa, Qb); break;
			case ADDITIVE_SCALING: normalize_column<ADDITIVE_SCALING, MID_THREADS>(col_t, k0, k1, h0, n_end, Sa, Sb, Qa, Qb); break;
			default: normalize_column<MULTIPLICATIVE_SCALING, MID_THREADS>(col_t, k0, k1, h0, n_end, Sa, Sb, Qa, Qb); break;
			}
		}
		// In the tail thread (pair, half) takes pixel `half` of the pair: it keeps its own partial sums of that pixel and
		// hands the other pixel's to its partner
		keep_s = half ? Sb : Sa; keep_q = half ? Qb : Qa;
		x_s[tid] = half ? Sa : Sb; x_q[tid] = half ? Qa : Qb;
		if (half == 0) {
			s_out[pair] = out_idx | (active ? 0u : 0x80000000u);
			s_layer[pair] = (unsigned char) layer;
		}
		}

		// 64 rows of the column into registers and back; rows at or beyond n_alloc are the virtual sentinels
		auto load_64 = [&](unsigned int (&v)[64], int c0) {
			const unsigned int *blk = col_t + c0 * MID_THREADS;
			if (c0 + 64 <= n_alloc) {
#pragma unroll
				for (int k = 0; k < 64; k++) v[k] = blk[k * MID_THREADS];
			} else {
#pragma unroll
				for (int k = 0; k < 64; k++) v[k] = c0 + k < n_alloc ? blk[k * MID_THREADS] : 0xFFFFFFFFu;
			}
		};
		auto store_64 = [&](const unsigned int (&v)[64], int c0) {
			unsigned int *blk = col_t + c0 * MID_THREADS;
			if (c0 + 64 <= n_alloc) {
#pragma unroll
				for (int k = 0; k < 64; k++) blk[k * MID_THREADS] = v[k];
			} else {
#pragma unroll
				for (int k = 0; k < 64; k++) if (c0 + k < n_alloc) blk[k * MID_THREADS] = v[k];
			}
		};
		// half-cleaners at distances d_first, d_first / 2, ... 64 over [first, first + len)
		auto half_cleaners = [&](int first, int len, int d_first) {
#pragma unroll 1
			for (int d = d_first; d >= 64; d >>= 1) {
#pragma unroll 1
				for (int sb = first; sb < first + len; sb += 2 * d) {
					// pairs (i, i + d), i in [sb, sb + d): those whose partner is a virtual sentinel (i + d >= n_alloc) are no-ops
					const int i_end = min(sb + d, n_alloc - d);
#pragma unroll 2
					for (int i = sb; i < i_end; i++) {
						unsigned int *lo = col_t + i * MID_THREADS;
						unsigned int *hi = lo + d * MID_THREADS;
						const unsigned int x = *lo, y = *hi;
						*lo = __vminu2(x, y);
						*hi = __vmaxu2(x, y);
					}
				}
			}
		};

		if (need_sort) {
			// One loop over the merge levels with ONE instance of the 64-row register step (the kernel is instruction-fetch
			// bound: every further inlined copy of the networks costs more than the branch).  "Level 32" is step 3, the runs
			// of 64 through the sorting network; levels 64 .. HALF / 2 are step 4a, bitonic merges inside my half (first stage:
			// element k of the lower run against element L-1-k of the upper run); level HALF is 4b, whose first stage the
			// pair's two threads share.
#pragma unroll 1
			for (int L = 32; L <= HALF; L <<= 1) {
				if (L >= 64 && L < HALF) {
#pragma unroll 1
					for (int base = h0; base < h0 + HALF; base += 2 * L) {
						// pairs (base + k, base + 2L-1 - k): no-ops while the upper element is a virtual sentinel (>= n_alloc)
#pragma unroll 2
						for (int k = max(0, base + 2 * L - n_alloc); k < L; k++) {
							unsigned int *lo = col_t + (base + k) * MID_THREADS;
							unsigned int *hi = col_t + (base + 2 * L - 1 - k) * MID_THREADS;
							const unsigned int x = *lo, y = *hi;
							*lo = __vminu2(x, y);
							*hi = __vmaxu2(x, y);
						}
					}
				} else if (L == HALF) {
					__syncthreads();
#pragma unroll 2
					for (int j = max(half * (HALF / 2), NPAD - n_alloc); j < (half + 1) * (HALF / 2); j++) {	// (j < NPAD - n_alloc: the partner is a virtual sentinel)
						unsigned int *lo = col_t + j * MID_THREADS;
						unsigned int *hi = col_t + (NPAD - 1 - j) * MID_THREADS;
						const unsigned int x = *lo, y = *hi;
						*lo = __vminu2(x, y);
						*hi = __vmaxu2(x, y);
					}
					__syncthreads();
				}
				if (L >= 64) half_cleaners(h0, HALF, L >> 1);	// (my half is bitonic now)
				// the 64-row blocks of my half: sorted (level 32) or taken through the last six stages of the bitonic merge
				const int c_end = L == 32 ? n_end : min(h0 + HALF, n_alloc);
#pragma unroll 1
				for (int c0 = h0; c0 < c_end; c0 += 64) {
					unsigned int v[64];
					load_64(v, c0);
					if (L == 32) sort_network<64>(v);
					else bitonic_merge_64(v);
					store_64(v, c0);
				}
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
				const unsigned long long S = (unsigned long long) keep_s + x_s[tid ^ MID_THREADS], Q = keep_q + x_q[tid ^ MID_THREADS];
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
	const int n_alloc = std::min(NPAD, (b.n_frames + 7) & ~7);	// rows of the column that are stored (the rest: virtual sentinels)
	auto smem_for = [](int rows, int frames) {	// the kernel's dynamic shared memory: columns, 2 coefficients per frame, sums exchange
		return (size_t) rows * MID_THREADS * sizeof(unsigned int) + (size_t) 2 * ((frames + 3) & ~3) * sizeof(double)
				+ (size_t) 2 * MID_THREADS * (sizeof(unsigned long long) + sizeof(unsigned int));
	};
	static int pad = -1;	// SSC_MID_SMEM_PAD: extra dynamic shared memory per block (occupancy experiments)
	if (pad < 0) { const char *e = getenv("SSC_MID_SMEM_PAD"); pad = e ? atoi(e) : 0; if (pad < 0) pad = 0; }
	const size_t smem = std::min(smem_for(n_alloc, b.n_frames) + (size_t) pad, smem_for(NPAD, NPAD));
	if (cudaFuncSetAttribute(k_stack_mid<NPAD, MID_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem_for(NPAD, NPAD)) != cudaSuccess)
		return SSC_ERR_CUDA;
	const unsigned long long tiles = (total + MID_THREADS - 1) / MID_THREADS;
	int dev = 0, sms = 148, per_sm = 1;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_stack_mid<NPAD, MID_THREADS>, 2 * MID_THREADS, smem) != cudaSuccess || per_sm < 1)
		per_sm = 1;
	unsigned long long blocks = (unsigned long long) sms * per_sm * 4;	// looping blocks, a few per resident slot
	if (blocks > tiles) blocks = tiles;
	k_stack_mid<NPAD, MID_THREADS><<<(unsigned int) blocks, 2 * MID_THREADS, smem, stream>>>(p, (unsigned int) tiles, is_median, n_alloc);
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
