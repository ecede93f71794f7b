// ssc_stack_fast.cu -- K2/K3 for stacks of up to 64 frames without registration
// shifts: the shape of the headline workloads (32-frame median, 64-frame
// sigma-clipped mean).
//
// One thread owns TWO horizontally adjacent pixels.  Their samples travel
// packed as u16x2 in one 32-bit register per frame -- which is exactly how two
// adjacent WORDs sit in memory, so a frame costs one coalesced 4-byte load per
// thread (128 B per warp) and no repacking.  The NP (16/32/64) registers are
// sorted by a Batcher odd-even merge network of VIMNMX.U16x2 min/max pairs:
// every compare-exchange orders both pixels at once, every index is a
// compile-time constant, nothing leaves the register file.  The sorted columns
// are then parked in shared memory (one conflict-free 4-byte store per frame)
// where the iterative, data-dependent part of the rejection (ssc_column.cuh)
// indexes them dynamically.
//
// Normalization (stacking.c:1635-1652) keeps the reference's individually
// rounded FP64 operations but avoids the slow int<->double conversion units:
// WORD -> double through the 2^52 mantissa trick, round_to_WORD through an
// add in round-toward-zero mode onto 1.5*2^52 (floor in the low word) and one
// VIMNMX.RELU clamp.  Both are exact; frames whose coefficients would leave
// the trick's domain (|value| >= 2^31) take the conversion-based form instead.
#include <cstdlib>
#include "ssc_column.cuh"
#include "ssc_kernels.h"
#include "ssc_network.cuh"
#include "ssc_leftover.cuh"

namespace ssc {


template <int NP, bool FULL, int KIND>
__device__ __forceinline__ void normalize_all(unsigned int (&v)[NP], int N, const double *s_coef,
		unsigned long long &Qa, unsigned long long &Qb) {
#pragma unroll
	for (int f = 0; f < NP; f++) {
		if (FULL || f < N) {
			double off = 0.0, mul = 1.0, sc = 1.0;
			if constexpr (KIND == ADDITIVE || KIND == ADDITIVE_SCALING) off = s_coef[f];
			if constexpr (KIND == MULTIPLICATIVE || KIND == MULTIPLICATIVE_SCALING) mul = s_coef[64 + f];
			if constexpr (KIND == ADDITIVE_SCALING || KIND == MULTIPLICATIVE_SCALING) sc = s_coef[128 + f];
			v[f] = normalize_pair_fast<KIND>(v[f], off, mul, sc, Qa, Qb);
		}
	}
}

#ifndef SSC_SIGMA_REG_PASSES
#define SSC_SIGMA_REG_PASSES 4	// clipping passes unrolled in registers before a pixel goes to the leftover list
#endif
enum { MODE_MEDIAN = 0, MODE_NOREJ = 1, MODE_SIGMA_REG = 2, MODE_QUEUE = 3 };

// FULL: n_frames == NP (no sentinel padding, no per-frame bound checks in the unrolled code)
// MODE: which tail follows the network (one per instantiation: the kernel is instruction-fetch
//       sensitive, dead alternatives are not carried along)
constexpr int FAST_THREADS = 128;

template <int NP, bool FULL, int MODE>
__global__ void __launch_bounds__(FAST_THREADS, 4)
k_stack_fast(StackParams p, LeftoverList list, unsigned int n_tiles) {
	extern __shared__ unsigned int cols[];		// [NP][FAST_THREADS] packed columns
	__shared__ double s_coef[3 * 64];
	__shared__ double s_unbias[65], s_inv_n[65], s_inv_n2[65];	// small-n tables (ssc_device.cuh)
	__shared__ int s_fast_norm;
	__shared__ unsigned long long blk_counters[SSC_COUNTERS];
	__shared__ unsigned int s_qcount;			// work queue of pixels whose first pass rejects something
	__shared__ unsigned short s_qitem[2 * FAST_THREADS], s_res[2 * FAST_THREADS];
	__shared__ unsigned char s_layer[FAST_THREADS];
	__shared__ unsigned int s_S[2 * FAST_THREADS];
	__shared__ unsigned long long s_Q[2 * FAST_THREADS];

	const ssc_band &b = p.band;
	const int W = b.width, N = FULL ? NP : b.n_frames, tid = threadIdx.x;
	const int W2 = W >> 1;
	const bool norm = p.coeff != nullptr && p.normalize != NO_NORM;

	if (tid < SSC_COUNTERS) blk_counters[tid] = 0;
	if (tid == 0) s_fast_norm = 1;
	if (tid >= 1 && tid <= 64) {
		s_unbias[tid] = tid > 1 ? ddiv((double) tid, (double) (tid - 1)) : 0.0;
		s_inv_n[tid] = ddiv(1.0, (double) tid);
		s_inv_n2[tid] = ddiv(1.0, (double) (tid * tid));
	}
	const SmallN tables{s_unbias, s_inv_n, s_inv_n2};
	__syncthreads();
	if (norm && tid < N) {
		const double off = p.coeff[tid], mul = p.coeff[N + tid], sc = p.coeff[2 * N + tid];
		s_coef[tid] = off;
		s_coef[64 + tid] = mul;
		s_coef[128 + tid] = sc;
		// domain of the 2^52 tricks: every intermediate stays below 2^31 in magnitude
		const bool additive = p.normalize == ADDITIVE || p.normalize == ADDITIVE_SCALING;
		const double bound = additive ? dadd(dmul(fabs(sc), 65535.0), fabs(off)) : dmul(dmul(fabs(sc), 65535.0), fabs(mul));
		if (!(bound < 2147483000.0)) s_fast_norm = 0;	// also catches NaN / inf
	}
	__syncthreads();

	const unsigned long long total = (unsigned long long) b.layers * b.out_rows * W2;
	// Blocks loop over tiles (the launcher bounds the grid): the tables, the coefficient copy and
	// their barriers above are paid once per block instead of once per tile.
	for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
	unsigned long long g = tile * FAST_THREADS + tid;
	const bool active = g < total;
	if (!active) g = total - 1;
	const int x2 = (int) (g % W2);
	const int rr = (int) ((g / W2) % b.out_rows);
	const int layer = (int) (g / ((unsigned long long) W2 * b.out_rows));
	const int R = b.out_row0 + rr;				// result storage row (stacking.c:731,1600)
	const int srow = b.row_order == SSC_ROWS_SER ? b.height - 1 - R : R;
	// per-thread offset inside a frame (32 bits: a frame holds < 2^32 pixel pairs) on top of a
	// block-uniform per-frame base, so the 64 address computations stay off the ALU pipe
	const unsigned int toff = (unsigned int) ((((size_t) layer * b.layer_stride + (size_t) (srow - b.buf_row0) * W) >> 1) + x2);
	const unsigned int out_idx = (unsigned int) ((size_t) layer * p.out_layer_stride + (size_t) (R - p.out_buf_row0) * W + 2 * x2);
	const unsigned int *frames32 = reinterpret_cast<const unsigned int *>(b.frames);
	const size_t fstride2 = b.frame_stride >> 1;

	// ---- gather: one packed load per frame, all issued before first use ----
	unsigned int v[NP];
	const char *ptr = reinterpret_cast<const char *>(frames32 + toff);
	const size_t fstride_bytes = fstride2 * sizeof(unsigned int);
#pragma unroll
	for (int f = 0; f < NP; f++) {
		v[f] = (FULL || f < N) ? __ldg(reinterpret_cast<const unsigned int *>(ptr)) : 0xFFFFFFFFu;
		ptr += fstride_bytes;
	}

	// ---- normalize (stacking.c:1635-1652; median path :750-764) ----
	unsigned long long Qa = 0, Qb = 0;	// exact sums of squares of the (normalized) samples
	bool have_q = false;
	if (norm) {
		if (s_fast_norm) {
			have_q = true;
			switch (p.normalize) {
			case ADDITIVE: normalize_all<NP, FULL, ADDITIVE>(v, N, s_coef, Qa, Qb); break;
			case MULTIPLICATIVE: normalize_all<NP, FULL, MULTIPLICATIVE>(v, N, s_coef, Qa, Qb); break;
			case ADDITIVE_SCALING: normalize_all<NP, FULL, ADDITIVE_SCALING>(v, N, s_coef, Qa, Qb); break;
			default: normalize_all<NP, FULL, MULTIPLICATIVE_SCALING>(v, N, s_coef, Qa, Qb); break;
			}
		} else {
			// coefficients outside the tricks' domain: conversion-based form, rolled loop over the
			// shared-memory copy (rare: keep it out of the straight-line code)
#pragma unroll
			for (int f = 0; f < NP; f++) cols[f * FAST_THREADS + tid] = v[f];
			for (int f = 0; f < N; f++) {
				unsigned int w = cols[f * FAST_THREADS + tid];
				unsigned int pa = normalize_sample(w & 0xFFFFu, p.normalize, s_coef[f], s_coef[64 + f], s_coef[128 + f]);
				unsigned int pb = normalize_sample(w >> 16, p.normalize, s_coef[f], s_coef[64 + f], s_coef[128 + f]);
				cols[f * FAST_THREADS + tid] = pa | (pb << 16);
			}
#pragma unroll
			for (int f = 0; f < NP; f++) v[f] = cols[f * FAST_THREADS + tid];
		}
	}

	unsigned int res[2];
	WORD *dst = p.out + out_idx;

	if constexpr (MODE == MODE_MEDIAN) {
		sort_network<NP>(v);			// sentinels (0xFFFF) sort behind the N samples
		unsigned int lo, hi;
		if (FULL) {				// stacking.c:765-767: truncating median, straight from registers
			lo = v[NP / 2 - 1];
			hi = v[NP / 2];
		} else {
#pragma unroll
			for (int f = 0; f < NP; f++) cols[f * FAST_THREADS + tid] = v[f];
			lo = cols[((N - 1) >> 1) * FAST_THREADS + tid];
			hi = cols[(N >> 1) * FAST_THREADS + tid];
		}
		res[0] = ((lo & 0xFFFFu) + (hi & 0xFFFFu)) >> 1;
		res[1] = ((lo >> 16) + (hi >> 16)) >> 1;
		if (active) *reinterpret_cast<unsigned int *>(dst) = res[0] | (res[1] << 16);
		continue;
	} else {
		// exact integer sums of the (normalized) samples; the sum fits 32 bits and comes straight
		// from the packed words (IDP.2A: acc + lo16 * 1 + hi16 * 0), the sum of squares needs 64
		unsigned int Sa = 0, Sb = 0;
#pragma unroll
		for (int f = 0; f < NP; f++) {
			if (FULL || f < N) {
				Sa = __dp2a_lo(v[f], 0x0001u, Sa);
				Sb = __dp2a_lo(v[f], 0x0100u, Sb);
			}
		}
		if (!have_q && MODE != MODE_NOREJ) {
#pragma unroll
			for (int f = 0; f < NP; f++) {
				if (FULL || f < N) {
					const unsigned int pa = v[f] & 0xFFFFu, pb = v[f] >> 16;
					Qa += (unsigned long long) pa * pa;
					Qb += (unsigned long long) pb * pb;
				}
			}
		}
		unsigned long long S[2] = {Sa, Sb}, Q[2] = {Qa, Qb};

		if constexpr (MODE == MODE_NOREJ) {
			// stacking.c:1790-1794 with nothing rejected: no ordering needed at all
			res[0] = mean_to_word<true>(S[0], N, &tables);
			res[1] = mean_to_word<true>(S[1], N, &tables);
			if (active) *reinterpret_cast<unsigned int *>(dst) = res[0] | (res[1] << 16);
			continue;
		} else if constexpr (MODE == MODE_SIGMA_REG) {
			// ---- sigma clipping (stacking.c:1674-1695) entirely in registers.  In a sorted stack the
			// low flags are a prefix and the high flags a suffix, the survivors stay [nl, NP - nh), and
			// each pass only looks at the sums, the two middle survivors and the next candidates at
			// both ends -- all of which sit at STATIC register indices as long as few samples go:
			// candidates v[0..2] below, v[NP-1..NP-6] above, medians within v[NP/2-3..NP/2+1].  Both
			// pixels of the pair run side by side.  Whoever needs more (a third low, a sixth high, a
			// fifth pass) is handed to the leftover list.  With NP >= 32 and at most 7 rejections the
			// reference's "N - r <= 4" break (:1684) cannot be reached here.
			static_assert(FULL && NP >= 32, "register-resident clipping needs a full stack of >= 32 frames");
			sort_network<NP>(v);
			int nl[2] = {0, 0}, nh[2] = {0, 0};
			unsigned int rl[2] = {0, 0}, rh[2] = {0, 0};
			bool done[2] = {false, false}, punt[2] = {false, false};
#pragma unroll 1
			for (int it = 0; it <= SSC_SIGMA_REG_PASSES; it++) {
				if (it == SSC_SIGMA_REG_PASSES) {	// one more pass would be needed
					punt[0] = !done[0];
					punt[1] = !done[1];
					break;
				}
#pragma unroll
				for (int h = 0; h < 2; h++) {
					if (done[h] || punt[h]) continue;
					const int Nc = NP - nl[h] - nh[h];
					const double sigma = sd_from_sums<true>(S[h], Q[h], Nc, &tables);
					const int lhs = nl[h] + ((Nc - 1) >> 1), rhs = nl[h] + (Nc >> 1);
					unsigned int ml = 0, mr = 0;
#pragma unroll
					for (int k = NP / 2 - 3; k <= NP / 2 + 1; k++) {
						const unsigned int e = h ? v[k] >> 16 : v[k] & 0xFFFFu;
						if (lhs == k) ml = e;
						if (rhs == k) mr = e;
					}
					// sigma_clipping (:1152,1156) in integers: the median is M2 / 2 with M2 an integer and the
					// sample e an integer, so (median - e) is exact in FP64 and
					//   median - e > tl  <=>  M2 - 2e > 2 tl  <=>  M2 - 2e > floor(2 tl)      (tl >= 0)
					// and likewise on the high side: one float-to-int conversion per pass and side instead
					// of a conversion, a subtraction and a compare per candidate.
					const int M2 = lhs == rhs ? 2 * (int) ml : (int) (ml + mr);
					const int TL = __double2int_rd(fmin(dmul(2.0, dmul(p.sig_low, sigma)), 1073741824.0));
					const int TH = __double2int_rd(fmin(dmul(2.0, dmul(p.sig_high, sigma)), 1073741824.0));
					int n = 0;
#pragma unroll
					for (int k = 0; k < 3; k++) {		// low candidates, in scan order
						const unsigned int e = h ? v[k] >> 16 : v[k] & 0xFFFFu;
						if (nl[h] == k && M2 - 2 * (int) e > TL) {
							nl[h]++; n++; rl[h]++;
							S[h] -= e; Q[h] -= (unsigned long long) e * e;
						}
					}
#pragma unroll
					for (int k = 0; k < 6; k++) {		// high candidates, from the top
						const unsigned int e = h ? v[NP - 1 - k] >> 16 : v[NP - 1 - k] & 0xFFFFu;
						if (nh[h] == k && 2 * (int) e - M2 > TH) {
							nh[h]++; n++; rh[h]++;
							S[h] -= e; Q[h] -= (unsigned long long) e * e;
						}
					}
					if (nl[h] == 3 || nh[h] == 6) punt[h] = true;	// the next candidate is out of the window
					else if (n == 0) done[h] = true;		// do { } while (n > 0 && N > 3)
				}
				if ((done[0] || punt[0]) && (done[1] || punt[1])) break;
			}
#pragma unroll
			for (int h = 0; h < 2; h++) {
				res[h] = mean_to_word<true>(S[h], NP - nl[h] - nh[h], &tables);
				punt[h] = punt[h] && active;
			}
			// ---- finished pixels: results and counters ----
			if (active) {
				if (!punt[0] && !punt[1]) {
					*reinterpret_cast<unsigned int *>(dst) = res[0] | (res[1] << 16);
				} else {
					if (!punt[0]) dst[0] = (WORD) res[0];
					if (!punt[1]) dst[1] = (WORD) res[1];
				}
				const unsigned int cl = (punt[0] ? 0 : rl[0]) + (punt[1] ? 0 : rl[1]);
				const unsigned int ch = (punt[0] ? 0 : rh[0]) + (punt[1] ? 0 : rh[1]);
				if (cl) atomicAdd(&blk_counters[layer * 2 + 0], (unsigned long long) cl);
				if (ch) atomicAdd(&blk_counters[layer * 2 + 1], (unsigned long long) ch);
			}
			// ---- the others: to the leftover list (one atomic per warp), or in place if it is full ----
			bool in_place = false;
#pragma unroll
			for (int h = 0; h < 2; h++) {
				const unsigned int mask = __ballot_sync(0xffffffffu, punt[h]);
				if (mask == 0) continue;
				const int lane = tid & 31;
				unsigned int base = 0;
				if (lane == __ffs(mask) - 1) base = atomicAdd(list.count, (unsigned int) __popc(mask));
				base = __shfl_sync(0xffffffffu, base, __ffs(mask) - 1);
				if (punt[h]) {
					const unsigned int slot = base + __popc(mask & ((1u << lane) - 1));
					if (slot < list.capacity) {
						uint4 *d = list.cols + (size_t) slot * (NP / 8);
#pragma unroll
						for (int q = 0; q < NP / 8; q++) {
							const unsigned int sel = h ? 0x7632u : 0x5410u;
							d[q] = make_uint4(__byte_perm(v[8 * q], v[8 * q + 1], sel), __byte_perm(v[8 * q + 2], v[8 * q + 3], sel),
									__byte_perm(v[8 * q + 4], v[8 * q + 5], sel), __byte_perm(v[8 * q + 6], v[8 * q + 7], sel));
						}
						list.meta[slot] = make_uint2(out_idx + h, (unsigned int) layer);
						punt[h] = false;
					} else {
						in_place = true;
					}
				}
			}
			if (__syncthreads_or(in_place)) {	// list full (adversarial data): finish in place, slowly
#pragma unroll
				for (int f = 0; f < NP; f++) cols[f * FAST_THREADS + tid] = v[f];
#pragma unroll 1
				for (int h = 0; h < 2; h++) {
					if (!punt[h]) continue;
					Column c{reinterpret_cast<WORD *>(cols) + 2 * tid + h, 2 * FAST_THREADS};
					unsigned long long s0 = 0, q0 = 0;
					for (int f = 0; f < NP; f++) { const unsigned int e = c.get(f); s0 += e; q0 += sq(e); }
					RejectState ist{0, 0, 0, 0};
					dst[h] = (WORD) reject_sorted_and_mean<true>(c, NP, SIGMA, p.sig_low, p.sig_high, s0, q0, ist, &tables);
					if (ist.rej_lo) atomicAdd(&blk_counters[layer * 2 + 0], (unsigned long long) ist.rej_lo);
					if (ist.rej_hi) atomicAdd(&blk_counters[layer * 2 + 1], (unsigned long long) ist.rej_hi);
					if (ist.guard) atomicAdd(&blk_counters[6], 1ull);
					if (ist.cap) atomicAdd(&blk_counters[7], 1ull);
				}
			}
		} else {	// MODE_QUEUE: every other rejection / stack size, through the block's work queue
			sort_network<NP>(v);
#pragma unroll
			for (int f = 0; f < NP; f++)
				if (FULL || f < N) cols[f * FAST_THREADS + tid] = v[f];
			WORD *cols16 = reinterpret_cast<WORD *>(cols);

			// ---- phase 1: can the first pass flag anything at all?  Flags are monotone in the
			// sample value (low flags a prefix, high flags a suffix of the sorted stack), so when
			// neither the smallest nor the largest sample is flagged nothing is, the reference's
			// do-while ends after one pass (n == 0) and the result is the plain mean.
			// (with N <= 4 the reference's "N - r <= 4" break is taken at the very first sample:
			// that pixel must go through the full pass so that the event is counted)
			const bool quick = (p.rejection == SIGMA && N > 4) || p.rejection == SIGMEDIAN || p.rejection == PERCENTILE;
			const unsigned int lo_w = v[0];
			const unsigned int hi_w = FULL ? v[NP - 1] : cols[(N - 1) * FAST_THREADS + tid];
			const unsigned int ml_w = FULL ? v[(NP - 1) / 2] : cols[((N - 1) >> 1) * FAST_THREADS + tid];
			const unsigned int mr_w = FULL ? v[NP / 2] : cols[(N >> 1) * FAST_THREADS + tid];
			bool pending[2];
#pragma unroll
			for (int h = 0; h < 2; h++) {
				const unsigned int lo = h ? lo_w >> 16 : lo_w & 0xFFFFu, hi = h ? hi_w >> 16 : hi_w & 0xFFFFu;
				const unsigned int ml = h ? ml_w >> 16 : ml_w & 0xFFFFu, mr = h ? mr_w >> 16 : mr_w & 0xFFFFu;
				const double median = (N & 1) ? (double) (int) ml : dmul((double) (int) (ml + mr), 0.5);
				bool flagged = true;
				if (quick) {
					if (p.rejection == PERCENTILE) {
						flagged = clip_percentile(lo, p.sig_low, p.sig_high, median) != 0
								|| clip_percentile(hi, p.sig_low, p.sig_high, median) != 0;
					} else {
						const double sigma = sd_from_sums<true>(S[h], Q[h], N, &tables);
						flagged = clip_sigma(lo, p.sig_low, p.sig_high, sigma, median) != 0
								|| clip_sigma(hi, p.sig_low, p.sig_high, sigma, median) != 0;
					}
				}
				pending[h] = flagged && active;	// padding threads of the last block own no pixel
				res[h] = mean_to_word<true>(S[h], N, &tables);
			}
			if (tid == 0) s_qcount = 0;
			s_layer[tid] = (unsigned char) layer;
			__syncthreads();
#pragma unroll
			for (int h = 0; h < 2; h++) {
				if (pending[h]) {
					const unsigned int slot = atomicAdd(&s_qcount, 1u);
					s_qitem[slot] = (unsigned short) (2 * tid + h);
					s_S[2 * tid + h] = (unsigned int) S[h];
					s_Q[2 * tid + h] = Q[h];
				}
			}
			__syncthreads();

			// ---- phase 2: the pending pixels, densely packed over the block's threads (their sorted
			// columns sit in shared memory, any thread can run them) ----
			const unsigned int n_items = s_qcount;
			for (unsigned int i = tid; i < n_items; i += FAST_THREADS) {
				const unsigned int item = s_qitem[i];
				Column c{cols16 + item, 2 * FAST_THREADS};
				RejectState ist{0, 0, 0, 0};
				const unsigned int r = reject_sorted_and_mean<true>(c, N, p.rejection, p.sig_low, p.sig_high,
						s_S[item], s_Q[item], ist, &tables);
				s_res[item] = (unsigned short) r;
				const int il = s_layer[item >> 1];
				if (ist.rej_lo) atomicAdd(&blk_counters[il * 2 + 0], (unsigned long long) ist.rej_lo);
				if (ist.rej_hi) atomicAdd(&blk_counters[il * 2 + 1], (unsigned long long) ist.rej_hi);
				if (ist.guard) atomicAdd(&blk_counters[6], 1ull);
				if (ist.cap) atomicAdd(&blk_counters[7], 1ull);
			}
			__syncthreads();
			if (pending[0]) res[0] = s_res[2 * tid];
			if (pending[1]) res[1] = s_res[2 * tid + 1];
			if (active) *reinterpret_cast<unsigned int *>(dst) = res[0] | (res[1] << 16);
		}
	}
	if (MODE == MODE_QUEUE || MODE == MODE_SIGMA_REG) __syncthreads();	// shared-memory columns / queue are reused by the next tile
	}	// tiles
	__syncthreads();
	if (p.counters && tid < SSC_COUNTERS && blk_counters[tid])
		atomicAdd(&p.counters[tid], blk_counters[tid]);
}

// The register-network kernel applies when two adjacent pixels can be fetched as one
// aligned 32-bit word from every frame and nothing shifts the frames against each other.
bool fast_path_eligible(const StackParams &p, int is_median) {
	const ssc_band &b = p.band;
	if (b.n_frames < 2 || b.n_frames > 64) return false;
	if (!is_median && (b.shiftx || b.shifty)) return false;
	if ((b.width & 1) || (b.frame_stride & 1) || (b.layer_stride & 1)) return false;
	if (((uintptr_t) b.frames & 3) || ((uintptr_t) p.out & 3) || (p.out_layer_stride & 1)) return false;
	if ((size_t) b.layers * p.out_layer_stride >= 0xFFFFFFF0ull) return false;	// 32-bit result indices
	return true;
}

template <int NP, bool FULL, int MODE>
static int launch_fast_cfg(const StackParams &p, cudaStream_t stream) {
	const ssc_band &b = p.band;
	const unsigned long long total = (unsigned long long) b.layers * b.out_rows * (b.width >> 1);
	if (total == 0) return SSC_OK;
	const size_t smem = (size_t) NP * FAST_THREADS * sizeof(unsigned int);
	if (cudaFuncSetAttribute(k_stack_fast<NP, FULL, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem) != cudaSuccess)
		return SSC_ERR_CUDA;
	const unsigned long long tiles = (total + FAST_THREADS - 1) / FAST_THREADS;
	if (tiles > 0x7FFFFFFFull) return SSC_ERR_ARG;
	unsigned long long blocks = tiles;
	{	// a bounded grid of looping blocks: 16 per resident slot measured best (SSC_FAST_WAVES overrides)
		int dev = 0, sms = 148, per_sm = 1;
		cudaGetDevice(&dev);
		cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
		if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_stack_fast<NP, FULL, MODE>, FAST_THREADS, smem) != cudaSuccess || per_sm < 1)
			per_sm = 1;
		static int waves = 0;
		if (!waves) { const char *e = getenv("SSC_FAST_WAVES"); waves = e ? atoi(e) : 16; if (waves < 1) waves = 1; }
		const unsigned long long cap = (unsigned long long) sms * per_sm * waves;
		if (blocks > cap) blocks = cap;
	}
	LeftoverList list{nullptr, 0, nullptr, nullptr};
	if (MODE == MODE_SIGMA_REG) {
		int rc = leftover_list(NP, total * 2, list, stream);
		if (rc) return rc;
	}
	k_stack_fast<NP, FULL, MODE><<<(unsigned int) blocks, FAST_THREADS, smem, stream>>>(p, list, (unsigned int) tiles);
	if (cudaGetLastError() != cudaSuccess) return SSC_ERR_CUDA;
	if constexpr (MODE == MODE_SIGMA_REG) return launch_leftover<NP>(p, list, stream);
	return SSC_OK;
}

template <int NP>
static int launch_fast_np(const StackParams &p, int is_median, cudaStream_t stream) {
	const bool full = p.band.n_frames == NP;
	if (is_median)
		return full ? launch_fast_cfg<NP, true, MODE_MEDIAN>(p, stream) : launch_fast_cfg<NP, false, MODE_MEDIAN>(p, stream);
	if (p.rejection == NO_REJEC)
		return full ? launch_fast_cfg<NP, true, MODE_NOREJ>(p, stream) : launch_fast_cfg<NP, false, MODE_NOREJ>(p, stream);
	if constexpr (NP >= 32) {
		// (non-negative parameters: a sample above the median is then never "low", one below never "high")
		if (full && p.rejection == SIGMA && p.sig_low >= 0.0 && p.sig_high >= 0.0)
			return launch_fast_cfg<NP, true, MODE_SIGMA_REG>(p, stream);
	}
	return full ? launch_fast_cfg<NP, true, MODE_QUEUE>(p, stream) : launch_fast_cfg<NP, false, MODE_QUEUE>(p, stream);
}

int launch_stack_fast(const StackParams &p, int is_median, cudaStream_t stream) {
	const int n = p.band.n_frames;
	if (n <= 16) return launch_fast_np<16>(p, is_median, stream);
	if (n <= 32) return launch_fast_np<32>(p, is_median, stream);
	return launch_fast_np<64>(p, is_median, stream);
}

}  // namespace ssc
