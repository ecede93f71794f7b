// ssc_stack_warp.cu -- K2/K3, one WARP per pixel pair: linear-fit clipping for 8..1024 frames
// (BASELINE.json's cfg5: 1000 frames) and the median of 513..1024 frames, without registration shifts.
//
// The linear fit (stacking.c:1750-1784) is the one rejection whose cost per iteration is proportional
// to the whole column: a least-squares line through the SORTED samples, the mean absolute residual,
// one flag per sample.  The reference (gsl_fit_linear) evaluates it as 4 N strictly sequential
// running-mean updates; one thread per pixel doing the same waits on a single FP64 dependency chain
// (408 ms for cfg5 in round 1's form).  Here a warp owns the column:
//   1. lane l loads the pair's samples of frames l, l+32, ... (4 bytes per frame; the warps of the grid work
//      on consecutive pairs at the same time, so the sectors are shared through the L2) and normalizes them
//      (exact FP64);
//   2. the column is sorted in registers as packed u16x2 (both pixels at once): a Batcher network on
//      each lane's R samples, then bitonic merges across lanes with shuffles (first stage against the
//      reversed partner, so every run stays ascending) -- lane l ends up with sorted samples
//      [l R, (l+1) R);
//   3. per pixel, every clipping iteration is two passes over the lane's R samples plus a few warp
//      reductions: the line comes from EXACT integer sums (sum y, sum i y: slope and intercept carry
//      one rounding instead of 4 N), the residuals are summed in FP64, and each sample is compared
//      with the two thresholds widened and narrowed by a safety margin.  If every comparison is decided
//      outside the margin -- which covers the rounding of both formulations by three orders of
//      magnitude -- the flags are the ones the reference's arithmetic produces, and the result is
//      bit-identical to it (oracle/siril_oracle.c restates both forms; tests/test_oracle_kat.py).
//      A PASS with a comparison inside the margin (an exact tie, a vanishing sigma) takes its line and sigma from
//      gsl_fit_linear's literal recurrences instead (ssc_column.cuh:fit_line_literal, one lane, out of line), and its
//      flags from line_clipping itself, so no result ever depends on the screening arithmetic;
//   4. the reference's `if (N - r <= 4) break;` (stacking.c:1771) is applied from prefix counts of the
//      flags: the samples behind the break keep flag 0, the event is counted (guard_fired).
#include <algorithm>
#include "ssc_column.cuh"
#include "ssc_kernels.h"
#include "ssc_leftover.cuh"
#include "ssc_network.cuh"

namespace ssc {

constexpr int WARP_PAIRS = 8;		// warps per block = pixel pairs per tile
constexpr unsigned int FULL = 0xffffffffu;

// ---------------------------------------------------------------------------------------------
// sorting: R packed samples per lane, 32 R per warp
// ---------------------------------------------------------------------------------------------
template <int R>
__device__ __forceinline__ void lane_sort(unsigned int (&v)[R]) {
	if constexpr (R == 32 || R == 16) {
		sort_network<R>(v);
	} else {
#pragma unroll
		for (int k = 2; k <= R; k <<= 1)
#pragma unroll
			for (int j = k >> 1; j > 0; j >>= 1)
#pragma unroll
				for (int i = 0; i < R; i++) {
					const int l = i ^ j;
					if (l > i) {
						const unsigned int lo = __vminu2(v[i], v[l]), hi = __vmaxu2(v[i], v[l]);
						if ((i & k) == 0) { v[i] = lo; v[l] = hi; } else { v[i] = hi; v[l] = lo; }
					}
				}
	}
}

template <int R>
__device__ __forceinline__ void warp_sort(unsigned int (&v)[R], const int lane) {
	lane_sort<R>(v);
	// (the five levels are ONE rolled loop: unrolled they were 3 k of the kernel's 14 k instructions, streamed through the
	// instruction cache once per pixel pair)
#pragma unroll 1
	for (int m = 1; m <= 16; m <<= 1) {	// runs of m lanes -> runs of 2 m lanes
		{	// element k of the lower run against element 2 m R - 1 - k of the upper run
			unsigned int q[R];
#pragma unroll
			for (int r = 0; r < R; r++) q[r] = __shfl_xor_sync(FULL, v[R - 1 - r], 2 * m - 1);
			const bool upper = (lane & m) != 0;
#pragma unroll
			for (int r = 0; r < R; r++) v[r] = upper ? __vmaxu2(v[r], q[r]) : __vminu2(v[r], q[r]);
		}
#pragma unroll 1
		for (int d = m >> 1; d >= 1; d >>= 1) {	// half-cleaners between lanes
			const bool upper = (lane & d) != 0;
#pragma unroll
			for (int r = 0; r < R; r++) {
				const unsigned int q = __shfl_xor_sync(FULL, v[r], d);
				v[r] = upper ? __vmaxu2(v[r], q) : __vminu2(v[r], q);
			}
		}
#pragma unroll
		for (int d = R >> 1; d >= 1; d >>= 1)	// half-cleaners inside the lane
#pragma unroll
			for (int i = 0; i < R; i++)
				if ((i & d) == 0) {
					const unsigned int lo = __vminu2(v[i], v[i + d]);
					v[i + d] = __vmaxu2(v[i], v[i + d]);
					v[i] = lo;
				}
	}
}

__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
	for (int o = 16; o >= 1; o >>= 1) x = dadd(x, __shfl_xor_sync(FULL, x, o));	// a + b == b + a: every lane gets the same bits
	return x;
}

__device__ __forceinline__ int warp_sum(int x) {
	return (int) __reduce_add_sync(FULL, (unsigned int) x);
}

// acc += x where cond != 0: a predicated add, never a branch (the per-sample loops below are straight-line code)
__device__ __forceinline__ void add_if(double &acc, const double x, const unsigned int cond) {
	asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p add.rn.f64 %0, %0, %1;\n\t}" : "+d"(acc) : "d"(x), "r"(cond));
}

// inclusive prefix sum over the lanes
__device__ __forceinline__ int warp_scan(int x, const int lane) {
#pragma unroll
	for (int o = 1; o <= 16; o <<= 1) {
		const int t = __shfl_up_sync(FULL, x, o);
		if (lane >= o) x += t;
	}
	return x;
}

// One clipping pass with the reference's own arithmetic -- the rare pass the screening below cannot decide (a comparison
// inside the margin, a vanishing sigma).  The survivors are compacted in order, ONE lane runs gsl_fit_linear's literal
// recurrences and the residual sum in the reference's order (ssc_column.cuh:fit_line_literal: 3 n strictly sequential
// rounded updates), then every lane flags its own samples with line_clipping (stacking.c:1170-1183) at their ranks.
// Out of line and rolled, reading the samples from the warp's shared-memory copy: the screening is straight-line code
// on registers and must not notice this path.  y: this lane's samples (y[r * 32]), alive: which of them survive,
// base: rank of the first of them among the survivors.
static __device__ __noinline__ void literal_pass(const double *y, unsigned int alive, int base, int n_cur, int lane, int rows,
		double slow, double shigh, FitTables fit, WORD *comp, unsigned int *masks) {
	int k = base;
	for (int r = 0; r < rows; r++)
		if (alive & (1u << r)) comp[k++] = (WORD) __double2int_rn(y[r * 32]);
	__syncwarp();
	double a = 0.0, b = 0.0, sigma = 0.0;
	if (lane == 0) fit_line_literal(Column{comp, 1}, n_cur, &fit, a, b, sigma);
	a = __shfl_sync(FULL, a, 0); b = __shfl_sync(FULL, b, 0); sigma = __shfl_sync(FULL, sigma, 0);
	const Reciprocal inv_sigma(sigma);	// line_clipping divides every residual by sigma (:1175,1179)
	unsigned int fm = 0, lom = 0;
	k = base;
	for (int r = 0; r < rows; r++) {
		if (alive & (1u << r)) {
			const int fl = clip_line(inv_sigma, (unsigned int) __double2int_rn(y[r * 32]), slow, shigh, k, a, b);
			if (fl) fm |= 1u << r;
			if (fl < 0) lom |= 1u << r;
			k++;
		}
	}
	__syncwarp();
	masks[0] = fm; masks[1] = lom;
}

// ---------------------------------------------------------------------------------------------
// linear-fit clipping of one pixel (half `sel` of the packed column).  Always finishes the pixel (true); the
// hand-back list of the earlier form stays for the operator layer's protocol but is no longer fed from here.
// ---------------------------------------------------------------------------------------------
template <int R>
__device__ __forceinline__ bool fit_pixel(const unsigned int (&v)[R], const int sel, const int N, const int lane,
		const double slow, const double shigh, const FitTables &fit, double *__restrict__ y, WORD *__restrict__ comp,
		unsigned int &result, unsigned int &rej_lo, unsigned int &rej_hi, int &guard, int &cap) {
	// y: this lane's R samples as doubles, in the warp's own piece of shared memory (y[r * 32]: conflict-free;
	// in registers they would cost 2 R of them next to the packed column)
	unsigned int alive = 0, Sy = 0, Wl = 0;	// Sy = sum of my live samples, Wl = sum of (rank inside the lane) * sample
	int cnt = 0;
#pragma unroll
	for (int r = 0; r < R; r++) {
		const unsigned int yi = sel ? v[r] >> 16 : v[r] & 0xFFFFu;
		y[r * 32] = word_to_double(yi);
		if (lane * R + r < N) {
			alive |= 1u << r;
			Wl += (unsigned int) cnt * yi;
			Sy += yi;
			cnt++;
		}
	}
	int n_cur = N, r_cum = 0, iter = 0;
	rej_lo = rej_hi = 0;
	for (;;) {
		const int base = warp_scan(cnt, lane) - cnt;	// index of my first live sample in the sorted survivors
		// sum y and sum i y of the survivors: exact integers below 2^53, in doubles
		const double sy = warp_sum((double) Sy);
		const double sxy = warp_sum(__fma_rn((double) base, (double) Sy, (double) Wl));
		const double n = (double) n_cur;
		// slope = (n Sxy - Sx Sy) / (n Sxx - Sx^2), Sx = n (n-1) / 2; numerator exact, one rounding for the quotient
		const double num = dsub(dmul(n, sxy), dmul(dmul(dmul(n, dsub(n, 1.0)), 0.5), sy));
		const double a = dmul(num, fit.inv_den[n_cur]);
		const double b = dsub(dmul(sy, fit.inv[n_cur]), dmul(dmul(dsub(n, 1.0), 0.5), a));
		const double pred0 = __fma_rn(a, (double) base, b);
		double pred = pred0, t = 0.0;
#pragma unroll 8
		for (int r = 0; r < R; r++) {
			const unsigned int on = alive & (1u << r);
			add_if(t, fabs(dsub(y[r * 32], pred)), on);
			add_if(pred, a, on);
		}
		const double sigma = dmul(warp_sum(t), fit.inv[n_cur]);
		bool literal = !(sigma > 0.05);	// (nearly) perfect line: the reference divides by this sigma
		const double tl = dmul(slow, sigma), th = dmul(shigh, sigma);
		const double ml = dadd(dmul(fabs(tl), 1e-7), 1e-6), mh = dadd(dmul(fabs(th), 1e-7), 1e-6);
		const double tl_s = dadd(tl, ml), tl_m = dsub(tl, ml), th_s = dadd(th, mh), th_m = dsub(th, mh);
		unsigned int fm = 0, lom = 0, near = 0;
		pred = pred0;
#pragma unroll 8
		for (int r = 0; r < R; r++) {
			const unsigned int on = alive & (1u << r);
			const double res = dsub(y[r * 32], pred);
			add_if(pred, a, on);
			const bool lo_s = -res > tl_s, lo_m = -res > tl_m, hi_s = res > th_s, hi_m = res > th_m;
			near |= ((lo_s != lo_m) | (!lo_s & (hi_s != hi_m))) ? on : 0u;
			lom |= lo_s ? on : 0u;				// line_clipping tests the low side first (:1175)
			fm |= (lo_s | hi_s) ? on : 0u;
		}
		literal = literal || __any_sync(FULL, near != 0);
		if (literal) {
			// ITS flags come from the reference's own arithmetic; the passes before and after keep the screening form, and
			// every flag is the reference's either way.  (The first form handed the whole pixel to a second kernel that
			// redid all its passes literally: a 5 ms chain per launch for a few hundred pixels.)
			unsigned int masks[2];
			literal_pass(y, alive, base, n_cur, lane, R, slow, shigh, fit, comp, masks);
			fm = masks[0]; lom = masks[1];
		}
		// the reference's scan: r++ per flagged sample, `if (N - r <= 4) break;` after every sample
		const int c = __popc(fm);
		const int e = warp_scan(c, lane) - c;
		const int flagged = __shfl_sync(FULL, e + c, 31);
		const int need = n_cur - 4 - r_cum;	// the break follows the need-th flagged sample of this pass
		if (need <= 0) {			// already true: only the very first sample is looked at
			fm = (base == 0 && cnt > 0) ? fm & (alive & (0u - alive)) : 0u;
			if (n_cur > 1) guard = 1;
		} else if (flagged >= need) {
			int hit = 0;
			if (e < need && need <= e + c) {	// the need-th flagged sample is mine: where does it sit?
				unsigned int tt = fm;
				for (int k = 1; k < need - e; k++) tt &= tt - 1;
				const unsigned int bit = tt & (0u - tt);
				hit = base + __popc(alive & (bit - 1)) < n_cur - 1;
			}
			if (__any_sync(FULL, hit)) guard = 1;
			int keep = need - e;
			keep = keep < 0 ? 0 : (keep > c ? c : keep);
			unsigned int tt = fm, kept = 0;
			for (int k = 0; k < keep; k++) { kept |= tt & (0u - tt); tt &= tt - 1; }
			fm = kept;
		}
		lom &= fm;
		const int removed = warp_sum(__popc(fm));
		rej_lo += __popc(lom);
		rej_hi += __popc(fm & ~lom);
		r_cum += removed;
		if (fm) {	// my survivors changed: their sums again
			alive &= ~fm;
			Sy = 0; Wl = 0; cnt = 0;
#pragma unroll
			for (int r = 0; r < R; r++) {
				const unsigned int on = (alive >> r) & 1u;
				const unsigned int yi = (sel ? v[r] >> 16 : v[r] & 0xFFFFu) * on;
				Wl += (unsigned int) cnt * yi;
				Sy += yi;
				cnt += (int) on;
			}
		}
		n_cur -= removed;
		if (++iter >= SSC_ITER_CAP) { cap = 1; break; }
		if (!(removed > 0 && n_cur > 3)) break;
	}
	rej_lo = (unsigned int) warp_sum((int) rej_lo);
	rej_hi = (unsigned int) warp_sum((int) rej_hi);
	result = mean_to_word<false>((unsigned long long) (unsigned int) warp_sum((int) Sy), n_cur);	// stacking.c:1790-1794
	return true;
}

// ---------------------------------------------------------------------------------------------
// the kernel: every warp on its own -- no block-wide step anywhere, because the number of clipping
// iterations differs from pixel to pixel by a factor of four and a tile barrier made every warp wait for
// the slowest of its block (30 % of the stall samples in the first form, profiles/r02_ncu_cfg5_warp_v2.txt).
// Warp g takes pixel pairs g, g + (warps of the grid), ...: at any time the warps of the whole GPU work on
// a run of consecutive pairs, so the 32-byte sectors a warp touches for its 4 bytes per frame are shared
// with seven other warps through the L2 (DRAM bytes stay the algorithmic ones).
// ---------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(32 * WARP_PAIRS, 2)
k_stack_warp(StackParams p, unsigned int n_pairs, int is_median, LeftoverList list) {
	extern __shared__ __align__(16) double s_fit[];		// [warps][R][32] doubles: the fit's copy of a pixel's samples
	const ssc_band &b = p.band;
	const int W = b.width, W2 = W >> 1, N = b.n_frames, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const bool norm = p.coeff != nullptr && p.normalize != NO_NORM;
	double *s_y = s_fit + (size_t) warp * 32 * R + lane;
	WORD *s_comp = reinterpret_cast<WORD *>(s_fit + (size_t) WARP_PAIRS * 32 * R) + (size_t) warp * 32 * R;	// survivors in order (literal passes)
	const unsigned long long fstride = (unsigned long long) b.frame_stride * sizeof(WORD);
	const unsigned int warps_total = gridDim.x * WARP_PAIRS;
	unsigned long long cnt[SSC_COUNTERS] = {0, 0, 0, 0, 0, 0, 0, 0};	// lane 0: this warp's share of the counters

	for (unsigned int pair = blockIdx.x * WARP_PAIRS + warp; pair < n_pairs; pair += warps_total) {
		const unsigned int row = pair / (unsigned int) W2;
		const int x2 = (int) (pair - row * (unsigned int) W2);
		const int layer = (int) (row / (unsigned int) b.out_rows);
		const int rr = (int) (row - (unsigned int) layer * (unsigned int) b.out_rows);
		const int Rr = b.out_row0 + rr;				// result storage row (stacking.c:731,1600)
		const int srow = b.row_order == SSC_ROWS_SER ? b.height - 1 - Rr : Rr;
		const char *src = reinterpret_cast<const char *>(reinterpret_cast<const unsigned int *>(b.frames)
				+ (((size_t) layer * b.layer_stride + (size_t) (srow - b.buf_row0) * W) >> 1) + x2);

		// ---- 1. my pair's column: lane l takes frames l, l + 32, ... (sentinels sort behind the samples) ----
		unsigned int v[R];
#pragma unroll
		for (int r = 0; r < R; r++) {
			const int f = r * 32 + lane;
			v[r] = f < N ? __ldg(reinterpret_cast<const unsigned int *>(src + (unsigned long long) f * fstride)) : 0xFFFFFFFFu;
		}
		if (norm) {	// normalization (stacking.c:1635-1652, :750-764), the conversion-based exact form: a ROLLED loop over
			// a shared-memory copy of the column (unrolled per register it was 7 k instructions of a 13 k kernel,
			// and instruction fetch stalled every launch whether it normalized or not)
			unsigned int *s_w = reinterpret_cast<unsigned int *>(s_fit) + (size_t) warp * 32 * R * 2;	// the warp's fit area
#pragma unroll
			for (int r = 0; r < R; r++) s_w[r * 32 + lane] = v[r];
			__syncwarp();
#pragma unroll 1
			for (int f = lane; f < N; f += 32) {
				const unsigned int w = s_w[f];
				const double off = p.coeff[f], mul = p.coeff[N + f], sc = p.coeff[2 * N + f];
				s_w[f] = normalize_sample(w & 0xFFFFu, p.normalize, off, mul, sc) | (normalize_sample(w >> 16, p.normalize, off, mul, sc) << 16);
			}
			__syncwarp();
#pragma unroll
			for (int r = 0; r < R; r++) v[r] = s_w[r * 32 + lane];
			__syncwarp();
		}

		// ---- 2. sort: lane l ends up with sorted samples [l R, (l + 1) R) of both pixels ----
		warp_sort<R>(v, lane);

		const size_t out_idx = (size_t) layer * p.out_layer_stride + (size_t) (Rr - p.out_buf_row0) * W + 2 * (size_t) x2;
		if (is_median) {					// stacking.c:765-767, double -> WORD truncation
			const int m1 = (N - 1) >> 1, m2 = N >> 1;
			unsigned int a = 0, c = 0;
#pragma unroll
			for (int r = 0; r < R; r++) {
				if (r == (m1 & (R - 1))) a = v[r];
				if (r == (m2 & (R - 1))) c = v[r];
			}
			a = __shfl_sync(FULL, a, m1 / R);
			c = __shfl_sync(FULL, c, m2 / R);
			if (lane == 0) {
				const unsigned int lo = ((a & 0xFFFFu) + (c & 0xFFFFu)) >> 1, hi = ((a >> 16) + (c >> 16)) >> 1;
				*reinterpret_cast<unsigned int *>(p.out + out_idx) = lo | (hi << 16);
			}
			continue;
		}
		// ---- 3. linear-fit clipping, one pixel of the pair after the other ----
#pragma unroll 1
		for (int sel = 0; sel < 2; sel++) {
			unsigned int res = 0, rej_lo = 0, rej_hi = 0;
			int guard = 0, cap = 0;
			const bool done = fit_pixel<R>(v, sel, N, lane, p.sig_low, p.sig_high, p.fit, s_y, s_comp, res, rej_lo, rej_hi, guard, cap);
			__syncwarp();
			if (lane == 0) {
				if (done) {
					p.out[out_idx + sel] = (WORD) res;
					cnt[layer * 2 + 0] += rej_lo;
					cnt[layer * 2 + 1] += rej_hi;
					cnt[6] += guard;
					cnt[7] += cap;
				} else {
					const unsigned int slot = atomicAdd(list.count, 1u);
					if (slot < list.capacity)
						list.ids[slot] = (unsigned int) (((size_t) (layer + p.hb_layer0) * (p.hb_defer ? p.hb_rows : b.out_rows)
								+ rr + p.hb_row0) * W + 2 * x2 + sel);
				}
			}
		}
	}
	if (p.counters && lane == 0)
#pragma unroll
		for (int k = 0; k < SSC_COUNTERS; k++)
			if (cnt[k]) atomicAdd(&p.counters[k], cnt[k]);
}

// ---------------------------------------------------------------------------------------------
// The pixels the screening could not decide (an exact tie of a residual with a threshold, a vanishing sigma):
// one warp per listed pixel reloads and sorts the column the same way, then lane 0 runs the complete
// LINEARFIT of ssc_column.cuh -- gsl_fit_linear's literal running means -- on a shared-memory copy.  (The general
// kernel's list mode did this with a serial shared-memory bitonic sort per thread: ~8 ms per launch for 1000
// frames however few pixels were listed.)
// ---------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(32 * WARP_PAIRS)
k_fit_handback(StackParams p, LeftoverList list) {
	extern __shared__ __align__(16) WORD s_cols[];		// [warps][32 R] the pixel's sorted column
	const ssc_band &b = p.band;
	const int W = b.width, N = b.n_frames, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const bool norm = p.coeff != nullptr && p.normalize != NO_NORM;
	WORD *col = s_cols + (size_t) warp * 32 * R;
	unsigned int n_items = *list.count;
	if (n_items > list.capacity) n_items = list.capacity;
	for (unsigned int item = blockIdx.x * WARP_PAIRS + warp; item < n_items; item += gridDim.x * WARP_PAIRS) {
		const unsigned int pix = list.ids[item];
		const unsigned int row = pix / (unsigned int) W;
		const int x = (int) (pix - row * (unsigned int) W);
		const int layer = (int) (row / (unsigned int) b.out_rows);
		const int rr = (int) (row - (unsigned int) layer * (unsigned int) b.out_rows);
		const int Rr = b.out_row0 + rr;
		const int srow = b.row_order == SSC_ROWS_SER ? b.height - 1 - Rr : Rr;
		const WORD *src = b.frames + (size_t) layer * b.layer_stride + (size_t) (srow - b.buf_row0) * W + x;
		unsigned int v[R], sum = 0;
#pragma unroll
		for (int r = 0; r < R; r++) {
			const int f = r * 32 + lane;
			unsigned int s = 0xFFFFu;				// sentinels sort behind the samples
			if (f < N) {
				s = __ldg(src + (size_t) f * b.frame_stride);
				if (norm) s = normalize_sample(s, p.normalize, p.coeff[f], p.coeff[N + f], p.coeff[2 * N + f]);
				sum += s;
			}
			v[r] = s;						// (the upper halves stay 0 through the packed min / max)
		}
		warp_sort<R>(v, lane);
#pragma unroll
		for (int r = 0; r < R; r++) col[lane * R + r] = (WORD) v[r];
		const unsigned long long S = (unsigned long long) (unsigned int) warp_sum((int) sum);
		__syncwarp();
		if (lane == 0) {
			RejectState st{0, 0, 0, 0};
			const unsigned int res = reject_sorted_and_mean<false>(Column{col, 1}, N, LINEARFIT, p.sig_low, p.sig_high, S, 0ull, st,
					nullptr, &p.fit);
			p.out[(size_t) layer * p.out_layer_stride + (size_t) (Rr - p.out_buf_row0) * W + x] = (WORD) res;
			if (p.counters) {
				if (st.rej_lo) atomicAdd(&p.counters[layer * 2 + 0], (unsigned long long) st.rej_lo);
				if (st.rej_hi) atomicAdd(&p.counters[layer * 2 + 1], (unsigned long long) st.rej_hi);
				if (st.guard) atomicAdd(&p.counters[6], 1ull);
				if (st.cap) atomicAdd(&p.counters[7], 1ull);
			}
		}
		__syncwarp();
	}
}

template <int R>
static int launch_fit_handback_r(const StackParams &p, const LeftoverList &list, cudaStream_t stream) {
	const size_t smem = (size_t) WARP_PAIRS * 32 * R * sizeof(WORD);
	int dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	k_fit_handback<R><<<sms * 2, 32 * WARP_PAIRS, smem, stream>>>(p, list);
	return cudaGetLastError() == cudaSuccess ? SSC_OK : SSC_ERR_CUDA;
}

// p.fit must hold the tables (launch_stack / launch_deferred_handback fill them)
int launch_fit_handback(const StackParams &p, const LeftoverList &list, cudaStream_t stream) {
	const int n = p.band.n_frames;
	if (n <= 32) return launch_fit_handback_r<1>(p, list, stream);
	if (n <= 64) return launch_fit_handback_r<2>(p, list, stream);
	if (n <= 128) return launch_fit_handback_r<4>(p, list, stream);
	if (n <= 256) return launch_fit_handback_r<8>(p, list, stream);
	if (n <= 512) return launch_fit_handback_r<16>(p, list, stream);
	return launch_fit_handback_r<32>(p, list, stream);
}

// Same shape conditions as the other packed-pair kernels; linear fit from 8 frames on (fewer: the
// general kernel), the median above 512 frames (below, ssc_stack_fast.cu / ssc_stack_mid.cu are the
// better sorters).
bool warp_path_eligible(const StackParams &p, int is_median) {
	const ssc_band &b = p.band;
	if (b.n_frames > 1024) return false;
	if (is_median ? b.n_frames <= 512 : (p.rejection != LINEARFIT || b.n_frames < 8)) return false;
	if (!is_median && (b.shiftx || b.shifty)) return false;
	if ((b.width & 1) || (b.frame_stride & 1) || (b.layer_stride & 1)) return false;
	if (((uintptr_t) b.frames & 3) || ((uintptr_t) p.out & 3) || (p.out_layer_stride & 1)) return false;
	if ((unsigned long long) b.layers * b.out_rows * b.width >= 0xFFFFFFF0ull) return false;	// 32-bit pixel ids
	return true;
}

template <int R>
static int launch_warp(const StackParams &p, int is_median, cudaStream_t stream) {
	const ssc_band &b = p.band;
	const unsigned long long pairs = (unsigned long long) (b.width >> 1) * b.layers * b.out_rows;
	if (pairs == 0) return SSC_OK;
	if (pairs >= 0xFFFFFFF0ull) return SSC_ERR_ARG;
	const bool norm = p.coeff != nullptr && p.normalize != NO_NORM;
	const size_t smem = (is_median && !norm) ? 0 : (size_t) WARP_PAIRS * 32 * R * (sizeof(double) + (is_median ? 0 : sizeof(WORD)));	// the fit's doubles (+ the compacted survivors of a literal pass) / the normalization's copy
	if (cudaFuncSetAttribute(k_stack_warp<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem) != cudaSuccess)
		return SSC_ERR_CUDA;
	LeftoverList list{nullptr, 0, nullptr, nullptr};
	const unsigned long long pixels = p.hb_defer ? p.hb_pixels : (unsigned long long) b.layers * b.out_rows * b.width;
	if (pixels >= 0xFFFFFFF0ull) return SSC_ERR_ARG;
	if (!is_median) {
		int rc = leftover_ids(pixels, list, stream, p.hb_defer == 1);
		if (rc) return rc;
	}
	int dev = 0, sms = 148, per_sm = 1;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_stack_warp<R>, 32 * WARP_PAIRS, smem) != cudaSuccess || per_sm < 1)
		per_sm = 1;
	unsigned long long blocks = (unsigned long long) sms * per_sm;	// resident warps only: every warp strides over the pairs
	const unsigned long long need = (pairs + WARP_PAIRS - 1) / WARP_PAIRS;
	if (blocks > need) blocks = need;
	k_stack_warp<R><<<(unsigned int) blocks, 32 * WARP_PAIRS, smem, stream>>>(p, (unsigned int) pairs, is_median, list);
	if (cudaGetLastError() != cudaSuccess) return SSC_ERR_CUDA;
	if (!is_median && !p.hb_defer)	// pixels the screening could not decide: gsl_fit_linear's literal recurrences (ssc_column.cuh)
		return launch_fit_handback(p, list, stream);
	return SSC_OK;
}

int launch_stack_warp(const StackParams &p, int is_median, cudaStream_t stream) {
	const int n = p.band.n_frames;
	if (n <= 32) return launch_warp<1>(p, is_median, stream);
	if (n <= 64) return launch_warp<2>(p, is_median, stream);
	if (n <= 128) return launch_warp<4>(p, is_median, stream);
	if (n <= 256) return launch_warp<8>(p, is_median, stream);
	if (n <= 512) return launch_warp<16>(p, is_median, stream);
	return launch_warp<32>(p, is_median, stream);
}

}  // namespace ssc
