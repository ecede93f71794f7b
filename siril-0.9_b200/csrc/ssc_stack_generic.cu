// ssc_stack_generic.cu -- K2/K3, general form: any frame count up to
// SSC_GENERIC_MAX_FRAMES, every rejection mode, registration shifts.
//
// One thread owns one pixel column.  The column's N samples live in shared
// memory as col[i * blockDim.x + threadIdx.x]: for a fixed i a warp touches 32
// consecutive WORDs, and because the stride between a thread's own elements is
// a multiple of 128 bytes any data-dependent index (sorting, compaction) keeps
// every lane in its own bank -- no conflicts whatever the access pattern.
//
// This is the reference-complete path; the register-resident sorting-network
// kernels in ssc_stack_fast.cu take over for the shapes they cover and hand
// the rare history-dependent pixels back to this one.
#include "ssc_device.cuh"
#include "ssc_kernels.h"

namespace ssc {

struct Column {
	WORD *base;	// &col[threadIdx.x]
	int stride;	// blockDim.x
	__device__ __forceinline__ unsigned int get(int i) const { return base[i * stride]; }
	__device__ __forceinline__ void set(int i, unsigned int v) { base[i * stride] = (WORD) v; }
};

// In-place ascending bitonic sort of n_pad (power of two) elements.
__device__ void column_sort(Column c, int n_pad) {
	for (int k = 2; k <= n_pad; k <<= 1) {
		for (int j = k >> 1; j > 0; j >>= 1) {
			for (int i = 0; i < n_pad; i++) {
				int l = i ^ j;
				if (l > i) {
					unsigned int a = c.get(i), b = c.get(l);
					bool up = (i & k) == 0;
					if ((a > b) == up) {
						c.set(i, b);
						c.set(l, a);
					}
				}
			}
		}
	}
}

__device__ __forceinline__ void column_pad(Column c, int n, int n_pad) {
	for (int i = n; i < n_pad; i++) c.set(i, 0xFFFFu);
}

// gsl_stats_ushort_median_from_sorted_data (stacking.c:767,1662,1678,...)
__device__ __forceinline__ double column_median(Column c, int n) {
	int lhs = (n - 1) >> 1, rhs = n >> 1;
	unsigned int a = c.get(lhs);
	if (lhs == rhs) return (double) (int) a;
	return ddiv((double) (int) (a + c.get(rhs)), 2.0);
}

__device__ __forceinline__ void column_sums(Column c, int n, unsigned long long &S, unsigned long long &Q) {
	unsigned long long s = 0, q = 0;
	for (int i = 0; i < n; i++) {
		unsigned int v = c.get(i);
		s += v;
		q += (unsigned long long) (v * v);	// v*v < 2^32
	}
	S = s;
	Q = q;
}

struct RejectState {
	unsigned int rej_lo, rej_hi;
	int guard, cap;
};

// One flag-and-remove pass of SIGMA / WINSORIZED (stacking.c:1680-1693, 1731-1747):
// flags are consumed on the fly, so the reference's rejected[] array is not
// needed; samples the reference's loop does not reach after its
// "if (N - r <= 4) break;" keep flag 0 (see siril_oracle.c for why).
__device__ __forceinline__ int sigma_pass(Column c, int N, double slow, double shigh, double sigma,
		double median, int &r, RejectState &st) {
	int out = 0, f = 0;
	for (; f < N; f++) {
		unsigned int v = c.get(f);
		int fl = clip_sigma(v, slow, shigh, sigma, median);
		if (fl) {
			r++;
			if (fl < 0) st.rej_lo++; else st.rej_hi++;
		} else {
			c.set(out++, v);
		}
		if (N - r <= 4) {
			if (f < N - 1) st.guard = 1;
			f++;
			break;
		}
	}
	for (; f < N; f++) c.set(out++, c.get(f));
	return N - out;
}

// stacking.c:1656-1794 for one column; returns the output sample.
__device__ unsigned int reject_and_mean(Column c, int N, int n_pad, int rejection, double slow,
		double shigh, RejectState &st) {
	double sigma, median;
	int n, r = 0, iter = 0;
	unsigned long long S, Q;

	switch (rejection) {
	case PERCENTILE: {	// stacking.c:1660-1673
		column_pad(c, N, n_pad);
		column_sort(c, n_pad);
		median = column_median(c, N);
		int cur = N, out = 0;
		for (int f = 0; f < N; f++) {
			unsigned int v = c.get(f);
			int fl = clip_percentile(v, slow, shigh, median);
			if (fl < 0) st.rej_lo++; else if (fl > 0) st.rej_hi++;
			if (fl != 0 && cur > 1) cur--;
			else c.set(out++, v);
		}
		N = cur;
		break;
	}
	case SIGMA:		// stacking.c:1674-1695
		column_pad(c, N, n_pad);
		column_sort(c, n_pad);	// survivors of a sorted stack stay sorted: one sort serves all iterations
		do {
			column_sums(c, N, S, Q);
			sigma = sd_from_sums(S, Q, N);
			median = column_median(c, N);
			n = sigma_pass(c, N, slow, shigh, sigma, median, r, st);
			N -= n;
			if (++iter >= SSC_ITER_CAP) { st.cap = 1; break; }
		} while (n > 0 && N > 3);
		break;
	case SIGMEDIAN:		// stacking.c:1696-1709
		do {
			column_sums(c, N, S, Q);
			sigma = sd_from_sums(S, Q, N);
			column_pad(c, N, n_pad);
			column_sort(c, n_pad);
			median = column_median(c, N);
			unsigned int med_w = round_to_word(median);
			n = 0;
			for (int f = 0; f < N; f++) {
				int fl = clip_sigma(c.get(f), slow, shigh, sigma, median);
				if (fl) {
					if (fl < 0) st.rej_lo++; else st.rej_hi++;
					c.set(f, med_w);
					n++;
				}
			}
			if (++iter >= SSC_ITER_CAP) { st.cap = 1; break; }
		} while (n > 0 && N > 3);
		break;
	case WINSORIZED:	// stacking.c:1710-1749
		column_pad(c, N, n_pad);
		column_sort(c, n_pad);
		do {
			column_sums(c, N, S, Q);
			sigma = sd_from_sums(S, Q, N);
			median = column_median(c, N);
			// w_stack is never materialised: clamping a sorted stack keeps it sorted and
			// only ever rewrites a prefix (to one value) and a suffix (to one value), so
			// w = [vlo x clo] ++ s[clo .. N-chi) ++ [vhi x chi] with exact running sums.
			int clo = 0, chi = 0, inner = 0;
			unsigned int vlo = 0, vhi = 0;
			unsigned long long Sm = S, Qm = Q;
			double sigma0;
			const int lhs = (N - 1) >> 1, rhs = N >> 1;
			do {
				double m0 = dsub(median, dmul(1.5, sigma));
				double m1 = dadd(median, dmul(1.5, sigma));
				unsigned int r0 = round_to_word(m0), r1 = round_to_word(m1);
				if (clo > 0) {		// Winsorized(), stacking.c:1163-1168, on the prefix value
					if ((double) (int) vlo < m0) vlo = r0;
					else if ((double) (int) vlo > m1) vlo = r1;
				}
				if (chi > 0) {
					if ((double) (int) vhi < m0) vhi = r0;
					else if ((double) (int) vhi > m1) vhi = r1;
				}
				while (clo + chi < N) {
					unsigned int v = c.get(clo);
					if (!((double) (int) v < m0)) break;
					Sm -= v; Qm -= (unsigned long long) (v * v);
					clo++; vlo = r0;
				}
				while (clo + chi < N) {
					unsigned int v = c.get(N - 1 - chi);
					if (!((double) (int) v > m1)) break;
					Sm -= v; Qm -= (unsigned long long) (v * v);
					chi++; vhi = r1;
				}
				unsigned int wl = lhs < clo ? vlo : (lhs >= N - chi ? vhi : c.get(lhs));
				unsigned int wr = rhs < clo ? vlo : (rhs >= N - chi ? vhi : c.get(rhs));
				median = lhs == rhs ? (double) (int) wl : ddiv((double) (int) (wl + wr), 2.0);
				unsigned long long Sw = Sm + (unsigned long long) clo * vlo + (unsigned long long) chi * vhi;
				unsigned long long Qw = Qm + (unsigned long long) clo * (vlo * vlo)
						+ (unsigned long long) chi * (vhi * vhi);
				sigma0 = sigma;
				sigma = dmul(1.134, sd_from_sums(Sw, Qw, N));
				if (++inner >= SSC_ITER_CAP) { st.cap = 1; break; }
			} while (ddiv(fabs(dsub(sigma, sigma0)), sigma0) > 0.0005);
			n = sigma_pass(c, N, slow, shigh, sigma, median, r, st);
			N -= n;
			if (++iter >= SSC_ITER_CAP) { st.cap = 1; break; }
		} while (n > 0 && N > 3);
		break;
	case LINEARFIT:		// stacking.c:1750-1784
		column_pad(c, N, n_pad);
		column_sort(c, n_pad);
		do {
			// gsl_fit_linear with x_i = i (GSL fit/linear.c), literal running means
			double m_x = 0.0, m_y = 0.0, m_dx2 = 0.0, m_dxdy = 0.0;
			for (int i = 0; i < N; i++) {
				double ip1 = dadd((double) i, 1.0);
				m_x = dadd(m_x, ddiv(dsub((double) i, m_x), ip1));
				m_y = dadd(m_y, ddiv(dsub((double) (int) c.get(i), m_y), ip1));
			}
			for (int i = 0; i < N; i++) {
				double ip1 = dadd((double) i, 1.0);
				double dx = dsub((double) i, m_x);
				double dy = dsub((double) (int) c.get(i), m_y);
				m_dx2 = dadd(m_dx2, ddiv(dsub(dmul(dx, dx), m_dx2), ip1));
				m_dxdy = dadd(m_dxdy, ddiv(dsub(dmul(dx, dy), m_dxdy), ip1));
			}
			double a = ddiv(m_dxdy, m_dx2);		// slope
			double b = dsub(m_y, dmul(m_x, a));	// intercept
			sigma = 0.0;
			for (int f = 0; f < N; f++)
				sigma = dadd(sigma, fabs(dsub((double) (int) c.get(f), dadd(dmul(a, (double) f), b))));
			sigma = ddiv(sigma, (double) N);
			int out = 0, f = 0;
			for (; f < N; f++) {
				unsigned int v = c.get(f);
				int fl = clip_line(v, slow, shigh, sigma, f, a, b);
				if (fl) {
					r++;
					if (fl < 0) st.rej_lo++; else st.rej_hi++;
				} else {
					c.set(out++, v);
				}
				if (N - r <= 4) {
					if (f < N - 1) st.guard = 1;
					f++;
					break;
				}
			}
			for (; f < N; f++) c.set(out++, c.get(f));
			n = N - out;
			N -= n;
			if (++iter >= SSC_ITER_CAP) { st.cap = 1; break; }
		} while (n > 0 && N > 3);
		break;
	default:
		break;
	}

	unsigned long long sum = 0;	// stacking.c:1790-1794 (the double sum of WORDs is exact)
	for (int f = 0; f < N; f++) sum += c.get(f);
	return round_to_word(ddiv((double) (long long) sum, (double) N));
}

// Which pixels a launch covers: either the dense range of the band's output
// rows, or an explicit list (the fast kernels' hand-back of guard pixels).
struct PixelSource {
	const unsigned int *list;	// packed pixel ids or nullptr
	const unsigned int *list_count;	// device count for list mode
	unsigned long long total;	// dense mode: layers*out_rows*width
};

__global__ void k_stack_generic(StackParams p, PixelSource src, int n_pad, int is_median) {
	extern __shared__ WORD smem_col[];
	__shared__ unsigned long long blk_counters[SSC_COUNTERS];
	if (threadIdx.x < SSC_COUNTERS) blk_counters[threadIdx.x] = 0;
	__syncthreads();

	const ssc_band &b = p.band;
	const int W = b.width, N = b.n_frames;
	unsigned long long total = src.list ? (unsigned long long) *src.list_count : src.total;
	unsigned long long g = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x;

	if (g < total) {
		unsigned long long pix = src.list ? src.list[g] : g;
		int x = (int) (pix % W);
		int rr = (int) ((pix / W) % b.out_rows);
		int layer = (int) (pix / ((unsigned long long) W * b.out_rows));
		int R = b.out_row0 + rr;		// result storage row
		int t = b.height - 1 - R;		// the reference's top-down block row (stacking.c:731,1600)
		Column c{smem_col + threadIdx.x, (int) blockDim.x};
		const bool have_reg = !is_median && (b.shiftx || b.shifty);	// stack_median ignores shifts
		const WORD *layer_base = b.frames + (size_t) layer * b.layer_stride;

		for (int f = 0; f < N; f++) {
			int sx = (have_reg && b.shiftx) ? b.shiftx[f] : 0;
			int sy = (have_reg && b.shifty) ? b.shifty[f] : 0;
			unsigned int v;
			if (sx && (x - sx >= W || x - sx < 0)) {
				v = 0;			// stacking.c:1628-1632: literal 0, not normalized
			} else {
				int ty = t + sy;	// stacking.c:1550-1569
				unsigned int raw = 0;	// rows outside the image are zero-filled (:1572-1576)
				if (ty >= 0 && ty < b.height) {
					int srow = b.row_order == SSC_ROWS_SER ? ty : b.height - 1 - ty;
					raw = layer_base[(size_t) f * b.frame_stride + (size_t) (srow - b.buf_row0) * W + (x - sx)];
				}
				if (p.coeff)
					v = normalize_sample(raw, p.normalize, p.coeff[f], p.coeff[N + f], p.coeff[2 * N + f]);
				else
					v = raw;
			}
			c.set(f, v);
		}

		unsigned int result;
		RejectState st{0, 0, 0, 0};
		if (is_median) {			// stacking.c:765-767, double -> WORD truncation
			column_pad(c, N, n_pad);
			column_sort(c, n_pad);
			int lhs = (N - 1) >> 1, rhs = N >> 1;
			result = (c.get(lhs) + c.get(rhs)) >> 1;
		} else {
			result = reject_and_mean(c, N, n_pad, p.rejection, p.sig_low, p.sig_high, st);
		}
		p.out[(size_t) layer * p.out_layer_stride + (size_t) (R - p.out_buf_row0) * W + x] = (WORD) result;

		if (st.rej_lo) atomicAdd(&blk_counters[layer * 2 + 0], (unsigned long long) st.rej_lo);
		if (st.rej_hi) atomicAdd(&blk_counters[layer * 2 + 1], (unsigned long long) st.rej_hi);
		if (st.guard) atomicAdd(&blk_counters[6], 1ull);
		if (st.cap) atomicAdd(&blk_counters[7], 1ull);
	}
	__syncthreads();
	if (p.counters && threadIdx.x < SSC_COUNTERS && blk_counters[threadIdx.x])
		atomicAdd(&p.counters[threadIdx.x], blk_counters[threadIdx.x]);
}

int launch_stack_generic(const StackParams &p, const unsigned int *list, const unsigned int *list_count,
		unsigned int list_capacity, int is_median, cudaStream_t stream) {
	const ssc_band &b = p.band;
	if (b.n_frames < 2 || b.n_frames > SSC_GENERIC_MAX_FRAMES) return SSC_ERR_ARG;
	int n_pad = 2;
	while (n_pad < b.n_frames) n_pad <<= 1;
	// largest block whose columns fit: prefer two resident CTAs (<= 100 KB each)
	int block = 256;
	while (block > 32 && (size_t) n_pad * block * sizeof(WORD) > 100 * 1024) block >>= 1;
	size_t smem = (size_t) n_pad * block * sizeof(WORD);
	if (smem > 200 * 1024) return SSC_ERR_ARG;
	if (cudaFuncSetAttribute(k_stack_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
		return SSC_ERR_CUDA;
	PixelSource src;
	src.list = list;
	src.list_count = list_count;
	src.total = (unsigned long long) b.layers * b.out_rows * b.width;
	unsigned long long work = list ? list_capacity : src.total;
	if (work == 0) return SSC_OK;
	unsigned int grid = (unsigned int) ((work + block - 1) / block);
	k_stack_generic<<<grid, block, smem, stream>>>(p, src, n_pad, is_median);
	return cudaGetLastError() == cudaSuccess ? SSC_OK : SSC_ERR_CUDA;
}

}  // namespace ssc
