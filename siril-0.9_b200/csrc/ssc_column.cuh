// ssc_column.cuh -- the per-pixel part of stack_mean_with_rejection
// (src/stacking/stacking.c:1656-1794) on one SORTED pixel column held in shared
// memory, shared by the general kernel (ssc_stack_generic.cu, bitonic sort in
// shared memory) and the register-network kernel (ssc_stack_fast.cu).
//
// The reference re-sorts and physically removes samples in every iteration.
// Here the column is sorted once; because sigma / winsorized / percentile
// clipping are monotone in the sample value, what they reject from a sorted
// stack is a prefix and a suffix, so the survivors stay the contiguous range
// [o, o+N) and the exact integer sums S = sum x, Q = sum x^2 of the survivors
// are maintained by subtracting what leaves.  The reference's
// "if (N - r <= 4) break;" (stacking.c:1684,1737,1771) is honoured sample by
// sample in scan order; when it cuts a high suffix short the leftovers are
// moved down so the range stays contiguous.  Linear-fit clipping can reject
// interior samples and compacts like the reference does.
#pragma once
#include "ssc_device.cuh"
#include "ssc_kernels.h"

namespace ssc {

struct Column {
	WORD *base;	// first element of this thread's column
	int stride;	// distance between consecutive elements, in WORDs (a multiple of 64: no bank conflicts)
	__device__ __forceinline__ unsigned int get(int i) const { return base[i * stride]; }
	__device__ __forceinline__ void set(int i, unsigned int v) { base[i * stride] = (WORD) v; }
};

struct RejectState {
	unsigned int rej_lo, rej_hi;
	int guard, cap;
};

__device__ __forceinline__ unsigned long long sq(unsigned int v) { return (unsigned long long) (v * v); }	// v*v < 2^32

// gsl_stats_ushort_median_from_sorted_data on c[o .. o+n)
__device__ __forceinline__ double column_median(Column c, int o, int n) {
	int lhs = (n - 1) >> 1, rhs = n >> 1;
	unsigned int a = c.get(o + lhs);
	if (lhs == rhs) return (double) (int) a;
	return dmul((double) (int) (a + c.get(o + rhs)), 0.5);	// x / 2.0 == x * 0.5 exactly
}

// fabs(sigma - sigma0) / sigma0 > 0.0005 (stacking.c:1728) without the division wherever the answer is not within
// 1e-9 of the limit (the quotient's own rounding is 1e-16): there the reference's expression is evaluated
__device__ __forceinline__ bool relative_change_exceeds(double sigma, double sigma0) {
	const double diff = fabs(dsub(sigma, sigma0)), lim = dmul(0.0005, sigma0);
	if (diff > dmul(lim, 1.000000001)) return true;
	if (diff < dmul(lim, 0.999999999)) return false;
	return ddiv(diff, sigma0) > 0.0005;
}

// One flag-and-remove pass of SIGMA / WINSORIZED (stacking.c:1680-1693, 1731-1747)
// on the sorted survivors c[o .. o+N).  Returns the number of samples removed.
__device__ __forceinline__ int sigma_pass_sorted(Column c, int &o, int &N, double slow, double shigh,
		double sigma, double median, int &r, unsigned long long &S, unsigned long long &Q,
		RejectState &st) {
	const double tl = dmul(slow, sigma), th = dmul(shigh, sigma);
	int nl = 0, n = 0;
	bool stop = false;
	while (nl < N) {			// low rejections: a prefix
		unsigned int v = c.get(o + nl);
		if (!(dsub(median, (double) (int) v) > tl)) break;
		r++; st.rej_lo++; S -= v; Q -= sq(v); nl++;
		if (N - r <= 4) {		// the reference's break, taken right after this sample
			stop = true;
			if (nl - 1 < N - 1) st.guard = 1;
			break;
		}
	}
	n = nl;
	if (!stop && nl < N) {
		int b = N;			// high rejections: a suffix [b, N)
		unsigned long long Sh = 0, Qh = 0;
		while (b > nl) {
			unsigned int v = c.get(o + b - 1);
			if (!(dsub((double) (int) v, median) > th)) break;
			Sh += v; Qh += sq(v); b--;
		}
		int ch = N - b;
		if (b > nl && N - r <= 4) {	// break taken at the first kept sample: highs never reached
			stop = true;
			if (nl < N - 1) st.guard = 1;
		}
		if (!stop && ch > 0) {
			int tstar = N - 4 - r;	// the break is taken after this many high samples
			if (tstar < 1) tstar = 1;
			if (ch < tstar) {
				r += ch; st.rej_hi += ch; n += ch; S -= Sh; Q -= Qh;
			} else {
				r += tstar; st.rej_hi += tstar; n += tstar;
				if (ch == tstar) {
					S -= Sh; Q -= Qh;
				} else {	// samples above the break survive: close the gap
					st.guard = 1;
					for (int t = 0; t < tstar; t++) { unsigned int v = c.get(o + b + t); S -= v; Q -= sq(v); }
					for (int t = b + tstar; t < N; t++) c.set(o + t - tstar, c.get(o + t));
				}
			}
		}
	}
	o += nl;
	N -= n;
	return n;
}

// The straight line of LINEARFIT (stacking.c:1750-1768) through the N sorted samples c[0 .. N): gsl_fit_linear with
// x_i = i (GSL fit/linear.c) as its literal running means -- 4 N individually rounded updates -- then the mean absolute
// residual in the reference's summation order.  a = slope, b = intercept.  With the tables (FitTables) the
// data-independent recurrences (mean of x, mean of dx^2) are looked up and every division by i + 1 becomes
// div_exact() with RN(1 / (i + 1)): same bits, ~5x less work.
static __device__ __forceinline__ void fit_line_literal(Column c, int N, const FitTables *fit, double &a, double &b, double &sigma) {
	double m_x = 0.0, m_y = 0.0, m_dx2 = 0.0, m_dxdy = 0.0;
	const bool tab = fit && N <= fit->n_max;
	if (tab) {
		m_x = fit->mx[N];
		m_dx2 = fit->mdx2[N];
		for (int i = 0; i < N; i++)
			m_y = dadd(m_y, div_exact(dsub((double) (int) c.get(i), m_y), (double) (i + 1), fit->inv[i + 1]));
		for (int i = 0; i < N; i++) {
			double dx = dsub((double) i, m_x);
			double dy = dsub((double) (int) c.get(i), m_y);
			m_dxdy = dadd(m_dxdy, div_exact(dsub(dmul(dx, dy), m_dxdy), (double) (i + 1), fit->inv[i + 1]));
		}
	} else {
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
	}
	a = ddiv(m_dxdy, m_dx2);		// slope
	b = dsub(m_y, dmul(m_x, a));		// intercept
	sigma = 0.0;
	for (int f = 0; f < N; f++)
		sigma = dadd(sigma, fabs(dsub((double) (int) c.get(f), dadd(dmul(a, (double) f), b))));
	sigma = ddiv(sigma, (double) N);
}

// stacking.c:1656-1794 for one column whose N samples are sorted ascending in
// c[0..N) with S = sum, Q = sum of squares.  Returns the output sample.
// tables: optional small-n tables (ssc_device.cuh) that replace the divisions, bit for bit.
template <bool SMALLN>
static __device__ __noinline__ unsigned int reject_sorted_and_mean(Column c, int N, int rejection, double slow,
		double shigh, unsigned long long S, unsigned long long Q, RejectState &st,
		const SmallN *tables = nullptr, const FitTables *fit = nullptr) {
	double sigma, median;
	int n, r = 0, iter = 0, o = 0;

	switch (rejection) {
	case PERCENTILE: {	// stacking.c:1660-1673: one pass; removal stops when a single sample is left
		median = column_median(c, 0, N);
		int nl = 0, b = N;
		while (nl < N) {
			unsigned int v = c.get(nl);
			if (!(ddiv(dsub(median, (double) (int) v), median) > slow)) break;
			st.rej_lo++; S -= v; nl++;
		}
		while (b > nl) {
			unsigned int v = c.get(b - 1);
			if (!(ddiv(dsub((double) (int) v, median), median) > shigh)) break;
			st.rej_hi++; S -= v; b--;
		}
		if (b == nl) {		// everything flagged: the last sample is never removed (N > 1 test, :1667)
			S = c.get(N - 1);
			N = 1;
		} else {
			N = b - nl;
		}
		break;
	}
	case SIGMA:		// stacking.c:1674-1695
		do {
			sigma = sd_from_sums<SMALLN>(S, Q, N, tables);
			median = column_median(c, o, N);
			n = sigma_pass_sorted(c, o, N, slow, shigh, sigma, median, r, S, Q, st);
			if (++iter >= SSC_ITER_CAP) { st.cap = 1; break; }
		} while (n > 0 && N > 3);
		break;
	case SIGMEDIAN:		// stacking.c:1696-1709: rejected samples become round_to_WORD(median), N stays
		do {
			sigma = sd_from_sums<SMALLN>(S, Q, N, tables);
			median = column_median(c, 0, N);
			const unsigned int m = round_to_word(median);
			const double tl = dmul(slow, sigma), th = dmul(shigh, sigma);
			int cl = 0, b = N;
			while (cl < N) {
				unsigned int v = c.get(cl);
				if (!(dsub(median, (double) (int) v) > tl)) break;
				st.rej_lo++; S -= v; Q -= sq(v); cl++;
			}
			while (b > cl) {
				unsigned int v = c.get(b - 1);
				if (!(dsub((double) (int) v, median) > th)) break;
				st.rej_hi++; S -= v; Q -= sq(v); b--;
			}
			const int ch = N - b;
			n = cl + ch;
			if (n > 0) {
				S += (unsigned long long) n * m;
				Q += (unsigned long long) n * sq(m);
				// keep the stack sorted (what the reference's next quicksort_s restores):
				// kept samples below m slide down by cl, the n copies of m follow, the rest slides up by ch
				int pos = cl;
				while (pos < b && c.get(pos) < m) pos++;
				for (int t = cl; t < pos; t++) c.set(t - cl, c.get(t));
				for (int t = b - 1; t >= pos; t--) c.set(t + ch, c.get(t));
				for (int t = pos - cl; t < pos + ch; t++) c.set(t, m);
			}
			if (++iter >= SSC_ITER_CAP) { st.cap = 1; break; }
		} while (n > 0 && N > 3);
		break;
	case WINSORIZED:	// stacking.c:1710-1749
		do {
			sigma = sd_from_sums<SMALLN>(S, Q, N, tables);
			median = column_median(c, o, N);
			// w_stack is never materialised: clamping a sorted stack keeps it sorted and only
			// ever rewrites a prefix (to one value) and a suffix (to one value), so
			// w = [vlo x clo] ++ s[clo .. N-chi) ++ [vhi x chi] with exact running sums.
			int clo = 0, chi = 0, inner = 0;
			unsigned int vlo = 0, vhi = 0;
			unsigned long long Sm = S, Qm = Q;
			double sigma0;
			const int lhs = (N - 1) >> 1, rhs = N >> 1;
			do {
				const double m0 = dsub(median, dmul(1.5, sigma));
				const double m1 = dadd(median, dmul(1.5, sigma));
				const unsigned int r0 = round_to_word(m0), r1 = round_to_word(m1);
				if (clo > 0) {		// Winsorized(), stacking.c:1163-1168, on the prefix value
					if ((double) (int) vlo < m0) vlo = r0;
					else if ((double) (int) vlo > m1) vlo = r1;
				}
				if (chi > 0) {
					if ((double) (int) vhi < m0) vhi = r0;
					else if ((double) (int) vhi > m1) vhi = r1;
				}
				while (clo + chi < N) {
					unsigned int v = c.get(o + clo);
					if (!((double) (int) v < m0)) break;
					Sm -= v; Qm -= sq(v);
					clo++; vlo = r0;
				}
				while (clo + chi < N) {
					unsigned int v = c.get(o + N - 1 - chi);
					if (!((double) (int) v > m1)) break;
					Sm -= v; Qm -= sq(v);
					chi++; vhi = r1;
				}
				const unsigned int wl = lhs < clo ? vlo : (lhs >= N - chi ? vhi : c.get(o + lhs));
				const unsigned int wr = rhs < clo ? vlo : (rhs >= N - chi ? vhi : c.get(o + rhs));
				median = lhs == rhs ? (double) (int) wl : dmul((double) (int) (wl + wr), 0.5);
				const unsigned long long Sw = Sm + (unsigned long long) clo * vlo + (unsigned long long) chi * vhi;
				const unsigned long long Qw = Qm + (unsigned long long) clo * sq(vlo) + (unsigned long long) chi * sq(vhi);
				sigma0 = sigma;
				sigma = dmul(1.134, sd_from_sums<SMALLN>(Sw, Qw, N, tables));
				if (++inner >= SSC_ITER_CAP) { st.cap = 1; break; }
			} while (relative_change_exceeds(sigma, sigma0));
			n = sigma_pass_sorted(c, o, N, slow, shigh, sigma, median, r, S, Q, st);
			if (++iter >= SSC_ITER_CAP) { st.cap = 1; break; }
		} while (n > 0 && N > 3);
		break;
	case LINEARFIT:		// stacking.c:1750-1784
		do {
			double a, b;
			fit_line_literal(c, N, fit, a, b, sigma);	// gsl_fit_linear's literal running means, mean absolute residual
			const Reciprocal inv_sigma(sigma);	// line_clipping divides every residual by sigma (:1175,1179)
			int out = 0, f = 0;
			for (; f < N; f++) {
				unsigned int v = c.get(f);
				int fl = clip_line(inv_sigma, v, slow, shigh, f, a, b);
				if (fl) {
					r++; S -= v;
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
	// stacking.c:1790-1794: the double sum of WORDs is exact, S is that sum
	return mean_to_word<SMALLN>(S, N, tables);
}

}  // namespace ssc
