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
#include "ssc_column.cuh"
#include "ssc_kernels.h"
#include <mutex>
#include "ssc_leftover.cuh"
#include "ssc_network.cuh"
#include <cstdlib>

namespace ssc {

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
	// (one pixel per thread when the grid covers the image; list mode loops with a bounded grid)
	for (unsigned long long g = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; g < total;
			g += (unsigned long long) gridDim.x * blockDim.x) {
		unsigned long long pix = src.list ? src.list[g] : g;
		int x = (int) (pix % W);
		int rr = (int) ((pix / W) % b.out_rows);
		int layer = (int) (pix / ((unsigned long long) W * b.out_rows));
		int R = b.out_row0 + rr;		// result storage row
		int t = b.height - 1 - R;		// the reference's top-down block row (stacking.c:731,1600)
		Column c{smem_col + threadIdx.x, (int) blockDim.x};
		const bool have_reg = !is_median && (b.shiftx || b.shifty);	// stack_median ignores shifts
		const WORD *layer_base = b.frames + (size_t) layer * b.layer_stride;

		unsigned long long S = 0, Q = 0;
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
			S += v;
			Q += sq(v);
		}

		unsigned int result;
		RejectState st{0, 0, 0, 0};
		if (is_median || p.rejection != NO_REJEC) {
			for (int i = N; i < n_pad; i++) c.set(i, 0xFFFFu);	// sentinels sort behind the samples
			column_sort(c, n_pad);
		}
		if (is_median) {			// stacking.c:765-767, double -> WORD truncation
			int lhs = (N - 1) >> 1, rhs = N >> 1;
			result = (c.get(lhs) + c.get(rhs)) >> 1;
		} else {
			result = reject_sorted_and_mean<false>(c, N, p.rejection, p.sig_low, p.sig_high, S, Q, st, nullptr, &p.fit);
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
	unsigned long long blocks = (work + block - 1) / block;
	if (list) {	// the list is usually short and its length only known on the device
		int dev = 0, sms = 148;
		cudaGetDevice(&dev);
		cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
		if (blocks > (unsigned long long) sms * 8) blocks = (unsigned long long) sms * 8;
	}
	if (blocks > 0x7FFFFFFFull) return SSC_ERR_ARG;
	unsigned int grid = (unsigned int) blocks;
	k_stack_generic<<<grid, block, smem, stream>>>(p, src, n_pad, is_median);
	return cudaGetLastError() == cudaSuccess ? SSC_OK : SSC_ERR_CUDA;
}

// ---------------------------------------------------------------------------
// FitTables (ssc_kernels.h): built once per device with the literal recurrences of gsl_fit_linear
// for x_i = i, so that looking them up gives the very bits the per-pixel loops would compute.
// ---------------------------------------------------------------------------
__global__ void k_fit_tables(double *inv, double *mx, double *mdx2, double *inv_den, double *unbias, double *inv_n2, int n_max) {
	const int n = blockIdx.x * blockDim.x + threadIdx.x;
	if (n > n_max) return;
	if (n == 0) { inv[0] = 0.0; mx[0] = 0.0; mdx2[0] = 0.0; inv_den[0] = 0.0; unbias[0] = 0.0; inv_n2[0] = 0.0; return; }
	unbias[n] = n > 1 ? ddiv((double) n, (double) (n - 1)) : 0.0;
	inv_n2[n] = ddiv(1.0, dmul((double) n, (double) n));
	{
		const double n2 = dmul((double) n, (double) n);	// n^2 (n^2 - 1) / 12 is an integer below 2^41: exact
		inv_den[n] = n > 1 ? ddiv(1.0, ddiv(dmul(n2, dsub(n2, 1.0)), 12.0)) : 0.0;
	}
	inv[n] = ddiv(1.0, (double) n);
	double m_x = 0.0, m_dx2 = 0.0;
	for (int i = 0; i < n; i++) m_x = dadd(m_x, ddiv(dsub((double) i, m_x), dadd((double) i, 1.0)));
	for (int i = 0; i < n; i++) {
		const double dx = dsub((double) i, m_x);
		m_dx2 = dadd(m_dx2, ddiv(dsub(dmul(dx, dx), m_dx2), dadd((double) i, 1.0)));
	}
	mx[n] = m_x;
	mdx2[n] = m_dx2;
}

static double *g_fit_store[64];	// per device: [6][SSC_GENERIC_MAX_FRAMES + 1]

static int fit_tables(FitTables &fit, cudaStream_t stream) {
	int dev = 0;
	cudaGetDevice(&dev);
	if (dev < 0 || dev >= 64) return SSC_ERR_ARG;
	const int n_max = SSC_GENERIC_MAX_FRAMES, stride = n_max + 1;
	if (!g_fit_store[dev]) {
		double *t = nullptr;
		if (cudaMalloc(&t, (size_t) 6 * stride * sizeof(double)) != cudaSuccess) return SSC_ERR_CUDA;
		// built on the launching stream: ordered before the kernel that reads it; later launches on
		// other streams of this device come after this call returned, hence the synchronisation
		k_fit_tables<<<(stride + 127) / 128, 128, 0, stream>>>(t, t + stride, t + 2 * stride, t + 3 * stride, t + 4 * stride, t + 5 * stride, n_max);
		if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(stream) != cudaSuccess) { cudaFree(t); return SSC_ERR_CUDA; }
		g_fit_store[dev] = t;
	}
	fit.inv = g_fit_store[dev];
	fit.mx = g_fit_store[dev] + stride;
	fit.mdx2 = g_fit_store[dev] + 2 * stride;
	fit.inv_den = g_fit_store[dev] + 3 * stride;
	fit.unbias = g_fit_store[dev] + 4 * stride;
	fit.inv_n2 = g_fit_store[dev] + 5 * stride;
	fit.n_max = n_max;
	return SSC_OK;
}

// ---------------------------------------------------------------------------
// The cheap form of the normalization (ssc_network.cuh) is only used on frames for which it has been
// compared with the reference's form, normalize_sample(), on ALL 65536 sample values: one CTA per frame,
// bad[f] = 1 if any value differs (or the coefficients leave the 2^52 forms' domain).  ~3 us per launch.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_verify_cheap_norm(const double *__restrict__ coeff, int n, int mode, unsigned int *__restrict__ bad) {
	const int f = blockIdx.x;
	const double off = coeff[f], mul = coeff[n + f], sc = coeff[2 * n + f];
	int differs = !norm_domain_ok(mode, off, mul, sc);
	if (!differs) {
		for (unsigned int v = threadIdx.x; v < 65536u; v += 256u) {
			bool hi_ok;
			differs |= cheap_sample_dyn(v, mode, off, mul, sc, hi_ok) != normalize_sample(v, mode, off, mul, sc);
			differs |= !hi_ok;
		}
	}
	differs = __syncthreads_or(differs);
	if (threadIdx.x == 0) bad[f] = differs ? 1u : 0u;
}

static unsigned int *g_norm_bad_store[64];	// per device: [SSC_GENERIC_MAX_FRAMES]

// verifies the launch's coefficients on its stream; p.norm_bad stays nullptr (exact form) when switched off
static int verify_cheap_norm(StackParams &p, cudaStream_t stream) {
	static int enabled = -1;
	if (enabled < 0) { const char *e = getenv("SSC_CHEAP_NORM"); enabled = e ? atoi(e) != 0 : 1; }
	p.norm_bad = nullptr;
	const int n = p.band.n_frames;
	if (!enabled || g_kernel_path == 2 || !p.coeff || p.normalize == NO_NORM || n < 1 || n > SSC_GENERIC_MAX_FRAMES) return SSC_OK;
	int dev = 0;
	cudaGetDevice(&dev);
	if (dev < 0 || dev >= 64) return SSC_ERR_ARG;
	if (!g_norm_bad_store[dev] && cudaMalloc(&g_norm_bad_store[dev], SSC_GENERIC_MAX_FRAMES * sizeof(unsigned int)) != cudaSuccess)
		return SSC_ERR_CUDA;
	k_verify_cheap_norm<<<n, 256, 0, stream>>>(p.coeff, n, p.normalize, g_norm_bad_store[dev]);
	if (cudaGetLastError() != cudaSuccess) return SSC_ERR_CUDA;
	p.norm_bad = g_norm_bad_store[dev];
	return SSC_OK;
}

// frees the per-device scratch the kernels keep between launches (ssc_release_workspace)
void release_kernel_scratch(int dev) {
	if (dev < 0 || dev >= 64) return;
	if (g_fit_store[dev]) { cudaFree(g_fit_store[dev]); g_fit_store[dev] = nullptr; }
	if (g_norm_bad_store[dev]) { cudaFree(g_norm_bad_store[dev]); g_norm_bad_store[dev] = nullptr; }
	LeftoverIdStore &ids = g_leftover_id_store[dev];
	if (ids.ids) cudaFree(ids.ids);
	if (ids.count) cudaFree(ids.count);
	ids = LeftoverIdStore{};
}

int g_kernel_path = 0;

// The kernels' per-device scratch (sigma hand-back list, linear-fit spill area and tables) is shared by
// every launch on a device.  Launches are therefore chained: the host side runs under a per-device
// mutex and a launch on another stream than the previous one first waits for that launch's event, so two
// streams or two host threads can never see each other's hand-back list or spilled columns.
struct ScratchOwner {
	std::mutex mu;
	cudaEvent_t done = nullptr;
	cudaStream_t last = nullptr;
	bool any = false;
};
static ScratchOwner g_scratch_owner[64];

static int launch_stack_unchained(const StackParams &p_in, int is_median, cudaStream_t stream) {
	StackParams p = p_in;
	p.fit = FitTables{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0};
	// the per-n tables: gsl_fit_linear's data-independent recurrences for the linear fit, and the reciprocals that
	// take the divisions out of sigma / mean for 65..1024 frames (ssc_stack_mid.cu)
	if (!is_median && (p.rejection == LINEARFIT || (p.rejection != NO_REJEC && g_kernel_path != 1 && mid_path_eligible(p, is_median)))) {
		int rc = fit_tables(p.fit, stream);
		if (rc) return rc;
	}
	p.norm_bad = nullptr;
	if (g_kernel_path != 1 && warp_path_eligible(p, is_median))
		return launch_stack_warp(p, is_median, stream);
	if (g_kernel_path != 1 && fast_path_eligible(p, is_median))
		return launch_stack_fast(p, is_median, stream);
	if (g_kernel_path != 1 && mid_path_eligible(p, is_median)) {
		// this kernel normalizes with the cheap form where k_verify_cheap_norm allows it
		int rc = verify_cheap_norm(p, stream);
		if (rc) return rc;
		return launch_stack_mid(p, is_median, stream);
	}
	return launch_stack_generic(p, nullptr, nullptr, 0, is_median, stream);
}

// what launches with hb_defer listed (ssc_stack_warp.cu), finished in one go: p is the whole piece
int launch_deferred_handback(const StackParams &p_in, cudaStream_t stream) {
	int dev = 0;
	cudaGetDevice(&dev);
	if (dev < 0 || dev >= 64) return SSC_ERR_ARG;
	const LeftoverIdStore &st = g_leftover_id_store[dev];
	if (!st.count || !st.ids) return SSC_OK;	// nothing was ever listed on this device
	StackParams p = p_in;
	p.hb_defer = 0;
	int rc = fit_tables(p.fit, stream);
	if (rc) return rc;
	const unsigned long long pixels = (unsigned long long) p.band.layers * p.band.out_rows * p.band.width;
	LeftoverList list{st.count, (unsigned int) (pixels < st.capacity ? pixels : st.capacity), st.ids, st.count + 1};
	return launch_fit_handback(p, list, stream);
}

int launch_stack(const StackParams &p, int is_median, cudaStream_t stream) {
	int dev = 0;
	cudaGetDevice(&dev);
	if (dev < 0 || dev >= 64) return SSC_ERR_ARG;
	ScratchOwner &own = g_scratch_owner[dev];
	std::lock_guard<std::mutex> lk(own.mu);
	if (!own.done && cudaEventCreateWithFlags(&own.done, cudaEventDisableTiming) != cudaSuccess) return SSC_ERR_CUDA;
	if (own.any && own.last != stream && cudaStreamWaitEvent(stream, own.done, 0) != cudaSuccess) return SSC_ERR_CUDA;
	const int rc = launch_stack_unchained(p, is_median, stream);
	if (cudaEventRecord(own.done, stream) != cudaSuccess) return rc ? rc : SSC_ERR_CUDA;
	own.last = stream;
	own.any = true;
	return rc;
}

}  // namespace ssc
