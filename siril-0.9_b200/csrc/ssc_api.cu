// ssc_api.cu -- the C ABI of libsiril_stack_cuda (include/siril_stack_cuda.h):
// the device layer (thin, asynchronous wrappers over the kernel launchers) and
// the operator layer (stack_method-compatible entry points that move a bound
// in-memory sequence through the GPUs band by band).
#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "ssc_kernels.h"

namespace {

thread_local std::string g_err;
std::mutex g_mu;

struct Bound {
	ssc_sequence_view view;
	std::vector<const WORD *> frames;
	std::vector<int> sx, sy;
	std::vector<double> exposure;
	bool have_reg = false;
};
std::map<const void *, std::shared_ptr<const Bound>> g_bound;	// a running operator keeps its own reference
ssc_callbacks g_cb = {nullptr, nullptr, nullptr};
ssc_result g_result;
bool g_has_result = false;
std::vector<int> g_devices = {0};

int fail(int code, const char *fmt, ...) {
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	g_err = buf;
	return code;
}

#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
	return fail(SSC_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); } while (0)

ssc_callbacks callbacks() {	// a snapshot: ssc_set_callbacks may run on another thread
	std::lock_guard<std::mutex> lk(g_mu);
	return g_cb;
}

void log_msg(const char *fmt, ...) {
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	const ssc_callbacks cb = callbacks();
	if (cb.log) cb.log(buf);
}

void progress(const char *text, double frac) {
	const ssc_callbacks cb = callbacks();
	if (cb.progress) cb.progress(text, frac);
}

bool keep_running() {
	const ssc_callbacks cb = callbacks();
	return cb.should_continue ? cb.should_continue() != 0 : true;
}

int check_band(const ssc_band *b) {
	if (!b || !b->frames) return fail(SSC_ERR_ARG, "null band");
	if (b->n_frames < 1 || b->width < 1 || b->height < 1 || b->layers < 1 || b->layers > 3)
		return fail(SSC_ERR_ARG, "bad band geometry");
	if (b->out_rows < 0 || b->out_row0 < 0 || b->out_row0 + b->out_rows > b->height)
		return fail(SSC_ERR_ARG, "bad output rows");
	if (b->buf_row0 < 0 || b->buf_rows < 1 || b->buf_row0 + b->buf_rows > b->height)
		return fail(SSC_ERR_ARG, "bad buffer rows");
	return SSC_OK;
}

}  // namespace

// ===========================================================================
// DEVICE layer
// ===========================================================================
extern "C" {

const char *ssc_last_error(void) { return g_err.c_str(); }
const char *ssc_version(void) { return "siril-stack-cuda 0.1 (sm_100a)"; }

int ssc_set_kernel_path(int mode) {
	if (mode < 0 || mode > 1) return fail(SSC_ERR_ARG, "kernel path must be 0 (automatic) or 1 (general kernel only)");
	ssc::g_kernel_path = mode;
	return SSC_OK;
}

int ssc_device_count(void) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
	return n;
}

int ssc_dev_stack_median(const ssc_band *band, int normalize, const double *coeff, WORD *out,
		size_t out_layer_stride, int out_buf_row0, void *stream) {
	int rc = check_band(band);
	if (rc) return rc;
	ssc::StackParams p{};
	p.band = *band;
	p.normalize = normalize;
	p.coeff = normalize == NO_NORM ? nullptr : coeff;
	p.rejection = NO_REJEC;
	p.out = out;
	p.out_layer_stride = out_layer_stride;
	p.out_buf_row0 = out_buf_row0;
	p.counters = nullptr;
	rc = ssc::launch_stack(p, 1, (cudaStream_t) stream);
	if (rc) return fail(rc, "stack_median launch failed (%s)", cudaGetErrorString(cudaGetLastError()));
	return SSC_OK;
}

int ssc_dev_stack_mean(const ssc_band *band, int normalize, const double *coeff, int type_of_rejection,
		double sig_low, double sig_high, WORD *out, size_t out_layer_stride, int out_buf_row0,
		uint64_t *counters, void *stream) {
	int rc = check_band(band);
	if (rc) return rc;
	if (type_of_rejection < NO_REJEC || type_of_rejection > LINEARFIT)
		return fail(SSC_ERR_ARG, "unknown rejection %d", type_of_rejection);
	ssc::StackParams p{};
	p.band = *band;
	p.normalize = normalize;
	p.coeff = normalize == NO_NORM ? nullptr : coeff;
	p.rejection = type_of_rejection;
	p.sig_low = sig_low;
	p.sig_high = sig_high;
	p.out = out;
	p.out_layer_stride = out_layer_stride;
	p.out_buf_row0 = out_buf_row0;
	p.counters = reinterpret_cast<unsigned long long *>(counters);
	rc = ssc::launch_stack(p, 0, (cudaStream_t) stream);
	if (rc) return fail(rc, "stack_mean launch failed (%s)", cudaGetErrorString(cudaGetLastError()));
	return SSC_OK;
}

int ssc_dev_sum_accumulate(const ssc_band *band, uint32_t *acc, unsigned long long *maxim, void *stream) {
	int rc = check_band(band);
	if (rc) return rc;
	rc = ssc::launch_sum_accumulate(*band, acc, maxim, (cudaStream_t) stream);
	if (rc) return fail(rc, "sum_accumulate launch failed");
	return SSC_OK;
}

int ssc_dev_sum_finish(const uint32_t *acc, size_t count, const unsigned long long *maxim, WORD *out,
		void *stream) {
	int rc = ssc::launch_sum_finish(acc, count, maxim, out, (cudaStream_t) stream);
	if (rc) return fail(rc, "sum_finish launch failed");
	return SSC_OK;
}

int ssc_dev_stack_extreme(const ssc_band *band, int take_max, WORD *out, unsigned int *extreme, void *stream) {
	int rc = check_band(band);
	if (rc) return rc;
	if (!out || !extreme) return fail(SSC_ERR_ARG, "null output");
	rc = ssc::launch_extreme(*band, take_max != 0, out, extreme, (cudaStream_t) stream);
	if (rc) return fail(rc, "addmax/addmin launch failed");
	return SSC_OK;
}

int ssc_dev_histogram(const ssc_band *band, uint32_t *hist, void *stream) {
	int rc = check_band(band);
	if (rc) return rc;
	rc = ssc::launch_histogram(*band, hist, (cudaStream_t) stream);
	if (rc) return fail(rc, "histogram launch failed");
	return SSC_OK;
}

int ssc_dev_ikss(const uint32_t *hist, int n_frames, double *stats, uint32_t *scratch, void *stream) {
	int rc = ssc::launch_ikss(hist, n_frames, stats, scratch, (cudaStream_t) stream);
	if (rc) return fail(rc, "ikss launch failed");
	return SSC_OK;
}

int ssc_dev_norm_coeffs(const double *stats, int n_frames, int ref_image, int mode, double *coeff,
		int *status, void *stream) {
	int rc = ssc::launch_norm_coeffs(stats, n_frames, ref_image, mode, coeff, status, (cudaStream_t) stream);
	if (rc) return fail(rc, "norm_coeffs launch failed");
	return SSC_OK;
}

int ssc_dev_synth_frames(WORD *frames, int n_frames, int layers, int rows, int width, int first_frame,
		int row0, int full_height, int kind, uint64_t seed, void *stream) {
	int rc = ssc::launch_synth(frames, n_frames, layers, rows, width, first_frame, row0, full_height, kind,
			seed, (cudaStream_t) stream);
	if (rc) return fail(rc, "synth launch failed");
	return SSC_OK;
}

// ===========================================================================
// OPERATOR layer
// ===========================================================================

int ssc_bind_sequence(const sequence *seq, const ssc_sequence_view *view) {
	if (!seq || !view || !view->frames) return fail(SSC_ERR_ARG, "null sequence view");
	if (view->number < 1 || view->rx < 1 || view->ry < 1 || (view->nb_layers != 1 && view->nb_layers != 3))
		return fail(SSC_ERR_ARG, "bad sequence geometry");
	auto bp = std::make_shared<Bound>();
	Bound &b = *bp;
	b.view = *view;
	b.frames.assign(view->frames, view->frames + view->number);
	b.have_reg = view->shiftx || view->shifty;
	if (b.have_reg) {
		b.sx.assign(view->number, 0);
		b.sy.assign(view->number, 0);
		if (view->shiftx) b.sx.assign(view->shiftx, view->shiftx + view->number);
		if (view->shifty) b.sy.assign(view->shifty, view->shifty + view->number);
	}
	if (view->exposure) b.exposure.assign(view->exposure, view->exposure + view->number);
	std::lock_guard<std::mutex> lk(g_mu);
	g_bound[seq] = std::move(bp);
	return SSC_OK;
}

int ssc_unbind_sequence(const sequence *seq) {
	std::lock_guard<std::mutex> lk(g_mu);
	return g_bound.erase(seq) ? SSC_OK : SSC_ERR_UNBOUND;
}

void ssc_set_callbacks(const ssc_callbacks *cb) {
	std::lock_guard<std::mutex> lk(g_mu);
	if (cb) g_cb = *cb; else g_cb = ssc_callbacks{nullptr, nullptr, nullptr};
}

int ssc_set_devices(const int *device_ids, int count) {
	int n = ssc_device_count();
	if (count < 1 || !device_ids) return fail(SSC_ERR_ARG, "empty device list");
	for (int i = 0; i < count; i++) {
		if (device_ids[i] < 0 || device_ids[i] >= n) return fail(SSC_ERR_CUDA, "device %d not present", device_ids[i]);
		for (int j = 0; j < i; j++)	// the per-device workspace is keyed by device id: one band per device
			if (device_ids[j] == device_ids[i]) return fail(SSC_ERR_ARG, "device %d listed twice", device_ids[i]);
	}
	std::lock_guard<std::mutex> lk(g_mu);
	g_devices.assign(device_ids, device_ids + count);
	return SSC_OK;
}

int ssc_take_result(ssc_result *out) {
	std::lock_guard<std::mutex> lk(g_mu);
	if (!out) return fail(SSC_ERR_ARG, "null result");
	if (!g_has_result) return fail(SSC_ERR_NO_RESULT, "no finished stack to take");
	*out = g_result;
	g_has_result = false;
	memset(&g_result, 0, sizeof(g_result));
	return SSC_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------
// engine
// ---------------------------------------------------------------------------
namespace {

enum OpKind { OP_SUM, OP_MEDIAN, OP_MEAN, OP_ADDMAX, OP_ADDMIN };
// the whole-frame operators (stack_summing, stack_addmax, stack_addmin) share geometry and frame selection
inline bool frame_walker(OpKind op) { return op == OP_SUM || op == OP_ADDMAX || op == OP_ADDMIN; }

struct Job {
	OpKind op;
	std::shared_ptr<const Bound> seq;	// stays alive through unbind / re-bind on another thread
	std::vector<int> images;	// sequence image index per stack position
	int normalize = NO_NORM, rejection = NO_REJEC;
	double sig[2] = {0, 0};
	int ref_pos = 0;
	bool want_coeff_only = false;
	double *coeff_out = nullptr;	// host [3][N]
};

struct Piece { int row0, rows; };	// result storage rows

// Device buffers are kept between operator calls (a 9 GB cudaMalloc/cudaFree pair costs far more
// than the stack itself); ssc_release_workspace() returns them.
struct Buf {
	void *p = nullptr;
	size_t cap = 0;
	int ensure(size_t bytes) {
		if (bytes <= cap) return SSC_OK;
		if (p) cudaFree(p);
		p = nullptr; cap = 0;
		if (cudaMalloc(&p, bytes) != cudaSuccess) return SSC_ERR_CUDA;
		cap = bytes;
		return SSC_OK;
	}
	void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct DevCache {
	cudaStream_t stream = nullptr;
	Buf frames, out, hist, scratch, acc, stats, coeff, status, shifts, counters, maxim;
	void release() {
		Buf *all[] = {&frames, &out, &hist, &scratch, &acc, &stats, &coeff, &status, &shifts, &counters, &maxim};
		for (Buf *b : all) b->release();
		if (stream) cudaStreamDestroy(stream);
		stream = nullptr;
	}
};
std::map<int, DevCache> g_cache;	// by device id; operator calls are serialised (g_op_mu)
std::mutex g_op_mu;
void *g_pinned = nullptr;		// pinned bounce buffer for the result rows
size_t g_pinned_cap = 0;

struct Dev {
	int id = 0;
	cudaStream_t stream = nullptr;
	std::vector<Piece> pieces;
	int max_rows = 0, max_buf_rows = 0;
	WORD *frames = nullptr, *out = nullptr;
	uint32_t *hist = nullptr, *scratch = nullptr, *acc = nullptr;
	double *stats = nullptr, *coeff = nullptr;
	int *status = nullptr, *shifts = nullptr;
	unsigned long long *counters = nullptr, *maxim = nullptr;
	bool resident = false;
};

// the source storage rows a piece needs: its own rows plus the reach of the y shifts
void source_rows(const Job &job, const Piece &pc, int H, int &buf_row0, int &buf_rows) {
	int min_sy = 0, max_sy = 0;
	if (job.op != OP_MEDIAN && job.seq->have_reg)
		for (int idx : job.images) {
			min_sy = std::min(min_sy, job.seq->sy[idx]);
			max_sy = std::max(max_sy, job.seq->sy[idx]);
		}
	int lo, hi;	// [lo, hi) storage rows
	const bool ser = job.seq->view.row_order == SSC_ROWS_SER;
	if (frame_walker(job.op)) {
		// fits row y reads fits row y - sy; fits row = storage row (FITS) or H-1-row (SER)
		int flo = pc.row0 - max_sy, fhi = pc.row0 + pc.rows - min_sy;
		if (ser) { lo = H - fhi; hi = H - flo; } else { lo = flo; hi = fhi; }
	} else if (!ser) {
		lo = pc.row0 - max_sy; hi = pc.row0 + pc.rows - min_sy;	// storage row R - sy
	} else {
		lo = H - (pc.row0 + pc.rows) + min_sy; hi = H - pc.row0 + max_sy;	// storage row (H-1-R) + sy
	}
	lo = std::max(lo, 0);
	hi = std::min(hi, H);
	if (hi <= lo) { lo = 0; hi = 1; }
	buf_row0 = lo;
	buf_rows = hi - lo;
}

int upload_piece(const Job &job, Dev &d, const Piece &pc, ssc_band &band) {
	const ssc_sequence_view &v = job.seq->view;
	const int W = v.rx, H = v.ry, C = v.nb_layers, N = (int) job.images.size();
	int buf_row0, buf_rows;
	source_rows(job, pc, H, buf_row0, buf_rows);
	band.frames = d.frames;
	band.layer_stride = (size_t) buf_rows * W;
	band.frame_stride = band.layer_stride * C;
	band.n_frames = N;
	band.layers = C;
	band.width = W;
	band.height = H;
	band.buf_row0 = buf_row0;
	band.buf_rows = buf_rows;
	band.out_row0 = pc.row0;
	band.out_rows = pc.rows;
	band.row_order = v.row_order;
	band.shiftx = (job.op != OP_MEDIAN && job.seq->have_reg) ? d.shifts : nullptr;
	band.shifty = (job.op != OP_MEDIAN && job.seq->have_reg) ? d.shifts + N : nullptr;
	const cudaMemcpyKind kind = v.frames_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
	for (int f = 0; f < N; f++) {
		const WORD *src = job.seq->frames[job.images[f]];
		if (!src) return fail(SSC_ERR_GENERIC, "frame %d has no data", job.images[f]);
		for (int c = 0; c < C; c++)
			CU(cudaMemcpyAsync(d.frames + (size_t) f * band.frame_stride + (size_t) c * band.layer_stride,
					src + ((size_t) c * H + buf_row0) * W, band.layer_stride * sizeof(WORD), kind, d.stream));
		if ((f & 15) == 15 && !keep_running()) return fail(SSC_ERR_GENERIC, "cancelled");
	}
	return SSC_OK;
}

int run_job(const Job &job) {
	using clk = std::chrono::steady_clock;
	const auto t_begin = clk::now();
	static const bool trace = getenv("SSC_TRACE") != nullptr;	// phase timeline on stderr (debug aid)
	auto mark = [&](const char *what) {
		if (trace) fprintf(stderr, "[ssc trace] %8.3f ms  %s\n", std::chrono::duration<double, std::milli>(clk::now() - t_begin).count(), what);
	};
	const ssc_sequence_view &v = job.seq->view;
	const int W = v.rx, H = v.ry, C = v.nb_layers, N = (int) job.images.size();
	const size_t plane = (size_t) W * H;
	const bool need_norm = job.normalize != NO_NORM;

	if (ssc_device_count() < 1) return fail(SSC_ERR_CUDA, "no CUDA device available");
	std::vector<int> ids;
	{
		std::lock_guard<std::mutex> lk(g_mu);
		ids = g_devices;
	}
	if ((int) ids.size() > H) ids.resize(H);
	const int G = (int) ids.size();
	std::vector<Dev> devs(G);
	std::lock_guard<std::mutex> op_lock(g_op_mu);	// one operator call in flight (as in Siril: com.run_thread)

	// --- plan: contiguous row bands per device (remainder rows to the first ones, as
	// stacking.c:620-625 does), each split further if the frames do not fit in HBM ---
	int row = 0;
	for (int g = 0; g < G; g++) {
		Dev &d = devs[g];
		d.id = ids[g];
		const int rows = H / G + (g < H % G ? 1 : 0);
		CU(cudaSetDevice(d.id));
		size_t free_b = 0, total_b = 0;
		mark("before cudaMemGetInfo");
		CU(cudaMemGetInfo(&free_b, &total_b));
		mark("after cudaMemGetInfo");
		{
			auto it = g_cache.find(d.id);	// what this library already holds on the device is reusable
			if (it != g_cache.end()) free_b += it->second.frames.cap + it->second.out.cap + it->second.hist.cap
					+ it->second.scratch.cap + it->second.acc.cap;
		}
		// what the library itself keeps on the device besides frames and result: histograms and IKSS scratch,
		// 4 bytes per result pixel for the hand-back lists of the register / warp kernels, and the extra
		// source rows the y shifts reach for (source_rows())
		size_t fixed = (size_t) N * (SSC_HIST_STRIDE + 65536) * 4 + (64u << 20);
		int reach = 0;
		if (job.op != OP_MEDIAN && job.seq->have_reg) {
			int lo = 0, hi = 0;
			for (int idx : job.images) { lo = std::min(lo, job.seq->sy[idx]); hi = std::max(hi, job.seq->sy[idx]); }
			reach = hi - lo;
		}
		fixed += (size_t) reach * N * C * W * sizeof(WORD);
		size_t budget = free_b > fixed ? (size_t) ((free_b - fixed) * 0.85) : 0;
		if (const char *e = getenv("SSC_HBM_BUDGET_BYTES")) {	// tests force the row split with it
			const unsigned long long cap = strtoull(e, nullptr, 10);
			if (cap > 0 && cap < budget) budget = (size_t) cap;
		}
		const size_t per_row = (size_t) N * C * W * sizeof(WORD) + (size_t) C * W * (sizeof(WORD) + 2 * sizeof(uint32_t));
		size_t fit_rows = per_row ? budget / per_row : 0;
		if (fit_rows < 1) return fail(SSC_ERR_GENERIC, "not enough device memory for one row of %d frames", N);
		const int n_pieces = (int) ((rows + fit_rows - 1) / fit_rows);
		for (int k = 0; k < n_pieces; k++) {
			int r = rows / n_pieces + (k < rows % n_pieces ? 1 : 0);
			d.pieces.push_back(Piece{row, r});
			row += r;
		}
		d.resident = n_pieces == 1;
		for (auto &pc : d.pieces) {
			int b0, br;
			source_rows(job, pc, H, b0, br);
			d.max_rows = std::max(d.max_rows, pc.rows);
			d.max_buf_rows = std::max(d.max_buf_rows, br);
		}
		DevCache &dc = g_cache[d.id];
		if (!dc.stream) CU(cudaStreamCreateWithFlags(&dc.stream, cudaStreamNonBlocking));
		d.stream = dc.stream;
#define SSC_ENSURE(buf, bytes) do { if (dc.buf.ensure(bytes)) return fail(SSC_ERR_CUDA, "device allocation of %zu bytes failed", (size_t) (bytes)); } while (0)
		SSC_ENSURE(frames, (size_t) N * C * d.max_buf_rows * W * sizeof(WORD));
		SSC_ENSURE(out, (size_t) C * d.max_rows * W * sizeof(WORD));
		SSC_ENSURE(counters, SSC_COUNTERS * sizeof(unsigned long long));
		SSC_ENSURE(coeff, (size_t) 3 * N * sizeof(double));
		d.frames = (WORD *) dc.frames.p; d.out = (WORD *) dc.out.p;
		d.counters = (unsigned long long *) dc.counters.p; d.coeff = (double *) dc.coeff.p;
		CU(cudaMemsetAsync(d.counters, 0, SSC_COUNTERS * sizeof(unsigned long long), d.stream));
		mark("buffers ensured");
		if (job.op == OP_SUM) {
			SSC_ENSURE(acc, (size_t) C * d.max_rows * W * sizeof(uint32_t));
			SSC_ENSURE(maxim, sizeof(unsigned long long));
			d.acc = (uint32_t *) dc.acc.p; d.maxim = (unsigned long long *) dc.maxim.p;
			CU(cudaMemsetAsync(d.maxim, 0, sizeof(unsigned long long), d.stream));
		}
		if (job.op == OP_ADDMAX || job.op == OP_ADDMIN) {
			SSC_ENSURE(maxim, sizeof(unsigned long long));
			d.maxim = (unsigned long long *) dc.maxim.p;
			const unsigned long long neutral = job.op == OP_ADDMAX ? 0ull : 0xFFFFull;	// low word = the kernel's unsigned int
			CU(cudaMemcpyAsync(d.maxim, &neutral, sizeof(neutral), cudaMemcpyHostToDevice, d.stream));
			CU(cudaStreamSynchronize(d.stream));
		}
		if (need_norm) {
			SSC_ENSURE(hist, (size_t) N * SSC_HIST_STRIDE * sizeof(uint32_t));
			d.hist = (uint32_t *) dc.hist.p;
			CU(cudaMemsetAsync(d.hist, 0, (size_t) N * SSC_HIST_STRIDE * sizeof(uint32_t), d.stream));
			if (g == 0) {
				SSC_ENSURE(scratch, (size_t) N * 65536 * sizeof(uint32_t));
				SSC_ENSURE(stats, (size_t) 4 * N * sizeof(double));
				SSC_ENSURE(status, sizeof(int));
				d.scratch = (uint32_t *) dc.scratch.p; d.stats = (double *) dc.stats.p; d.status = (int *) dc.status.p;
				CU(cudaMemsetAsync(d.status, 0, sizeof(int), d.stream));
			}
		}
		if (job.op != OP_MEDIAN && job.seq->have_reg) {
			std::vector<int> sh(2 * N);
			for (int f = 0; f < N; f++) {
				sh[f] = job.seq->sx[job.images[f]];
				sh[N + f] = job.seq->sy[job.images[f]];
			}
			SSC_ENSURE(shifts, 2 * N * sizeof(int));
			d.shifts = (int *) dc.shifts.p;
			CU(cudaMemcpyAsync(d.shifts, sh.data(), 2 * N * sizeof(int), cudaMemcpyHostToDevice, d.stream));
			CU(cudaStreamSynchronize(d.stream));	// sh goes out of scope
		}
	}

	// The result (malloc()ed: it becomes gfit.data, stacking.c:776-785) is allocated now and its pages
	// are touched on a helper thread while the frames upload: first-touch faults on 100+ MB would
	// otherwise sit between the download and the caller.
	WORD *result = job.want_coeff_only ? nullptr : (WORD *) malloc(plane * C * sizeof(WORD));
	struct ResultGuard { WORD *&p; ~ResultGuard() { free(p); } } rguard{result};
	struct Prefault {
		std::thread th;
		void join() { if (th.joinable()) th.join(); }
		~Prefault() { join(); }
	} prefault;
	if (result) {
		WORD *r = result;
		const size_t bytes = plane * C * sizeof(WORD);
		prefault.th = std::thread([r, bytes] { memset(r, 0, bytes); });
	}
	mark("planned, workspace ready");
	double t_h2d = 0, t_dev = 0, t_d2h = 0;
	auto t_phase = clk::now();
	auto phase_end = [&](double &acc_ms) -> int {	// sync all devices, book the wall time of this phase
		for (auto &d : devs) {
			CU(cudaSetDevice(d.id));
			CU(cudaStreamSynchronize(d.stream));
		}
		auto now = clk::now();
		acc_ms += std::chrono::duration<double, std::milli>(now - t_phase).count();
		t_phase = now;
		return SSC_OK;
	};
	{
		int rc = phase_end(t_dev);	// allocation / setup is booked as device time
		if (rc) return rc;
	}

	std::vector<ssc_band> bands(G);
	int rc;
	mark("setup synchronised");

	// --- normalization statistics: K4 per band, combine, K5, K6 ---
	std::vector<double> h_coeff((size_t) 3 * N);
	if (need_norm) {
		log_msg("Computing normalization...\n");
		progress("Computing normalization...", 0.0);
		size_t max_pieces = 0;
		for (auto &d : devs) max_pieces = std::max(max_pieces, d.pieces.size());
		for (size_t k = 0; k < max_pieces; k++) {
			for (int g = 0; g < G; g++) {
				Dev &d = devs[g];
				if (k >= d.pieces.size()) continue;
				CU(cudaSetDevice(d.id));
				if ((rc = upload_piece(job, d, d.pieces[k], bands[g]))) return rc;
			}
			if ((rc = phase_end(t_h2d))) return rc;
			for (int g = 0; g < G; g++) {
				Dev &d = devs[g];
				if (k >= d.pieces.size()) continue;
				CU(cudaSetDevice(d.id));
				ssc_band hb = bands[g];
				// histogram each storage row exactly once: the piece's own rows
				hb.out_row0 = v.row_order == SSC_ROWS_SER ? H - (d.pieces[k].row0 + d.pieces[k].rows) : d.pieces[k].row0;
				hb.out_rows = d.pieces[k].rows;
				if ((rc = ssc::launch_histogram(hb, d.hist, d.stream))) return fail(rc, "histogram launch failed");
			}
			if ((rc = phase_end(t_dev))) return rc;
		}
		// combine on device 0 (exact integer sums), then K5 + K6 there, broadcast the coefficients
		Dev &d0 = devs[0];
		CU(cudaSetDevice(d0.id));
		if (G > 1) {
			std::vector<uint32_t> h_sum((size_t) N * SSC_HIST_STRIDE), h_tmp((size_t) N * SSC_HIST_STRIDE);
			CU(cudaMemcpy(h_sum.data(), d0.hist, h_sum.size() * 4, cudaMemcpyDeviceToHost));
			for (int g = 1; g < G; g++) {
				CU(cudaSetDevice(devs[g].id));
				CU(cudaMemcpy(h_tmp.data(), devs[g].hist, h_tmp.size() * 4, cudaMemcpyDeviceToHost));
				for (size_t i = 0; i < h_sum.size(); i++) h_sum[i] += h_tmp[i];
			}
			CU(cudaSetDevice(d0.id));
			CU(cudaMemcpy(d0.hist, h_sum.data(), h_sum.size() * 4, cudaMemcpyHostToDevice));
		}
		if ((rc = ssc::launch_ikss(d0.hist, N, d0.stats, d0.scratch, d0.stream))) return fail(rc, "ikss launch failed");
		if ((rc = ssc::launch_norm_coeffs(d0.stats, N, job.ref_pos, job.normalize, d0.coeff, d0.status, d0.stream)))
			return fail(rc, "norm_coeffs launch failed");
		int h_status = 0;
		CU(cudaMemcpyAsync(h_coeff.data(), d0.coeff, h_coeff.size() * sizeof(double), cudaMemcpyDeviceToHost, d0.stream));
		CU(cudaMemcpyAsync(&h_status, d0.status, sizeof(int), cudaMemcpyDeviceToHost, d0.stream));
		CU(cudaStreamSynchronize(d0.stream));
		if (h_status) {
			log_msg("Error: no data computed.\n");
			return fail(SSC_ERR_GENERIC, "a frame has no non-zero pixel: no statistics");
		}
		for (int g = 1; g < G; g++) {
			CU(cudaSetDevice(devs[g].id));
			CU(cudaMemcpyAsync(devs[g].coeff, h_coeff.data(), h_coeff.size() * sizeof(double), cudaMemcpyHostToDevice, devs[g].stream));
		}
		if ((rc = phase_end(t_dev))) return rc;
		progress(nullptr, 1.0);
	} else {
		for (int i = 0; i < N; i++) { h_coeff[i] = 0.0; h_coeff[N + i] = 1.0; h_coeff[2 * N + i] = 1.0; }
	}
	if (job.coeff_out) memcpy(job.coeff_out, h_coeff.data(), h_coeff.size() * sizeof(double));
	if (job.want_coeff_only) return SSC_OK;
	if (!keep_running()) return fail(SSC_ERR_GENERIC, "cancelled");

	// --- the stack itself ---
	if (!result) return fail(SSC_ERR_GENERIC, "result allocation failed");

	mark("normalization done");
	log_msg("Starting stacking...\n");
	unsigned long long h_max = 0;
	size_t max_pieces = 0;
	for (auto &d : devs) max_pieces = std::max(max_pieces, d.pieces.size());
	const int sum_passes = job.op == OP_SUM ? 2 : 1;	// sum: all pieces accumulate the max before any is scaled
	for (int pass = 0; pass < sum_passes; pass++) {
		for (size_t k = 0; k < max_pieces; k++) {
			bool uploaded = false;
			for (int g = 0; g < G; g++) {
				Dev &d = devs[g];
				if (k >= d.pieces.size()) continue;
				CU(cudaSetDevice(d.id));
				const bool have = d.resident && (need_norm || pass > 0);
				if (!have) {
					if ((rc = upload_piece(job, d, d.pieces[k], bands[g]))) return rc;
					uploaded = true;
				}
			}
			if (uploaded && (rc = phase_end(t_h2d))) return rc;
			for (int g = 0; g < G; g++) {
				Dev &d = devs[g];
				if (k >= d.pieces.size()) continue;
				CU(cudaSetDevice(d.id));
				const Piece &pc = d.pieces[k];
				ssc::StackParams p{};
				p.band = bands[g];
				p.normalize = job.normalize;
				p.coeff = need_norm ? d.coeff : nullptr;
				p.rejection = job.rejection;
				p.sig_low = job.sig[0];
				p.sig_high = job.sig[1];
				p.out = d.out;
				p.out_layer_stride = (size_t) pc.rows * W;
				p.out_buf_row0 = pc.row0;
				p.counters = d.counters;
				if (job.op == OP_SUM) {
					if (pass == 0) {
						if ((rc = ssc::launch_sum_accumulate(bands[g], d.acc, d.maxim, d.stream))) return fail(rc, "sum launch failed");
					} else {
						if (!d.resident || max_pieces > 1) {
							// accumulators of other pieces were overwritten: recompute this piece's sums
							unsigned long long *scratch_max = d.counters;	// discard: the global max is already known
							if ((rc = ssc::launch_sum_accumulate(bands[g], d.acc, scratch_max, d.stream))) return fail(rc, "sum launch failed");
						}
						if ((rc = ssc::launch_sum_finish(d.acc, (size_t) C * pc.rows * W, d.maxim, d.out, d.stream))) return fail(rc, "sum finish failed");
					}
				} else if (job.op == OP_ADDMAX || job.op == OP_ADDMIN) {
					if ((rc = ssc::launch_extreme(bands[g], job.op == OP_ADDMAX, d.out, reinterpret_cast<unsigned int *>(d.maxim), d.stream)))
						return fail(rc, "addmax/addmin launch failed");
				} else {
					if ((rc = ssc::launch_stack(p, job.op == OP_MEDIAN, d.stream)))
						return fail(rc, "stack launch failed: %s", cudaGetErrorString(cudaGetLastError()));
				}
			}
			if ((rc = phase_end(t_dev))) return rc;
			if (job.op != OP_SUM || pass == 1) {
				// device -> pinned bounce buffer (full PCIe speed), then into the malloc()ed result
				if (g_pinned_cap < plane * C * sizeof(WORD)) {
					if (g_pinned) cudaFreeHost(g_pinned);
					g_pinned = nullptr; g_pinned_cap = 0;
					if (cudaHostAlloc(&g_pinned, plane * C * sizeof(WORD), cudaHostAllocDefault) == cudaSuccess)
						g_pinned_cap = plane * C * sizeof(WORD);
				}
				WORD *landing = g_pinned ? (WORD *) g_pinned : result;
				prefault.join();	// (long done: the upload took far longer than touching the pages)
				for (int g = 0; g < G; g++) {
					Dev &d = devs[g];
					if (k >= d.pieces.size()) continue;
					CU(cudaSetDevice(d.id));
					const Piece &pc = d.pieces[k];
					for (int c = 0; c < C; c++)
						CU(cudaMemcpyAsync(landing + (size_t) c * plane + (size_t) pc.row0 * W,
								d.out + (size_t) c * pc.rows * W, (size_t) pc.rows * W * sizeof(WORD),
								cudaMemcpyDeviceToHost, d.stream));
				}
				for (auto &d : devs) {
					CU(cudaSetDevice(d.id));
					CU(cudaStreamSynchronize(d.stream));
				}
				if (landing != result) {
					std::vector<std::thread> th;
					for (int g = 0; g < G; g++) {
						Dev &d = devs[g];
						if (k >= d.pieces.size()) continue;
						const Piece pc = d.pieces[k];
						for (int c = 0; c < C; c++) {
							const size_t off = (size_t) c * plane + (size_t) pc.row0 * W, cnt = (size_t) pc.rows * W;
							const int parts = 4;
							for (int q = 0; q < parts; q++) {
								const size_t a = cnt * q / parts, z = cnt * (q + 1) / parts;
								th.emplace_back([=] { memcpy(result + off + a, landing + off + a, (z - a) * sizeof(WORD)); });
							}
						}
					}
					for (auto &t : th) t.join();
				}
				if ((rc = phase_end(t_d2h))) return rc;
			}
			progress(nullptr, (double) (k + 1) / (double) max_pieces);
			if (!keep_running()) return fail(SSC_ERR_GENERIC, "cancelled");
		}
		if (job.op == OP_SUM && pass == 0) {
			// the only exchange of this operator: the global max (stacking.c:311-313,328-330)
			for (auto &d : devs) {
				unsigned long long m = 0;
				CU(cudaSetDevice(d.id));
				CU(cudaMemcpy(&m, d.maxim, sizeof(m), cudaMemcpyDeviceToHost));
				h_max = std::max(h_max, m);
			}
			for (auto &d : devs) {
				CU(cudaSetDevice(d.id));
				CU(cudaMemcpy(d.maxim, &h_max, sizeof(h_max), cudaMemcpyHostToDevice));
			}
		}
	}

	mark("stack + download done");
	ssc_result res;
	memset(&res, 0, sizeof(res));
	for (auto &d : devs) {
		unsigned long long h_cnt[SSC_COUNTERS];
		CU(cudaSetDevice(d.id));
		CU(cudaMemcpy(h_cnt, d.counters, sizeof(h_cnt), cudaMemcpyDeviceToHost));
		if (job.op == OP_MEAN) {
			for (int i = 0; i < 6; i++) res.irej[i / 2][i % 2] += h_cnt[i];
			res.guard_fired += h_cnt[6];
			res.cap_hit += h_cnt[7];
		}
	}
	res.data = result;
	result = nullptr;	// ownership moves to the result
	res.rx = W; res.ry = H;
	res.naxis = C == 3 ? 3 : 2;
	res.naxes[0] = W; res.naxes[1] = H; res.naxes[2] = C;
	res.pdata[0] = res.data;
	res.pdata[1] = C == 3 ? res.data + plane : res.data;
	res.pdata[2] = C == 3 ? res.data + 2 * plane : res.data;
	for (int idx : job.images)
		if (!job.seq->exposure.empty()) res.exposure += job.seq->exposure[idx];
	if (job.op == OP_SUM) res.hi = h_max > 65535ull ? 65535 : (WORD) h_max;	// round_to_WORD(maxim), :324
	if (job.op == OP_ADDMAX || job.op == OP_ADDMIN) {	// gfit.hi = round_to_WORD(maxim / minim), :947, :1103
		unsigned int e = job.op == OP_ADDMAX ? 0u : 0xFFFFu;
		for (auto &d : devs) {
			unsigned long long m = 0;
			CU(cudaSetDevice(d.id));
			CU(cudaMemcpy(&m, d.maxim, sizeof(m), cudaMemcpyDeviceToHost));
			e = job.op == OP_ADDMAX ? std::max(e, (unsigned int) m) : std::min(e, (unsigned int) m);
		}
		res.hi = (WORD) e;
	}
	res.t_h2d_ms = t_h2d;
	res.t_device_ms = t_dev;
	res.t_d2h_ms = t_d2h;
	res.t_total_ms = std::chrono::duration<double, std::milli>(clk::now() - t_begin).count();

	if (job.op == OP_MEAN) {
		const double nb_tot = (double) W * H * N;	// stacking.c:1811-1817
		for (int c = 0; c < C; c++)
			log_msg("Pixel rejection in channel #%d: %.3lf%% - %.3lf%%\n", c,
					res.irej[c][0] / nb_tot * 100.0, res.irej[c][1] / nb_tot * 100.0);
	}
	{
		std::lock_guard<std::mutex> lk(g_mu);
		if (g_has_result) free(g_result.data);
		g_result = res;
		g_has_result = true;
	}
	mark("result published");
	return SSC_OK;
}

int make_job(struct stacking_args *args, OpKind op, Job &job) {
	if (!args || !args->seq) return fail(SSC_ERR_ARG, "null stacking_args");
	{
		std::lock_guard<std::mutex> lk(g_mu);
		auto it = g_bound.find(args->seq);
		if (it == g_bound.end()) return fail(SSC_ERR_UNBOUND, "sequence not bound (ssc_bind_sequence)");
		job.seq = it->second;
	}
	job.op = op;
	const int number = job.seq->view.number;
	if (frame_walker(op)) {
		// stack_summing / addmax / addmin walk the whole sequence through the filter (stacking.c:223-231, 855-863)
		if (args->nb_images_to_stack <= 1) {
			log_msg("No frame selected for stacking (select at least 2). Aborting.\n");
			return fail(SSC_ERR_GENERIC, "fewer than 2 frames");
		}
		for (int j = 0; j < number; j++)
			if (!args->filtering_criterion || args->filtering_criterion(args->seq, j, args->filtering_parameter))
				job.images.push_back(j);
		if (job.images.size() < 1) return fail(SSC_ERR_GENERIC, "no frame passes the filter");
	} else {
		if (args->nb_images_to_stack < 2) {	// stacking.c:390-393,1217-1220
			log_msg("Select at least two frames for stacking. Aborting.\n");
			return fail(SSC_ERR_GENERIC, "fewer than 2 frames");
		}
		if (args->nb_images_to_stack > number) return fail(SSC_ERR_ARG, "more frames than the sequence holds");
		for (int i = 0; i < args->nb_images_to_stack; i++) {
			int idx = args->image_indices ? args->image_indices[i] : i;
			if (idx < 0 || idx >= number) return fail(SSC_ERR_ARG, "image index %d out of range", idx);
			job.images.push_back(idx);
		}
		job.normalize = args->normalize;
		if (job.normalize < NO_NORM || job.normalize > MULTIPLICATIVE_SCALING)
			// the "default:" labels: stack_median groups it with the additive branch (stacking.c:753-757),
			// stack_mean_with_rejection with NO_NORM (:1636-1638: raw samples, whatever coefficients
			// compute_normalization's own default: label produced)
			job.normalize = op == OP_MEAN ? NO_NORM : ADDITIVE_SCALING;
		if (op == OP_MEAN) {
			job.rejection = args->type_of_rejection;
			if (job.rejection < NO_REJEC || job.rejection > LINEARFIT) job.rejection = NO_REJEC;	// "default:" label
			job.sig[0] = args->sig[0];
			job.sig[1] = args->sig[1];
		}
		int ref = job.seq->view.reference_image;
		job.ref_pos = ref == -1 ? 0 : ref;	// stacking.c:143-145 (used as an index into the stack list)
		if (job.ref_pos < 0 || job.ref_pos >= (int) job.images.size())
			return fail(SSC_ERR_GENERIC, "reference image %d is not in the stack list", ref);
		if ((int) job.images.size() > SSC_GENERIC_MAX_FRAMES)
			return fail(SSC_ERR_ARG, "more than %d frames are not supported yet", SSC_GENERIC_MAX_FRAMES);
	}
	return SSC_OK;
}

int finish(struct stacking_args *args, int rc, const char *what) {
	if (rc) {
		log_msg("Stacking failed.\n");
		progress(what, 0.0);
		// keep the reference's codes; library-specific failures collapse to -1 for a stack_method caller
	}
	if (args) args->retval = rc;
	return rc;
}

}  // namespace

extern "C" {

int ssc_release_workspace(void) {
	std::lock_guard<std::mutex> op_lock(g_op_mu);
	for (auto &kv : g_cache) {
		cudaSetDevice(kv.first);
		kv.second.release();
		ssc::release_kernel_scratch(kv.first);
	}
	g_cache.clear();
	if (g_pinned) cudaFreeHost(g_pinned);
	g_pinned = nullptr;
	g_pinned_cap = 0;
	return SSC_OK;
}

int ssc_stack_summing(struct stacking_args *args) {
	Job job;
	int rc = make_job(args, OP_SUM, job);
	if (!rc) rc = run_job(job);
	return finish(args, rc, "Stacking failed. Check the log.");
}

int ssc_stack_addmax(struct stacking_args *args) {
	Job job;
	int rc = make_job(args, OP_ADDMAX, job);
	if (!rc) rc = run_job(job);
	return finish(args, rc, "Stacking failed. Check the log.");
}

int ssc_stack_addmin(struct stacking_args *args) {
	Job job;
	int rc = make_job(args, OP_ADDMIN, job);
	if (!rc) rc = run_job(job);
	return finish(args, rc, "Stacking failed. Check the log.");
}

int ssc_stack_median(struct stacking_args *args) {
	Job job;
	int rc = make_job(args, OP_MEDIAN, job);
	if (!rc) rc = run_job(job);
	if (!rc) log_msg("Median stacking complete. %d have been stacked.\n", (int) job.images.size());
	return finish(args, rc, "Median stacking failed. Check the log.");
}

int ssc_stack_mean_with_rejection(struct stacking_args *args) {
	Job job;
	int rc = make_job(args, OP_MEAN, job);
	if (!rc) rc = run_job(job);
	return finish(args, rc, "Rejection stacking failed. Check the log.");
}

int ssc_compute_normalization(struct stacking_args *args, norm_coeff *coeff, normalization mode) {
	if (!coeff || !coeff->offset || !coeff->mul || !coeff->scale) return fail(SSC_ERR_ARG, "null norm_coeff");
	Job job;
	int rc = make_job(args, OP_MEAN, job);
	if (rc) return rc == SSC_ERR_GENERIC ? 1 : rc;
	const int N = (int) job.images.size();
	job.normalize = (mode < NO_NORM || mode > MULTIPLICATIVE_SCALING) ? ADDITIVE_SCALING : mode;	// "default:" label, :96
	job.want_coeff_only = true;
	std::vector<double> c((size_t) 3 * N);
	job.coeff_out = c.data();
	if (mode == NO_NORM) {		// stacking.c:130-137
		for (int i = 0; i < N; i++) { coeff->offset[i] = 0.0; coeff->mul[i] = 1.0; coeff->scale[i] = 1.0; }
		return 0;
	}
	rc = run_job(job);
	if (rc) return rc == SSC_ERR_GENERIC ? 1 : rc;	// the reference returns 1 on failure (:161,189)
	for (int i = 0; i < N; i++) {
		coeff->offset[i] = c[i];
		coeff->mul[i] = c[N + i];
		coeff->scale[i] = c[2 * N + i];
	}
	return 0;
}

}  // extern "C"
