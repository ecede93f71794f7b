// ssc_api.cu -- the C ABI of libsiril_stack_cuda (include/siril_stack_cuda.h):
// the device layer (thin, asynchronous wrappers over the kernel launchers) and
// the operator layer (stack_method-compatible entry points that move a bound
// in-memory sequence through the GPUs band by band).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <future>
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
	if (mode < 0 || mode > 2) return fail(SSC_ERR_ARG, "kernel path must be 0 (automatic), 1 (general kernel only) or 2 (automatic, exact normalization only)");
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
	if (view->frame_format < SSC_FRAMES_U16 || view->frame_format > SSC_FRAMES_SER_RGB) return fail(SSC_ERR_ARG, "unknown frame format");
	if (view->frame_format != SSC_FRAMES_U16 && view->frames_on_device) return fail(SSC_ERR_ARG, "device-resident frames must be native WORDs");
	if (view->frame_format == SSC_FRAMES_SER_RGB && (view->nb_layers != 3 || (view->ser_bytes_per_sample != 1 && view->ser_bytes_per_sample != 2)))
		return fail(SSC_ERR_ARG, "SER colour frames have three planes of 1 or 2 bytes per sample");
	if ((view->pp_offset || view->pp_dark || view->pp_flat) && view->frames_on_device) return fail(SSC_ERR_ARG, "preprocess masters go with host frames");
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
// One launch of a stacking kernel: a layer of a row chunk of a piece.  The frames of a chunk are uploaded,
// stacked and downloaded as a unit, so the three stages of consecutive units overlap.
struct Unit { int piece, layer, row0, rows; };

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
	cudaStream_t copy = nullptr, comp = nullptr, out = nullptr;	// upload, kernels, download
	cudaEvent_t t0 = nullptr, t1 = nullptr;				// the upload's duration
	Buf frames, outb, hist, scratch, acc, stats, coeff, status, shifts, counters, maxim, flags, staging, masters;
	std::vector<cudaEvent_t> events;
	// the row split of the previous call and what it was made for: cudaMemGetInfo takes 10-35 ms on a busy device,
	// a repeated call of the same shape keeps its plan (the buffers are still there)
	std::vector<long long> plan_key;
	std::vector<int> plan_piece_rows;
	ssc_peer_group *group = nullptr;	// this device's end of the NVLink exchanges (several devices only)
	std::vector<int> group_ids;
	int group_frames = 0;
	void release() {
		Buf *all[] = {&frames, &outb, &hist, &scratch, &acc, &stats, &coeff, &status, &shifts, &counters, &maxim, &flags, &staging, &masters};
		for (Buf *b : all) b->release();
		for (cudaEvent_t e : events) cudaEventDestroy(e);
		events.clear();
		if (group) ssc_peer_destroy(group);
		group = nullptr;
		for (cudaStream_t *s : {&copy, &comp, &out}) { if (*s) cudaStreamDestroy(*s); *s = nullptr; }
		for (cudaEvent_t *e : {&t0, &t1}) { if (*e) cudaEventDestroy(*e); *e = nullptr; }
	}
	int event(size_t i, cudaEvent_t &e) {
		while (events.size() <= i) {
			cudaEvent_t n;
			if (cudaEventCreateWithFlags(&n, cudaEventDisableTiming) != cudaSuccess) return SSC_ERR_CUDA;
			events.push_back(n);
		}
		e = events[i];
		return SSC_OK;
	}
};
std::map<int, DevCache> g_cache;	// by device id; operator calls are serialised (g_op_mu)
std::mutex g_op_mu;
void *g_pinned = nullptr;		// pinned bounce buffer for the result rows
size_t g_pinned_cap = 0;

struct Dev {
	int id = 0, index = 0;
	DevCache *dc = nullptr;
	std::vector<Piece> pieces;
	std::vector<Unit> units;
	int max_rows = 0, max_buf_rows = 0;
	WORD *frames = nullptr, *out = nullptr;
	uint32_t *hist = nullptr, *scratch = nullptr, *acc = nullptr, *flags = nullptr;
	double *stats = nullptr, *coeff = nullptr;
	int *status = nullptr, *shifts = nullptr;
	unsigned long long *counters = nullptr, *maxim = nullptr;	// counters: [layers][SSC_COUNTERS]
	ssc_peer_group *group = nullptr;
	unsigned char *staging = nullptr;	// raw bytes of one unit of every frame (formats that are not decoded in place)
	WORD *masters = nullptr;		// preprocess masters of the current piece: [offset, dark, flat][C][buf_rows][W]
	int rc = SSC_OK;
	std::string err;
	double t_h2d = 0, t_dev = 0, t_d2h = 0;
};

// the source storage rows a range of result rows needs: its own rows plus the reach of the y shifts
void source_rows(const Job &job, int row0, int rows, int H, int &buf_row0, int &buf_rows) {
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
		int flo = row0 - max_sy, fhi = row0 + rows - min_sy;
		if (ser) { lo = H - fhi; hi = H - flo; } else { lo = flo; hi = fhi; }
	} else if (!ser) {
		lo = row0 - max_sy; hi = row0 + rows - min_sy;	// storage row R - sy
	} else {
		lo = H - (row0 + rows) + min_sy; hi = H - row0 + max_sy;	// storage row (H-1-R) + sy
	}
	lo = std::max(lo, 0);
	hi = std::min(hi, H);
	if (hi <= lo) { lo = 0; hi = 1; }
	buf_row0 = lo;
	buf_rows = hi - lo;
}

// what a device thread needs to know about the call
struct Plan {
	const Job *job;
	int W, H, C, N, G;
	bool need_norm, shifted, strided;
	size_t host_stride;		// bytes between consecutive frames on the host when they are equally spaced
	WORD *result, *landing;
	size_t plane;
	const std::atomic<bool> *cancel;
	std::shared_future<void> result_ready;	// the result's pages have been touched (see run_job)
	std::chrono::steady_clock::time_point t_begin;
	bool trace;
};

// the device-resident band of a piece
void piece_band(const Plan &pl, const Dev &d, const Piece &pc, ssc_band &band) {
	const ssc_sequence_view &v = pl.job->seq->view;
	int b0, br;
	source_rows(*pl.job, pc.row0, pc.rows, pl.H, b0, br);
	band.frames = d.frames;
	band.layer_stride = (size_t) br * pl.W;
	band.frame_stride = band.layer_stride * pl.C;
	band.n_frames = pl.N;
	band.layers = pl.C;
	band.width = pl.W;
	band.height = pl.H;
	band.buf_row0 = b0;
	band.buf_rows = br;
	band.out_row0 = pc.row0;
	band.out_rows = pc.rows;
	band.row_order = v.row_order;
	band.shiftx = pl.shifted ? d.shifts : nullptr;
	band.shifty = pl.shifted ? d.shifts + pl.N : nullptr;
}

// `bytes` bytes of every frame, starting `src_off` bytes into it, to dst + f * dst_pitch: one pitched copy when
// the frames are equally spaced in host memory, one copy per frame otherwise
int copy_frames(const Plan &pl, void *dst, size_t dst_pitch, size_t src_off, size_t bytes, cudaStream_t st) {
	const cudaMemcpyKind kind = pl.job->seq->view.frames_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
	if (pl.strided) {
		const char *src = (const char *) pl.job->seq->frames[pl.job->images[0]] + src_off;
		return cudaMemcpy2DAsync(dst, dst_pitch, src, pl.host_stride, bytes, (size_t) pl.N, kind, st) == cudaSuccess ? SSC_OK : SSC_ERR_CUDA;
	}
	for (int f = 0; f < pl.N; f++) {
		const char *src = (const char *) pl.job->seq->frames[pl.job->images[f]];
		if (!src) return SSC_ERR_GENERIC;
		if (cudaMemcpyAsync((char *) dst + (size_t) f * dst_pitch, src + src_off, bytes, kind, st) != cudaSuccess) return SSC_ERR_CUDA;
	}
	return SSC_OK;
}

// upload the source rows of result rows [row0, row0 + rows) of one layer of every frame into the piece's band,
// decode them from the file's format and run preprocess on them -- all on the copy stream, so whatever waits
// for the unit's event sees calibrated WORDs
int upload_rows(const Plan &pl, Dev &d, const ssc_band &band, int layer, int row0, int rows, cudaStream_t st) {
	const ssc_sequence_view &v = pl.job->seq->view;
	int s0, sr;
	source_rows(*pl.job, row0, rows, pl.H, s0, sr);
	const size_t count = (size_t) sr * pl.W;		// samples per frame and layer
	WORD *dst = d.frames + (size_t) layer * band.layer_stride + (size_t) (s0 - band.buf_row0) * pl.W;
	const size_t pitch = band.frame_stride * sizeof(WORD);
	int first_layer = layer, n_layers = 1;
	switch (v.frame_format) {
	case SSC_FRAMES_U16:
	case SSC_FRAMES_FITS16:
	case SSC_FRAMES_RAW16_BE: {
		int rc = copy_frames(pl, dst, pitch, ((size_t) layer * pl.H + s0) * pl.W * sizeof(WORD), count * sizeof(WORD), st);
		if (rc) return rc;
		if (v.frame_format != SSC_FRAMES_U16
				&& (rc = ssc::launch_decode_inplace(dst, band.frame_stride, pl.N, count, v.frame_format == SSC_FRAMES_FITS16 ? 1 : 2, st)))
			return rc;
		break;
	}
	case SSC_FRAMES_RAW8: {
		int rc = copy_frames(pl, d.staging, count, ((size_t) layer * pl.H + s0) * pl.W, count, st);
		if (rc) return rc;
		if ((rc = ssc::launch_decode_raw(d.staging, count, dst, band.frame_stride, band.layer_stride, pl.N, count, 1, 1, 0, 0, 0, st))) return rc;
		break;
	}
	default: {	// SSC_FRAMES_SER_RGB: the three colours of a row are interleaved -- they arrive with layer 0
		if (layer != 0) return SSC_OK;
		const size_t raw = count * 3 * v.ser_bytes_per_sample;
		int rc = copy_frames(pl, d.staging, raw, (size_t) s0 * pl.W * 3 * v.ser_bytes_per_sample, raw, st);
		if (rc) return rc;
		if ((rc = ssc::launch_decode_raw(d.staging, raw, dst, band.frame_stride, band.layer_stride, pl.N, count, v.ser_bytes_per_sample, 3,
				v.ser_big_endian, v.ser_bgr, 0, st))) return rc;
		n_layers = 3;
		break;
	}
	}
	if (d.masters) {	// preprocess (siril.c:945-962) on what was just decoded
		const size_t mplane = (size_t) pl.C * band.layer_stride;	// one master of this piece
		for (int c = first_layer; c < first_layer + n_layers; c++) {
			const size_t at = (size_t) c * band.layer_stride + (size_t) (s0 - band.buf_row0) * pl.W;
			int rc = ssc::launch_preprocess(d.frames + at, band.frame_stride, pl.N, count, v.pp_offset ? d.masters + at : nullptr,
					v.pp_dark ? d.masters + mplane + at : nullptr, v.pp_flat ? d.masters + 2 * mplane + at : nullptr, v.pp_flat_level, st);
			if (rc) return rc;
		}
	}
	return SSC_OK;
}

// the preprocess masters' rows of a piece, once per piece (on the copy stream, ahead of the piece's units)
int upload_masters(const Plan &pl, Dev &d, const ssc_band &band, cudaStream_t st) {
	if (!d.masters) return SSC_OK;
	const ssc_sequence_view &v = pl.job->seq->view;
	const WORD *src[3] = {v.pp_offset, v.pp_dark, v.pp_flat};
	const size_t mplane = (size_t) pl.C * band.layer_stride;
	for (int m = 0; m < 3; m++) {
		if (!src[m]) continue;
		for (int c = 0; c < pl.C; c++)
			if (cudaMemcpyAsync(d.masters + m * mplane + (size_t) c * band.layer_stride, src[m] + ((size_t) c * pl.H + band.buf_row0) * pl.W,
					band.layer_stride * sizeof(WORD), cudaMemcpyHostToDevice, st) != cudaSuccess) return SSC_ERR_CUDA;
	}
	return SSC_OK;
}

// the band of one unit: a single layer, a row chunk
void unit_band(const ssc_band &piece, const Unit &u, ssc_band &b) {
	b = piece;
	b.frames = piece.frames + (size_t) u.layer * piece.layer_stride;
	b.layers = 1;
	b.out_row0 = u.row0;
	b.out_rows = u.rows;
}

#define DCU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { d.err = std::string(#call) + ": " + cudaGetErrorString(e_); return SSC_ERR_CUDA; } } while (0)
#define DRC(call, what) do { int r_ = (call); if (r_) { d.err = what; return r_; } } while (0)

// K4 for one range of result rows of a piece (each storage row is counted exactly once: the rows themselves)
int histogram_rows(const Plan &pl, Dev &d, const ssc_band &piece, int row0, int rows, bool layer0_only, cudaStream_t st) {
	ssc_band hb = piece;
	hb.out_row0 = piece.row_order == SSC_ROWS_SER ? pl.H - (row0 + rows) : row0;
	hb.out_rows = rows;
	uint32_t *hist = d.hist;
	if (d.group && ssc::peer_local_histograms(d.group, pl.N, &hist, st)) return SSC_ERR_CUDA;
	return ssc::launch_histogram_ex(hb, hist, nullptr, d.flags, layer0_only, st);
}

// K5 + K6 once this device's counts are complete; the coefficients end up in d.coeff on every device
int statistics_and_coefficients(const Plan &pl, Dev &d, cudaStream_t st) {
	const Job &job = *pl.job;
	if (d.group) {
		DRC(ssc::peer_signal_histograms(d.group, st), "histogram signal failed");
		DRC(ssc_dev_ikss_gather(d.group, pl.N, d.scratch, st), ssc_peer_last_error());
		DRC(ssc_dev_norm_coeffs_gathered(d.group, pl.N, job.ref_pos, job.normalize, d.coeff, d.status, nullptr, st), ssc_peer_last_error());
	} else {
		DRC(ssc::launch_ikss(d.hist, pl.N, d.stats, d.scratch, st), "ikss launch failed");
		DRC(ssc::launch_norm_coeffs(d.stats, pl.N, job.ref_pos, job.normalize, d.coeff, d.status, st), "norm_coeffs launch failed");
	}
	return SSC_OK;
}

// one unit through its kernel (everything but stack_summing's second pass)
int launch_unit(const Plan &pl, Dev &d, const ssc_band &piece, const Piece &pc, const Unit &u, int hb_defer, cudaStream_t st) {
	const Job &job = *pl.job;
	ssc_band ub;
	unit_band(piece, u, ub);
	const size_t layer_out = (size_t) pc.rows * pl.W;
	if (job.op == OP_SUM)
		return ssc::launch_sum_accumulate(ub, d.acc + (size_t) u.layer * layer_out + (size_t) (u.row0 - pc.row0) * pl.W, d.maxim, st);
	if (job.op == OP_ADDMAX || job.op == OP_ADDMIN)
		return ssc::launch_extreme(ub, job.op == OP_ADDMAX, d.out + (size_t) u.layer * layer_out + (size_t) (u.row0 - pc.row0) * pl.W,
				reinterpret_cast<unsigned int *>(d.maxim), st);
	ssc::StackParams p{};
	p.band = ub;
	p.normalize = job.normalize;
	p.coeff = pl.need_norm ? d.coeff : nullptr;
	p.rejection = job.rejection;
	p.sig_low = job.sig[0];
	p.sig_high = job.sig[1];
	p.out = d.out + (size_t) u.layer * layer_out;
	p.out_layer_stride = layer_out;
	p.out_buf_row0 = pc.row0;
	p.counters = d.counters + (size_t) u.layer * SSC_COUNTERS;	// the kernel sees layer 0: one counter set per layer
	p.hb_defer = hb_defer;		// (linear fit: undecided pixels are listed in the piece's geometry and finished once)
	p.hb_layer0 = u.layer;
	p.hb_row0 = u.row0 - pc.row0;
	p.hb_rows = pc.rows;
	p.hb_pixels = (unsigned long long) pl.C * pc.rows * pl.W;
	return ssc::launch_stack(p, job.op == OP_MEDIAN, st);
}

// download result rows [row0, row0 + rows) of a layer of a piece into the landing buffer
int download_rows(const Plan &pl, Dev &d, const Piece &pc, int layer, int row0, int rows, cudaStream_t st) {
	DCU(cudaMemcpyAsync(pl.landing + (size_t) layer * pl.plane + (size_t) row0 * pl.W,
			d.out + ((size_t) layer * pc.rows + (row0 - pc.row0)) * pl.W, (size_t) rows * pl.W * sizeof(WORD),
			cudaMemcpyDeviceToHost, st));
	return SSC_OK;
}

void land(const Plan &pl, int layer, int row0, int rows) {	// pinned landing buffer -> the malloc()ed result
	if (pl.landing == pl.result) return;
	pl.result_ready.wait();
	const size_t off = (size_t) layer * pl.plane + (size_t) row0 * pl.W;
	memcpy(pl.result + off, pl.landing + off, (size_t) rows * pl.W * sizeof(WORD));
}

double ms_since(std::chrono::steady_clock::time_point t0) {
	return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

// ---------------------------------------------------------------------------------------------
// One device's share of the call, on its own host thread.  `phase` splits the work where the devices
// have to meet on the host: 0 = everything up to stack_summing's maximum, 1 = the rest of stack_summing.
//
// Resident band (one piece): the frames go up chunk by chunk -- with normalization layer 0 of all
// chunks first -- and everything else runs behind the copy front on other streams: K4 on every layer-0
// chunk as it lands, K5 / K6 (with the NVLink exchange when several devices work), K3 / K2 on every unit
// as soon as it is complete, its result rows down while the next unit is stacked.  The end-to-end time
// is the upload plus one unit's kernel and download.
// Several pieces (frames larger than the HBM budget): piece after piece, first all of them for the
// statistics, then again for the stack.
// ---------------------------------------------------------------------------------------------
int device_run(const Plan &pl, Dev &d, int phase) {
	const Job &job = *pl.job;
	DevCache &dc = *d.dc;
	DCU(cudaSetDevice(d.id));
	const bool resident = d.pieces.size() == 1;
	const auto t_begin = std::chrono::steady_clock::now();
	auto dmark = [&](const char *what) {
		if (pl.trace) fprintf(stderr, "[ssc trace] %8.3f ms  device %d: %s\n", ms_since(pl.t_begin), d.id, what);
	};
	dmark("thread started");
	size_t ev_next = 0;
	auto new_event = [&](cudaEvent_t &e) -> int { return dc.event(ev_next++, e); };

	if (job.op == OP_SUM && phase == 1) {	// K1b with the global maximum (already in d.maxim)
		for (size_t k = 0; k < d.pieces.size(); k++) {
			const Piece &pc = d.pieces[k];
			ssc_band piece;
			piece_band(pl, d, pc, piece);
			if (!resident) {	// the accumulators of other pieces were overwritten: this piece's sums again
				DRC(upload_masters(pl, d, piece, dc.comp), "upload failed");
				for (int c = 0; c < pl.C; c++) DRC(upload_rows(pl, d, piece, c, pc.row0, pc.rows, dc.comp), "upload failed");
				for (const Unit &u : d.units)
					if (u.piece == (int) k) DRC(ssc::launch_sum_accumulate([&] { ssc_band ub; unit_band(piece, u, ub); return ub; }(),
							d.acc + ((size_t) u.layer * pc.rows + (u.row0 - pc.row0)) * pl.W, d.counters, dc.comp), "sum launch failed");
			}
			DRC(ssc::launch_sum_finish(d.acc, (size_t) pl.C * pc.rows * pl.W, d.maxim, d.out, dc.comp), "sum finish failed");
			for (int c = 0; c < pl.C; c++) DRC(download_rows(pl, d, pc, c, pc.row0, pc.rows, dc.comp), "download failed");
			DCU(cudaStreamSynchronize(dc.comp));
			for (int c = 0; c < pl.C; c++) land(pl, c, pc.row0, pc.rows);
		}
		d.t_d2h += ms_since(t_begin);
		return SSC_OK;
	}

	// ---- statistics over several pieces (or the statistics alone, ssc_compute_normalization): a pass of its own ----
	const bool stats_only = d.units.empty();
	if (pl.need_norm && (!resident || stats_only)) {
		for (size_t k = 0; k < d.pieces.size(); k++) {
			const Piece &pc = d.pieces[k];
			ssc_band piece;
			piece_band(pl, d, pc, piece);
			const auto t0 = std::chrono::steady_clock::now();
			DRC(upload_masters(pl, d, piece, dc.comp), "upload failed");
			for (int c = 0; c < pl.C; c++) DRC(upload_rows(pl, d, piece, c, pc.row0, pc.rows, dc.comp), "upload failed");
			DCU(cudaStreamSynchronize(dc.comp));
			d.t_h2d += ms_since(t0);
			DRC(histogram_rows(pl, d, piece, pc.row0, pc.rows, false, dc.comp), "histogram launch failed");
			DCU(cudaStreamSynchronize(dc.comp));	// the next piece overwrites the frames
			if (*pl.cancel) return SSC_ERR_GENERIC;
		}
		DRC(statistics_and_coefficients(pl, d, dc.comp), d.err.c_str());
		DCU(cudaStreamSynchronize(dc.comp));
	}
	if (stats_only) return SSC_OK;

	// ---- the units ----
	std::vector<cudaEvent_t> up(d.units.size()), done(d.units.size()), down(d.units.size());
	for (size_t k = 0; k < d.pieces.size(); k++) {
		const Piece &pc = d.pieces[k];
		ssc_band piece;
		piece_band(pl, d, pc, piece);
		std::vector<size_t> mine;	// this piece's units, in upload order
		for (size_t i = 0; i < d.units.size(); i++) if (d.units[i].piece == (int) k) mine.push_back(i);
		if (resident && pl.need_norm)	// layer 0 first: the statistics only need that one
			std::stable_sort(mine.begin(), mine.end(), [&](size_t a, size_t b) { return (d.units[a].layer != 0) < (d.units[b].layer != 0); });
		const auto t_up = std::chrono::steady_clock::now();
		if (resident) DCU(cudaEventRecord(dc.t0, dc.copy));
		DRC(upload_masters(pl, d, piece, dc.copy), "upload failed");
		for (size_t i : mine) {	// the copy stream runs ahead of everything else
			const Unit &u = d.units[i];
			DRC(upload_rows(pl, d, piece, u.layer, u.row0, u.rows, dc.copy), "upload failed");
			DRC(new_event(up[i]), "event");
			DCU(cudaEventRecord(up[i], dc.copy));
		}
		if (resident) DCU(cudaEventRecord(dc.t1, dc.copy));
		dmark("uploads enqueued");
		if (resident && pl.need_norm) {
			// K4 behind the copy front, layer 0 chunk by chunk
			for (size_t i : mine) {
				const Unit &u = d.units[i];
				if (u.layer != 0) continue;
				DCU(cudaStreamWaitEvent(dc.comp, up[i], 0));
				DRC(histogram_rows(pl, d, piece, u.row0, u.rows, true, dc.comp), "histogram launch failed");
			}
			if (pl.C > 1) {
				// get_normalized_value (utils.c:454-459) looks at all layers, but only for a frame whose layer 0
				// stays <= 255 in this band can the others matter: ask, and only then wait for the whole upload
				std::vector<uint32_t> h_flags(pl.N);
				DCU(cudaMemcpyAsync(h_flags.data(), d.flags, (size_t) pl.N * sizeof(uint32_t), cudaMemcpyDeviceToHost, dc.comp));
				DCU(cudaStreamSynchronize(dc.comp));
				bool all_set = true;
				for (uint32_t f : h_flags) all_set = all_set && f != 0;
				if (!all_set) {
					DCU(cudaStreamWaitEvent(dc.comp, up[mine.back()], 0));
					ssc_band hb = piece;
					hb.out_row0 = piece.row_order == SSC_ROWS_SER ? pl.H - (pc.row0 + pc.rows) : pc.row0;
					hb.out_rows = pc.rows;
					uint32_t *hist = d.hist;
					if (d.group) DRC(ssc::peer_local_histograms(d.group, pl.N, &hist, dc.comp), "histogram launch failed");
					DRC(ssc::launch_other_layers(hb, hist, nullptr, d.flags, dc.comp), "histogram launch failed");
				}
			}
			DRC(statistics_and_coefficients(pl, d, dc.comp), d.err.c_str());
			dmark("statistics enqueued");
		}
		if (!resident) {	// (several pieces: the stack of this piece starts when its upload is complete)
			DCU(cudaStreamSynchronize(dc.copy));
			d.t_h2d += ms_since(t_up);
		}
		// Linear fit through the warp kernel: the few pixels its screening cannot decide go through gsl_fit_linear's
		// literal recurrences, ~8 ms per launch however few they are -- listed by every unit, finished once per piece,
		// and the piece's rows go down after that
		bool defer = false;
		if (job.op == OP_MEAN && job.rejection == LINEARFIT && ssc::g_kernel_path != 1 && !mine.empty()) {
			ssc::StackParams probe{};
			unit_band(piece, d.units[mine[0]], probe.band);
			probe.rejection = LINEARFIT;
			probe.out = d.out;
			probe.out_layer_stride = (size_t) pc.rows * pl.W;
			defer = ssc::warp_path_eligible(probe, 0);
		}
		const bool streamed_out = job.op != OP_SUM && !defer;	// stack_summing writes after the global maximum is known
		std::vector<cudaEvent_t> tk0, tk1;		// SSC_TRACE: when every unit's kernels ran
		const bool unit_trace = pl.trace && resident;	// (dc.t0: the moment the first upload began)
		for (size_t i : mine) {
			const Unit &u = d.units[i];
			DCU(cudaStreamWaitEvent(dc.comp, up[i], 0));
			if (unit_trace) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, dc.comp); tk0.push_back(e); }
			DRC(launch_unit(pl, d, piece, pc, u, defer ? (i == mine[0] ? 2 : 1) : 0, dc.comp), "stack launch failed");
			if (unit_trace) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, dc.comp); tk1.push_back(e); }
			if (streamed_out) {
				DRC(new_event(done[i]), "event");
				DCU(cudaEventRecord(done[i], dc.comp));
				DCU(cudaStreamWaitEvent(dc.out, done[i], 0));
				DRC(download_rows(pl, d, pc, u.layer, u.row0, u.rows, dc.out), "download failed");
				DRC(new_event(down[i]), "event");
				DCU(cudaEventRecord(down[i], dc.out));
			}
		}
		if (defer) {
			ssc::StackParams p{};
			p.band = piece;
			p.normalize = job.normalize;
			p.coeff = pl.need_norm ? d.coeff : nullptr;
			p.rejection = job.rejection;
			p.sig_low = job.sig[0];
			p.sig_high = job.sig[1];
			p.out = d.out;
			p.out_layer_stride = (size_t) pc.rows * pl.W;
			p.out_buf_row0 = pc.row0;
			p.counters = d.counters;	// real layer numbers here: layer c's counts land in set 0, slots 2 c and 2 c + 1
			DRC(ssc::launch_deferred_handback(p, dc.comp), "hand-back launch failed");
			for (int c = 0; c < pl.C; c++) DRC(download_rows(pl, d, pc, c, pc.row0, pc.rows, dc.comp), "download failed");
		}
		dmark("units enqueued");
		if (streamed_out) {
			for (size_t i : mine) {	// every unit's rows into the caller's image as they arrive (under the upload)
				DCU(cudaEventSynchronize(down[i]));
				land(pl, d.units[i].layer, d.units[i].row0, d.units[i].rows);
			}
		}
		DCU(cudaStreamSynchronize(dc.comp));
		if (defer) for (int c = 0; c < pl.C; c++) land(pl, c, pc.row0, pc.rows);
		dmark("results landed");
		if (unit_trace) {
			for (size_t k = 0; k < tk0.size(); k++) {
				float a = 0, b = 0;
				cudaEventElapsedTime(&a, dc.t0, tk0[k]);
				cudaEventElapsedTime(&b, tk0[k], tk1[k]);
				fprintf(stderr, "[ssc trace]   device %d unit %zu (layer %d rows %d+%d): kernels start %.3f ms after the first upload began, run %.3f ms\n",
						d.id, k, d.units[mine[k]].layer, d.units[mine[k]].row0, d.units[mine[k]].rows, a, b);
				cudaEventDestroy(tk0[k]); cudaEventDestroy(tk1[k]);
			}
		}
		if (resident) {	// how long the upload took, and what was left to do when it ended
			DCU(cudaStreamSynchronize(dc.copy));
			float ms = 0.0f;
			if (!mine.empty() && cudaEventElapsedTime(&ms, dc.t0, dc.t1) == cudaSuccess) d.t_h2d += ms;
			d.t_dev += std::max(0.0, ms_since(t_up) - (double) ms);
		}
		if (*pl.cancel) return SSC_ERR_GENERIC;
	}
	return SSC_OK;
}

int run_job(const Job &job) {
	using clk = std::chrono::steady_clock;
	const auto t_begin = clk::now();
	static const bool trace = getenv("SSC_TRACE") != nullptr;	// phase timeline on stderr (debug aid)
	static clk::time_point t_last_exit = t_begin;
	struct ExitStamp { clk::time_point &t; ~ExitStamp() { t = clk::now(); } } exit_stamp{t_last_exit};
	if (trace) fprintf(stderr, "[ssc trace] %8.3f ms since the previous operator call returned\n",
			std::chrono::duration<double, std::milli>(t_begin - t_last_exit).count());
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

	std::atomic<bool> cancel{false};
	Plan pl{};
	pl.job = &job;
	pl.W = W; pl.H = H; pl.C = C; pl.N = N; pl.G = G;
	pl.need_norm = need_norm;
	pl.shifted = job.op != OP_MEDIAN && job.seq->have_reg;
	pl.plane = plane;
	pl.cancel = &cancel;
	pl.t_begin = t_begin;
	pl.trace = trace;
	{	// equally spaced host frames (one big array, as a sequence read into memory usually is): pitched copies
		pl.strided = N > 1;
		const char *f0 = (const char *) job.seq->frames[job.images[0]], *f1 = N > 1 ? (const char *) job.seq->frames[job.images[1]] : f0;
		if (!f0) return fail(SSC_ERR_GENERIC, "frame %d has no data", job.images[0]);
		const ptrdiff_t st = f1 - f0;
		for (int f = 0; f < N && pl.strided; f++) {
			const char *p = (const char *) job.seq->frames[job.images[f]];
			if (!p) return fail(SSC_ERR_GENERIC, "frame %d has no data", job.images[f]);
			if (p - f0 != st * f) pl.strided = false;
		}
		if (st <= 0) pl.strided = false;
		pl.host_stride = (size_t) st;
	}

	// --- plan: contiguous row bands per device (remainder rows to the first ones, as
	// stacking.c:620-625 does), each split further if the frames do not fit in HBM ---
	int reach = 0;
	if (pl.shifted) {
		int lo = 0, hi = 0;
		for (int idx : job.images) { lo = std::min(lo, job.seq->sy[idx]); hi = std::max(hi, job.seq->sy[idx]); }
		reach = hi - lo;
	}
	int row = 0;
	for (int g = 0; g < G; g++) {
		Dev &d = devs[g];
		d.id = ids[g];
		d.index = g;
		const int rows = H / G + (g < H % G ? 1 : 0);
		CU(cudaSetDevice(d.id));
		DevCache &dc = g_cache[d.id];
		d.dc = &dc;
		const char *budget_env = getenv("SSC_HBM_BUDGET_BYTES");
		const std::vector<long long> key = {N, C, W, H, rows, reach, (long long) job.op, (long long) job.rejection, need_norm,
				budget_env ? atoll(budget_env) : 0ll};
		if (dc.plan_key == key && !dc.plan_piece_rows.empty() && dc.frames.p) {
			for (int r : dc.plan_piece_rows) { d.pieces.push_back(Piece{row, r}); row += r; }
		} else {
			size_t free_b = 0, total_b = 0;
			CU(cudaMemGetInfo(&free_b, &total_b));
			mark("plan: cudaMemGetInfo");
			free_b += dc.frames.cap + dc.outb.cap + dc.hist.cap + dc.scratch.cap + dc.acc.cap;	// what this library already holds is reusable
			// what the library itself keeps on the device besides frames and result: histograms and IKSS scratch,
			// 4 bytes per result pixel for the hand-back lists of the register / warp kernels, and the extra
			// source rows the y shifts reach for (source_rows())
			size_t fixed = (size_t) N * (SSC_HIST_STRIDE + 65536) * 4 + (64u << 20);
			fixed += (size_t) reach * N * C * W * sizeof(WORD);
			size_t budget = free_b > fixed ? (size_t) ((free_b - fixed) * 0.85) : 0;
			if (budget_env) {	// tests force the row split with it
				const unsigned long long cap = strtoull(budget_env, nullptr, 10);
				if (cap > 0 && cap < budget) budget = (size_t) cap;
			}
			const size_t per_row = (size_t) N * C * W * sizeof(WORD) + (size_t) C * W * (sizeof(WORD) + 2 * sizeof(uint32_t));
			size_t fit_rows = per_row ? budget / per_row : 0;
			if (fit_rows < 1) return fail(SSC_ERR_GENERIC, "not enough device memory for one row of %d frames", N);
			const int n_pieces = (int) ((rows + fit_rows - 1) / fit_rows);
			dc.plan_piece_rows.clear();
			for (int k = 0; k < n_pieces; k++) {
				int r = rows / n_pieces + (k < rows % n_pieces ? 1 : 0);
				d.pieces.push_back(Piece{row, r});
				dc.plan_piece_rows.push_back(r);
				row += r;
			}
			dc.plan_key = key;
		}
		for (size_t k = 0; k < d.pieces.size(); k++) {
			const Piece &pc = d.pieces[k];
			int b0, br;
			source_rows(job, pc.row0, pc.rows, H, b0, br);
			d.max_rows = std::max(d.max_rows, pc.rows);
			d.max_buf_rows = std::max(d.max_buf_rows, br);
			// units: per layer; without registration shifts also in row chunks of ~64 MB of samples, so that the
			// upload, the kernels and the download of consecutive chunks overlap
			int chunk = pc.rows;
			if (!pl.shifted && !v.frames_on_device) {
				const size_t row_bytes = (size_t) N * W * sizeof(WORD);
				size_t chunk_bytes = 64u << 20;
				if (const char *e = getenv("SSC_CHUNK_BYTES")) {	// tests force many small chunks with it
					const unsigned long long cb = strtoull(e, nullptr, 10);
					if (cb > 0) chunk_bytes = (size_t) cb;
				}
				chunk = (int) std::max<size_t>(getenv("SSC_CHUNK_BYTES") ? 1 : 8, chunk_bytes / std::max<size_t>(row_bytes, 1));
				// (a unit should still fill the GPU: at least 256 k pixels per launch however many frames there are)
				if (!getenv("SSC_CHUNK_BYTES")) chunk = std::max(chunk, (262144 + W - 1) / W);
				const int n_chunks = std::min(64, std::max(1, (pc.rows + chunk - 1) / chunk));
				chunk = (pc.rows + n_chunks - 1) / n_chunks;
			}
			for (int r0 = 0; r0 < pc.rows; r0 += chunk)	// chunk-major: a chunk's layers are consecutive
				for (int c = 0; c < C; c++)
					d.units.push_back(Unit{(int) k, c, pc.row0 + r0, std::min(chunk, pc.rows - r0)});
		}
		for (cudaStream_t *s : {&dc.copy, &dc.comp, &dc.out})
			if (!*s) CU(cudaStreamCreateWithFlags(s, cudaStreamNonBlocking));
		for (cudaEvent_t *e : {&dc.t0, &dc.t1})
			if (!*e) CU(cudaEventCreate(e));
#define SSC_ENSURE(buf, bytes) do { if (dc.buf.ensure(bytes)) return fail(SSC_ERR_CUDA, "device allocation of %zu bytes failed", (size_t) (bytes)); } while (0)
		SSC_ENSURE(frames, (size_t) N * C * d.max_buf_rows * W * sizeof(WORD));
		SSC_ENSURE(outb, (size_t) C * d.max_rows * W * sizeof(WORD));
		SSC_ENSURE(counters, (size_t) 3 * SSC_COUNTERS * sizeof(unsigned long long));
		SSC_ENSURE(coeff, (size_t) 3 * N * sizeof(double));
		SSC_ENSURE(maxim, sizeof(unsigned long long));
		d.frames = (WORD *) dc.frames.p; d.out = (WORD *) dc.outb.p;
		d.counters = (unsigned long long *) dc.counters.p; d.coeff = (double *) dc.coeff.p;
		d.maxim = (unsigned long long *) dc.maxim.p;
		mark("plan: buffers ensured");
		CU(cudaMemsetAsync(d.counters, 0, (size_t) 3 * SSC_COUNTERS * sizeof(unsigned long long), dc.comp));
		{
			const unsigned long long neutral = job.op == OP_ADDMIN ? 0xFFFFull : 0ull;	// low word = the extreme kernels' unsigned int
			CU(cudaMemcpyAsync(d.maxim, &neutral, sizeof(neutral), cudaMemcpyHostToDevice, dc.comp));
			CU(cudaStreamSynchronize(dc.comp));
		}
		if (v.frame_format == SSC_FRAMES_RAW8 || v.frame_format == SSC_FRAMES_SER_RGB) {
			int max_src = 1;	// source rows of the largest unit
			for (const Unit &u : d.units) { int s0, sr; source_rows(job, u.row0, u.rows, H, s0, sr); max_src = std::max(max_src, sr); }
			for (const Piece &pc : d.pieces) { int s0, sr; source_rows(job, pc.row0, pc.rows, H, s0, sr); if (d.pieces.size() > 1 || job.want_coeff_only) max_src = std::max(max_src, sr); }
			const size_t per_sample = v.frame_format == SSC_FRAMES_SER_RGB ? 3 * (size_t) v.ser_bytes_per_sample : 1;
			SSC_ENSURE(staging, (size_t) N * max_src * W * per_sample);
			d.staging = (unsigned char *) dc.staging.p;
		}
		if (v.pp_offset || v.pp_dark || v.pp_flat) {
			SSC_ENSURE(masters, (size_t) 3 * C * d.max_buf_rows * W * sizeof(WORD));
			d.masters = (WORD *) dc.masters.p;
		}
		mark("plan: neutral element set");
		if (job.op == OP_SUM) {
			SSC_ENSURE(acc, (size_t) C * d.max_rows * W * sizeof(uint32_t));
			d.acc = (uint32_t *) dc.acc.p;
		}
		if (need_norm) {
			SSC_ENSURE(flags, (size_t) N * sizeof(uint32_t));
			SSC_ENSURE(status, sizeof(int));
			d.flags = (uint32_t *) dc.flags.p; d.status = (int *) dc.status.p;
			CU(cudaMemsetAsync(d.flags, 0, (size_t) N * sizeof(uint32_t), dc.comp));
			CU(cudaMemsetAsync(d.status, 0, sizeof(int), dc.comp));
			if (G == 1) {
				SSC_ENSURE(hist, (size_t) N * SSC_HIST_STRIDE * sizeof(uint32_t));
				SSC_ENSURE(scratch, (size_t) N * 65536 * sizeof(uint32_t));
				SSC_ENSURE(stats, (size_t) 4 * N * sizeof(double));
				d.hist = (uint32_t *) dc.hist.p; d.scratch = (uint32_t *) dc.scratch.p; d.stats = (double *) dc.stats.p;
				CU(cudaMemsetAsync(d.hist, 0, (size_t) N * SSC_HIST_STRIDE * sizeof(uint32_t), dc.comp));
			} else {
				SSC_ENSURE(scratch, (size_t) ((N + G - 1) / G) * 65536 * sizeof(uint32_t));
				d.scratch = (uint32_t *) dc.scratch.p;
			}
		}
		if (pl.shifted) {
			std::vector<int> sh(2 * N);
			for (int f = 0; f < N; f++) {
				sh[f] = job.seq->sx[job.images[f]];
				sh[N + f] = job.seq->sy[job.images[f]];
			}
			SSC_ENSURE(shifts, 2 * N * sizeof(int));
			d.shifts = (int *) dc.shifts.p;
			CU(cudaMemcpyAsync(d.shifts, sh.data(), 2 * N * sizeof(int), cudaMemcpyHostToDevice, dc.comp));
			CU(cudaStreamSynchronize(dc.comp));	// sh goes out of scope
		}
	}
	// several devices: every device's end of the NVLink exchanges (kept between calls of the same shape)
	if (G > 1) {
		bool rebuild = false;
		for (auto &d : devs) rebuild = rebuild || !d.dc->group || d.dc->group_ids != ids || d.dc->group_frames < N;
		if (rebuild) {
			for (auto &kv : g_cache)
				if (kv.second.group) { cudaSetDevice(kv.first); ssc_peer_destroy(kv.second.group); kv.second.group = nullptr; }
			for (auto &d : devs) {
				CU(cudaSetDevice(d.id));
				if (ssc_peer_create(G, d.index, N, &d.dc->group)) return fail(SSC_ERR_CUDA, "peer group: %s", ssc_peer_last_error());
				d.dc->group_ids = ids;
				d.dc->group_frames = N;
			}
			for (auto &d : devs) {
				for (auto &o : devs)
					if (ssc_peer_attach_local(d.dc->group, o.index, o.dc->group)) return fail(SSC_ERR_CUDA, "peer group: %s", ssc_peer_last_error());
				if (ssc_peer_connect(d.dc->group)) return fail(SSC_ERR_CUDA, "peer group: %s", ssc_peer_last_error());
			}
		}
		for (auto &d : devs) d.group = d.dc->group;
	}
	mark("planned, workspace ready");

	// The result (malloc()ed: it becomes gfit.data, stacking.c:776-785) is allocated now and its pages
	// are touched on a helper thread while the frames upload: first-touch faults on 100+ MB would
	// otherwise sit between the download and the caller.
	WORD *result = job.want_coeff_only ? nullptr : (WORD *) malloc(plane * C * sizeof(WORD));
	struct ResultGuard { WORD *&p; ~ResultGuard() { free(p); } } rguard{result};
	if (!job.want_coeff_only && !result) return fail(SSC_ERR_GENERIC, "result allocation failed");
	if (result) {
		WORD *r = result;
		const size_t bytes = plane * C * sizeof(WORD);
		pl.result_ready = std::async(std::launch::async, [r, bytes] { memset(r, 0, bytes); }).share();
		// device -> pinned bounce buffer (full PCIe speed), then into the malloc()ed result
		if (g_pinned_cap < bytes) {
			if (g_pinned) cudaFreeHost(g_pinned);
			g_pinned = nullptr; g_pinned_cap = 0;
			if (cudaHostAlloc(&g_pinned, bytes, cudaHostAllocPortable) == cudaSuccess) g_pinned_cap = bytes;
		}
	}
	struct ReadyGuard { std::shared_future<void> &f; ~ReadyGuard() { if (f.valid()) f.wait(); } } ready_guard{pl.result_ready};
	pl.result = result;
	pl.landing = g_pinned ? (WORD *) g_pinned : result;
	std::vector<double> h_coeff((size_t) 3 * N);
	for (int i = 0; i < N; i++) { h_coeff[i] = 0.0; h_coeff[N + i] = 1.0; h_coeff[2 * N + i] = 1.0; }

	if (need_norm) {
		log_msg("Computing normalization...\n");
		progress("Computing normalization...", 0.0);
	}
	if (job.want_coeff_only)	// ssc_compute_normalization: K4 over every piece, K5, K6 -- no unit is ever launched
		for (auto &d : devs) d.units.clear();

	// --- every device on its own host thread; the caller's cancel hook is polled here, on the calling thread ---
	auto run_phase = [&](int phase) -> int {
		std::atomic<int> running{G};
		std::vector<std::thread> th;
		for (int g = 0; g < G; g++)
			th.emplace_back([&, g] {
				devs[g].rc = device_run(pl, devs[g], phase);
				running--;
			});
		while (running.load() > 0) {
			std::this_thread::sleep_for(std::chrono::milliseconds(2));
			if (!keep_running()) cancel = true;	// the device threads stop at their next piece
		}
		for (auto &t : th) t.join();
		int bad = SSC_OK;
		for (auto &d : devs)
			if (d.rc && !bad) bad = fail(d.rc, "device %d: %s", d.id, d.err.empty() ? (cancel ? "cancelled" : "failed") : d.err.c_str());
		if (bad)	// the devices' exchange epochs may have drifted apart: new peer groups next time
			for (auto &d : devs)
				if (d.dc->group) { cudaSetDevice(d.id); cudaDeviceSynchronize(); ssc_peer_destroy(d.dc->group); d.dc->group = nullptr; }
		return bad;
	};
	if (!job.want_coeff_only) log_msg("Starting stacking...\n");
	int rc = run_phase(0);
	if (rc) return rc;
	mark("phase 0 done");

	if (need_norm) {
		Dev &d0 = devs[0];
		int h_status = 0;
		CU(cudaSetDevice(d0.id));
		CU(cudaMemcpy(h_coeff.data(), d0.coeff, h_coeff.size() * sizeof(double), cudaMemcpyDeviceToHost));
		for (auto &d : devs) {
			int s = 0;
			CU(cudaSetDevice(d.id));
			CU(cudaMemcpy(&s, d.status, sizeof(int), cudaMemcpyDeviceToHost));
			h_status |= s;
			if (d.group && ssc_peer_status(d.group, d.dc->comp)) return fail(SSC_ERR_CUDA, "device %d waited for a peer that never arrived", d.id);
		}
		if (h_status) {
			log_msg("Error: no data computed.\n");
			return fail(SSC_ERR_GENERIC, "a frame has no non-zero pixel: no statistics");
		}
		progress(nullptr, 1.0);
	}
	if (job.coeff_out) memcpy(job.coeff_out, h_coeff.data(), h_coeff.size() * sizeof(double));
	if (job.want_coeff_only) return SSC_OK;
	if (!keep_running()) return fail(SSC_ERR_GENERIC, "cancelled");

	unsigned long long h_max = 0;
	if (job.op == OP_SUM) {
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
		if ((rc = run_phase(1))) return rc;
	}
	{
		size_t max_pieces = 0;
		for (auto &d : devs) max_pieces = std::max(max_pieces, d.pieces.size());
		for (size_t k = 0; k < max_pieces; k++) progress(nullptr, (double) (k + 1) / (double) max_pieces);
	}
	if (!keep_running()) return fail(SSC_ERR_GENERIC, "cancelled");

	mark("stack + download done");
	ssc_result res;
	memset(&res, 0, sizeof(res));
	for (auto &d : devs) {
		unsigned long long h_cnt[3 * SSC_COUNTERS];
		CU(cudaSetDevice(d.id));
		CU(cudaMemcpy(h_cnt, d.counters, sizeof(h_cnt), cudaMemcpyDeviceToHost));
		if (job.op == OP_MEAN) {
			for (int c = 0; c < C; c++) {	// one counter set per layer (every unit's kernel saw a single layer);
				// a deferred hand-back pass knows the real layers and counts layer c in set 0, slots 2 c and 2 c + 1
				res.irej[c][0] += h_cnt[c * SSC_COUNTERS + 0] + (c > 0 ? h_cnt[2 * c + 0] : 0);
				res.irej[c][1] += h_cnt[c * SSC_COUNTERS + 1] + (c > 0 ? h_cnt[2 * c + 1] : 0);
				res.guard_fired += h_cnt[c * SSC_COUNTERS + 6];
				res.cap_hit += h_cnt[c * SSC_COUNTERS + 7];
			}
		}
		res.t_h2d_ms = std::max(res.t_h2d_ms, d.t_h2d);
		res.t_device_ms = std::max(res.t_device_ms, d.t_dev);
		res.t_d2h_ms = std::max(res.t_d2h_ms, d.t_d2h);
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
