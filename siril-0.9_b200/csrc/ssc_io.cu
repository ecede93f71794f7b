// ssc_io.cu -- the data formats either side of the stacking path: the frames
// arrive as they lie in the files and are decoded on the GPU behind the copy, the result leaves as a
// FITS file.
//
//   FITS (src/io/image_format_fits.c): Siril writes and stacks BITPIX 16 images with BZERO 32768 --
//       big-endian signed samples, physical value = sample + 32768 -- or BITPIX 8; planes and rows in
//       the order of fits.data (rows bottom-up: readfits / read_opened_fits_partial hand the stackers
//       the very row order of the file, image_format_fits.c:599-632).  Decoding is a byte swap and a
//       sign flip per sample, in place on the uploaded rows.
//   SER (src/io/ser.c): a 178-byte header, then frames of 8- or 16-bit samples, mono or three
//       interleaved colours (RGB / BGR); 8-bit samples are widened, big-endian ones swapped
//       (ser_manage_endianess_and_depth, ser.c:476-491 -- the header's endianness flag is used with
//       the opposite meaning of the specification, ser.h:31-42), interleaved colours go to planes
//       (ser.c:953-964).  CFA frames are read as mono (ser.c:789-791 without the debayer box).
//   preprocess (src/core/siril.c:945-962): offset and dark subtraction saturate at 0 (imoper OPER_SUB
//       through round_to_WORD, :169-173), the flat division is coef * (g / flat) in double with
//       flat 0 -> 1 (fdiv, :252-275); applied to the decoded rows before anything else sees them.
//   savefits (image_format_fits.c:652-739, save_fits_header :741-): header cards as cfitsio writes
//       them, the samples encoded on the GPU.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "ssc_device.cuh"
#include "ssc_kernels.h"

namespace ssc {

__device__ __forceinline__ unsigned int fits16_pair(unsigned int w) {	// two samples: swap bytes, flip the sign bit
	return __byte_perm(w, 0, 0x2301) ^ 0x80008000u;
}

// rows of every frame in place: base + f * frame_stride, `count` samples each (count even, base 4-byte aligned)
__global__ void k_fits16_inplace(WORD *base, size_t frame_stride, size_t count, int big_endian_signed) {
	unsigned int *p = reinterpret_cast<unsigned int *>(base + (size_t) blockIdx.y * frame_stride);
	const size_t n2 = count >> 1;
	for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t) gridDim.x * blockDim.x) {
		const unsigned int w = p[i];
		p[i] = big_endian_signed == 2 ? __byte_perm(w, 0, 0x2301) : fits16_pair(w);	// 2: plain byte swap (SER big endian)
	}
	if ((count & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
		WORD *q = base + (size_t) blockIdx.y * frame_stride + count - 1;
		const unsigned int v = *q, s = ((v >> 8) | (v << 8)) & 0xFFFFu;
		*q = (WORD) (big_endian_signed == 2 ? s : s ^ 0x8000u);
	}
}

// raw rows from the staging buffer into planes: bytes per sample 1 or 2, 1 or 3 interleaved colours
// raw: [frames][rows][width][colours] samples; dst plane c of frame f at dst + f*frame_stride + c*layer_stride
__global__ void k_decode_raw(const unsigned char *__restrict__ raw, size_t raw_frame_bytes, WORD *__restrict__ dst,
		size_t frame_stride, size_t layer_stride, size_t count, int bytes, int colours, int big_endian, int bgr,
		int signed_bzero) {
	const unsigned char *src = raw + (size_t) blockIdx.y * raw_frame_bytes;
	WORD *out = dst + (size_t) blockIdx.y * frame_stride;
	for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < count * colours; i += (size_t) gridDim.x * blockDim.x) {
		unsigned int v;
		if (bytes == 1) v = src[i];
		else {
			const unsigned int lo = src[2 * i], hi = src[2 * i + 1];
			v = big_endian ? (lo << 8) | hi : (hi << 8) | lo;
			if (signed_bzero) v ^= 0x8000u;
		}
		const size_t px = i / colours;
		int c = (int) (i - px * colours);
		if (bgr) c = 2 - c;						// ser.c:954-957
		out[(size_t) c * layer_stride + px] = (WORD) v;
	}
}

__global__ void k_encode_fits16(const WORD *__restrict__ src, WORD *__restrict__ dst, size_t count) {
	for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t) gridDim.x * blockDim.x) {
		const unsigned int v = src[i] ^ 0x8000u;
		dst[i] = (WORD) (((v >> 8) | (v << 8)) & 0xFFFFu);
	}
}

// preprocess (siril.c:945-962) on rows of every frame: masters hold the same rows, [count] samples each
__global__ void k_preprocess(WORD *base, size_t frame_stride, size_t count, const WORD *__restrict__ offset,
		const WORD *__restrict__ dark, const WORD *__restrict__ flat, double level) {
	WORD *p = base + (size_t) blockIdx.y * frame_stride;
	for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t) gridDim.x * blockDim.x) {
		int g = p[i];
		if (offset) g = max(g - (int) offset[i], 0);		// round_to_WORD(g - b) of two WORDs (:169-173)
		if (dark) g = max(g - (int) dark[i], 0);
		if (flat) {						// fdiv, :262-273
			unsigned int f = flat[i];
			if (f == 0) f = 1;
			g = (int) round_to_word(dmul(level, ddiv((double) g, (double) (int) f)));
		}
		p[i] = (WORD) g;
	}
}

static unsigned int grid_for(size_t count) {
	const size_t want = (count + 255) / 256;
	return (unsigned int) (want < 148 * 8 ? (want ? want : 1) : 148 * 8);
}

int launch_decode_inplace(WORD *base, size_t frame_stride, int n_frames, size_t count, int mode, cudaStream_t st) {
	if (count == 0) return SSC_OK;
	dim3 grid(grid_for(count >> 1), n_frames);
	k_fits16_inplace<<<grid, 256, 0, st>>>(base, frame_stride, count, mode);
	return cudaGetLastError() == cudaSuccess ? SSC_OK : SSC_ERR_CUDA;
}

int launch_decode_raw(const void *raw, size_t raw_frame_bytes, WORD *dst, size_t frame_stride, size_t layer_stride,
		int n_frames, size_t count, int bytes, int colours, int big_endian, int bgr, int signed_bzero, cudaStream_t st) {
	if (count == 0) return SSC_OK;
	dim3 grid(grid_for(count * colours), n_frames);
	k_decode_raw<<<grid, 256, 0, st>>>((const unsigned char *) raw, raw_frame_bytes, dst, frame_stride, layer_stride, count,
			bytes, colours, big_endian, bgr, signed_bzero);
	return cudaGetLastError() == cudaSuccess ? SSC_OK : SSC_ERR_CUDA;
}

int launch_preprocess(WORD *base, size_t frame_stride, int n_frames, size_t count, const WORD *offset, const WORD *dark,
		const WORD *flat, float level, cudaStream_t st) {
	if (count == 0 || (!offset && !dark && !flat)) return SSC_OK;
	dim3 grid(grid_for(count), n_frames);
	k_preprocess<<<grid, 256, 0, st>>>(base, frame_stride, count, offset, dark, flat, (double) level);
	return cudaGetLastError() == cudaSuccess ? SSC_OK : SSC_ERR_CUDA;
}

}  // namespace ssc

using namespace ssc;

namespace {
thread_local char g_io_err[256];
int io_fail(int code, const char *fmt, const char *a = "") {
	snprintf(g_io_err, sizeof(g_io_err), fmt, a);
	return code;
}

// one 80-character header card: "KEYWORD = value / comment"
bool card_value(const char *card, const char *key, std::string &value) {
	char k[9];
	memcpy(k, card, 8);
	k[8] = 0;
	for (int i = 7; i >= 0 && k[i] == ' '; i--) k[i] = 0;
	if (strcmp(k, key) != 0 || card[8] != '=') return false;
	std::string v(card + 10, 70);
	if (!v.empty() && v[v.find_first_not_of(' ') == std::string::npos ? 0 : v.find_first_not_of(' ')] == '\'') {
		const size_t a = v.find('\''), b = v.find('\'', a + 1);
		value = v.substr(a + 1, b == std::string::npos ? std::string::npos : b - a - 1);
	} else {
		const size_t slash = v.find('/');
		if (slash != std::string::npos) v = v.substr(0, slash);
		const size_t a = v.find_first_not_of(' '), b = v.find_last_not_of(' ');
		value = a == std::string::npos ? "" : v.substr(a, b - a + 1);
	}
	return true;
}

void put_card(std::string &h, const char *key, const char *value, const char *comment) {
	char card[81];
	snprintf(card, sizeof(card), "%-8s= %20s / %-47s", key, value, comment);
	card[80] = 0;
	std::string c(card);
	c.resize(80, ' ');
	h += c;
}
}  // namespace

extern "C" {

const char *ssc_io_last_error(void) { return g_io_err; }

// readfits' view of a file (image_format_fits.c:40-130): geometry, sample type, where the data unit starts
int ssc_fits_parse_header(const void *bytes, size_t len, ssc_fits_info *info) {
	if (!bytes || !info) return io_fail(SSC_ERR_ARG, "null argument");
	memset(info, 0, sizeof(*info));
	info->bscale = 1.0;
	info->naxes[0] = info->naxes[1] = info->naxes[2] = 1;
	const char *p = (const char *) bytes;
	bool end = false;
	size_t off = 0;
	for (; off + 80 <= len && !end; off += 80) {
		const char *card = p + off;
		std::string v;
		if (!strncmp(card, "END     ", 8)) { end = true; break; }
		if (card_value(card, "BITPIX", v)) info->bitpix = atoi(v.c_str());
		else if (card_value(card, "NAXIS", v)) info->naxis = atoi(v.c_str());
		else if (card_value(card, "NAXIS1", v)) info->naxes[0] = atol(v.c_str());
		else if (card_value(card, "NAXIS2", v)) info->naxes[1] = atol(v.c_str());
		else if (card_value(card, "NAXIS3", v)) info->naxes[2] = atol(v.c_str());
		else if (card_value(card, "BZERO", v)) info->bzero = atof(v.c_str());
		else if (card_value(card, "BSCALE", v)) info->bscale = atof(v.c_str());
		else if (card_value(card, "EXPTIME", v) || card_value(card, "EXPOSURE", v)) info->exposure = atof(v.c_str());
		else if (card_value(card, "MIPS-HI", v)) info->hi = (WORD) atoi(v.c_str());
		else if (card_value(card, "MIPS-LO", v)) info->lo = (WORD) atoi(v.c_str());
	}
	if (!end) return io_fail(SSC_ERR_READ, "no END card in the FITS header");
	info->data_offset = ((off + 80 + 2879) / 2880) * 2880;
	if (info->naxis < 2 || info->naxis > 3 || info->naxes[0] < 1 || info->naxes[1] < 1) return io_fail(SSC_ERR_SIZE, "not a 2D / 3D image");
	if (info->naxis == 3 && info->naxes[2] != 1 && info->naxes[2] != 3) return io_fail(SSC_ERR_SIZE, "1 or 3 layers expected");	// :60-66
	if (info->naxis == 2) info->naxes[2] = 1;
	if (info->bitpix == 16 && info->bzero == 32768.0 && info->bscale == 1.0) info->frame_format = SSC_FRAMES_FITS16;
	else if (info->bitpix == 8 && info->bzero == 0.0 && info->bscale == 1.0) info->frame_format = SSC_FRAMES_RAW8;
	else return io_fail(SSC_ERR_GENERIC, "only BITPIX 16 with BZERO 32768 and BITPIX 8 are decoded on the device");
	info->data_bytes = (size_t) info->naxes[0] * info->naxes[1] * info->naxes[2] * (info->bitpix / 8);
	if (info->data_offset + info->data_bytes > len && len > 0 && len != (size_t) -1) return io_fail(SSC_ERR_READ, "FITS file shorter than its data unit");
	return SSC_OK;
}

// ser_read_header (ser.c:291-350): the integers of the header are little endian
int ssc_ser_parse_header(const void *bytes, size_t file_size, ssc_ser_info *info) {
	if (!bytes || !info) return io_fail(SSC_ERR_ARG, "null argument");
	if (file_size < 178) return io_fail(SSC_ERR_READ, "SER file shorter than its header");
	const unsigned char *h = (const unsigned char *) bytes;
	auto le32 = [&](int o) { return (int) (h[o] | (h[o + 1] << 8) | (h[o + 2] << 16) | ((unsigned int) h[o + 3] << 24)); };
	memset(info, 0, sizeof(*info));
	info->color_id = le32(18);
	info->big_endian = le32(22) != 0;		// SER_BIG_ENDIAN = 1 (ser.h:31-42: the flag's inverted use)
	info->width = le32(26);
	info->height = le32(30);
	info->bit_depth = le32(34);
	info->frame_count = le32(38);
	info->bytes_per_sample = info->bit_depth <= 8 ? 1 : 2;			// :327-329
	info->planes = (info->color_id == 100 || info->color_id == 101) ? 3 : 1;	// :331-334
	if (info->width < 1 || info->height < 1) return io_fail(SSC_ERR_SIZE, "bad SER geometry");
	info->frame_bytes = (size_t) info->width * info->height * info->planes * info->bytes_per_sample;
	info->data_offset = 178;
	if (info->frame_count == 0)		// ser_fix_broken_file: the count that fits the file
		info->frame_count = (int) ((file_size - 178) / info->frame_bytes);
	if (178 + info->frame_bytes * (size_t) info->frame_count > file_size) return io_fail(SSC_ERR_READ, "SER file shorter than its frames");
	info->frame_format = info->planes == 3 ? SSC_FRAMES_SER_RGB : (info->bytes_per_sample == 1 ? SSC_FRAMES_RAW8
			: (info->big_endian ? SSC_FRAMES_RAW16_BE : SSC_FRAMES_U16));
	return SSC_OK;
}

// savefits (image_format_fits.c:652-739) for the USHORT_IMG images the stackers produce: BITPIX 16, BZERO 32768,
// the cards of save_fits_header that a stacking result carries.  data: host, planar [layers][ry][rx].
int ssc_savefits(const char *path, const WORD *data, int rx, int ry, int layers, WORD hi, WORD lo, double exposure,
		int overwrite) {
	if (!path || !data || rx < 1 || ry < 1 || (layers != 1 && layers != 3)) return io_fail(SSC_ERR_ARG, "bad image");
	int n_dev = 0;
	if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev < 1) return io_fail(SSC_ERR_CUDA, "no CUDA device available");
	if (!overwrite) {
		FILE *t = fopen(path, "rb");
		if (t) { fclose(t); return io_fail(SSC_ERR_GENERIC, "%s exists", path); }
	}
	std::string h;
	char num[32];
	put_card(h, "SIMPLE", "T", "file does conform to FITS standard");
	put_card(h, "BITPIX", "16", "number of bits per data pixel");
	snprintf(num, sizeof(num), "%d", layers == 3 ? 3 : 2);
	put_card(h, "NAXIS", num, "number of data axes");
	snprintf(num, sizeof(num), "%d", rx);
	put_card(h, "NAXIS1", num, "length of data axis 1");
	snprintf(num, sizeof(num), "%d", ry);
	put_card(h, "NAXIS2", num, "length of data axis 2");
	if (layers == 3) put_card(h, "NAXIS3", "3", "length of data axis 3");
	put_card(h, "EXTEND", "T", "FITS dataset may contain extensions");
	put_card(h, "BZERO", "32768", "offset data range to that of unsigned short");
	put_card(h, "BSCALE", "1", "default scaling factor");
	if (hi) {	// "may not be initialized" (:747)
		snprintf(num, sizeof(num), "%u", (unsigned) hi);
		put_card(h, "MIPS-HI", num, "Upper visualization cutoff");
		snprintf(num, sizeof(num), "%u", (unsigned) lo);
		put_card(h, "MIPS-LO", num, "Lower visualization cutoff");
	}
	if (exposure > 0.0) {
		snprintf(num, sizeof(num), "%.15G", exposure);
		put_card(h, "EXPTIME", num, "Exposure time [s]");
	}
	put_card(h, "PROGRAM", "'libsiril_stack_cuda'", "Software that created this HDU");
	{ std::string e("END"); e.resize(80, ' '); h += e; }
	h.resize((h.size() + 2879) / 2880 * 2880, ' ');

	const size_t count = (size_t) rx * ry * layers, bytes = count * sizeof(WORD);
	WORD *d_src = nullptr, *d_dst = nullptr;
	std::vector<WORD> enc(count);
	cudaError_t e = cudaMalloc(&d_src, bytes);
	if (e == cudaSuccess) e = cudaMalloc(&d_dst, bytes);
	if (e == cudaSuccess) e = cudaMemcpy(d_src, data, bytes, cudaMemcpyHostToDevice);
	if (e == cudaSuccess) {
		k_encode_fits16<<<grid_for(count), 256>>>(d_src, d_dst, count);
		e = cudaGetLastError();
	}
	if (e == cudaSuccess) e = cudaMemcpy(enc.data(), d_dst, bytes, cudaMemcpyDeviceToHost);
	cudaFree(d_src);
	cudaFree(d_dst);
	if (e != cudaSuccess) return io_fail(SSC_ERR_CUDA, "encoding failed: %s", cudaGetErrorString(e));

	remove(path);	// "Delete old file if it already exists" (:671)
	FILE *f = fopen(path, "wb");
	if (!f) return io_fail(SSC_ERR_GENERIC, "cannot create %s", path);
	bool ok = fwrite(h.data(), 1, h.size(), f) == h.size() && fwrite(enc.data(), 1, bytes, f) == bytes;
	const size_t pad = (2880 - bytes % 2880) % 2880;
	if (ok && pad) { std::vector<char> z(pad, 0); ok = fwrite(z.data(), 1, pad, f) == pad; }
	ok = fclose(f) == 0 && ok;
	return ok ? SSC_OK : io_fail(SSC_ERR_GENERIC, "short write to %s", path);
}

// the device-layer forms: decode rows already resident on the device, in place / from a raw copy
int ssc_dev_decode_fits16(WORD *samples, size_t count, void *stream) {
	return launch_decode_inplace(samples, 0, 1, count, 1, (cudaStream_t) stream) ? io_fail(SSC_ERR_CUDA, "decode launch failed") : SSC_OK;
}

int ssc_dev_decode_ser(const void *raw, WORD *planes, size_t layer_stride, size_t pixels, int bytes_per_sample, int big_endian,
		int color_id, void *stream) {
	const int colours = (color_id == 100 || color_id == 101) ? 3 : 1;
	return launch_decode_raw(raw, 0, planes, 0, layer_stride, 1, pixels, bytes_per_sample, colours, big_endian, color_id == 101, 0,
			(cudaStream_t) stream) ? io_fail(SSC_ERR_CUDA, "decode launch failed") : SSC_OK;
}

int ssc_dev_preprocess(WORD *samples, size_t count, const WORD *offset, const WORD *dark, const WORD *flat, float level, void *stream) {
	return launch_preprocess(samples, 0, 1, count, offset, dark, flat, level, (cudaStream_t) stream) ? io_fail(SSC_ERR_CUDA, "preprocess launch failed") : SSC_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// After the stack: statistics(fit, layer, NULL, STATS_BASIC, STATS_ZERO_NULLCHECK)
// (src/algos/statistics.c:207-326), which noise() (src/core/siril.c:1740-1771) and the stacking
// summary run on the result.  Everything but the noise estimate follows from the layer's exact histogram
// (K4 again): the count of non-zero pixels, their mean and sigma from the integer sums of k h[k] and
// k^2 h[k], min, max, and siril_stats_ushort_median (statistics.c:47-64) -- the first bin where the
// running count exceeds n / 2.  The noise estimate is cfitsio's (fits_img_stats_ushort, "noise1": per
// row, the 5-sigma clipped standard deviation of the differences of consecutive non-null pixels, three
// clipping passes; the median over the rows, divided by sqrt 2).  cfitsio is not part of
// /root/reference: that estimator is restated from its published source, "parity unpinned".
// ---------------------------------------------------------------------------------------------
namespace ssc {

__global__ void __launch_bounds__(1024)
k_basic_stats(const uint32_t *__restrict__ hist, int nullcheck, double total, ssc_imstats *__restrict__ out) {
	const uint32_t *h = hist + (size_t) blockIdx.x * SSC_HIST_STRIDE;
	__shared__ unsigned long long s_n[32], s_s[32], s_q[32];
	__shared__ unsigned int s_min, s_max, s_wsum[32], s_median;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if (tid == 0) { s_min = 0xFFFFFFFFu; s_max = 0; s_median = 0; }
	__syncthreads();
	// thread t owns bins [64 t, 64 t + 64)
	unsigned long long n = 0, s = 0, q = 0;
	unsigned int lo = 0xFFFFFFFFu, hi = 0;
	for (int k = 0; k < 64; k++) {
		const unsigned int bin = (unsigned int) tid * 64 + k;
		const unsigned long long c = (nullcheck && bin == 0) ? 0ull : h[bin];
		if (c) { n += c; s += c * bin; q += c * bin * bin; lo = min(lo, bin); hi = max(hi, bin); }
	}
	const unsigned int mine = (unsigned int) n;	// (a layer holds fewer than 2^32 pixels)
	for (int o = 16; o > 0; o >>= 1) {
		n += __shfl_xor_sync(0xffffffffu, n, o);
		s += __shfl_xor_sync(0xffffffffu, s, o);
		q += __shfl_xor_sync(0xffffffffu, q, o);
		lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
		hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
	}
	if (lane == 0) { s_n[warp] = n; s_s[warp] = s; s_q[warp] = q; atomicMin(&s_min, lo); atomicMax(&s_max, hi); }
	// exclusive prefix of the per-thread counts: where does the running count pass n / 2 ?
	unsigned int inc = mine;
	for (int o = 1; o < 32; o <<= 1) { const unsigned int u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
	if (lane == 31) s_wsum[warp] = inc;
	__syncthreads();
	unsigned long long N = 0, S = 0, Q = 0;
	unsigned int before = 0;
	for (int w = 0; w < 32; w++) { N += s_n[w]; S += s_s[w]; Q += s_q[w]; if (w < warp) before += s_wsum[w]; }
	const unsigned int excl = before + inc - mine;
	const double half = dmul((double) (long long) N, 0.5);		// (double) n * 0.5, statistics.c:58
	if ((double) excl <= half && (double) (excl + mine) > half) {	// the crossing is inside my bins
		unsigned int run = excl;
		for (int k = 0; k < 64; k++) {
			const unsigned int bin = (unsigned int) tid * 64 + k;
			run += (nullcheck && bin == 0) ? 0u : h[bin];
			if ((double) run > half) { s_median = bin; break; }
		}
	}
	__syncthreads();
	if (tid == 0) {
		ssc_imstats r;
		r.total = total;
		r.ngoodpix = (double) (long long) N;
		const double nn = (double) (long long) N;
		const double mean = N ? ddiv((double) (long long) S, nn) : 0.0;
		r.mean = mean;
		r.sigma = N > 1 ? dsqrt(dsub(ddiv(__ull2double_rn(Q), nn), dmul(mean, mean))) : 0.0;	// FnMeanSigma
		r.median = (double) s_median;
		r.min = N ? (double) s_min : 0.0;
		r.max = (double) s_max;
		r.bgnoise = 0.0;
		r.normValue = 0.0;
		out[blockIdx.x] = r;
	}
}

// one block per row: the 5-sigma clipped standard deviation of the first differences (FnNoise1)
constexpr int NOISE_THREADS = 256, NOISE_PER_THREAD = 64;	// rows up to 16384 pixels
__global__ void __launch_bounds__(NOISE_THREADS)
k_row_noise(const WORD *__restrict__ img, int width, int nullcheck, double *__restrict__ row_sd) {
	const WORD *row = img + (size_t) blockIdx.x * width;
	const int tid = threadIdx.x;
	const int per = (width + NOISE_THREADS - 1) / NOISE_THREADS;	// consecutive pixels per thread
	const int x0 = tid * per, x1 = min(width, x0 + per);
	__shared__ int s_last[NOISE_THREADS];		// my last valid pixel's value, -1: none
	__shared__ double s_a[NOISE_THREADS], s_b[NOISE_THREADS];
	__shared__ int s_c[NOISE_THREADS];
	int last = -1;
	for (int x = x0; x < x1; x++) { const int v = row[x]; if (!nullcheck || v != 0) last = v; }
	s_last[tid] = last;
	__syncthreads();
	int prev = -1;					// the valid pixel before my first one
	for (int t = tid - 1; t >= 0 && prev < 0; t--) prev = s_last[t];
	float d[NOISE_PER_THREAD];
	unsigned long long alive = 0;
	int nd = 0;
	for (int x = x0; x < x1 && nd < NOISE_PER_THREAD; x++) {
		const int v = row[x];
		if (nullcheck && v == 0) continue;
		if (prev >= 0) { d[nd] = (float) (prev - v); alive |= 1ull << nd; nd++; }
		prev = v;
	}
	auto reduce = [&](double &sum, double &sum2, int &cnt) {
		double a = 0.0, b = 0.0;
		int c = 0;
#pragma unroll 8
		for (int k = 0; k < NOISE_PER_THREAD; k++)
			if (k < nd && (alive >> k & 1ull)) { const double v = (double) d[k]; a += v; b += v * v; c++; }
		s_a[tid] = a; s_b[tid] = b; s_c[tid] = c;
		__syncthreads();
		for (int o = NOISE_THREADS / 2; o > 0; o >>= 1) {
			if (tid < o) { s_a[tid] += s_a[tid + o]; s_b[tid] += s_b[tid + o]; s_c[tid] += s_c[tid + o]; }
			__syncthreads();
		}
		sum = s_a[0]; sum2 = s_b[0]; cnt = s_c[0];
		__syncthreads();
	};
	double sum, sum2;
	int cnt;
	reduce(sum, sum2, cnt);
	double sd = -1.0;				// rows with fewer than two differences do not count
	if (cnt >= 2) {
		double mean = sum / cnt;
		sd = sqrt(fmax(sum2 / cnt - mean * mean, 0.0));
		if (sd > 0.0)
			for (int iter = 0; iter < 3; iter++) {
				const int before = cnt;
				for (int k = 0; k < nd; k++)
					if ((alive >> k & 1ull) && !(fabs((double) d[k] - mean) < 5.0 * sd)) alive &= ~(1ull << k);
				reduce(sum, sum2, cnt);
				if (cnt == before) break;
				if (cnt > 1) { mean = sum / cnt; sd = sqrt(fmax(sum2 / cnt - mean * mean, 0.0)); }
				else if (cnt == 1) { mean = sum; sd = 0.0; }
				else { mean = 0.0; sd = 0.0; }
			}
	}
	if (tid == 0) row_sd[blockIdx.x] = sd;
}

}  // namespace ssc

extern "C" int ssc_image_statistics(const WORD *data, int rx, int ry, int layers, int nullcheck, ssc_imstats *stats) {
	if (!data || !stats || rx < 1 || ry < 1 || layers < 1 || layers > 3) return io_fail(SSC_ERR_ARG, "bad image");
	if (rx > NOISE_THREADS * NOISE_PER_THREAD) return io_fail(SSC_ERR_ARG, "rows longer than 16384 pixels are not supported");
	int n_dev = 0;
	if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev < 1) return io_fail(SSC_ERR_CUDA, "no CUDA device available");
	const size_t plane = (size_t) rx * ry;
	WORD *d_img = nullptr;
	uint32_t *d_hist = nullptr;
	ssc_imstats *d_stats = nullptr;
	double *d_rows = nullptr;
	std::vector<double> rows((size_t) layers * ry);
	cudaError_t e = cudaMalloc(&d_img, plane * layers * sizeof(WORD));
	if (e == cudaSuccess) e = cudaMalloc(&d_hist, (size_t) layers * SSC_HIST_STRIDE * sizeof(uint32_t));
	if (e == cudaSuccess) e = cudaMalloc(&d_stats, layers * sizeof(ssc_imstats));
	if (e == cudaSuccess) e = cudaMalloc(&d_rows, rows.size() * sizeof(double));
	if (e == cudaSuccess) e = cudaMemcpy(d_img, data, plane * layers * sizeof(WORD), cudaMemcpyHostToDevice);
	if (e == cudaSuccess) e = cudaMemset(d_hist, 0, (size_t) layers * SSC_HIST_STRIDE * sizeof(uint32_t));
	int rc = SSC_OK;
	if (e == cudaSuccess) {
		ssc_band b{};	// every layer as a one-layer "frame": K4 histograms layer 0 of every frame
		b.frames = d_img;
		b.frame_stride = plane;
		b.layer_stride = plane;
		b.n_frames = layers;
		b.layers = 1;
		b.width = rx;
		b.height = ry;
		b.buf_rows = b.out_rows = ry;
		rc = launch_histogram(b, d_hist, nullptr);
		if (!rc) {
			k_basic_stats<<<layers, 1024>>>(d_hist, nullcheck, (double) plane, d_stats);
			k_row_noise<<<layers * ry, NOISE_THREADS>>>(d_img, rx, nullcheck, d_rows);
			e = cudaGetLastError();
		}
	}
	if (e == cudaSuccess && !rc) e = cudaMemcpy(stats, d_stats, layers * sizeof(ssc_imstats), cudaMemcpyDeviceToHost);
	if (e == cudaSuccess && !rc) e = cudaMemcpy(rows.data(), d_rows, rows.size() * sizeof(double), cudaMemcpyDeviceToHost);
	cudaFree(d_img); cudaFree(d_hist); cudaFree(d_stats); cudaFree(d_rows);
	if (rc) return io_fail(rc, "histogram launch failed");
	if (e != cudaSuccess) return io_fail(SSC_ERR_CUDA, "statistics failed: %s", cudaGetErrorString(e));
	WORD mx = 0;	// get_normalized_value (utils.c:454-459): 255 when no layer has a sample above it
	for (int c = 0; c < layers; c++) mx = std::max<WORD>(mx, (WORD) stats[c].max);
	for (int c = 0; c < layers; c++) {
		std::vector<double> sd;
		for (int y = 0; y < ry; y++) if (rows[(size_t) c * ry + y] >= 0.0) sd.push_back(rows[(size_t) c * ry + y]);
		double noise = 0.0;
		if (sd.size() == 1) noise = sd[0];
		else if (sd.size() > 1) {	// the median over the rows
			std::sort(sd.begin(), sd.end());
			noise = (sd[(sd.size() - 1) / 2] + sd[sd.size() / 2]) / 2.0;
		}
		stats[c].bgnoise = 0.70710678 * noise;
		stats[c].normValue = mx <= 255 ? 255.0 : 65535.0;
		if (stats[c].ngoodpix < 1.0) return io_fail(SSC_ERR_GENERIC, "no data computed");	// statistics() returns NULL (:247-250)
	}
	return SSC_OK;
}
