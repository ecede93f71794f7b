// ssc_kernels.h -- host-callable launchers of the CUDA kernels (internal; the
// public surface is include/siril_stack_cuda.h).
#pragma once
#include <cuda_runtime.h>
#include "../../include/siril_stack_cuda.h"

#define SSC_GENERIC_MAX_FRAMES 2048

namespace ssc {

// Tables that take the divisions out of the linear-fit rejection without changing a bit
// (ssc_column.cuh, LINEARFIT): device arrays indexed by n = 1..n_max.
struct FitTables {
	const double *inv;	// RN(1 / n)
	const double *mx;	// gsl_fit_linear's running mean of x_i = i after n steps (data independent)
	const double *mdx2;	// its running mean of (i - mx[n])^2 after n steps
	const double *inv_den;	// RN(12 / (n^2 (n^2 - 1))): 1 / (n Sxx - Sx^2) of x_i = i (ssc_stack_warp.cu)
	const double *unbias;	// n / (n - 1), the reference's own quotient in gsl_stats_ushort_sd's last step
	const double *inv_n2;	// RN(1 / n^2): with inv they are ssc_device.cuh's SmallN tables for n <= 1024 (ssc_stack_mid.cu)
	int n_max;		// 0: no tables, literal divisions
};

struct StackParams {
	ssc_band band;
	int normalize;
	const double *coeff;		// device double[3][N] or nullptr
	// device uint[N] written by k_verify_cheap_norm, or nullptr: norm_bad[f] == 0 means that the cheap form of the
	// normalization (ssc_network.cuh) equals the reference's for every sample value of frame f.  Filled in by
	// launch_stack() -- callers leave it alone.
	const unsigned int *norm_bad;
	int rejection;
	double sig_low, sig_high;
	WORD *out;
	size_t out_layer_stride;
	int out_buf_row0;
	unsigned long long *counters;	// device uint64[SSC_COUNTERS] or nullptr
	FitTables fit;			// filled by launch_stack() for LINEARFIT
	// Deferred hand-back (the operator layer stacks a band unit by unit): 0 = every launch finishes its own
	// undecided pixels; 1 / 2 = the warp kernel only lists them (2: the list starts empty), with pixel ids in the
	// geometry of the whole piece (layer + hb_layer0, row + hb_row0 of hb_rows rows, hb_pixels in all), and
	// launch_deferred_handback() finishes the whole list once -- the literal fit costs ~8 ms per launch however few
	// pixels it gets
	int hb_defer, hb_layer0, hb_row0, hb_rows;
	unsigned long long hb_pixels;
};

// ssc_stack_generic.cu
int launch_stack_generic(const StackParams &p, const unsigned int *list, const unsigned int *list_count,
		unsigned int list_capacity, int is_median, cudaStream_t stream);

// ssc_stack_fast.cu
bool fast_path_eligible(const StackParams &p, int is_median);
int launch_stack_fast(const StackParams &p, int is_median, cudaStream_t stream);
// ssc_stack_mid.cu: 65..256 frames, register network on runs of 64 + shared-memory merges
bool mid_path_eligible(const StackParams &p, int is_median);
int launch_stack_mid(const StackParams &p, int is_median, cudaStream_t stream);
// ssc_stack_warp.cu: one warp per pixel pair -- linear fit (8..1024 frames), median of 513..1024 frames
bool warp_path_eligible(const StackParams &p, int is_median);
int launch_stack_warp(const StackParams &p, int is_median, cudaStream_t stream);
// finishes what launches with hb_defer listed; p describes the whole piece (its band, result and counters)
int launch_deferred_handback(const StackParams &p, cudaStream_t stream);
// picks the register-network kernel when the shape allows it, the general kernel otherwise
int launch_stack(const StackParams &p, int is_median, cudaStream_t stream);
void release_kernel_scratch(int dev);	// per-device lists and tables the kernels keep between launches
extern int g_kernel_path;	// 0 automatic, 1 general kernel only, 2 automatic without the cheap normalization (ssc_set_kernel_path)

// ssc_sum.cu
int launch_extreme(const ssc_band &b, bool take_max, WORD *out, unsigned int *extreme, cudaStream_t stream);
int launch_sum_accumulate(const ssc_band &b, uint32_t *acc, unsigned long long *maxim, cudaStream_t stream);
int launch_sum_finish(const uint32_t *acc, size_t count, const unsigned long long *maxim, WORD *out,
		cudaStream_t stream);

// ssc_stats.cu
int launch_histogram(const ssc_band &b, uint32_t *hist, cudaStream_t stream);
// the same with one histogram pointer per frame (device array; may point into peer GPUs' memory)
int launch_histogram_to(const ssc_band &b, uint32_t *hist, uint32_t *const *frame_hist, cudaStream_t stream);
int launch_histogram_ex(const ssc_band &b, uint32_t *hist, uint32_t *const *frame_hist, uint32_t *local_gt255,
		bool layer0_only, cudaStream_t stream);
int launch_other_layers(const ssc_band &b, uint32_t *hist, uint32_t *const *frame_hist, const uint32_t *local_gt255,
		cudaStream_t stream);

// ssc_peer.cu, for the operator layer: the pieces of the exchanges it sequences itself
int peer_local_histograms(ssc_peer_group *g, int n_frames, uint32_t **hist, cudaStream_t stream);	// where this rank counts
int peer_signal_histograms(ssc_peer_group *g, cudaStream_t stream);	// "this rank's counts are all in"
int launch_ikss(const uint32_t *hist, int n_frames, double *stats, uint32_t *scratch, cudaStream_t stream);
int launch_norm_coeffs(const double *stats, int n_frames, int ref_image, int mode, double *coeff,
		int *status, cudaStream_t stream);

// ssc_io.cu: decoding of the file formats and preprocess, on rows of every frame of a band
// mode 1: FITS BITPIX 16 + BZERO 32768 (byte swap, sign flip), 2: big-endian unsigned (byte swap)
int launch_decode_inplace(WORD *base, size_t frame_stride, int n_frames, size_t count, int mode, cudaStream_t st);
int launch_decode_raw(const void *raw, size_t raw_frame_bytes, WORD *dst, size_t frame_stride, size_t layer_stride,
		int n_frames, size_t count, int bytes, int colours, int big_endian, int bgr, int signed_bzero, cudaStream_t st);
int launch_preprocess(WORD *base, size_t frame_stride, int n_frames, size_t count, const WORD *offset, const WORD *dark,
		const WORD *flat, float level, cudaStream_t st);

// ssc_synth.cu
int launch_synth(WORD *frames, int n_frames, int layers, int rows, int width, int first_frame,
		int row0, int full_height, int kind, uint64_t seed, cudaStream_t stream);

}  // namespace ssc
