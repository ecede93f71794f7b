// ssc_synth.cu -- counter-based synthetic frame generator for
// the benchmark and the full-size property tests.  Every sample is a pure
// function of (seed, kind, frame, layer, y, x), so any row band on any GPU
// regenerates exactly the same pixels without materialising its neighbours.
#include "ssc_device.cuh"
#include "ssc_kernels.h"

namespace ssc {

__device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
	z += 0x9E3779B97F4A7C15ull;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}

__device__ __forceinline__ float u01(uint32_t v) { return ((float) v + 0.5f) * (1.0f / 4294967296.0f); }

// kind 0: deep-sky field (background + sky gradient + star grid + noise + outliers + dead samples)
// kind 1: planetary disc (limb-darkened disc, strong noise, no stars)
__global__ void k_synth(WORD *__restrict__ frames, int n_frames, int layers, int rows, int width,
		int first_frame, int row0, int height, int kind, uint64_t seed) {
	const size_t plane = (size_t) rows * width;
	const size_t total = plane * layers;
	for (size_t g = (size_t) blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (size_t) gridDim.x * blockDim.x) {
		const int x = (int) (g % width);
		const int y = row0 + (int) ((g / width) % rows);
		const int c = (int) (g / plane);
		const float chan_gain = layers == 1 ? 1.0f : (c == 0 ? 0.8f : (c == 1 ? 1.0f : 0.6f));
		float fixed_signal = 0.0f, sky = 0.0f;
		if (kind == 0) {
			sky = 300.0f * (float) x / (float) (width > 1 ? width - 1 : 1);
			// one star per 96x96 cell, gaussian PSF sigma 2 px, peak 2000..90000 (some saturate)
			const int cx = x / 96, cy = y / 96;
			for (int dy = -1; dy <= 1; dy++)
				for (int dx = -1; dx <= 1; dx++) {
					const int ux = cx + dx, uy = cy + dy;
					if (ux < 0 || uy < 0) continue;
					uint64_t h = splitmix64(seed ^ 0x5354415253ull ^ ((uint64_t) uy << 24) ^ (uint64_t) ux);
					float sx = (ux + u01((uint32_t) h)) * 96.0f, sy = (uy + u01((uint32_t) (h >> 32))) * 96.0f;
					uint64_t h2 = splitmix64(h);
					float peak = 2000.0f + 88000.0f * u01((uint32_t) h2) * u01((uint32_t) (h2 >> 32));
					float r2 = (x - sx) * (x - sx) + (y - sy) * (y - sy);
					if (r2 < 400.0f) fixed_signal += peak * __expf(-r2 * 0.125f);
				}
		} else {
			const float rad = 0.33f * (float) (width < height ? width : height);
			const float rx = ((float) x - 0.5f * width) / rad, ry = ((float) y - 0.5f * height) / rad;
			const float r2 = rx * rx + ry * ry;
			fixed_signal = r2 < 1.0f ? 8000.0f + 32000.0f * sqrtf(1.0f - r2) : 0.0f;
		}
		for (int k = 0; k < n_frames; k++) {
			const int kf = first_frame + k;
			uint64_t h = splitmix64(seed ^ ((uint64_t) kind << 56) ^ ((uint64_t) kf << 40) ^ ((uint64_t) c << 36)
					^ ((uint64_t) y << 18) ^ (uint64_t) x);
			uint64_t h2 = splitmix64(h);
			const float u1 = u01((uint32_t) h), u2 = u01((uint32_t) (h >> 32));
			const float gauss = sqrtf(-2.0f * __logf(u1)) * __cosf(6.2831853f * u2);
			float v;
			if (kind == 0) {
				const float bg = 1000.0f + 37.0f * (float) (kf % 17);
				const float gain = 1.0f + 0.02f * (float) ((kf * 7) % 11 - 5);
				const float sigma = 40.0f + (float) (kf % 5);
				v = (bg * gain + sky + fixed_signal * gain) * chan_gain + sigma * gauss;
			} else {
				const float gain = 1.0f + 0.01f * (float) ((kf * 7) % 11 - 5);
				v = 200.0f + fixed_signal * gain * chan_gain + 200.0f * gauss;
			}
			const float uo = u01((uint32_t) h2), uz = u01((uint32_t) (h2 >> 32));
			if (uo < 1e-3f) v += 5000.0f + 45000.0f * (uo * 1e3f);	// cosmic ray / hot sample
			v = fminf(fmaxf(rintf(v), 0.0f), 65535.0f);
			if (uz < 1e-4f) v = 0.0f;				// dead sample
			frames[(size_t) k * total + g] = (WORD) v;
		}
	}
}

int launch_synth(WORD *frames, int n_frames, int layers, int rows, int width, int first_frame,
		int row0, int full_height, int kind, uint64_t seed, cudaStream_t stream) {
	if (n_frames < 1 || layers < 1 || rows < 1 || width < 1 || row0 < 0 || row0 + rows > full_height)
		return SSC_ERR_ARG;
	size_t total = (size_t) layers * rows * width;
	unsigned int grid = (unsigned int) ((total + 255) / 256 < 148 * 32 ? (total + 255) / 256 : 148 * 32);
	k_synth<<<grid, 256, 0, stream>>>(frames, n_frames, layers, rows, width, first_frame, row0, full_height,
			kind, seed);
	return cudaGetLastError() == cudaSuccess ? SSC_OK : SSC_ERR_CUDA;
}

}  // namespace ssc
