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
#include "ssc_column.cuh"
#include "ssc_kernels.h"
#include "ssc_sortnet.h"

namespace ssc {

constexpr int FAST_THREADS = 128;

__device__ __forceinline__ double word_to_double(unsigned int p) {
	return dsub(__hiloint2double(0x43300000, (int) p), 4503599627370496.0);	// (2^52 + p) - 2^52, exact
}

// round_to_WORD (utils.c:68-74) for |y| < 2^31: floor(y + 0.5) read from the low word of
// (y + 0.5) +rz 1.5*2^52, then max(min(., 65535), 0).  Equals the reference for every such y:
// y <= 0 -> floor <= 0 -> 0;  0 < y <= 65535 -> trunc(y + 0.5);  y > 65535 -> >= 65535 -> 65535.
__device__ __forceinline__ unsigned int round_to_word_fast(double y) {
	double u = __dadd_rz(dadd(y, 0.5), 6755399441055744.0);
	return (unsigned int) __vimin_s32_relu(__double2loint(u), 65535);
}

template <int NP>
__device__ __forceinline__ void sort_network(unsigned int (&v)[NP]) {
#define SSC_CE(a, b) { unsigned int lo_ = __vminu2(v[a], v[b]); v[b] = __vmaxu2(v[a], v[b]); v[a] = lo_; }
	if constexpr (NP == 16) { SSC_SORTNET_16 }
	else if constexpr (NP == 32) { SSC_SORTNET_32 }
	else { SSC_SORTNET_64 }
#undef SSC_CE
}

template <int NP>
__global__ void __launch_bounds__(FAST_THREADS, 4)
k_stack_fast(StackParams p, int is_median) {
	extern __shared__ unsigned int cols[];		// [NP][FAST_THREADS] packed sorted columns
	__shared__ double s_coef[3 * 64];
	__shared__ int s_fast_norm;
	__shared__ unsigned long long blk_counters[SSC_COUNTERS];

	const ssc_band &b = p.band;
	const int W = b.width, N = b.n_frames, tid = threadIdx.x;
	const int W2 = W >> 1;
	const bool norm = p.coeff != nullptr && p.normalize != NO_NORM;

	if (tid < SSC_COUNTERS) blk_counters[tid] = 0;
	if (tid == 0) s_fast_norm = 1;
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
	unsigned long long g = (unsigned long long) blockIdx.x * FAST_THREADS + tid;
	const bool active = g < total;
	if (!active) g = total - 1;
	const int x2 = (int) (g % W2);
	const int rr = (int) ((g / W2) % b.out_rows);
	const int layer = (int) (g / ((unsigned long long) W2 * b.out_rows));
	const int R = b.out_row0 + rr;				// result storage row (stacking.c:731,1600)
	const int srow = b.row_order == SSC_ROWS_SER ? b.height - 1 - R : R;
	const unsigned int *src = reinterpret_cast<const unsigned int *>(
			b.frames + (size_t) layer * b.layer_stride + (size_t) (srow - b.buf_row0) * W) + x2;
	const size_t fstride2 = b.frame_stride >> 1;

	// ---- gather: one packed load per frame, all issued before first use ----
	unsigned int v[NP];
#pragma unroll
	for (int f = 0; f < NP; f++)
		v[f] = f < N ? __ldg(src + (size_t) f * fstride2) : 0xFFFFFFFFu;

	// ---- normalize (stacking.c:1635-1652; median path :750-764) and exact sums ----
	unsigned long long Sa = 0, Sb = 0, Qa = 0, Qb = 0;
	if (norm) {
		const bool fast = s_fast_norm != 0;
		const int mode = p.normalize;
#pragma unroll
		for (int f = 0; f < NP; f++) {
			if (f < N) {
				unsigned int pa = v[f] & 0xFFFFu, pb = v[f] >> 16;
				const double off = s_coef[f], mul = s_coef[64 + f], sc = s_coef[128 + f];
				if (fast) {
					double da = word_to_double(pa), db = word_to_double(pb);
					if (mode == ADDITIVE_SCALING || mode == MULTIPLICATIVE_SCALING) {
						da = dmul(da, sc);
						db = dmul(db, sc);
					}		// ADDITIVE / MULTIPLICATIVE: scale is exactly 1.0, x * 1.0 == x
					if (mode == ADDITIVE || mode == ADDITIVE_SCALING) {
						da = dsub(da, off);
						db = dsub(db, off);
					} else {
						da = dmul(da, mul);
						db = dmul(db, mul);
					}
					pa = round_to_word_fast(da);
					pb = round_to_word_fast(db);
				} else {
					pa = normalize_sample(pa, mode, off, mul, sc);
					pb = normalize_sample(pb, mode, off, mul, sc);
				}
				Sa += pa; Sb += pb; Qa += sq(pa); Qb += sq(pb);
				v[f] = pa | (pb << 16);
			}
		}
	} else {
#pragma unroll
		for (int f = 0; f < NP; f++) {
			if (f < N) {
				unsigned int pa = v[f] & 0xFFFFu, pb = v[f] >> 16;
				Sa += pa; Sb += pb; Qa += sq(pa); Qb += sq(pb);
			}
		}
	}

	unsigned int ra, rb;
	RejectState sta{0, 0, 0, 0}, stb{0, 0, 0, 0};
	if (!is_median && p.rejection == NO_REJEC) {
		// stacking.c:1790-1794 with nothing rejected: no ordering needed at all
		ra = round_to_word(ddiv((double) (long long) Sa, (double) N));
		rb = round_to_word(ddiv((double) (long long) Sb, (double) N));
	} else {
		sort_network<NP>(v);			// sentinels (0xFFFF) sort behind the N samples
		if (is_median && N == NP) {		// stacking.c:765-767: truncating median, straight from registers
			unsigned int lo = v[NP / 2 - 1], hi = v[NP / 2];
			ra = ((lo & 0xFFFFu) + (hi & 0xFFFFu)) >> 1;
			rb = ((lo >> 16) + (hi >> 16)) >> 1;
		} else {
#pragma unroll
			for (int f = 0; f < NP; f++)
				if (f < N) cols[f * FAST_THREADS + tid] = v[f];
			Column ca{reinterpret_cast<WORD *>(cols) + 2 * tid, 2 * FAST_THREADS};
			Column cb{reinterpret_cast<WORD *>(cols) + 2 * tid + 1, 2 * FAST_THREADS};
			if (is_median) {
				const int lhs = (N - 1) >> 1, rhs = N >> 1;
				ra = (ca.get(lhs) + ca.get(rhs)) >> 1;
				rb = (cb.get(lhs) + cb.get(rhs)) >> 1;
			} else {
				ra = reject_sorted_and_mean(ca, N, p.rejection, p.sig_low, p.sig_high, Sa, Qa, sta);
				rb = reject_sorted_and_mean(cb, N, p.rejection, p.sig_low, p.sig_high, Sb, Qb, stb);
			}
		}
	}

	if (active) {
		unsigned int *dst = reinterpret_cast<unsigned int *>(
				p.out + (size_t) layer * p.out_layer_stride + (size_t) (R - p.out_buf_row0) * W) + x2;
		*dst = ra | (rb << 16);
		const unsigned int rl = sta.rej_lo + stb.rej_lo, rh = sta.rej_hi + stb.rej_hi;
		if (rl) atomicAdd(&blk_counters[layer * 2 + 0], (unsigned long long) rl);
		if (rh) atomicAdd(&blk_counters[layer * 2 + 1], (unsigned long long) rh);
		if (sta.guard + stb.guard) atomicAdd(&blk_counters[6], (unsigned long long) (sta.guard + stb.guard));
		if (sta.cap + stb.cap) atomicAdd(&blk_counters[7], (unsigned long long) (sta.cap + stb.cap));
	}
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
	return true;
}

template <int NP>
static int launch_fast_np(const StackParams &p, int is_median, cudaStream_t stream) {
	const ssc_band &b = p.band;
	const unsigned long long total = (unsigned long long) b.layers * b.out_rows * (b.width >> 1);
	if (total == 0) return SSC_OK;
	const size_t smem = (size_t) NP * FAST_THREADS * sizeof(unsigned int);
	if (cudaFuncSetAttribute(k_stack_fast<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem) != cudaSuccess)
		return SSC_ERR_CUDA;
	const unsigned int grid = (unsigned int) ((total + FAST_THREADS - 1) / FAST_THREADS);
	k_stack_fast<NP><<<grid, FAST_THREADS, smem, stream>>>(p, is_median);
	return cudaGetLastError() == cudaSuccess ? SSC_OK : SSC_ERR_CUDA;
}

int launch_stack_fast(const StackParams &p, int is_median, cudaStream_t stream) {
	const int n = p.band.n_frames;
	if (n <= 16) return launch_fast_np<16>(p, is_median, stream);
	if (n <= 32) return launch_fast_np<32>(p, is_median, stream);
	return launch_fast_np<64>(p, is_median, stream);
}

}  // namespace ssc
