// ssc_stack_pipe.cu -- K3 for the headline shape (a full stack of 32 or 64 frames, sigma
// clipping, no registration shifts) as a warp-specialised pipeline.
//
// The per-pixel work has three parts that load different pipes: fetching and normalizing the
// samples (HBM latency, FP64 + conversion units), ordering them (the ALU's min/max), and the
// iterative rejection (latency-bound FP64 chains).  A warp that does all three in sequence keeps
// each pipe idle most of the time, and the 64 live registers of the network cap occupancy so that
// other warps cannot fill the gaps.  Here the parts run in different warps of a persistent block:
//
//   producer warps   cp.async (LDGSTS) the next tile of 32 pixel pairs x NP frames straight into a
//                    shared-memory slot, normalize it in place in a rolled loop (a few hundred
//                    bytes of code), accumulate the exact sums, and publish the slot (mbarrier);
//   sorter warps     take a published slot into registers, run the packed-u16x2 Batcher network,
//                    finish the clipping in registers (ssc_stack_fast.cu explains the scheme),
//                    write the results, hand back the slot; pixels that need more than the register
//                    window go to the leftover list.
//
// Each sorter owns two slots (double buffering); one producer feeds two sorters.  Synchronisation
// is per slot and per warp pair (full / empty mbarriers, 32 arrivals each): no block-wide barrier
// in the steady state.  Waits are bounded: a protocol error traps instead of hanging the GPU.
#include "ssc_column.cuh"
#include "ssc_kernels.h"
#include "ssc_network.cuh"
#include "ssc_leftover.cuh"

namespace ssc {

constexpr int PIPE_PRODUCERS = 5;
constexpr int PIPE_SORTERS = 2 * PIPE_PRODUCERS;
constexpr int PIPE_WARPS = PIPE_PRODUCERS + PIPE_SORTERS;
constexpr int PIPE_THREADS = PIPE_WARPS * 32;
constexpr int PIPE_SLOTS = 2;		// per sorter
#ifndef SSC_PIPE_PASSES
#define SSC_PIPE_PASSES 4
#endif
#ifndef SSC_PIPE_GROUP_SYNC
#define SSC_PIPE_GROUP_SYNC 1
#endif

__device__ __forceinline__ unsigned int smem_u32(const void *p) { return (unsigned int) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned int count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
	asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.release.cta.shared::cta.b64 st, [%0];\n}\n" :: "r"(smem_u32(bar)) : "memory");
}
// wait until the phase with the given parity has completed; bounded (trap on a protocol error)
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned int parity) {
	unsigned int done = 0;
	for (int spin = 0; spin < (1 << 20); spin++) {
		asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
				: "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
		if (done) return;
		__nanosleep(spin < 64 ? 32 : 256);	// do not burn issue slots the other roles need
	}
	asm volatile("trap;\n");
}

__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gmem_src) {
	asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" :: "r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int PENDING>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(PENDING) : "memory"); }

struct PairRef {
	unsigned int toff;	// offset of the pixel pair inside a frame, in 32-bit words
	unsigned int out_idx;	// element index of the pair's first pixel in the result buffer
	int layer;
};

__device__ __forceinline__ PairRef locate_pair(const StackParams &p, unsigned long long g) {
	const ssc_band &b = p.band;
	const int W = b.width, W2 = W >> 1;
	const int x2 = (int) (g % W2);
	const int rr = (int) ((g / W2) % b.out_rows);
	const int layer = (int) (g / ((unsigned long long) W2 * b.out_rows));
	const int R = b.out_row0 + rr;				// result storage row (stacking.c:731,1600)
	const int srow = b.row_order == SSC_ROWS_SER ? b.height - 1 - R : R;
	PairRef r;
	r.toff = (unsigned int) ((((size_t) layer * b.layer_stride + (size_t) (srow - b.buf_row0) * W) >> 1) + x2);
	r.out_idx = (unsigned int) ((size_t) layer * p.out_layer_stride + (size_t) (R - p.out_buf_row0) * W + 2 * x2);
	r.layer = layer;
	return r;
}

// one slot: the tile's packed samples [NP][32 lanes], then the lanes' exact sums
template <int NP>
struct __align__(16) PipeSlot {
	unsigned int w[NP * 32];
	unsigned long long q[2][32];	// sum of squares, pixel a / b of every lane
	unsigned int s[2][32];		// sum
};

// normalization of the lane's NP words in place + sums (rolled: small code)
template <int NP, int KIND>
__device__ __forceinline__ void produce_normalized(PipeSlot<NP> &slot, int lane, const double *s_coef) {
	unsigned int Sa = 0, Sb = 0;
	unsigned long long Qa = 0, Qb = 0;
#pragma unroll 4
	for (int f = 0; f < NP; f++) {
		unsigned int w = slot.w[f * 32 + lane];
		if constexpr (KIND != NO_NORM) {
			double off = 0.0, mul = 1.0, sc = 1.0;
			if constexpr (KIND == ADDITIVE || KIND == ADDITIVE_SCALING) off = s_coef[f];
			if constexpr (KIND == MULTIPLICATIVE || KIND == MULTIPLICATIVE_SCALING) mul = s_coef[64 + f];
			if constexpr (KIND == ADDITIVE_SCALING || KIND == MULTIPLICATIVE_SCALING) sc = s_coef[128 + f];
			w = normalize_pair_fast<KIND>(w, off, mul, sc, Qa, Qb);
			slot.w[f * 32 + lane] = w;
		} else {
			const unsigned int pa = w & 0xFFFFu, pb = w >> 16;
			Qa += (unsigned long long) pa * pa;
			Qb += (unsigned long long) pb * pb;
		}
		Sa = __dp2a_lo(w, 0x0001u, Sa);
		Sb = __dp2a_lo(w, 0x0100u, Sb);
	}
	slot.s[0][lane] = Sa; slot.s[1][lane] = Sb;
	slot.q[0][lane] = Qa; slot.q[1][lane] = Qb;
}

template <int NP>
__global__ void __launch_bounds__(PIPE_THREADS, 1)
k_stack_pipe(StackParams p, LeftoverList list, unsigned int n_tiles) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	PipeSlot<NP> *slots = reinterpret_cast<PipeSlot<NP> *>(smem_raw);		// [PIPE_SORTERS][PIPE_SLOTS]
	__shared__ unsigned long long bar_full[PIPE_SORTERS][PIPE_SLOTS], bar_empty[PIPE_SORTERS][PIPE_SLOTS];
	__shared__ double s_coef[3 * 64];
	__shared__ double s_unbias[65], s_inv_n[65], s_inv_n2[65];
	__shared__ unsigned long long blk_counters[SSC_COUNTERS];
	__shared__ int s_slow_norm;

	const ssc_band &b = p.band;
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const bool norm = p.coeff != nullptr && p.normalize != NO_NORM;

	if (tid < SSC_COUNTERS) blk_counters[tid] = 0;
	if (tid == 0) s_slow_norm = 0;
	__syncthreads();
	if (tid >= 1 && tid <= 64) {
		s_unbias[tid] = tid > 1 ? ddiv((double) tid, (double) (tid - 1)) : 0.0;
		s_inv_n[tid] = ddiv(1.0, (double) tid);
		s_inv_n2[tid] = ddiv(1.0, (double) (tid * tid));
	}
	if (norm && tid < NP) {
		const double off = p.coeff[tid], mul = p.coeff[NP + tid], sc = p.coeff[2 * NP + tid];
		s_coef[tid] = off;
		s_coef[64 + tid] = mul;
		s_coef[128 + tid] = sc;
		// domain of the 2^52 rounding trick (ssc_network.cuh): every intermediate below 2^31
		const bool additive = p.normalize == ADDITIVE || p.normalize == ADDITIVE_SCALING;
		const double bound = additive ? dadd(dmul(fabs(sc), 65535.0), fabs(off)) : dmul(dmul(fabs(sc), 65535.0), fabs(mul));
		if (!(bound < 2147483000.0)) s_slow_norm = 1;	// also catches NaN / inf
	}
	if (tid < PIPE_SORTERS * PIPE_SLOTS) {
		mbar_init(&bar_full[tid / PIPE_SLOTS][tid % PIPE_SLOTS], 32);
		mbar_init(&bar_empty[tid / PIPE_SLOTS][tid % PIPE_SLOTS], 32);
	}
	const SmallN tables{s_unbias, s_inv_n, s_inv_n2};
	__syncthreads();
	const int slow_norm = norm && s_slow_norm;

	const unsigned long long total = (unsigned long long) b.layers * b.out_rows * (b.width >> 1);
	const unsigned int sorter_stride = gridDim.x * PIPE_SORTERS;

	if (warp < PIPE_PRODUCERS) {
		// ===================================== PRODUCER =====================================
		// feeds sorters 2*warp and 2*warp+1; job j = (round r, which sorter a): tile and slot follow
		const char *frames_bytes = reinterpret_cast<const char *>(b.frames);
		const size_t fstride_bytes = b.frame_stride * sizeof(WORD);
		const unsigned int first[2] = {blockIdx.x * PIPE_SORTERS + 2 * warp, blockIdx.x * PIPE_SORTERS + 2 * warp + 1};
		auto job_tile = [&](unsigned int j) { return (unsigned long long) first[j & 1] + (unsigned long long) (j >> 1) * sorter_stride; };
		auto issue = [&](unsigned int j) {	// cp.async the lane's NP words of job j into its slot
			const int a = j & 1, r = j >> 1;
			PipeSlot<NP> &slot = slots[(2 * warp + a) * PIPE_SLOTS + (r & 1)];
			mbar_wait(&bar_empty[2 * warp + a][r & 1], ((r >> 1) & 1) ^ 1);	// the slot's previous use was consumed
			unsigned long long g = job_tile(j) * 32 + lane;
			if (g >= total) g = total - 1;
			const PairRef ref = locate_pair(p, g);
			const char *src = frames_bytes + (size_t) ref.toff * 4;
#pragma unroll 4
			for (int f = 0; f < NP; f++) {
				cp_async4(&slot.w[f * 32 + lane], src);
				src += fstride_bytes;
			}
			cp_async_commit();
		};
		unsigned int n_jobs = 0;	// jobs of this producer: both sorters' tile sequences interleaved
		while (job_tile(n_jobs) < n_tiles) n_jobs++;	// (sequences are monotone; a's run ends no later than it starts again)
		// note: job_tile(j) for odd j can exceed n_tiles one job before the even ones do; such a job is skipped below
		if (n_jobs > 0) issue(0);
		for (unsigned int j = 0; j < n_jobs; j++) {
			const bool more = j + 1 < n_jobs;
			if (more) issue(j + 1);				// next tile's loads fly while this one is normalized
			if (more) cp_async_wait<1>(); else cp_async_wait<0>();
			const int a = j & 1, r = j >> 1;
			PipeSlot<NP> &slot = slots[(2 * warp + a) * PIPE_SLOTS + (r & 1)];
			if (slow_norm) {
				// coefficients outside the 2^52 trick's domain: conversion-based form (rare)
				unsigned int Sa = 0, Sb = 0;
				unsigned long long Qa = 0, Qb = 0;
				for (int f = 0; f < NP; f++) {
					const unsigned int w = slot.w[f * 32 + lane];
					const unsigned int pa = normalize_sample(w & 0xFFFFu, p.normalize, s_coef[f], s_coef[64 + f], s_coef[128 + f]);
					const unsigned int pb = normalize_sample(w >> 16, p.normalize, s_coef[f], s_coef[64 + f], s_coef[128 + f]);
					slot.w[f * 32 + lane] = pa | (pb << 16);
					Sa += pa; Sb += pb;
					Qa += (unsigned long long) pa * pa;
					Qb += (unsigned long long) pb * pb;
				}
				slot.s[0][lane] = Sa; slot.s[1][lane] = Sb;
				slot.q[0][lane] = Qa; slot.q[1][lane] = Qb;
			} else {
				switch (norm ? p.normalize : NO_NORM) {
				case ADDITIVE: produce_normalized<NP, ADDITIVE>(slot, lane, s_coef); break;
				case MULTIPLICATIVE: produce_normalized<NP, MULTIPLICATIVE>(slot, lane, s_coef); break;
				case ADDITIVE_SCALING: produce_normalized<NP, ADDITIVE_SCALING>(slot, lane, s_coef); break;
				case MULTIPLICATIVE_SCALING: produce_normalized<NP, MULTIPLICATIVE_SCALING>(slot, lane, s_coef); break;
				default: produce_normalized<NP, NO_NORM>(slot, lane, s_coef); break;
				}
			}
			mbar_arrive(&bar_full[2 * warp + a][r & 1]);		// publish (release)
		}
	} else {
		// ====================================== SORTER ======================================
		const int s = warp - PIPE_PRODUCERS;
		// sorters that share a scheduler (warp % 4) enter the network together: they then run the
		// same 17 KB of straight-line code at the same time and share its instruction fetches
		const int smsp = warp & 3;
		int group = 0;
		for (int w2 = PIPE_PRODUCERS; w2 < PIPE_WARPS; w2++) group += (w2 & 3) == smsp;
		const unsigned int rounds = (unsigned int) ((n_tiles + sorter_stride - 1) / sorter_stride);
		for (unsigned int r = 0; r < rounds; r++) {
			const unsigned long long tile = (unsigned long long) blockIdx.x * PIPE_SORTERS + s + (unsigned long long) r * sorter_stride;
			if (tile >= n_tiles) {		// no tile for this sorter in the last round: only keep the group's barrier balanced
				if (SSC_PIPE_GROUP_SYNC) asm volatile("bar.sync %0, %1;\n" :: "r"(1 + smsp), "r"(32 * group) : "memory");
				continue;
			}
			PipeSlot<NP> &slot = slots[s * PIPE_SLOTS + (r & 1)];
			unsigned long long g = tile * 32 + lane;
			const bool active = g < total;
			if (!active) g = total - 1;
			const PairRef ref = locate_pair(p, g);

			mbar_wait(&bar_full[s][r & 1], (r >> 1) & 1);		// acquire: the producer's writes are visible
			unsigned int v[NP];
#pragma unroll
			for (int f = 0; f < NP; f++) v[f] = slot.w[f * 32 + lane];
			unsigned long long S[2] = {slot.s[0][lane], slot.s[1][lane]};
			unsigned long long Q[2] = {slot.q[0][lane], slot.q[1][lane]};

			if (SSC_PIPE_GROUP_SYNC) asm volatile("bar.sync %0, %1;\n" :: "r"(1 + smsp), "r"(32 * group) : "memory");
			sort_network<NP>(v);

			// ---- sigma clipping in registers (the scheme of ssc_stack_fast.cu, MODE_SIGMA_REG) ----
			int nl[2] = {0, 0}, nh[2] = {0, 0};
			unsigned int rl[2] = {0, 0}, rh[2] = {0, 0};
			bool done[2] = {false, false}, punt[2] = {false, false};
#pragma unroll 1
			for (int it = 0; it <= SSC_PIPE_PASSES; it++) {
				if (it == SSC_PIPE_PASSES) {
					punt[0] = !done[0];
					punt[1] = !done[1];
					break;
				}
#pragma unroll
				for (int h = 0; h < 2; h++) {
					if (done[h] || punt[h]) continue;
					const int Nc = NP - nl[h] - nh[h];
					const double sigma = sd_from_sums<true>(S[h], Q[h], Nc, &tables);
					const int lhs = nl[h] + ((Nc - 1) >> 1), rhs = nl[h] + (Nc >> 1);
					unsigned int ml = 0, mr = 0;
#pragma unroll
					for (int k = NP / 2 - 3; k <= NP / 2 + 1; k++) {
						const unsigned int e = h ? v[k] >> 16 : v[k] & 0xFFFFu;
						if (lhs == k) ml = e;
						if (rhs == k) mr = e;
					}
					const int M2 = lhs == rhs ? 2 * (int) ml : (int) (ml + mr);
					const int TL = __double2int_rd(fmin(dmul(2.0, dmul(p.sig_low, sigma)), 1073741824.0));
					const int TH = __double2int_rd(fmin(dmul(2.0, dmul(p.sig_high, sigma)), 1073741824.0));
					int n = 0;
#pragma unroll
					for (int k = 0; k < 3; k++) {
						const unsigned int e = h ? v[k] >> 16 : v[k] & 0xFFFFu;
						if (nl[h] == k && M2 - 2 * (int) e > TL) {
							nl[h]++; n++; rl[h]++;
							S[h] -= e; Q[h] -= (unsigned long long) e * e;
						}
					}
#pragma unroll
					for (int k = 0; k < 6; k++) {
						const unsigned int e = h ? v[NP - 1 - k] >> 16 : v[NP - 1 - k] & 0xFFFFu;
						if (nh[h] == k && 2 * (int) e - M2 > TH) {
							nh[h]++; n++; rh[h]++;
							S[h] -= e; Q[h] -= (unsigned long long) e * e;
						}
					}
					if (nl[h] == 3 || nh[h] == 6) punt[h] = true;
					else if (n == 0) done[h] = true;
				}
				if ((done[0] || punt[0]) && (done[1] || punt[1])) break;
			}
			unsigned int res[2];
#pragma unroll
			for (int h = 0; h < 2; h++) {
				res[h] = mean_to_word<true>(S[h], NP - nl[h] - nh[h], &tables);
				punt[h] = punt[h] && active;
			}
			WORD *dst = p.out + ref.out_idx;
			if (active) {
				if (!punt[0] && !punt[1]) {
					*reinterpret_cast<unsigned int *>(dst) = res[0] | (res[1] << 16);
				} else {
					if (!punt[0]) dst[0] = (WORD) res[0];
					if (!punt[1]) dst[1] = (WORD) res[1];
				}
				const unsigned int cl = (punt[0] ? 0 : rl[0]) + (punt[1] ? 0 : rl[1]);
				const unsigned int ch = (punt[0] ? 0 : rh[0]) + (punt[1] ? 0 : rh[1]);
				if (cl) atomicAdd(&blk_counters[ref.layer * 2 + 0], (unsigned long long) cl);
				if (ch) atomicAdd(&blk_counters[ref.layer * 2 + 1], (unsigned long long) ch);
			}
			// ---- the others: leftover list (one atomic per warp), or in place in the slot if it is full ----
#pragma unroll
			for (int h = 0; h < 2; h++) {
				const unsigned int mask = __ballot_sync(0xffffffffu, punt[h]);
				if (mask == 0) continue;
				unsigned int base = 0;
				if (lane == __ffs(mask) - 1) base = atomicAdd(list.count, (unsigned int) __popc(mask));
				base = __shfl_sync(0xffffffffu, base, __ffs(mask) - 1);
				if (punt[h]) {
					const unsigned int li = base + __popc(mask & ((1u << lane) - 1));
					if (li < list.capacity) {
						uint4 *d = list.cols + (size_t) li * (NP / 8);
						const unsigned int sel = h ? 0x7632u : 0x5410u;
#pragma unroll
						for (int q = 0; q < NP / 8; q++)
							d[q] = make_uint4(__byte_perm(v[8 * q], v[8 * q + 1], sel), __byte_perm(v[8 * q + 2], v[8 * q + 3], sel),
									__byte_perm(v[8 * q + 4], v[8 * q + 5], sel), __byte_perm(v[8 * q + 6], v[8 * q + 7], sel));
						list.meta[li] = make_uint2(ref.out_idx + h, (unsigned int) ref.layer);
					} else {
						// list full (adversarial data): finish in place, the slot is still ours
#pragma unroll
						for (int f = 0; f < NP; f++) slot.w[f * 32 + lane] = v[f];
						Column c{reinterpret_cast<WORD *>(slot.w) + 2 * lane + h, 64};
						unsigned long long s0 = 0, q0 = 0;
						for (int f = 0; f < NP; f++) { const unsigned int e = c.get(f); s0 += e; q0 += sq(e); }
						RejectState ist{0, 0, 0, 0};
						dst[h] = (WORD) reject_sorted_and_mean<true>(c, NP, SIGMA, p.sig_low, p.sig_high, s0, q0, ist, &tables);
						if (ist.rej_lo) atomicAdd(&blk_counters[ref.layer * 2 + 0], (unsigned long long) ist.rej_lo);
						if (ist.rej_hi) atomicAdd(&blk_counters[ref.layer * 2 + 1], (unsigned long long) ist.rej_hi);
						if (ist.guard) atomicAdd(&blk_counters[6], 1ull);
						if (ist.cap) atomicAdd(&blk_counters[7], 1ull);
					}
				}
			}
			__syncwarp();
			mbar_arrive(&bar_empty[s][r & 1]);			// hand the slot back (release)
		}
	}
	__syncthreads();
	if (p.counters && tid < SSC_COUNTERS && blk_counters[tid])
		atomicAdd(&p.counters[tid], blk_counters[tid]);
}

bool pipe_path_eligible(const StackParams &p, int is_median) {
	const ssc_band &b = p.band;
	if (is_median || p.rejection != SIGMA) return false;
	if (b.n_frames != 32 && b.n_frames != 64) return false;
	if (!(p.sig_low >= 0.0) || !(p.sig_high >= 0.0)) return false;
	return fast_path_eligible(p, is_median);
}

template <int NP>
static int launch_pipe_np(const StackParams &p, cudaStream_t stream) {
	const ssc_band &b = p.band;
	const unsigned long long total = (unsigned long long) b.layers * b.out_rows * (b.width >> 1);
	if (total == 0) return SSC_OK;
	const unsigned long long tiles = (total + 31) / 32;
	if (tiles > 0xFFFFFFF0ull) return SSC_ERR_ARG;
	const size_t smem = sizeof(PipeSlot<NP>) * PIPE_SORTERS * PIPE_SLOTS;
	if (cudaFuncSetAttribute(k_stack_pipe<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem) != cudaSuccess)
		return SSC_ERR_CUDA;
	LeftoverList list{nullptr, 0, nullptr, nullptr, nullptr};
	int rc = leftover_list(NP, total * 2, list, stream);
	if (rc) return rc;
	int dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	const unsigned long long want = (tiles + PIPE_SORTERS - 1) / PIPE_SORTERS;
	const unsigned int grid = (unsigned int) (want < (unsigned long long) sms ? want : (unsigned long long) sms);
	k_stack_pipe<NP><<<grid, PIPE_THREADS, smem, stream>>>(p, list, (unsigned int) tiles);
	if (cudaGetLastError() != cudaSuccess) return SSC_ERR_CUDA;
	return launch_leftover<NP>(p, list, stream);
}

int launch_stack_pipe(const StackParams &p, cudaStream_t stream) {
	return p.band.n_frames == 32 ? launch_pipe_np<32>(p, stream) : launch_pipe_np<64>(p, stream);
}

}  // namespace ssc
