/*
 * fft_power.cuh - kernel 1 of the hot path: window multiply, forward FFT,
 * log-power, waterfall-row emission.
 *
 * Replaces, for any power-of-two N in 512..16384:
 *   lib/fosphor/fft.cl:397-466     fft1D_1024 (window on load, DFT, complex store)
 *   lib/fosphor/display.cl:133-146 pwr = log10(hypot(re, im)) + waterfall write
 * The reference stores the complex spectrum (8 B/sample) and re-reads it in its
 * display kernel; here the epilogue of the last FFT pass turns each bin into
 * log10|X| and writes it straight into the waterfall ring (4 B/sample), which
 * is also what kernel 2 (accumulate.cuh) consumes, from L2.
 *
 * Algorithm: mixed-radix Stockham autosort with per-thread register DFTs
 * (fft_regs.cuh).  N = R0 * R1 (two passes) or R0 * R1 * R1 (three passes).
 * Pass 0 reads global memory (coalesced float2, window folded in) and has no
 * twiddles; later passes take precomputed twiddles tw[t][k] from a table laid
 * out so a warp reads consecutive entries.  N = 1024 is R0 = R1 = 32: one
 * warp per spectrum, one padded shared-memory exchange, no block barrier.
 *
 * Input addressing: spectrum s starts at in + s * hop complex samples; hop = N
 * is the reference contract (pre-overlapped windows, cl.c:903-910), hop < N
 * subsumes the overlap block (lib/overlap_cc_impl.cc:64-79).
 */
#pragma once
#include "fft_regs.cuh"

namespace fosphor_b200 {

/* 0.5 * log10(2): pwr = log10(sqrt(p)) = 0.5 * log10(2) * log2(p) */
#define FOSPHOR_HALF_LOG10_2 0.15051499783199060f

template <int N_, int R0_, int R1_, int NPASS_>
struct FftPlan {
	static constexpr int N = N_, R0 = R0_, R1 = R1_, NPASS = NPASS_;
	static_assert(NPASS == 2 || NPASS == 3, "passes");
	static_assert((NPASS == 2 ? R0 * R1 : R0 * R1 * R1) == N, "radices");
	static constexpr int NB0 = N / R0;            /* butterflies in pass 0   */
	static constexpr int NB1 = N / R1;            /* ... in passes 1 (and 2) */
	static constexpr int T = NB1 < 32 ? 32 : NB1; /* threads per spectrum    */
	static constexpr int SPB = T == 32 ? 4 : 1;   /* spectra per CTA         */
	static constexpr int THREADS = T * SPB;
	static constexpr int PADSHIFT = ilog2c(R0);
	static constexpr int SM_ELEMS = N + (N >> PADSHIFT);
	static constexpr size_t SMEM = sizeof(float2) * (size_t)SM_ELEMS * SPB;
	/* twiddle table: pass 1 [R1][P1] with P1 = R0, then pass 2 [R1][P2], P2 = R0*R1 */
	static constexpr int TW1 = R1 * R0;
	static constexpr int TW2 = NPASS == 3 ? R1 * R0 * R1 : 0;
	static constexpr int TW_ELEMS = TW1 + TW2;
};

template <class P>
__device__ __forceinline__ int pad_idx(int a) { return a + (a >> P::PADSHIFT); }

__device__ __forceinline__ float log_power(float2 x)
{
	/* display.cl:136: log10(hypot(re, im)).  re^2+im^2 cannot overflow for
	 * |x| < 1.8e19, far above any windowed sum of [-1,1] IQ samples. */
	return __log2f(fmaf(x.x, x.x, x.y * x.y)) * FOSPHOR_HALF_LOG10_2;
}

template <class P>
__device__ __forceinline__ void sync_spectrum()
{
	if constexpr (P::T == 32)
		__syncwarp();
	else
		__syncthreads();
}

/* CPLX = false: the product kernel (log-power rows into the waterfall ring).
 * CPLX = true:  stores X[k] as cf32 [n_spectra][N]; used only by the
 *               stage-wise parity test of the transform (tests/test_fft_parity.py). */
template <class P, bool CPLX>
__global__ void __launch_bounds__(P::THREADS)
fft_power_kernel(const float2 *__restrict__ in, long long hop,
                 const float *__restrict__ win, const float2 *__restrict__ tw,
                 float *__restrict__ wf, int wf_pos, int wf_mask,
                 float2 *__restrict__ cplx_out, int n_spectra)
{
	constexpr int N = P::N, R0 = P::R0, R1 = P::R1;
	extern __shared__ float2 smem[];

	const int sub = threadIdx.x / P::T;      /* spectrum slot in this CTA */
	const int tid = threadIdx.x % P::T;
	const int s = blockIdx.x * P::SPB + sub;
	if (s >= n_spectra)
		return;                          /* whole warp / whole CTA exits */

	float2 *buf = smem + (size_t)sub * P::SM_ELEMS;
	const float2 *x = in + (long long)s * hop;
	float *row = CPLX ? nullptr : wf + (size_t)((wf_pos + s) & wf_mask) * N;
	float2 *crow = CPLX ? cplx_out + (size_t)s * N : nullptr;

	auto store_bin = [&](int idx, float2 v) {
		if constexpr (CPLX)
			crow[idx] = v;
		else
			row[idx] = log_power(v);
	};

	/* ---- pass 0: global -> registers -> smem (P = 1, no twiddles) ---- */
	for (int i = tid; i < P::NB0; i += P::T) {
		float2 v[R0];
#pragma unroll
		for (int t = 0; t < R0; t++) {
			const float2 a = x[i + t * P::NB0];
			const float w = __ldg(&win[i + t * P::NB0]);
			v[t] = make_float2(a.x * w, a.y * w);   /* fft.cl:416-417 */
		}
		dif<R0>(v);
		static_for<0, R0>([&](auto tc) {
			constexpr int t = decltype(tc)::value;
			buf[pad_idx<P>(i * R0 + t)] = v[brev<R0>(t)];
		});
	}
	sync_spectrum<P>();

	if (tid >= P::NB1)
		return;          /* only N = 512 (NB1 = 16 < warp); no barrier follows there */

	/* ---- pass 1 (P = R0) ---- */
	const int i = tid;
	constexpr int PP = R0;
	const int k = i & (PP - 1);
	float2 v[R1];
#pragma unroll
	for (int t = 0; t < R1; t++)
		v[t] = buf[pad_idx<P>(i + t * P::NB1)];
#pragma unroll
	for (int t = 1; t < R1; t++)
		v[t] = cmul(v[t], __ldg(&tw[t * PP + k]));
	dif<R1>(v);

	if constexpr (P::NPASS == 2) {
		/* last pass: j = i, output index i + t * NB1 (coalesced) */
		static_for<0, R1>([&](auto tc) {
			constexpr int t = decltype(tc)::value;
			store_bin(i + t * P::NB1, v[brev<R1>(t)]);
		});
	} else {
		static_assert(P::NPASS == 2 || P::NB1 == P::T, "3-pass plans keep every thread busy");
		const int j = (i - k) * R1 + k;
		sync_spectrum<P>();                     /* all reads of pass 1 done */
		static_for<0, R1>([&](auto tc) {
			constexpr int t = decltype(tc)::value;
			buf[pad_idx<P>(j + t * PP)] = v[brev<R1>(t)];
		});
		sync_spectrum<P>();

		/* ---- pass 2 (P = R0 * R1), last ---- */
		constexpr int P2 = R0 * R1;
		const int k2 = i & (P2 - 1);
#pragma unroll
		for (int t = 0; t < R1; t++)
			v[t] = buf[pad_idx<P>(i + t * P::NB1)];
#pragma unroll
		for (int t = 1; t < R1; t++)
			v[t] = cmul(v[t], __ldg(&tw[P::TW1 + t * P2 + k2]));
		dif<R1>(v);
		static_for<0, R1>([&](auto tc) {
			constexpr int t = decltype(tc)::value;
			store_bin(i + t * P::NB1, v[brev<R1>(t)]);
		});
	}
}

/* The supported sizes (BASELINE.json configs[4] sweep) */
using Plan512   = FftPlan<512,   16, 32, 2>;
using Plan1024  = FftPlan<1024,  32, 32, 2>;
using Plan2048  = FftPlan<2048,   8, 16, 3>;
using Plan4096  = FftPlan<4096,  16, 16, 3>;
using Plan8192  = FftPlan<8192,   8, 32, 3>;
using Plan16384 = FftPlan<16384, 16, 32, 3>;

} /* namespace fosphor_b200 */
