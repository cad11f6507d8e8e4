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
	/* exchange-buffer padding: one float2 every 2^PADSHIFT.  log2(R0) keeps the stride-R0 stores of
	 * pass 0 conflict free; at least 4 so that the pass-1 stores of the R0 = 8 plans (two runs of 8
	 * float2, R0*R1 apart) land in different halves of the 32 banks (ncu: 46 % of the shared
	 * wavefronts of N = 2048 / 8192 were conflict replays with a shift of 3) */
	static constexpr int PADSHIFT = ilog2c(R0) < 4 ? 4 : ilog2c(R0);
	/* second level, N = 8192 only (R0 = 8, R1 = 32): the two runs of a pass-1 store are R0*R1 = 256
	 * float2 apart, which the first level pads by 16 - the same banks again; 8 more float2 per 256
	 * move the second run to the other half of the banks (tests/test_fft_layout.py simulates every
	 * access of every plan) */
	static constexpr int PADSHIFT2 = (R0_ == 8 && R1_ == 32) ? 8 : 0;
	static constexpr int PAD2 = PADSHIFT2 ? 8 : 0;
	/* launch bound of the plain kernel = the CTAs/SM that shared memory allows (8192: 3 CTAs at <= 80 registers) */
	static constexpr int MIN_CTAS = R1_ == 64 ? 1 : (N_ <= 1024 ? 4 : (N_ == 2048 ? 8 : (N_ == 4096 ? 4 : (N_ == 8192 ? 3 : 1))));
	static constexpr int SM_ELEMS = N + (N >> PADSHIFT) + (PADSHIFT2 ? PAD2 * (N >> PADSHIFT2) : 0);
	static constexpr size_t SMEM = sizeof(float2) * (size_t)SM_ELEMS * SPB;
	/* twiddle table: pass 1 [R1][P1] with P1 = R0, then pass 2 [R1][P2], P2 = R0*R1 */
	static constexpr int TW1 = R1 * R0;
	static constexpr int TW2 = NPASS == 3 ? R1 * R0 * R1 : 0;
	static constexpr int TW_ELEMS = TW1 + TW2;
};

template <class P>
__device__ __forceinline__ int pad_idx(int a)
{
	if constexpr (P::PADSHIFT2 != 0)
		return a + (a >> P::PADSHIFT) + P::PAD2 * (a >> P::PADSHIFT2);
	else
		return a + (a >> P::PADSHIFT);
}

/* pad_idx(base + t * STRIDE) == pad_idx(base) + pad_step<P, STRIDE>(t) for every access pattern of
 * the kernels below: the low PADSHIFT bits of base and of t * STRIDE never carry (pass-1/2 reads:
 * STRIDE = NB1 is a multiple of 2^PADSHIFT; pass-1 stores: base = (i-k)*R1 + k with k < R0 and
 * STRIDE = R0 = 8 or 16; pass-0 stores: base = i*R0, STRIDE = 1, t < R0).  The compiler cannot see
 * that, and without it every element costs a shift and an add (12 % of the instructions of the
 * three-pass kernels, ncu); with it the element offsets are immediates. */
template <class P, int STRIDE>
__host__ __device__ constexpr int pad_step(int t)
{
	return t * STRIDE + ((t * STRIDE) >> P::PADSHIFT) +
	       (P::PADSHIFT2 ? P::PAD2 * ((t * STRIDE) >> P::PADSHIFT2) : 0);
}

__device__ __forceinline__ float log_power(float2 x)
{
	/* display.cl:136: log10(hypot(re, im)).  re^2+im^2 cannot overflow for
	 * |x| < 1.8e19, far above any windowed sum of [-1,1] IQ samples. */
	float l;
	/* MUFU.LG2 without the denormal pre-scaling of __log2f: |X|^2 < 1.2e-38
	 * (|X| < 1e-19) is treated as 0 -> -inf, far below any displayable level */
	asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(fmaf(x.x, x.x, x.y * x.y)));
	return l * FOSPHOR_HALF_LOG10_2;
}

/* fft.cl:416-417: the windowed sample is an f32 product.  __fmul_rn is never contracted into the
 * first butterfly's add (a plain * may or may not be, depending on the surrounding kernel), which
 * keeps every kernel variant bit-identical and matches the oracle's rounded products. */
__device__ __forceinline__ float2 win_mul(float2 x, float w)
{
	return make_float2(__fmul_rn(x.x, w), __fmul_rn(x.y, w));
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
__global__ void __launch_bounds__(P::THREADS, P::MIN_CTAS)
fft_power_kernel(const float2 *__restrict__ in, long long hop,
                 const float *__restrict__ win, const float2 *__restrict__ tw,
                 float *__restrict__ wf, int wf_pos, int wf_mask,
                 float2 *__restrict__ cplx_out, int n_spectra, int pf_dist)
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
	if constexpr (P::SPB == 1) {
		/* pull the spectrum of the CTA that will follow this one on the SM into L2 (pf_dist =
		 * resident CTAs; 0 = off / unaligned spectra): its pass-0 loads then hit L2 */
		if (tid == 0 && pf_dist > 0 && s + pf_dist < n_spectra)
			asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;"
			             ::"l"(in + (long long)(s + pf_dist) * hop), "r"((unsigned)(sizeof(float2) * N)) : "memory");
	}
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
			v[t] = win_mul(a, w);   /* fft.cl:416-417 */
		}
		dif<R0>(v);
		static_for<0, R0>([&](auto tc) {
			constexpr int t = decltype(tc)::value;
			buf[pad_idx<P>(i * R0) + pad_step<P, 1>(t)] = v[brev<R0>(t)];
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
		v[t] = buf[pad_idx<P>(i) + pad_step<P, P::NB1>(t)];
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
			buf[pad_idx<P>(j) + pad_step<P, PP>(t)] = v[brev<R1>(t)];
		});
		sync_spectrum<P>();

		/* ---- pass 2 (P = R0 * R1), last ---- */
		constexpr int P2 = R0 * R1;
		const int k2 = i & (P2 - 1);
#pragma unroll
		for (int t = 0; t < R1; t++)
			v[t] = buf[pad_idx<P>(i) + pad_step<P, P::NB1>(t)];
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

/* ------------------------------------------------------------------------ */
/* Streaming variant for the one-warp-per-spectrum plans (N = 512, 1024)      */
/* ------------------------------------------------------------------------ */
/*
 * Same arithmetic as fft_power_kernel (bit-identical results), different data
 * movement: warps are persistent and each one double-buffers its input with
 * the TMA bulk-copy engine (cp.async.bulk global -> shared, completion on an
 * mbarrier), so the 8 KB of the NEXT spectrum are in flight while the current
 * one is being transformed.  ncu of the plain kernel showed it latency bound
 * (long-scoreboard + LSU-queue stalls, 47 % issue utilisation, 32 % of HBM
 * peak); here no thread ever waits on a global load in steady state and the
 * load instructions leave the LSU queue altogether.  The buffer that held the
 * inputs of the current spectrum is reused as the padded exchange buffer
 * between the two passes, so a warp needs 2 x SM_ELEMS float2 of shared memory.
 * The window lives in registers for the lifetime of the warp.
 * Requires 16-byte aligned spectra: even hop and 16-byte aligned base.
 */
__device__ __forceinline__ unsigned smem_u32(const void *p)
{
	return (unsigned)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}

__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(unsigned bar, unsigned parity)
{
	unsigned ok;
	asm volatile("{\n\t.reg .pred p;\n\t"
	             "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
	             "selp.u32 %0, 1, 0, p;\n\t}"
	             : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
	return ok != 0;
}

__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

/* L2 residency hints (engine knob L2_HINTS): samples are read once (evict first), log-power rows are
 * read again by the accumulate kernel a chunk later (evict last) */
__device__ __forceinline__ unsigned long long l2_policy_evict_first()
{
	unsigned long long p;
	asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
	return p;
}

__device__ __forceinline__ unsigned long long l2_policy_evict_last()
{
	unsigned long long p;
	asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
	return p;
}

__device__ __forceinline__ void bulk_g2s_hint(unsigned dst, const void *src, unsigned bytes, unsigned bar,
                                              unsigned long long pol)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
	             ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
}

__device__ __forceinline__ void st_f32_hint(float *p, float v, unsigned long long pol)
{
	asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(pol) : "memory");
}

template <class P>
struct StreamCfg {
	static_assert(P::NPASS == 2 && P::T == 32, "one warp per spectrum plans only");
	/* N = 512 has 16 last-pass butterflies: a warp takes TWO spectra at a time so that
	 * no lane idles in the radix-32 pass (lanes 0-15 spectrum A, 16-31 spectrum B) */
	static constexpr int SPW = P::NB1 < 32 ? 32 / P::NB1 : 1;   /* spectra per warp and iteration */
	static constexpr int WARPS = 4;
	static constexpr int THREADS = WARPS * 32;
	static constexpr int CTAS_PER_SM = 3;
	static constexpr int BUF_ELEMS = P::SM_ELEMS * SPW;           /* >= SPW * N: holds inputs, then the exchange */
	static constexpr size_t BUF_BYTES_ALL = sizeof(float2) * (size_t)BUF_ELEMS * 2 * WARPS;
	static constexpr size_t SMEM = BUF_BYTES_ALL + 8 * 2 * WARPS;
	/* TWREG variant: + the window, shared by the CTA, every tap twice (w, w): one 64-bit load feeds
	 * one packed multiply (FMUL2) of the complex sample */
	static constexpr size_t SMEM_TWREG = SMEM + sizeof(float2) * P::N;
	static constexpr unsigned IN_BYTES = sizeof(float2) * P::N;   /* one spectrum */
};

/* TWREG = false: window slice in registers, pass-1 twiddles fetched (L1) per spectrum.
 * TWREG = true : pass-1 twiddles in registers for the lifetime of the warp, window
 *                read from shared memory - no global load at all in the steady state. */
template <class P, bool TWREG>
__global__ void __launch_bounds__(StreamCfg<P>::THREADS, StreamCfg<P>::CTAS_PER_SM)
fft_power_stream_kernel(const float2 *__restrict__ in, long long hop,
                        const float *__restrict__ win, const float2 *__restrict__ tw,
                        float *__restrict__ wf, int wf_pos, int wf_mask, int n_spectra, int l2_hints)
{
	using C = StreamCfg<P>;
	constexpr int N = P::N, R0 = P::R0, R1 = P::R1;
	extern __shared__ __align__(16) unsigned char smem_raw[];

	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int gw = blockIdx.x * C::WARPS + warp;
	const unsigned long long pol_in = l2_hints ? l2_policy_evict_first() : 0ull;
	const unsigned long long pol_out = l2_hints ? l2_policy_evict_last() : 0ull;
	const int G = gridDim.x * C::WARPS;

	float2 *bufs = reinterpret_cast<float2 *>(smem_raw) + (size_t)warp * 2 * C::BUF_ELEMS;
	unsigned long long *bars = reinterpret_cast<unsigned long long *>(
		smem_raw + sizeof(float2) * (size_t)C::BUF_ELEMS * 2 * C::WARPS) + warp * 2;
	const unsigned bar0 = smem_u32(bars), buf0 = smem_u32(bufs);
	constexpr unsigned BUF_BYTES = sizeof(float2) * C::BUF_ELEMS;

	if (lane == 0) {
		mbar_init(bar0, 1);
		mbar_init(bar0 + 8, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncwarp();

	/* window: registers (slice of this lane: pass-0 element lane + t*NB0) or shared memory */
	constexpr int SPW = C::SPW;
	const int q1 = lane / P::NB1, i1 = lane % P::NB1;    /* pass 1: spectrum slot and butterfly of this lane */
	float wreg[TWREG ? 1 : R0];
	const float2 *swin = reinterpret_cast<const float2 *>(smem_raw + C::SMEM);
	float2 twreg[TWREG ? R1 : 1];
	if constexpr (TWREG) {
		float2 *sw = reinterpret_cast<float2 *>(smem_raw + C::SMEM);
		for (int i = threadIdx.x; i < N; i += C::THREADS) {
			const float w = __ldg(&win[i]);
			sw[i] = make_float2(w, w);
		}
		__syncthreads();
		if (q1 < SPW) {
#pragma unroll
			for (int t = 1; t < R1; t++)
				twreg[t] = __ldg(&tw[t * R0 + (i1 & (R0 - 1))]);
		}
	} else {
		if (lane < P::NB0) {
#pragma unroll
			for (int t = 0; t < R0; t++)
				wreg[t] = __ldg(&win[lane + t * P::NB0]);
		}
	}

	/* unit u = spectra u*SPW .. u*SPW+SPW-1 (the last unit may be short) */
	const int n_units = (n_spectra + SPW - 1) / SPW;
	auto request = [&](int u, unsigned slot) {           /* lane 0 only */
		const int s0 = u * SPW;
		const int nv = n_spectra - s0 < SPW ? n_spectra - s0 : SPW;
		mbar_expect_tx(bar0 + 8 * slot, C::IN_BYTES * (unsigned)nv);
		for (int q = 0; q < nv; q++) {
			const unsigned dst = buf0 + slot * BUF_BYTES + (unsigned)q * (unsigned)(sizeof(float2) * P::SM_ELEMS);
			if (l2_hints)
				bulk_g2s_hint(dst, in + (long long)(s0 + q) * hop, C::IN_BYTES, bar0 + 8 * slot, pol_in);
			else
				bulk_g2s(dst, in + (long long)(s0 + q) * hop, C::IN_BYTES, bar0 + 8 * slot);
		}
	};

	int u = gw;
	if (u < n_units && lane == 0)
		request(u, 0u);
	unsigned phases = 0u;                /* bit b = parity to wait for on barrier b */

	for (int it = 0; u < n_units; it++, u += G) {
		const int b = it & 1;
		float2 *buf = bufs + (size_t)b * C::BUF_ELEMS;

		/* prefetch the next unit into the other buffer (last touched by this
		 * warp's generic-proxy loads/stores one iteration ago) */
		__syncwarp();
		if (lane == 0 && u + G < n_units) {
			asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
			request(u + G, (unsigned)(b ^ 1));
		}

		while (!mbar_try_wait(bar0 + 8 * b, (phases >> b) & 1u)) { }
		phases ^= 1u << b;

		/* ---- pass 0, one spectrum slot after the other (a short last unit transforms stale
		 * shared memory in its empty slot; nothing of it is stored) ---- */
#pragma unroll
		for (int q = 0; q < SPW; q++) {
			float2 *bq = buf + q * P::SM_ELEMS;
			float2 v0[R0];
			if (lane < P::NB0) {
#pragma unroll
				for (int t = 0; t < R0; t++) {
					const float2 x = bq[lane + t * P::NB0];
					if constexpr (TWREG)
						v0[t] = __fmul2_rn(x, swin[lane + t * P::NB0]);   /* fft.cl:416-417, both components at once */
					else
						v0[t] = win_mul(x, wreg[t]);
				}
			}
			__syncwarp();                   /* inputs consumed: the slot becomes the exchange buffer */
			if (lane < P::NB0) {
				dif<R0>(v0);
				static_for<0, R0>([&](auto tc) {
					constexpr int t = decltype(tc)::value;
					bq[pad_idx<P>(lane * R0) + pad_step<P, 1>(t)] = v0[brev<R0>(t)];
				});
			}
		}
		__syncwarp();

		/* ---- pass 1 (P = R0), last ---- */
		if (q1 < SPW) {
			const float2 *bq = buf + q1 * P::SM_ELEMS;
			const int s = u * SPW + q1;
			float *row = wf + (size_t)((wf_pos + s) & wf_mask) * N;
			const int k = i1 & (R0 - 1);
			float2 v[R1];
#pragma unroll
			for (int t = 0; t < R1; t++)
				v[t] = bq[pad_idx<P>(i1) + pad_step<P, P::NB1>(t)];
#pragma unroll
			for (int t = 1; t < R1; t++)
				v[t] = cmul(v[t], TWREG ? twreg[t] : __ldg(&tw[t * R0 + k]));
			dif<R1>(v);
			if (s < n_spectra) {
				if (l2_hints) {
					static_for<0, R1>([&](auto tc) {
						constexpr int t = decltype(tc)::value;
						st_f32_hint(&row[i1 + t * P::NB1], log_power(v[brev<R1>(t)]), pol_out);
					});
				} else {
					static_for<0, R1>([&](auto tc) {
						constexpr int t = decltype(tc)::value;
						row[i1 + t * P::NB1] = log_power(v[brev<R1>(t)]);
					});
				}
			}
		}
	}
}

/* ------------------------------------------------------------------------ */
/* Streaming variant for the three-pass plans (N = 2048, 4096, 8192)           */
/* ------------------------------------------------------------------------ */
/*
 * One CTA per spectrum in flight, persistent CTAs striding over the spectra.
 * Thread 0 asks the TMA engine for the NEXT spectrum (one cp.async.bulk of 8N
 * bytes, mbarrier completion) into the second shared buffer while the CTA
 * transforms the current one; the buffer that held the inputs becomes the padded
 * exchange buffer of the same spectrum (pass-0 results wait in registers across
 * one barrier so that no input is overwritten before it was read).  Same
 * arithmetic as fft_power_kernel: bit-identical output.  N = 16384 does not
 * fit two buffers in 227 KB and stays on the plain kernel.
 */
template <class P>
struct CtaStreamCfg {
	static_assert(P::NPASS == 3, "three-pass plans");
	static constexpr int THREADS = P::T;
	static constexpr int BF0 = P::NB0 / P::T;                     /* pass-0 butterflies per thread */
	static_assert(BF0 * P::T == P::NB0, "pass 0 divides evenly");
	static constexpr int BUF_ELEMS = P::SM_ELEMS;
	static constexpr size_t SMEM = sizeof(float2) * (size_t)BUF_ELEMS * 2 + 16;
	static constexpr unsigned IN_BYTES = sizeof(float2) * P::N;
	static constexpr bool FITS = SMEM <= 200 * 1024;
	static constexpr int CTAS_PER_SM = SMEM > 110 * 1024 ? 1 : (SMEM > 72 * 1024 ? 2 : (SMEM > 54 * 1024 ? 3 : 4));
};

template <class P>
__global__ void __launch_bounds__(CtaStreamCfg<P>::THREADS, CtaStreamCfg<P>::CTAS_PER_SM)
fft_power_cta_stream_kernel(const float2 *__restrict__ in, long long hop,
                            const float *__restrict__ win, const float2 *__restrict__ tw,
                            float *__restrict__ wf, int wf_pos, int wf_mask, int n_spectra)
{
	using C = CtaStreamCfg<P>;
	constexpr int N = P::N, R0 = P::R0, R1 = P::R1;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	float2 *bufs = reinterpret_cast<float2 *>(smem_raw);
	const unsigned buf0 = smem_u32(bufs);
	const unsigned bar0 = smem_u32(smem_raw + sizeof(float2) * (size_t)C::BUF_ELEMS * 2);
	constexpr unsigned BUF_BYTES = sizeof(float2) * C::BUF_ELEMS;
	const int tid = threadIdx.x;

	if (tid == 0) {
		mbar_init(bar0, 1);
		mbar_init(bar0 + 8, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
		if ((int)blockIdx.x < n_spectra) {
			mbar_expect_tx(bar0, C::IN_BYTES);
			bulk_g2s(buf0, in + (long long)blockIdx.x * hop, C::IN_BYTES, bar0);
		}
	}
	__syncthreads();

	unsigned phases = 0u;
	int it = 0;
	for (int s = blockIdx.x; s < n_spectra; s += gridDim.x, it++) {
		const int b = it & 1;
		float2 *buf = bufs + (size_t)b * C::BUF_ELEMS;

		/* everybody is done with the other buffer (previous spectrum): refill it */
		__syncthreads();
		if (tid == 0 && s + (int)gridDim.x < n_spectra) {
			const unsigned nb = (unsigned)(b ^ 1);
			asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
			mbar_expect_tx(bar0 + 8 * nb, C::IN_BYTES);
			bulk_g2s(buf0 + nb * BUF_BYTES, in + (long long)(s + gridDim.x) * hop, C::IN_BYTES, bar0 + 8 * nb);
		}
		while (!mbar_try_wait(bar0 + 8 * b, (phases >> b) & 1u)) { }
		phases ^= 1u << b;

		float *row = wf + (size_t)((wf_pos + s) & wf_mask) * N;

		/* ---- pass 0: inputs from shared memory, results parked in registers ---- */
		float2 v0[C::BF0][R0];
#pragma unroll
		for (int q = 0; q < C::BF0; q++) {
			const int i = tid + q * P::T;
#pragma unroll
			for (int t = 0; t < R0; t++) {
				const float2 x = buf[i + t * P::NB0];
				const float w = __ldg(&win[i + t * P::NB0]);
				v0[q][t] = win_mul(x, w);           /* fft.cl:416-417 */
			}
			dif<R0>(v0[q]);
		}
		__syncthreads();                        /* all inputs consumed: buf becomes the exchange */
#pragma unroll
		for (int q = 0; q < C::BF0; q++) {
			const int i = tid + q * P::T;
			static_for<0, R0>([&](auto tc) {
				constexpr int t = decltype(tc)::value;
				buf[pad_idx<P>(i * R0) + pad_step<P, 1>(t)] = v0[q][brev<R0>(t)];
			});
		}
		__syncthreads();

		/* ---- pass 1 (P = R0) ---- */
		const int i = tid;
		const int k = i & (R0 - 1);
		float2 v[R1];
#pragma unroll
		for (int t = 0; t < R1; t++)
			v[t] = buf[pad_idx<P>(i) + pad_step<P, P::NB1>(t)];
#pragma unroll
		for (int t = 1; t < R1; t++)
			v[t] = cmul(v[t], __ldg(&tw[t * R0 + k]));
		dif<R1>(v);
		const int j = (i - k) * R1 + k;
		__syncthreads();
		static_for<0, R1>([&](auto tc) {
			constexpr int t = decltype(tc)::value;
			buf[pad_idx<P>(j) + pad_step<P, R0>(t)] = v[brev<R1>(t)];
		});
		__syncthreads();

		/* ---- pass 2 (P = R0 * R1), last ---- */
		constexpr int P2 = R0 * R1;
		const int k2 = i & (P2 - 1);
#pragma unroll
		for (int t = 0; t < R1; t++)
			v[t] = buf[pad_idx<P>(i) + pad_step<P, P::NB1>(t)];
#pragma unroll
		for (int t = 1; t < R1; t++)
			v[t] = cmul(v[t], __ldg(&tw[P::TW1 + t * P2 + k2]));
		dif<R1>(v);
		static_for<0, R1>([&](auto tc) {
			constexpr int t = decltype(tc)::value;
			row[i + t * P::NB1] = log_power(v[brev<R1>(t)]);
		});
	}
}

/* ------------------------------------------------------------------------ */
/* Half-staged persistent variant for N = 16384                               */
/* ------------------------------------------------------------------------ */
/*
 * The exchange buffer of a 16384-point spectrum is 136 KB: one CTA per SM, no room for a
 * second input buffer, and the plain kernel serialises "load 128 KB" and "transform" (ncu:
 * 47 % issue, 35 % of HBM peak).  Here the CTA is persistent and the 64 KB of shared memory
 * that are left stage HALF of the next spectrum - the inputs of pass-0 butterflies 0..511,
 * sixteen 4 KB runs - through the TMA engine while the current spectrum is transformed.
 * The other half is loaded straight into registers at the top of the iteration (the whole
 * next spectrum was pulled into L2 one iteration earlier) and arrives while the staged half
 * is being transformed.  Same arithmetic as fft_power_kernel: bit-identical rows.
 */
template <class P>
struct HalfStageCfg {
	static_assert(P::NPASS == 3 && P::NB0 == 2 * P::T, "two pass-0 butterflies per thread");
	static constexpr int THREADS = P::T;
	static constexpr int STG_ELEMS = P::R0 * P::T;                 /* float2: [R0][T] */
	static constexpr unsigned RUN_BYTES = sizeof(float2) * P::T;  /* one contiguous run of the staged half */
	static constexpr size_t SMEM = sizeof(float2) * ((size_t)P::SM_ELEMS + STG_ELEMS) + 16;
};

template <class P>
__global__ void __launch_bounds__(HalfStageCfg<P>::THREADS, 1)
fft_power_half_stage_kernel(const float2 *__restrict__ in, long long hop,
                            const float *__restrict__ win, const float2 *__restrict__ tw,
                            float *__restrict__ wf, int wf_pos, int wf_mask, int n_spectra)
{
	using C = HalfStageCfg<P>;
	constexpr int N = P::N, R0 = P::R0, R1 = P::R1, T = P::T;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	float2 *buf = reinterpret_cast<float2 *>(smem_raw);                  /* padded exchange buffer */
	float2 *stg = buf + P::SM_ELEMS;                                     /* staged half: [R0][T] */
	const unsigned stg0 = smem_u32(stg);
	const unsigned bar = smem_u32(smem_raw + sizeof(float2) * ((size_t)P::SM_ELEMS + C::STG_ELEMS));
	const int tid = threadIdx.x;

	auto request = [&](int s) {             /* thread 0: staged half of spectrum s, and all of it into L2 */
		const float2 *x = in + (long long)s * hop;
		asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(x), "r"((unsigned)(sizeof(float2) * N)) : "memory");
		mbar_expect_tx(bar, C::RUN_BYTES * R0);
#pragma unroll 1
		for (int t = 0; t < R0; t++)
			bulk_g2s(stg0 + (unsigned)t * C::RUN_BYTES, x + t * P::NB0, C::RUN_BYTES, bar);
	};

	if (tid == 0) {
		mbar_init(bar, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
		if ((int)blockIdx.x < n_spectra)
			request(blockIdx.x);
	}
	__syncthreads();

	unsigned phase = 0u;
	for (int s = blockIdx.x; s < n_spectra; s += gridDim.x) {
		const float2 *x = in + (long long)s * hop;
		float *row = wf + (size_t)((wf_pos + s) & wf_mask) * N;

		/* ---- pass 0, second butterfly (i = tid + T): inputs straight from L2 into registers ---- */
		float2 vb[R0];
#pragma unroll
		for (int t = 0; t < R0; t++)
			vb[t] = __ldcg(&x[tid + T + t * P::NB0]);

		/* ---- pass 0, first butterfly (i = tid): inputs from the staged half ---- */
		while (!mbar_try_wait(bar, phase)) { }
		phase ^= 1u;
		float2 va[R0];
#pragma unroll
		for (int t = 0; t < R0; t++) {
			const float2 a = stg[t * T + tid];
			const float w = __ldg(&win[tid + t * P::NB0]);
			va[t] = win_mul(a, w);                   /* fft.cl:416-417 */
		}
		/* everybody has read the staged half, and everybody is past the pass-2 reads of the previous
		 * spectrum: refill the stage, overwrite the exchange buffer */
		__syncthreads();
		if (tid == 0 && s + (int)gridDim.x < n_spectra) {
			asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
			request(s + gridDim.x);
		}
		dif<R0>(va);
		static_for<0, R0>([&](auto tc) {
			constexpr int t = decltype(tc)::value;
			buf[pad_idx<P>(tid * R0) + pad_step<P, 1>(t)] = va[brev<R0>(t)];
		});
#pragma unroll
		for (int t = 0; t < R0; t++) {
			const float w = __ldg(&win[tid + T + t * P::NB0]);
			vb[t] = win_mul(vb[t], w);
		}
		dif<R0>(vb);
		static_for<0, R0>([&](auto tc) {
			constexpr int t = decltype(tc)::value;
			buf[pad_idx<P>((tid + T) * R0) + pad_step<P, 1>(t)] = vb[brev<R0>(t)];
		});
		__syncthreads();

		/* ---- pass 1 (P = R0) ---- */
		const int i = tid;
		const int k = i & (R0 - 1);
		float2 v[R1];
#pragma unroll
		for (int t = 0; t < R1; t++)
			v[t] = buf[pad_idx<P>(i) + pad_step<P, P::NB1>(t)];
#pragma unroll
		for (int t = 1; t < R1; t++)
			v[t] = cmul(v[t], __ldg(&tw[t * R0 + k]));
		dif<R1>(v);
		const int j = (i - k) * R1 + k;
		__syncthreads();
		static_for<0, R1>([&](auto tc) {
			constexpr int t = decltype(tc)::value;
			buf[pad_idx<P>(j) + pad_step<P, R0>(t)] = v[brev<R1>(t)];
		});
		__syncthreads();

		/* ---- pass 2 (P = R0 * R1), last ---- */
		constexpr int P2 = R0 * R1;
		const int k2 = i & (P2 - 1);
#pragma unroll
		for (int t = 0; t < R1; t++)
			v[t] = buf[pad_idx<P>(i) + pad_step<P, P::NB1>(t)];
#pragma unroll
		for (int t = 1; t < R1; t++)
			v[t] = cmul(v[t], __ldg(&tw[P::TW1 + t * P2 + k2]));
		dif<R1>(v);
		static_for<0, R1>([&](auto tc) {
			constexpr int t = decltype(tc)::value;
			row[i + t * P::NB1] = log_power(v[brev<R1>(t)]);
		});
	}
}

/* ------------------------------------------------------------------------ */
/* Grouped variant for the R0 x 32 x 32 plans (N = R0 * 1024: 8192, 16384)     */
/* ------------------------------------------------------------------------ */
/*
 * After the first (radix-R0) pass a spectrum falls apart into R0 independent 1024-point
 * problems: in the Stockham numbering used above, pass-1 butterfly i and pass-2 butterfly i'
 * exchange data only when i == i' (mod R0).  The plain kernel ignores that and runs both
 * exchanges through CTA-wide barriers (ncu of N = 16384: issue slots 52 % busy, stalls per issue
 * barrier 0.89 + mio_throttle 1.10 + short_scoreboard 0.72).  Here
 *   - warp g of a spectrum owns group g (butterflies i = R0*l + g, l = lane) in BOTH late passes,
 *     so everything between pass 0 and the output is warp-local: __syncwarp instead of four
 *     block barriers, and the group's 1024 values live in a private region with the classic
 *     33-stride transpose padding - every access of every pass is conflict free;
 *   - only pass 0 meets across warps (block barrier 1), and because group g produces exactly
 *     the outputs X[k], k == g (mod R0), the log-power values are parked as f32 in the group's
 *     own region and written out by all threads of the spectrum as coalesced 16-byte stores
 *     (barriers 2 and 3);
 *   - a CTA is 16 warps = 16 groups = 16/R0 spectra, persistent, one per SM; the inputs of the
 *     next spectrum are requested (plain coalesced loads of data pulled into L2 one iteration
 *     earlier) before the copy-out of the current one and land in registers meanwhile.
 * Same butterflies, same twiddle values, same operation order as fft_power_kernel with the
 * same plan: bit-identical rows.
 */
template <class P>
struct GroupedCfg {
	static_assert(P::NPASS == 3 && P::R1 == 32 && P::N == P::R0 * 1024, "R0 x 32 x 32 plans");
	static constexpr int R0 = P::R0;
	static constexpr int WARPS = 16;
	static constexpr int THREADS = WARPS * 32;
	static constexpr int SPC = WARPS / R0;                  /* spectra per CTA and iteration */
	static constexpr int TPS = 32 * R0;                     /* threads per spectrum */
	static constexpr int BF0 = 32 / R0;                     /* pass-0 butterflies per thread */
	/* float2 per group region: 1024 + 31 of transpose padding, rounded so that consecutive regions
	 * start two banks apart as f32 arrays (the copy-out reads R0 regions at once) */
	static constexpr int REGION = 1057;
	static constexpr size_t SMEM = sizeof(float2) * (size_t)REGION * WARPS;
	static constexpr int TWG_ELEMS = R0 * 32 * 32;          /* pass-2 twiddles regrouped [g][t][lane] */
};

template <class P>
__global__ void __launch_bounds__(GroupedCfg<P>::THREADS, 1)
fft_power_grouped_kernel(const float2 *__restrict__ in, long long hop,
                         const float *__restrict__ win, const float2 *__restrict__ tw,
                         const float2 *__restrict__ twg,
                         float *__restrict__ wf, int wf_pos, int wf_mask, int n_spectra, int prefetch)
{
	using C = GroupedCfg<P>;
	constexpr int N = P::N, R0 = C::R0, NB0 = P::NB0;
	static_assert(NB0 == 1024, "pass-0 butterflies per spectrum");
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int q = warp / R0, g = warp % R0;                  /* spectrum slot of this warp, its group */
	const int ts = g * 32 + lane;                            /* thread within the spectrum */
	float2 *slot = reinterpret_cast<float2 *>(smem_raw) + (size_t)q * R0 * C::REGION;
	float2 *mine = slot + (size_t)g * C::REGION;
	const int stride = gridDim.x * C::SPC;
	int s = blockIdx.x * C::SPC + q;

	auto slot_barrier = [&]() {                              /* the 32*R0 threads of this spectrum */
		if constexpr (C::SPC == 1)
			__syncthreads();
		else
			asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "n"(C::TPS) : "memory");
	};

	float2 v0[C::BF0][R0];
	auto load_inputs = [&](int sp) {
		const float2 *x = in + (long long)sp * hop;
#pragma unroll
		for (int u = 0; u < C::BF0; u++)
#pragma unroll
			for (int t = 0; t < R0; t++)
				v0[u][t] = __ldcg(&x[ts + u * C::TPS + t * NB0]);
	};

	if (s < n_spectra)
		load_inputs(s);
	for (; s < n_spectra; s += stride) {
		/* the spectrum after next: into L2 now, so that its loads (issued one iteration from now,
		 * before the copy-out) find it there */
		if (prefetch && ts == 0 && s + 2 * stride < n_spectra)
			asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;"
			             ::"l"(in + (long long)(s + 2 * stride) * hop), "r"((unsigned)(sizeof(float2) * N)) : "memory");

		/* ---- pass 0 (radix R0, no twiddles): registers -> the R0 group regions ---- */
#pragma unroll
		for (int u = 0; u < C::BF0; u++) {
			const int i0 = ts + u * C::TPS;
#pragma unroll
			for (int t = 0; t < R0; t++)
				v0[u][t] = win_mul(v0[u][t], __ldg(&win[i0 + t * NB0]));   /* fft.cl:416-417 */
			dif<R0>(v0[u]);
		}
		slot_barrier();                  /* the copy-out of the previous spectrum has left the regions */
#pragma unroll
		for (int u = 0; u < C::BF0; u++) {
			const int i0 = ts + u * C::TPS;
			float2 *dst = slot + i0 + (i0 >> 5);
			static_for<0, R0>([&](auto tc) {
				constexpr int t = decltype(tc)::value;
				dst[t * C::REGION] = v0[u][brev<R0>(t)];     /* Stockham position R0*i0 + t: group t, element i0 */
			});
		}
		slot_barrier();

		/* ---- pass 1 (P = R0): butterfly i = R0*lane + g, warp-local ---- */
		float2 v[32];
#pragma unroll
		for (int t = 0; t < 32; t++)
			v[t] = mine[lane + 33 * t];                      /* element lane + 32 t of the group */
#pragma unroll
		for (int t = 1; t < 32; t++)
			v[t] = cmul(v[t], __ldg(&tw[t * R0 + g]));       /* k = i & (R0-1) = g: one value per warp */
		dif<32>(v);
		__syncwarp();
		static_for<0, 32>([&](auto tc) {
			constexpr int t = decltype(tc)::value;
			mine[33 * lane + t] = v[brev<32>(t)];            /* element 32 lane + t */
		});
		__syncwarp();

		/* ---- pass 2 (P = 32 R0), last: same butterflies, outputs X[R0 (lane + 32 t) + g] ---- */
#pragma unroll
		for (int t = 0; t < 32; t++)
			v[t] = mine[lane + 33 * t];
#pragma unroll
		for (int t = 1; t < 32; t++)
			v[t] = cmul(v[t], __ldg(&twg[(g * 32 + t) * 32 + lane]));
		dif<32>(v);
		__syncwarp();                        /* every lane has its inputs: the region becomes the f32 parking lot */
		{
			float *park = reinterpret_cast<float *>(mine);
			static_for<0, 32>([&](auto tc) {
				constexpr int t = decltype(tc)::value;
				park[lane + 32 * t] = log_power(v[brev<32>(t)]);   /* display.cl:136 */
			});
		}
		slot_barrier();

		/* inputs of this slot's next spectrum: in flight during the copy-out */
		if (s + stride < n_spectra)
			load_inputs(s + stride);

		/* ---- copy-out: X[k] sits in group k % R0 at k / R0 ---- */
		{
			float *row = wf + (size_t)((wf_pos + s) & wf_mask) * N;
			const float *parked = reinterpret_cast<const float *>(slot);
#pragma unroll
			for (int c = 0; c < N / (4 * C::TPS); c++) {
				const int k = 4 * (ts + c * C::TPS);
				float4 o;
				o.x = parked[((k + 0) % R0) * (2 * C::REGION) + (k + 0) / R0];
				o.y = parked[((k + 1) % R0) * (2 * C::REGION) + (k + 1) / R0];
				o.z = parked[((k + 2) % R0) * (2 * C::REGION) + (k + 2) / R0];
				o.w = parked[((k + 3) % R0) * (2 * C::REGION) + (k + 3) / R0];
				*reinterpret_cast<float4 *>(row + k) = o;    /* display.cl:141-146 */
			}
		}
	}
}

/* The supported sizes (BASELINE.json configs[4] sweep) */
using Plan512   = FftPlan<512,   16, 32, 2>;
using Plan1024  = FftPlan<1024,  32, 32, 2>;
using Plan2048  = FftPlan<2048,   8, 16, 3>;
using Plan4096  = FftPlan<4096,  16, 16, 3>;
using Plan8192  = FftPlan<8192,   8, 32, 3>;
using Plan16384 = FftPlan<16384, 16, 32, 3>;
/* two-pass alternatives with a radix-64 register pass: one shared-memory exchange and one twiddle
 * stage less than the three-pass plans (those are LSU bound: 77-93 % l1tex in ncu) */
using Plan2048R64 = FftPlan<2048, 32, 64, 2>;
using Plan4096R64 = FftPlan<4096, 64, 64, 2>;

} /* namespace fosphor_b200 */
