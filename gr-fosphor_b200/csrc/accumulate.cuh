/*
 * accumulate.cuh - kernel 2 of the hot path: per frequency column, fold the B
 * log-power rows of one call into the persistence state.
 *
 * Replaces the second half of the reference's display program:
 *   lib/fosphor/display.cl:149-150,186-214  live spectrum (weighted IIR)
 *   lib/fosphor/display.cl:160-178          bin mapping + hit counting
 *   lib/fosphor/display.cl:217-254          histogram rise / decay
 *   lib/fosphor/display.cl:257-310          max hold with decay
 * The reference runs this on a fixed 64 work-groups that each loop over the
 * whole batch (cl.c:945-948).  Here the batch is split over S CTAs per tile of
 * 32 columns: every CTA counts its rows into a shared-memory tile
 * hits[bin][lane] (lane == column, so a warp never has a bank conflict; the
 * 8 warps of the CTA meet only through shared atomics), flushes the non-zero
 * counts once to a global u32 array, and the LAST CTA of a tile to finish
 * (ticket counter) applies the rise/decay, live and max-hold updates.  With
 * S == 1 the global hit array is bypassed.
 *
 * The rise/decay closed form depends only on (hit count, B):
 *   a = hc/B; b = a/t0r; c = b + 1/t0d; d = b/c; e = (1-c)^B; hv' = (hv-d)e+d
 * so the host tabulates (d, e) for hc = 0..B and the live weights
 * (1-alpha)^(B-1-s) once per distinct B (engine.cu: BatchTables); the kernel
 * does a table lookup + FMA per state cell instead of two transcendental calls.
 */
#pragma once
#include <cfloat>
#include <cuda_runtime.h>

namespace fosphor_b200 {

constexpr int ACC_COLS = 32;     /* columns per tile == warp width */
constexpr int ACC_WARPS = 8;
constexpr int ACC_THREADS = ACC_WARPS * 32;
constexpr int REF_ROWS = 16;     /* display.cl:206-207 "sum / get_local_size(1)" */

struct AccumArgs {
	const float *wf;        /* waterfall ring [W][N] (kernel 1 output)       */
	float *hist;            /* histogram state [K][N]                        */
	float2 *spectrum;       /* live[N] then max[N], display order            */
	unsigned *ghits;        /* [K][N] cross-CTA hit counts (zero between calls) */
	float *part_live;       /* [S][N] per-split partial live sums            */
	float *part_max;        /* [S][N] per-split partial maxima               */
	unsigned *tickets;      /* [N/32] arrival counters (zero between calls)  */
	const float *weights;   /* [B]   (1-alpha)^(B-1-s)                       */
	const float2 *lut;      /* [B+1] (d, e) per hit count                    */
	int n, n_bins, wf_mask, wf_pos;
	int batch, splits, rows_per_split;
	float hscale, hofs;     /* cl.c:1087-1088 */
	float alpha, live_carry;/* live_carry = (1-alpha)^B, display.cl:210      */
	float mh_keep, mh_mix;  /* display.cl:303 */
};

__device__ __forceinline__ int map_bin(float x, int kmax)
{
	/* display.cl:161-165: (int)round(x) half away from zero, clamped to
	 * [0, K-1].  NaN and -inf -> 0, +inf -> K-1 (what the reference yields on
	 * the NVIDIA OpenCL runtime; fixed as the rule in DESIGN.md). */
	if (!(x > 0.0f))
		return 0;
	if (x >= (float)kmax)
		return kmax;
	const float fl = floorf(x);
	return (int)fl + ((x - fl) >= 0.5f ? 1 : 0);   /* <= kmax since x < kmax */
}

__global__ void __launch_bounds__(ACC_THREADS)
accumulate_kernel(const AccumArgs a)
{
	extern __shared__ unsigned sh_hits[];           /* [K][32] */
	__shared__ float sh_live[ACC_WARPS][32];
	__shared__ float sh_max[ACC_WARPS][32];
	__shared__ unsigned sh_ticket;

	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int tile = blockIdx.x, split = blockIdx.y;
	const int col = tile * ACC_COLS + lane;
	const int K = a.n_bins, N = a.n;

	for (int i = threadIdx.x; i < K * 32; i += ACC_THREADS)
		sh_hits[i] = 0;
	__syncthreads();

	/* ---- count: rows [row0, row1) of this call ---- */
	const int row0 = split * a.rows_per_split;
	const int row1 = min(a.batch, row0 + a.rows_per_split);
	float live = 0.0f, mx = -1000.0f;               /* display.cl:91,113 */
	const int kmax = K - 1;

#pragma unroll 4
	for (int s = row0 + warp; s < row1; s += ACC_WARPS) {
		const float pwr = __ldcg(&a.wf[(size_t)((a.wf_pos + s) & a.wf_mask) * N + col]);
		live = fmaf(pwr, __ldg(&a.weights[s]), live);       /* :149-150 */
		mx = fmaxf(mx, pwr);                                /* :139 */
		const int bin = map_bin(__fmul_rn(a.hscale, __fadd_rn(pwr, a.hofs)), kmax);
		atomicAdd(&sh_hits[bin * 32 + lane], 1u);           /* :170-177 */
	}
	sh_live[warp][lane] = live;
	sh_max[warp][lane] = mx;
	__syncthreads();

	if (warp == 0) {
		float sum = 0.0f, m = -1000.0f;
#pragma unroll
		for (int w = 0; w < ACC_WARPS; w++) {
			sum += sh_live[w][lane];
			m = fmaxf(m, sh_max[w][lane]);
		}
		sh_live[0][lane] = sum;
		sh_max[0][lane] = m;
		if (a.splits > 1) {
			a.part_live[(size_t)split * N + col] = sum;
			a.part_max[(size_t)split * N + col] = m;
		}
	}

	bool last = true;
	if (a.splits > 1) {
		/* flush non-zero counts once, then take a ticket */
		for (int bin = warp; bin < K; bin += ACC_WARPS) {
			const unsigned c = sh_hits[bin * 32 + lane];
			if (c)
				atomicAdd(&a.ghits[(size_t)bin * N + col], c);
		}
		__threadfence();
		__syncthreads();
		if (threadIdx.x == 0)
			sh_ticket = atomicAdd(&a.tickets[tile], 1u);
		__syncthreads();
		last = (sh_ticket == (unsigned)(a.splits - 1));
		if (!last)
			return;
		__threadfence();
		if (threadIdx.x == 0)
			a.tickets[tile] = 0;                /* ready for the next call */
	} else {
		__syncthreads();
	}

	/* ---- update (one CTA per tile gets here) ---- */
	for (int bin = warp; bin < K; bin += ACC_WARPS) {
		const size_t idx = (size_t)bin * N + col;
		unsigned hc;
		if (a.splits > 1) {
			hc = __ldcg(&a.ghits[idx]);
			if (hc)
				a.ghits[idx] = 0;
		} else {
			hc = sh_hits[bin * 32 + lane];
		}
		float hv = a.hist[idx];
		if (hv <= 0.01f && hc == 0)                     /* display.cl:237-238 */
			continue;
		const float2 de = __ldg(&a.lut[hc]);
		hv = __fadd_rn(__fmul_rn(__fsub_rn(hv, de.x), de.y), de.x);   /* :247 */
		hv = fminf(fmaxf(hv, 0.0f), 1.0f);                            /* :250 */
		a.hist[idx] = hv;
	}

	if (warp == 0) {
		float sum, bmax;
		if (a.splits > 1) {
			sum = 0.0f;
			bmax = -1000.0f;
			for (int sp = 0; sp < a.splits; sp++) {
				sum += __ldcg(&a.part_live[(size_t)sp * N + col]);
				bmax = fmaxf(bmax, __ldcg(&a.part_max[(size_t)sp * N + col]));
			}
		} else {
			sum = sh_live[0][lane];
			bmax = sh_max[0][lane];
		}

		const int half = N >> 1;
		const int i = col ^ half;                                 /* display.cl:201 */
		const float xpos = ((float)i / (float)half) - 1.0f;       /* :209 */

		/* live spectrum, display.cl:203-214 */
		float y = a.spectrum[i].y;
		if (!isfinite(y))
			y = sum / (float)REF_ROWS;
		y = __fadd_rn(__fmul_rn(y, a.live_carry), __fmul_rn(sum, a.alpha));
		a.spectrum[i] = make_float2(xpos, y);

		/* max hold with decay, display.cl:287-309 */
		float m = a.spectrum[N + i].y;
		if (!isfinite(m))
			m = -FLT_MAX;
		m = __fadd_rn(__fmul_rn(m, a.mh_keep), __fmul_rn(a.mh_mix, y));
		m = fmaxf(m, bmax);
		a.spectrum[N + i] = make_float2(xpos, m);
	}
}

/* first-use state, cl.c:406-465 */
__global__ void fill_kernel(float *p, size_t n, float v)
{
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
	     i += (size_t)gridDim.x * blockDim.x)
		p[i] = v;
}

/* max-hold trace (y values, display order) for the multi-GPU reduce */
__global__ void export_maxhold_kernel(const float2 *spectrum, int n, float *out)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n)
		out[i] = spectrum[n + i].y;
}

} /* namespace fosphor_b200 */
