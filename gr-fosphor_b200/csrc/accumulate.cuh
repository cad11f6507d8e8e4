/*
 * accumulate.cuh - kernel 2 of the hot path: fold the log-power rows of one or
 * more calls into the persistence state.
 *
 * Two implementations of the same arithmetic:
 *   accumulate_fused_kernel (second half of this file) - the default: ONE launch,
 *      state tile resident in shared memory, warp-specialised counters / updaters;
 *   count_kernel / count_tma_kernel + update_kernel (first half) - the any-shape
 *      fallback (batches that are not multiples of 16 rows, 2B < K), hit counts
 *      through HBM as u16.
 *
 * Replaces the second half of the reference's display program:
 *   lib/fosphor/display.cl:149-150,186-214  live spectrum (weighted IIR)
 *   lib/fosphor/display.cl:160-178          bin mapping + hit counting
 *   lib/fosphor/display.cl:217-254          histogram rise / decay
 *   lib/fosphor/display.cl:257-310          max hold with decay
 * The reference runs all of this on a fixed 64 work-groups that each loop over
 * the whole batch and then over all bins (cl.c:945-948).  Here the work is cut
 * along its real dependencies (described for the split kernels; the fused
 * kernel keeps the same cut inside one CTA):
 *
 *  count_kernel  (embarrassingly parallel over tiles x slices)
 *      A slice is a run of rows of ONE call; a tile is 32 columns.  Each CTA
 *      counts its rows into a shared-memory tile hits[bin][lane] (lane ==
 *      column: a warp never has a bank conflict, the 8 warps of the CTA meet
 *      only through shared atomics), then stores the whole tile once, as u16,
 *      to cnt[slice][bin][col], plus the slice's partial live sum and maximum
 *      per column.  Nothing is read-modify-written in global memory, so there
 *      is nothing to clear between calls.
 *
 *  update_kernel + update_columns_kernel (parallel over the K x N state cells /
 *  the N columns)
 *      The only sequential dependence of the whole path is along the CALL axis
 *      inside one cell:  hv <- (hv - d) e + d  with (d, e) a function of that
 *      call's hit count.  One thread owns one cell (or one column for
 *      live / max-hold), sums the slices of each call and walks the calls in
 *      order.  Many calls are therefore folded by ONE launch with exactly the
 *      per-call semantics of display.cl:241-247,303.
 *
 * The rise/decay closed form depends only on (hit count, B):
 *   a = hc/B; b = a/t0r; c = b + 1/t0d; d = b/c; e = (1-c)^B; hv' = (hv-d)e+d
 * so the host tabulates (d, e) for hc = 0..B and the live weights
 * (1-alpha)^(B-1-s) once per distinct B (engine.cu: BatchTables).
 */
#pragma once
#include <cfloat>
#include <cuda.h>            /* CUtensorMap (types only; the encoder is fetched at run time) */
#include <cuda_runtime.h>

namespace fosphor_b200 {

constexpr int ACC_COLS = 32;     /* columns per tile == warp width */
constexpr int ACC_WARPS = 8;
constexpr int ACC_THREADS = ACC_WARPS * 32;
constexpr int UPD_THREADS = 256;
constexpr int REF_ROWS = 16;     /* display.cl:206-207 "sum / get_local_size(1)" */

struct AccumArgs {
	const float *wf;          /* waterfall ring [W][N] (kernel 1 output)        */
	float *hist;              /* histogram state [K][N]                         */
	float2 *spectrum;         /* live[N] then max[N], display order             */
	unsigned short *cnt;      /* [slices][K][N] hit counts of this chunk        */
	float *part_live;         /* [calls][row blocks per call][N] partial live sums */
	float *part_max;          /* [calls][row blocks per call][N] partial maxima    */
	const float *weights;     /* [B]   (1-alpha)^(B-1-s)                        */
	const float2 *lut;        /* [B+1] (d, e) per hit count                     */
	int n, n_bins, wf_mask, wf_pos;   /* wf_pos: ring row of the chunk's first spectrum */
	int batch;                /* B, spectra per call                            */
	int n_calls;              /* calls in this chunk                            */
	int splits;               /* slices per call                                */
	int rows_per_split;
	float hscale, hofs;       /* cl.c:1087-1088 */
	float alpha, live_carry;  /* live_carry = (1-alpha)^B, display.cl:210       */
	float mh_keep, mh_mix;    /* display.cl:303 */
	float rho_rg;             /* fused kernel: (1-alpha)^(rows per warp step) */
	int depth_log2;           /* fused kernel: log2 of the boxes in the stage ring */
	int lut_staged;           /* update_kernel: the (d, e) table fits its shared memory (else read from global) */
	int l2_hints;             /* fused kernel: rows are read for the last time (L2 evict-first) */
};

/* display.cl:161-165: bin = (int)round(histo_scale * (pwr + histo_ofs)), round
 * half away from zero, clamped to [0, K-1].  Returned as the BYTE OFFSET of the
 * bin row in the hits[K][32] u32 tile (bin * 128).
 *   round_half_away(x) = (floor(2x) + 1) >> 1        for x >= 0 (exact, ties included)
 * and 2x = (2*hscale) * (pwr + hofs) is the very same rounding as x (scaling by
 * two is exact), so one float->int conversion replaces rint + tie fix-up.
 * Negative x land on bin 0 through the lower clamp; the conversion saturates:
 * NaN -> 0, -inf -> bin 0, +inf -> bin K-1, which is what the reference yields
 * on the NVIDIA OpenCL runtime (golden case zeros_then_data) and is fixed as
 * the rule in DESIGN.md. */
__device__ __forceinline__ unsigned bin_row_offset(float pwr, float hofs, float hscale2, int kmax2)
{
	const int i = __float2int_rd(__fmul_rn(hscale2, __fadd_rn(pwr, hofs)));
	const int j = min(max(i, -1), kmax2) + 1;        /* 0 .. 2*kmax + 1 */
	return ((unsigned)j & ~1u) << 6;                 /* (j >> 1) * 128 */
}

/* Write the CTA's hit tile hits[K][32] (u32 in shared memory) to
 * cnt[slice][K][N] as u16: a half-warp packs one bin row (16 x 2 columns) into
 * 32-bit words, so each store instruction writes two bin rows of 64 bytes. */
__device__ __forceinline__ void store_tile(const AccumArgs &a, const unsigned *sh_hits, int slice, int tile)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int K = a.n_bins, N = a.n;
	const int sub = lane >> 4, pair = lane & 15;
	unsigned *dst = reinterpret_cast<unsigned *>(a.cnt + (size_t)slice * K * N + tile * ACC_COLS) + pair;
	const size_t row_words = (size_t)N / 2;
#pragma unroll 4
	for (int bin = 2 * warp + sub; bin < K; bin += 2 * ACC_WARPS) {
		const uint2 h = *reinterpret_cast<const uint2 *>(sh_hits + bin * 32 + 2 * pair);
		dst[(size_t)bin * row_words] = (h.x & 0xffffu) | (h.y << 16);
	}
}

constexpr int ROWBLOCK = 128;     /* canonical unit of the f32 live partial sums */
constexpr int BLK_GROUP = 8;      /* row blocks reduced per pass through shared memory */

constexpr int WARP_ROWS = ROWBLOCK / ACC_WARPS;   /* 16 consecutive rows of a block per warp */

struct RowCtx {
	const float *base;        /* wf + column */
	const float *weights;
	unsigned *my_hits;        /* sh_hits + lane */
	unsigned ring0, mask, n;
	float hscale2, hofs;      /* 2 * histo_scale, histo_ofs */
	int kmax2;                /* 2 * (K - 1) */
};

/* U consecutive rows s .. s+U-1 of one column per lane: all loads first, then
 * the arithmetic; no bounds checks inside. */
template <int U>
__device__ __forceinline__ void count_rows(const RowCtx &c, int s, float &live, float &mx)
{
	float pw[U], wt[U];
#pragma unroll
	for (int u = 0; u < U; u++) {
		const unsigned r = (c.ring0 + (unsigned)(s + u)) & c.mask;
		pw[u] = __ldcg(c.base + (size_t)(r * c.n));
		wt[u] = __ldg(c.weights + s + u);
	}
#pragma unroll
	for (int u = 0; u < U; u++) {
		live = fmaf(pw[u], wt[u], live);                                  /* display.cl:149-150 */
		mx = fmaxf(mx, pw[u]);                                            /* :139 */
		const unsigned off = bin_row_offset(pw[u], c.hofs, c.hscale2, c.kmax2);   /* :161-165 */
		atomicAdd(c.my_hits + (off >> 2), 1u);                            /* :170-177 */
	}
}

/* The f32 live-spectrum partial sums are formed per ROWBLOCK rows in a fixed
 * order (warp w takes rows 16w .. 16w+15 of the block in row order; the 8 warp
 * partials are added in warp order; update_kernel adds the blocks in row
 * order), so the result does not depend on how a call is cut into slices nor
 * on which of the two count kernels ran. */
__global__ void __launch_bounds__(ACC_THREADS)
count_kernel(const AccumArgs a)
{
	extern __shared__ unsigned sh_hits[];           /* [K][32] */
	__shared__ float sh_live[BLK_GROUP][ACC_WARPS][32];
	__shared__ float sh_max[BLK_GROUP][ACC_WARPS][32];

	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int tile = blockIdx.x, slice = blockIdx.y;
	const int call = slice / a.splits, split = slice - call * a.splits;
	const int col = tile * ACC_COLS + lane;
	const int K = a.n_bins, N = a.n;

	{
		uint4 *z = reinterpret_cast<uint4 *>(sh_hits);
		for (int i = threadIdx.x; i < K * 8; i += ACC_THREADS)
			z[i] = make_uint4(0u, 0u, 0u, 0u);
	}
	__syncthreads();

	/* rows [row0, row1) of this call; rows_per_split is a multiple of ROWBLOCK */
	const int row0 = split * a.rows_per_split;
	const int row1 = min(a.batch, row0 + a.rows_per_split);
	RowCtx c;
	c.base = a.wf + col;
	c.weights = a.weights;
	c.my_hits = sh_hits + lane;
	c.ring0 = (unsigned)(a.wf_pos + call * a.batch);
	c.mask = (unsigned)a.wf_mask;
	c.n = (unsigned)N;
	c.hscale2 = 2.0f * a.hscale;
	c.hofs = a.hofs;
	c.kmax2 = 2 * (K - 1);
	const int blocks_per_call = (a.batch + ROWBLOCK - 1) / ROWBLOCK;
	const size_t part_base = (size_t)call * blocks_per_call;

	for (int g0 = row0; g0 < row1; g0 += ROWBLOCK * BLK_GROUP) {
		const int g1 = min(row1, g0 + ROWBLOCK * BLK_GROUP);
		int nblk = 0;
		for (int b0 = g0; b0 < g1; b0 += ROWBLOCK, nblk++) {
			const int b1 = min(g1, b0 + ROWBLOCK);
			float live = 0.0f, mx = -1000.0f;       /* display.cl:91,113 */
			int s = b0 + warp * WARP_ROWS;
			const int e = min(b1, s + WARP_ROWS);
			for (; s + 8 <= e; s += 8)
				count_rows<8>(c, s, live, mx);
			for (; s + 2 <= e; s += 2)
				count_rows<2>(c, s, live, mx);
			if (s < e)
				count_rows<1>(c, s, live, mx);
			sh_live[nblk][warp][lane] = live;
			sh_max[nblk][warp][lane] = mx;
		}
		__syncthreads();
		if (warp < nblk) {
			float sum = 0.0f, m = -1000.0f;
#pragma unroll
			for (int w = 0; w < ACC_WARPS; w++) {
				sum += sh_live[warp][w][lane];
				m = fmaxf(m, sh_max[warp][w][lane]);
			}
			const size_t o = (part_base + (size_t)(g0 / ROWBLOCK + warp)) * N + col;
			a.part_live[o] = sum;
			a.part_max[o] = m;
		}
		__syncthreads();
	}

	store_tile(a, sh_hits, slice, tile);
}

/* ---- TMA-staged variant of count_kernel ----------------------------------- */
/*
 * The plain kernel above is latency bound (ncu: ~16 long-scoreboard stall cycles
 * per issue; the compiler sinks the row loads next to their uses and a warp ends
 * up with 3-4 loads in flight).  Here every warp runs its own little pipeline:
 * its 16 rows x 32 columns of a row block are exactly one 2-D tensor-map box
 * (2 KB), which lane 0 requests from the TMA engine TMA_DEPTH blocks ahead into
 * a private shared-memory ring, completion on the warp's own mbarriers.  The
 * rows are then counted from shared memory (lane == column, conflict free).
 * No block-wide barrier inside the loop, no global load instruction except the
 * 16 weights per block.  Same arithmetic and the same canonical summation order
 * as count_kernel: bit-identical results.  Needs ring position and batch to be
 * multiples of 16 rows.
 */
constexpr int TMA_ROWS = WARP_ROWS;   /* rows per tensor-map box == rows per warp per block */
constexpr int TMA_DEPTH = 2;          /* boxes per warp: one being counted, one in flight */
static_assert(TMA_ROWS == 16, "box is 16 rows x 32 columns");

__device__ __forceinline__ unsigned cnt_smem_u32(const void *p)
{
	return (unsigned)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap *tmap, int c0, int c1, unsigned bar)
{
	asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
	             ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

__device__ __forceinline__ void tma_load_2d_hint(unsigned dst, const CUtensorMap *tmap, int c0, int c1, unsigned bar,
                                                 unsigned long long pol)
{
	asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
	             ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(bar), "l"(pol) : "memory");
}

constexpr int CNT_MAXBLK = 4;         /* row blocks whose partials are kept before a flush */

struct __align__(128) CountStage {
	float rows[ACC_WARPS][TMA_DEPTH][TMA_ROWS][32];      /* 8 x 4 x 2 KB, TMA destinations */
	unsigned long long bar[ACC_WARPS][TMA_DEPTH];
	float live[CNT_MAXBLK][ACC_WARPS][32];
	float mx[CNT_MAXBLK][ACC_WARPS][32];
};

__global__ void __launch_bounds__(ACC_THREADS)
count_tma_kernel(const AccumArgs a, const __grid_constant__ CUtensorMap tmap)
{
	extern __shared__ __align__(128) unsigned char cnt_smem[];
	CountStage &st = *reinterpret_cast<CountStage *>(cnt_smem);
	unsigned *sh_hits = reinterpret_cast<unsigned *>(cnt_smem + sizeof(CountStage));   /* [K][32] */

	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int tile = blockIdx.x, slice = blockIdx.y;
	const int call = slice / a.splits, split = slice - call * a.splits;
	const int col = tile * ACC_COLS + lane;
	const int K = a.n_bins, N = a.n;

	const int row0 = split * a.rows_per_split;
	const int row1 = min(a.batch, row0 + a.rows_per_split);
	const int nblocks = (row1 - row0 + ROWBLOCK - 1) / ROWBLOCK;
	const unsigned ring0 = (unsigned)(a.wf_pos + call * a.batch);
	const unsigned mask = (unsigned)a.wf_mask;
	const unsigned bar0 = cnt_smem_u32(&st.bar[warp][0]);
	const unsigned stage0 = cnt_smem_u32(&st.rows[warp][0][0][0]);
	constexpr unsigned BOX_BYTES = TMA_ROWS * 32 * sizeof(float);

	/* the box of this warp in block blk starts at row row0 + blk*128 + 16*warp */
	auto my_rows = [&](int blk) { return row0 + blk * ROWBLOCK + warp * WARP_ROWS; };
	auto issue = [&](int blk) {          /* lane 0 only; box present iff its first row exists */
		const int s = my_rows(blk);
		if (s < row1) {
			const unsigned slot = (unsigned)(blk % TMA_DEPTH);
			asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
			             ::"r"(bar0 + 8u * slot), "r"(BOX_BYTES) : "memory");
			tma_load_2d(stage0 + slot * BOX_BYTES, &tmap, tile * ACC_COLS,
			            (int)((ring0 + (unsigned)s) & mask), bar0 + 8u * slot);
		}
	};

	if (lane == 0) {
#pragma unroll
		for (int d = 0; d < TMA_DEPTH; d++)
			asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8u * d));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
		for (int d = 0; d < TMA_DEPTH && d < nblocks; d++)
			issue(d);
	}
	{
		uint4 *z = reinterpret_cast<uint4 *>(sh_hits);
		for (int i = threadIdx.x; i < K * 8; i += ACC_THREADS)
			z[i] = make_uint4(0u, 0u, 0u, 0u);
	}
	__syncthreads();

	const int kmax2 = 2 * (K - 1);
	const float hscale2 = 2.0f * a.hscale;
	const unsigned hits_addr = cnt_smem_u32(sh_hits + lane);     /* + bin * 128 bytes */
	const int blocks_per_call = (a.batch + ROWBLOCK - 1) / ROWBLOCK;
	const size_t part_base = (size_t)call * blocks_per_call + (size_t)(row0 / ROWBLOCK);
	unsigned phases = 0u;                    /* bit d = parity to wait for on slot d */

	for (int g0 = 0; g0 < nblocks; g0 += CNT_MAXBLK) {
		const int g1 = min(nblocks, g0 + CNT_MAXBLK);
		for (int blk = g0; blk < g1; blk++) {
			const int s = my_rows(blk);
			float live = 0.0f, mx = -1000.0f;               /* display.cl:91,113 */
			if (s < row1) {
				const unsigned slot = (unsigned)(blk % TMA_DEPTH);
				const float wts = __ldg(&a.weights[s + (lane & (WARP_ROWS - 1))]);
				unsigned ok;
				do {
					asm volatile("{\n\t.reg .pred p;\n\t"
					             "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
					             "selp.u32 %0, 1, 0, p;\n\t}"
					             : "=r"(ok) : "r"(bar0 + 8u * slot), "r"((phases >> slot) & 1u) : "memory");
				} while (!ok);
				phases ^= 1u << slot;
				const float *rp = &st.rows[warp][slot][0][lane];
				/* all 16 rows into registers first: the shared-memory increments below
				 * may not be reordered against shared loads, and 16 independent
				 * dependency chains are what keeps the issue slots busy */
				float pw[WARP_ROWS];
#pragma unroll
				for (int r = 0; r < WARP_ROWS; r++)
					pw[r] = rp[r * 32];
#pragma unroll
				for (int r = 0; r < WARP_ROWS; r++) {
					live = fmaf(pw[r], __shfl_sync(0xffffffffu, wts, r), live);   /* display.cl:149-150 */
					mx = fmaxf(mx, pw[r]);                                        /* :139 */
					const unsigned off = bin_row_offset(pw[r], a.hofs, hscale2, kmax2);                /* :161-165 */
					asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(hits_addr + off) : "memory");     /* :170-177 */
				}
				/* Every lane has now USED its 16 values, so all shared loads of the slot
				 * have completed; only then may the TMA engine overwrite it.  (Issuing
				 * right after the loads were merely issued raced: the variant-equality
				 * test caught a histogram mismatch once in a few runs.) */
				__syncwarp();
				if (lane == 0 && blk + TMA_DEPTH < nblocks)
					issue(blk + TMA_DEPTH);
			}
			st.live[blk - g0][warp][lane] = live;
			st.mx[blk - g0][warp][lane] = mx;
		}
		__syncthreads();
		if (warp < g1 - g0) {
			float sum = 0.0f, m = -1000.0f;
#pragma unroll
			for (int w = 0; w < ACC_WARPS; w++) {
				sum += st.live[warp][w][lane];
				m = fmaxf(m, st.mx[warp][w][lane]);
			}
			const size_t o = (part_base + (size_t)(g0 + warp)) * N + col;
			a.part_live[o] = sum;
			a.part_max[o] = m;
		}
		if (g1 < nblocks)
			__syncthreads();
	}

	store_tile(a, sh_hits, slice, tile);
}

constexpr int UPD_CELLS = 4;      /* adjacent cells per thread (one 64-bit load of four u16 counts) */
constexpr int UPD_SLICES = 16;    /* slices whose loads are issued together */
constexpr int UPD_COLS = 32;      /* columns per live/max-hold block */
constexpr int UPD_PARTS = 64;     /* live/max partials (call x row block) staged per pass */

/* display.cl:237-250 without a branch: cells with hv <= 0.01 and no hit keep
 * their value exactly (the reference skips the write), everything else takes
 * hv <- clamp((hv - d) e + d) with (d, e) from the per-batch table. */
__device__ __forceinline__ float rise_decay(float hv, unsigned hc, const float2 *lut)
{
	const float2 de = lut[hc];
	const float nv = fminf(fmaxf(__fadd_rn(__fmul_rn(__fsub_rn(hv, de.x), de.y), de.x), 0.0f), 1.0f);
	return (hv <= 0.01f && hc == 0) ? hv : nv;
}

/* update_kernel: one thread per UPD_CELLS adjacent histogram cells (bin-major:
 * a warp covers 128 consecutive columns of one bin; N is a multiple of 4, so a
 * cell group never straddles two bins). */
__device__ __forceinline__ void update_cells(const AccumArgs &a, int block, float2 *sh_lut)
{
	const int K = a.n_bins, N = a.n;
	const size_t KN = (size_t)K * N;
	(void)N;

	{
		size_t cell = ((size_t)block * UPD_THREADS + threadIdx.x) * UPD_CELLS;
		const bool live_thread = cell < KN;
		if (!live_thread)
			cell = 0;
		/* state + first group of counts are requested before the LUT is staged */
		const float4 hv0 = *reinterpret_cast<const float4 *>(a.hist + cell);
		const int total = a.n_calls * a.splits;
		const uint2 *cnt = reinterpret_cast<const uint2 *>(a.cnt + cell);          /* 4 x u16 */
		const size_t stride = KN / 4;                                              /* in uint2 */
		uint2 w[UPD_SLICES];
#pragma unroll
		for (int u = 0; u < UPD_SLICES; u++)
			w[u] = (u < total) ? __ldcg(cnt + (size_t)u * stride) : make_uint2(0u, 0u);
		if (a.lut_staged) {
			for (int i = threadIdx.x; i <= a.batch; i += UPD_THREADS)
				sh_lut[i] = __ldg(&a.lut[i]);
			__syncthreads();
		}
		const float2 *lut = a.lut_staged ? sh_lut : a.lut;   /* batches beyond ~8 K rows: table too big to stage */
		if (!live_thread)
			return;
		float4 hv = hv0;
		unsigned hc0 = 0, hc1 = 0, hc2 = 0, hc3 = 0;
		int in_call = 0;
		const uint2 *p = cnt + (size_t)UPD_SLICES * stride;
		for (int j0 = 0; j0 < total; j0 += UPD_SLICES) {
			if (j0 > 0) {
#pragma unroll
				for (int u = 0; u < UPD_SLICES; u++) {
					w[u] = (j0 + u < total) ? __ldcg(p) : make_uint2(0u, 0u);
					p += stride;
				}
			}
#pragma unroll
			for (int u = 0; u < UPD_SLICES; u++) {
				if (j0 + u < total) {                   /* warp-uniform */
					hc0 += w[u].x & 0xffffu;
					hc1 += w[u].x >> 16;
					hc2 += w[u].y & 0xffffu;
					hc3 += w[u].y >> 16;
					if (++in_call == a.splits) {        /* all slices of this call summed */
						/* nothing to do for a warp whose cells are all empty and unhit
						 * (most of the plane): display.cl:237-238 */
						const bool idle = (hc0 | hc1 | hc2 | hc3) == 0 &&
						                  fmaxf(fmaxf(hv.x, hv.y), fmaxf(hv.z, hv.w)) <= 0.01f;
						if (!__all_sync(0xffffffffu, idle)) {
							hv.x = rise_decay(hv.x, hc0, lut);
							hv.y = rise_decay(hv.y, hc1, lut);
							hv.z = rise_decay(hv.z, hc2, lut);
							hv.w = rise_decay(hv.w, hc3, lut);
						}
						hc0 = hc1 = hc2 = hc3 = 0;
						in_call = 0;
					}
				}
			}
		}
		if (hv.x != hv0.x || hv.y != hv0.y || hv.z != hv0.z || hv.w != hv0.w)
			*reinterpret_cast<float4 *>(a.hist + cell) = hv;
	}
}

/* live IIR and max-hold.  A block owns UPD_COLS columns; the partials of the
 * chunk are fetched in parallel into shared memory (independent loads, `cap`
 * partials per pass), then one thread per column runs the serial recurrences
 * over the calls of the pass. */
__device__ __forceinline__ void update_columns(const AccumArgs &a, int block, int cap, float *sh_part)
{
	const int N = a.n;
	const int col0 = block * UPD_COLS;
	const int blocks_per_call = (a.batch + ROWBLOCK - 1) / ROWBLOCK;
	const int calls_per_group = cap / blocks_per_call;    /* >= 1: cap >= blocks_per_call (host) */
	const int col = col0 + threadIdx.x;
	const bool owner = threadIdx.x < UPD_COLS && col < N;
	const int half = N >> 1;
	const int i = (owner ? col : 0) ^ half;                   /* display.cl:201 */
	const float xpos = ((float)i / (float)half) - 1.0f;       /* :209 */
	float y = 0.0f, m = 0.0f;
	if (owner) {
		y = a.spectrum[i].y;
		m = a.spectrum[N + i].y;
	}
	for (int c0 = 0; c0 < a.n_calls; c0 += calls_per_group) {
		const int nc = min(calls_per_group, a.n_calls - c0);
		const int nparts = nc * blocks_per_call;
		const size_t pbase = (size_t)c0 * blocks_per_call;
		__syncthreads();                              /* previous group consumed */
		{
			/* thread (p0, c): partials p0, p0 + PSTEP, ... of column c; 2 x 8 loads in flight */
			const int c = threadIdx.x % UPD_COLS, p0 = threadIdx.x / UPD_COLS;
			constexpr int PSTEP = UPD_THREADS / UPD_COLS;
			const bool okc = col0 + c < N;
			const float *gl = a.part_live + pbase * N + col0 + c;
			const float *gm = a.part_max + pbase * N + col0 + c;
			for (int q0 = p0; q0 < nparts; q0 += 8 * PSTEP) {
				float vl[8], vm[8];
#pragma unroll
				for (int u = 0; u < 8; u++) {
					const int q = q0 + u * PSTEP;
					const bool ok = okc && q < nparts;
					vl[u] = ok ? __ldcg(gl + (size_t)q * N) : 0.0f;
					vm[u] = ok ? __ldcg(gm + (size_t)q * N) : -1000.0f;
				}
#pragma unroll
				for (int u = 0; u < 8; u++) {
					const int q = q0 + u * PSTEP;
					if (q < nparts) {
						sh_part[q * UPD_COLS + c] = vl[u];
						sh_part[(cap + q) * UPD_COLS + c] = vm[u];
					}
				}
			}
		}
		__syncthreads();
		if (owner) {
			const float *pl = sh_part + threadIdx.x;
			const float *pm = sh_part + cap * UPD_COLS + threadIdx.x;
			for (int c = 0; c < nc; c++) {
				float sum = 0.0f, bmax = -1000.0f;
				for (int b = 0; b < blocks_per_call; b++) {
					sum += pl[(c * blocks_per_call + b) * UPD_COLS];
					bmax = fmaxf(bmax, pm[(c * blocks_per_call + b) * UPD_COLS]);
				}
				/* live spectrum, display.cl:203-214 */
				if (!isfinite(y))
					y = sum / (float)REF_ROWS;
				y = __fadd_rn(__fmul_rn(y, a.live_carry), __fmul_rn(sum, a.alpha));
				/* max hold with decay, display.cl:287-309 */
				if (!isfinite(m))
					m = -FLT_MAX;
				m = __fadd_rn(__fmul_rn(m, a.mh_keep), __fmul_rn(a.mh_mix, y));
				m = fmaxf(m, bmax);
			}
		}
	}
	if (owner && a.n_calls > 0) {
		a.spectrum[i] = make_float2(xpos, y);
		a.spectrum[N + i] = make_float2(xpos, m);
	}
}

/* blocks [0, cell_blocks): histogram cells; blocks beyond: UPD_COLS columns each
 * of live / max-hold.  The two roles are independent and share the launch. */
__global__ void __launch_bounds__(UPD_THREADS)
update_kernel(const AccumArgs a, int cell_blocks, int cap)
{
	extern __shared__ __align__(16) unsigned char upd_smem[];
	if ((int)blockIdx.x < cell_blocks)
		update_cells(a, (int)blockIdx.x, reinterpret_cast<float2 *>(upd_smem));
	else
		update_columns(a, (int)blockIdx.x - cell_blocks, cap, reinterpret_cast<float *>(upd_smem));
}

/* ------------------------------------------------------------------------ */
/* Fused accumulate kernel: count + rise/decay + live IIR + max-hold            */
/* ------------------------------------------------------------------------ */
constexpr int ACC_VW = 16;        /* virtual warps: the unit of the row -> lane assignment */
constexpr int ACC_STAGE_ROWS = 2048; /* rows of the tile staged ahead by the producer warp (full-size CTA) */
constexpr int ACC_STAGE_ROWS_SLIM = 1024; /* ... by the slim CTA that shares its SM with FFT CTAs */
constexpr int ACC_XW = 2;          /* extra warps: one for the live / max-hold columns, one TMA producer */
constexpr unsigned ACC_WAIT_HINT_NS = 20000u; /* mbarrier.try_wait suspend-time hint */
constexpr int ACC_LUT_MAX = 4096; /* batches up to this keep the (d, e) table in shared memory */

template <int COLS>
__device__ __forceinline__ unsigned bin_cell_addr(float pwr, float hofs, float hscale2, unsigned kmax2, unsigned tile_col)
{
	/* see bin_row_offset().  Shared-memory byte address of hits[bin][col], COLS u32 per bin, for a tile
	 * aligned to its size; tile_col = tile base | col * 4.  The UNSIGNED floor conversion saturates:
	 * t < 0 (floor = -1 -> bin 0), NaN and -inf give 0 -> bin 0, +inf gives UINT_MAX -> kmax2 ->
	 * bin K-1; then bin = (i + 1) >> 1, times COLS * 4 bytes, OR-ed into the tile address (one LOP3). */
	const unsigned i = min(__float2uint_rd(__fmul_rn(hscale2, __fadd_rn(pwr, hofs))), kmax2);
	return (((i + 1u) * (COLS * 2)) & ~(unsigned)(COLS * 4 - 1)) | tile_col;
}

/* rows per virtual warp and call: contiguous runs, at least 64 rows so that the per-run
 * overhead (barrier probes, pointer set-up, partial stores) is paid once per 16+ warp steps;
 * small batches therefore use fewer than ACC_VW virtual warps */
__host__ __device__ inline int acc_rows_per_vwarp(int batch)
{
	const int units = (batch + 15) / 16;
	const int rv = 16 * ((units + ACC_VW - 1) / ACC_VW);
	return rv < 64 ? 64 : rv;
}

__device__ __forceinline__ void mbar_wait_parity(unsigned bar, unsigned parity)
{
	/* suspend-time hint (ns): the warp sleeps in the barrier unit instead of re-issuing the probe.
	 * ncu of cfg3 (B = 256) without it: one third of all executed instructions were probes of
	 * counter warps waiting for the updaters, taken from the very warps they were waiting for. */
	unsigned ok;
	do {
		asm volatile("{\n\t.reg .pred p;\n\t"
		             "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
		             "selp.u32 %0, 1, 0, p;\n\t}"
		             : "=r"(ok) : "r"(bar), "r"(parity), "r"(ACC_WAIT_HINT_NS) : "memory");
	} while (!ok);
}

__device__ __forceinline__ void mbar_arrive(unsigned bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

/* A hit tile starts at a multiple of its (power of two) size, so that the byte address of
 * hits[bin][col] is  tile | bin_offset | col*4  - one LOP3 - instead of an addition. */
__host__ __device__ inline unsigned acc_tile_bytes(int K, int cols)
{
	unsigned a = 128;
	while (a < (unsigned)(K * cols) * 4u)
		a <<= 1;
	return a;
}

/* Stage-ring safety.  A warp waits for box n on full[n % depth] by PARITY, which only tells the
 * current phase of that barrier from the previous one: the wait is correct only if box n - depth
 * has already landed when the warp starts waiting - otherwise it passes on the older box (wrong
 * data, a spurious arrival on empty[], and in the end a producer that waits forever).  That is
 * guaranteed when
 *   - the launch has no more boxes than the ring, or
 *   - depth is a multiple of the boxes of one synchronisation group: box n and box n - depth then sit
 *     at the same position of their groups, are consumed by the same warps, in program order, or
 *   - the ring spans two groups: box n - depth belongs to group g-2 or older, which every counter
 *     warp has finished before any warp may start group g (hits_free barrier).
 * Batches with more boxes per call than the ring holds (B = 32768: 128 boxes, ring of 32) violate all
 * three - warps 4..15 start out waiting two and three phases ahead - and take the plain-load path. */
__host__ __device__ inline bool acc_ring_safe(long long depth, long long boxes_per_group, long long boxes_total)
{
	return boxes_total <= depth || depth % boxes_per_group == 0 || depth >= 2 * boxes_per_group;
}

template <int COLS, int FW, int UW, int BOXR, int GC = 1>
struct FusedCfg {
	static_assert(GC == 1 || GC == 2 || GC == 4, "calls per synchronisation group");
	static_assert(COLS == 4 || COLS == 8 || COLS == 16 || COLS == 32, "tile width");
	static_assert(BOXR == 16 || BOXR == 64 || BOXR == 256, "box rows");
	static_assert(ACC_VW % FW == 0, "counter warps divide the virtual ones");
	static constexpr int RG = 32 / COLS;                 /* rows per warp step */
	static constexpr int VPW = ACC_VW / FW;              /* virtual warps per counter warp */
	static constexpr int THREADS = (FW + UW + ACC_XW) * 32;   /* counters, cell updaters, column warp, producer */
	static constexpr bool SLIM = FW + UW < ACC_VW;           /* co-resident variant: <= 48 registers, half the stage */
	static constexpr int MIN_CTAS = SLIM ? 2 : 1;        /* launch bound only: keeps the register count under 73 (it is 48) */
	static constexpr size_t BOX_BYTES = sizeof(float) * BOXR * COLS;
	static constexpr int DEPTH_MAX = 8192 / BOXR;        /* boxes in the stage ring: a power of two chosen per launch */
	static constexpr size_t BAR_BYTES = 8 * (2 * DEPTH_MAX + 4) + 96;   /* full[], empty[], 4 role barriers; keeps 128 B alignment */
	static constexpr size_t PART_BYTES = sizeof(float) * 2 * GC * 2 * ACC_VW * 32;
	/* everything but the stage ring */
	static size_t smem_fixed(int K, int batch)
	{
		size_t bar = (BAR_BYTES + 127) & ~(size_t)127;
		/* state tile, 2*GC hit tiles at tile-size granularity + the worst-case alignment gap */
		return bar + PART_BYTES + sizeof(float) * (size_t)K * COLS + (size_t)acc_tile_bytes(K, COLS) * (2 * GC + 1) +
		       (batch <= ACC_LUT_MAX ? sizeof(float2) * (size_t)(batch + 1) : 0) + 128;
	}
	/* The stage ring is what keeps HBM busy: one CTA per SM, so bytes in flight per SM = the ring
	 * (64 KB gave 3.7 TB/s at cfg2: 2.6 us of latency per 32-byte row segment box).  Take what the
	 * SM has left, up to DEPTH_MAX boxes; the slim variant shares its SM and stays at 1024 rows. */
	static int depth_log2(int K, int batch, size_t smem_limit)
	{
		const size_t fixed = smem_fixed(K, batch);
		int d = 0;
		const int cap = SLIM ? (1024 / BOXR > 1 ? 1024 / BOXR : 1) : DEPTH_MAX;
		while ((2 << d) <= cap && fixed + BOX_BYTES * (size_t)(2 << d) <= smem_limit)
			d++;
		return d;
	}
	static size_t smem(int K, int batch, bool tma, int dlog)
	{
		return (tma ? BOX_BYTES << dlog : 0) + smem_fixed(K, batch);
	}
};

/* Fused accumulate kernel.
 *
 * One CTA owns COLS adjacent frequency columns for the whole launch and walks
 * the calls of the chunk in order, so the per-cell recurrence along the call
 * axis (display.cl:241-247) never leaves the SM: the COLS x K histogram cells
 * of the tile live in shared memory from the first call to the last, the hit
 * counts of a call are built and consumed in shared memory, and the only
 * global traffic is the log-power rows (read once, through the TMA engine) and
 * one read + one write of the tile's state per LAUNCH.
 *
 * Warp roles (no block barrier after the prologue; everything meets through
 * mbarriers):
 *   producer (1 warp)  streams the tile's rows - contiguous in the ring across
 *                      call boundaries - as BOXR x COLS tensor-map boxes into a
 *                      ring of DEPTH slots (full[] / empty[] barriers);
 *   counters (FW)      turn rows into hit counts (shared-memory atomics on
 *                      hits[bin][col]) and live / max partials;
 *   updaters (UW + 1)  apply rise/decay (UW cell warps) and the live / max-hold
 *                      recurrences (one column warp) of call c while the
 *                      counters are already on call c+1 (hit tile and partials
 *                      double buffered by call parity: cnt_done[par] counters ->
 *                      updaters, hits_free[par] back).  The update of a call is
 *                      a dependent chain that shares issue slots with the
 *                      counters: it must stay shorter than a counter's share of
 *                      the call, hence many short updater warps.
 *
 * Row -> lane assignment: a warp step is RG = 32/COLS consecutive rows x COLS
 * columns; the rows of a call are dealt to ACC_VW = 16 virtual warps in
 * contiguous runs of Rv = 16 * ceil(B/16 / 16) rows; counter warp w takes
 * virtual warps w, w+FW, ...  The live-spectrum sum of a lane is a Horner
 * recurrence over its rows (acc = acc * (1-alpha)^RG + pwr) scaled by the
 * table weight of its last row; the 16 virtual-warp partials are added in
 * order, the row groups by an xor butterfly - a fixed order that depends on
 * (B, COLS) only, not on FW, BOXR, SUBR, the load path or the launch folding.
 *
 * SUBR = rows handled by one unrolled body (loads of SUBR/RG steps issued
 * together).  TMA = false: same arithmetic with plain loads (rows past the
 * batch masked), for batches / ring positions that do not align to a box.
 */
template <int COLS, int FW, int UW, int BOXR, int SUBR, int LOAD, int GC = 1>
__global__ void __launch_bounds__((FW + UW + ACC_XW) * 32, (FusedCfg<COLS, FW, UW, BOXR, GC>::MIN_CTAS))
accumulate_fused_kernel(const AccumArgs a, const __grid_constant__ CUtensorMap tmap)
{
	using C = FusedCfg<COLS, FW, UW, BOXR, GC>;
	constexpr int RG = C::RG;
	const unsigned dlog = (unsigned)a.depth_log2, DEPTH = 1u << dlog;   /* boxes in the stage ring */
	constexpr int SSTEPS = SUBR / RG;                    /* steps per unrolled body */
	static_assert(SUBR == 16 || SUBR == 64, "sub-block rows");
	/* LOAD: 0 plain loads by the counters; 1 TMA tensor-map boxes (one producer lane).
	 * (cp.async loader warps were tried instead of the TMA producer: 15 % slower.) */
	static_assert(LOAD == 0 || LOAD == 1, "load path");
	constexpr bool TMA = LOAD != 0;                      /* rows staged in shared memory */
	extern __shared__ __align__(128) unsigned char fz_smem[];

	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int col0 = blockIdx.x * COLS;
	const int K = a.n_bins, N = a.n, B = a.batch;
	const int cells = K * COLS;
	const int Rv = acc_rows_per_vwarp(B);

	/* carve shared memory */
	unsigned char *sp = fz_smem;
	float *stage = reinterpret_cast<float *>(sp);
	if (TMA) sp += C::BOX_BYTES << dlog;
	unsigned long long *bars = reinterpret_cast<unsigned long long *>(sp);
	sp += (C::BAR_BYTES + 127) & ~(size_t)127;
	float *parts = reinterpret_cast<float *>(sp);        /* [2 parity][GC calls][2 live/max][ACC_VW][32] */
	sp += C::PART_BYTES;
	float *hist_s = reinterpret_cast<float *>(sp);       /* [K][COLS] */
	sp += sizeof(float) * (size_t)cells;
	float2 *lut_s = reinterpret_cast<float2 *>(sp);      /* [B+1] (d, e) table when B <= ACC_LUT_MAX */
	if (B <= ACC_LUT_MAX)
		sp += (sizeof(float2) * (size_t)(B + 1) + 15) & ~(size_t)15;
	/* hit tiles [2 parity][GC calls], each [K][COLS] u32 at a multiple of the tile size (a power of two) */
	const unsigned tile_bytes = acc_tile_bytes(K, COLS);
	const unsigned tile_words = tile_bytes / 4;
	sp += (tile_bytes - (cnt_smem_u32(sp) & (tile_bytes - 1))) & (tile_bytes - 1);
	unsigned *hits = reinterpret_cast<unsigned *>(sp);

	const unsigned full0 = cnt_smem_u32(bars);                   /* full[DEPTH]  */
	const unsigned empty0 = full0 + 8u * DEPTH;                  /* empty[DEPTH] */
	const unsigned role_bar = empty0 + 8u * DEPTH;               /* cnt_done[0,1], hits_free[0,1] */

	if (threadIdx.x == 0) {
		/* virtual warps that consume one box */
		const int sharers = BOXR > Rv ? BOXR / Rv : 1;
		if (TMA)
			for (unsigned i = 0; i < DEPTH; i++) {
				asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(full0 + 8u * i), "r"(1));
				asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(empty0 + 8u * i), "r"(sharers));
			}
		asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(role_bar), "r"(FW));
		asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(role_bar + 8), "r"(FW));
		asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(role_bar + 16), "r"(UW + 1));
		asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(role_bar + 24), "r"(UW + 1));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	}
	__syncthreads();                /* barriers are live: the producer may start filling the stage ring ... */
	/* ... while the counters clear the hit tiles and the updaters bring the tile state and the table
	 * in: the counters start on the first call without waiting for the state (ncu of cfg4, 16 calls
	 * per CTA: 10 % of the samples sat on the state load).  Everything the other role reads of these
	 * buffers is ordered by the cnt_done / hits_free barriers. */
	if (warp < FW) {
		uint4 *z4 = reinterpret_cast<uint4 *>(hits);
		for (int g = threadIdx.x; g < 2 * GC * (int)tile_words / 4; g += FW * 32)
			z4[g] = make_uint4(0u, 0u, 0u, 0u);
		/* partials of virtual warps that small batches never visit: 0 / -1000 (display.cl:91,113) */
		for (int i = threadIdx.x; i < 2 * GC * 2 * ACC_VW * 32; i += FW * 32)
			parts[i] = ((i / (ACC_VW * 32)) & 1) ? -1000.0f : 0.0f;
		asm volatile("bar.sync 1, %0;" ::"n"(FW * 32) : "memory");
	} else if (warp < FW + UW) {
		constexpr int cpr = COLS / 4;                /* float4 groups per bin row */
		float4 *h4 = reinterpret_cast<float4 *>(hist_s);
		const int ut = threadIdx.x - FW * 32;
		for (int g = ut; g < cells / 4; g += UW * 32) {
			const int bin = g / cpr, c4 = (g % cpr) * 4;
			h4[g] = *reinterpret_cast<const float4 *>(a.hist + (size_t)bin * N + col0 + c4);
		}
		if (B <= ACC_LUT_MAX)
			for (int i = ut; i <= B; i += UW * 32)
				lut_s[i] = __ldg(&a.lut[i]);
		asm volatile("bar.sync 2, %0;" ::"n"(UW * 32) : "memory");
	}

	if (warp < FW) {
		/* ================= counter warps ================= */
		const int r = lane / COLS, cc = lane % COLS;
		const unsigned mask = (unsigned)a.wf_mask;
		const unsigned kmax2 = 2u * (unsigned)(K - 1), cc4 = 4u * (unsigned)cc;
		const float hscale2 = 2.0f * a.hscale;
		const float hofs = a.hofs;
		const float rho = a.rho_rg;                          /* (1-alpha)^RG */
		const float *wcol = a.wf + col0 + cc;
		const int CH = TMA ? (Rv < BOXR ? Rv : BOXR) : SUBR; /* rows per chunk (TMA: one box or one run) */

		/* The (call, virtual warp) pairs of a group of GC calls are dealt round robin to the counter
		 * warps: pair p = ci * nv + v, warp w takes p = w, w + FW, ...  With B = 256 and GC = 4 every
		 * warp counts one 64-row run of one of the four calls between two barrier hand-overs. */
		const int nv = (B + Rv - 1) / Rv;                    /* virtual warps in use, <= ACC_VW */
		const int n_groups = (a.n_calls + GC - 1) / GC;
		/* nv divides FW (every case but 16 virtual warps on the 8 counter warps of the slim CTA):
		 * a warp always meets the same virtual warp, whose run is then loop invariant */
		const bool vfixed = (FW % nv) == 0;
		const int lo0 = (warp % nv) * Rv;
		const int rows0 = min(B - lo0, Rv);
		const float wtail0 = rows0 > r ? __ldg(&a.weights[lo0 + r + RG * ((rows0 - 1 - r) / RG)]) : 0.0f;
		for (int grp = 0; grp < n_groups; grp++) {
			const int par = grp & 1;
			const int ncg = min(GC, a.n_calls - grp * GC);   /* calls of this group */
			if (grp >= 2)               /* the updaters have consumed (and cleared) this parity's tiles */
				mbar_wait_parity(role_bar + 16 + 8 * par, (unsigned)((grp >> 1) - 1) & 1u);
#pragma unroll 1
			for (int p = warp; p < ncg * nv; p += FW) {
				const int ci = GC == 1 ? 0 : p / nv, v = p - ci * nv;
				const int call = grp * GC + ci;
				const int tile = par * GC + ci;
				/* my run: rows [lo, lo + rows) of the call, and the table weight of my last row */
				int lo = lo0, rows = rows0;
				float wtail = wtail0;
				if (!vfixed) {
					lo = v * Rv;
					rows = min(B - lo, Rv);
					wtail = rows > r ? __ldg(&a.weights[lo + r + RG * ((rows - 1 - r) / RG)]) : 0.0f;
				}
				/* the rows of one warp step can collide in a bank (~3 wavefronts per increment, ncu); private
				 * replicas per row were tried (2 and 4): what they save here they cost twice in the update */
				const unsigned hb = cnt_smem_u32(hits + tile * tile_words) | cc4;   /* tile | column: the bin offset is OR-ed in */
				float *pl = parts + (size_t)tile * 2 * ACC_VW * 32;
				float acc = 0.0f, mx = -1000.0f;             /* display.cl:91,113 */
#pragma unroll 1
				for (int c0 = 0; c0 < rows; c0 += CH) {
					if (TMA) {
						/* rows of the launch are contiguous in the ring: box n holds rows n*BOXR .. */
						const unsigned g = (unsigned)call * (unsigned)B + (unsigned)(lo + c0);
						const unsigned n = g / BOXR, slot = n & (DEPTH - 1u);
						mbar_wait_parity(full0 + 8u * slot, (n >> dlog) & 1u);
						const float *bp = stage + (size_t)slot * (BOXR * COLS) + (g % BOXR) * COLS + lane;
#pragma unroll 1
						for (int s0 = 0; s0 < CH; s0 += SUBR, bp += SUBR * COLS) {
							float pw[SSTEPS];
#pragma unroll
							for (int i = 0; i < SSTEPS; i++)
								pw[i] = bp[i * 32];
#pragma unroll
							for (int i = 0; i < SSTEPS; i++) {
								acc = fmaf(acc, rho, pw[i]);                      /* display.cl:149-150 (Horner) */
								mx = fmaxf(mx, pw[i]);                            /* :139 */
								const unsigned cell = bin_cell_addr<COLS>(pw[i], hofs, hscale2, kmax2, hb);   /* :161-165 */
								asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(cell) : "memory");       /* :170-177 */
							}
						}
						/* every lane has USED its values: hand the box back */
						__syncwarp();
						if (lane == 0)
							mbar_arrive(empty0 + 8u * slot);
					} else {
						const unsigned ring = (unsigned)a.wf_pos + (unsigned)call * (unsigned)B;
						const int s0 = lo + c0;
						const int lim = lo + rows;
						float pw[SSTEPS];
#pragma unroll
						for (int i = 0; i < SSTEPS; i++) {
							const int row = s0 + RG * i + r;
							pw[i] = row < lim ? __ldcg(wcol + (size_t)((ring + (unsigned)row) & mask) * N) : 0.0f;
						}
#pragma unroll
						for (int i = 0; i < SSTEPS; i++) {
							if (s0 + RG * i + r < lim) {
								acc = fmaf(acc, rho, pw[i]);
								mx = fmaxf(mx, pw[i]);
								const unsigned cell = bin_cell_addr<COLS>(pw[i], hofs, hscale2, kmax2, hb);
								asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(cell) : "memory");
							}
						}
					}
				}
				pl[v * 32 + lane] = __fmul_rn(acc, wtail);
				pl[ACC_VW * 32 + v * 32 + lane] = mx;
			}
			__syncwarp();           /* orders every lane's increments and partials before the arrive */
			if (lane == 0)
				mbar_arrive(role_bar + 8 * par);
		}
	} else if (warp < FW + UW) {
		/* ================= cell updater warps ================= */
		const int ut = threadIdx.x - FW * 32;                /* 0 .. UW*32-1 */
		constexpr int UT = UW * 32;
		float4 *h4 = reinterpret_cast<float4 *>(hist_s);
		const float2 *lut = B <= ACC_LUT_MAX ? lut_s : a.lut;

		const int n_groups = (a.n_calls + GC - 1) / GC;
		for (int grp = 0; grp < n_groups; grp++) {
			const int par = grp & 1;
			const int ncg = min(GC, a.n_calls - grp * GC);       /* calls of this group */
			mbar_wait_parity(role_bar + 8 * par, (unsigned)(grp >> 1) & 1u);
			/* ---- rise / decay of the tile's cells, display.cl:217-254 ---- */
			uint4 *hc4 = reinterpret_cast<uint4 *>(hits + (size_t)par * GC * tile_words);
			constexpr int UNR = 2;
			for (int g0 = ut; g0 < cells / 4; g0 += UT * UNR) {
				uint4 hc[UNR][GC];
				float4 hv[UNR];
#pragma unroll
				for (int u = 0; u < UNR; u++) {
					const int g = g0 + u * UT;
					if (g < cells / 4) {
						hv[u] = h4[g];
#pragma unroll
						for (int ci = 0; ci < GC; ci++)
							hc[u][ci] = hc4[ci * (tile_words / 4) + g];
					} else {
						hv[u] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
						for (int ci = 0; ci < GC; ci++)
							hc[u][ci] = make_uint4(0u, 0u, 0u, 0u);
					}
				}
#pragma unroll
				for (int u = 0; u < UNR; u++) {
					const int g = g0 + u * UT;
					bool touched = false;
#pragma unroll
					for (int ci = 0; ci < GC; ci++) {
						if (ci < ncg) {      /* tiles past the last call are all zero and must not decay the cells */
							const bool hit = (hc[u][ci].x | hc[u][ci].y | hc[u][ci].z | hc[u][ci].w) != 0u;
							if (hit)
								hc4[ci * (tile_words / 4) + g] = make_uint4(0u, 0u, 0u, 0u);
							if (hit || fmaxf(fmaxf(hv[u].x, hv[u].y), fmaxf(hv[u].z, hv[u].w)) > 0.01f) {
								hv[u].x = rise_decay(hv[u].x, hc[u][ci].x, lut);
								hv[u].y = rise_decay(hv[u].y, hc[u][ci].y, lut);
								hv[u].z = rise_decay(hv[u].z, hc[u][ci].z, lut);
								hv[u].w = rise_decay(hv[u].w, hc[u][ci].w, lut);
								touched = true;
							}
						}
					}
					if (touched)
						h4[g] = hv[u];
				}
			}
			__syncwarp();           /* all lanes done with this parity's tiles */
			if (lane == 0)
				mbar_arrive(role_bar + 16 + 8 * par);
		}

		/* tile state out: each updater thread stores the groups it owns */
		{
			constexpr int cpr = COLS / 4;
			for (int g = ut; g < cells / 4; g += UT) {
				const int bin = g / cpr, c4 = (g % cpr) * 4;
				*reinterpret_cast<float4 *>(a.hist + (size_t)bin * N + col0 + c4) = h4[g];
			}
		}
	} else if (warp == FW + UW) {
		/* ================= column warp: live IIR and max hold, display.cl:186-214,257-310 ================= */
		const int cc = lane % COLS;
		const int half = N >> 1;
		const int di = (col0 + cc) ^ half;                   /* display.cl:201 */
		float y = 0.0f, m = 0.0f;
		if (lane < COLS) {
			y = a.spectrum[di].y;
			m = a.spectrum[N + di].y;
		}
		for (int call = 0; call < a.n_calls; call++) {
			const int grp = call / GC, ci = call % GC, par = grp & 1;
			const bool last_of_group = ci == GC - 1 || call == a.n_calls - 1;
			if (ci == 0)
				mbar_wait_parity(role_bar + 8 * par, (unsigned)(grp >> 1) & 1u);
			const float *pl = parts + (size_t)(par * GC + ci) * 2 * ACC_VW * 32;
			float pv[ACC_VW], pm[ACC_VW];
#pragma unroll
			for (int w = 0; w < ACC_VW; w++) {
				pv[w] = pl[w * 32 + lane];
				pm[w] = pl[ACC_VW * 32 + w * 32 + lane];
			}
			if (last_of_group) {
				__syncwarp();       /* partials are in registers: hand the buffers back early */
				if (lane == 0)
					mbar_arrive(role_bar + 16 + 8 * par);
			}
			float sum = 0.0f, bmax = -1000.0f;
#pragma unroll
			for (int w = 0; w < ACC_VW; w++) {
				sum += pv[w];
				bmax = fmaxf(bmax, pm[w]);
			}
#pragma unroll
			for (int o = COLS; o < 32; o <<= 1) {
				sum += __shfl_xor_sync(0xffffffffu, sum, o);
				bmax = fmaxf(bmax, __shfl_xor_sync(0xffffffffu, bmax, o));
			}
			if (!isfinite(y))
				y = sum / (float)REF_ROWS;
			y = __fadd_rn(__fmul_rn(y, a.live_carry), __fmul_rn(sum, a.alpha));
			if (!isfinite(m))
				m = -FLT_MAX;
			m = __fadd_rn(__fmul_rn(m, a.mh_keep), __fmul_rn(a.mh_mix, y));
			m = fmaxf(m, bmax);
		}
		if (lane < COLS && a.n_calls > 0) {
			const float xpos = ((float)di / (float)half) - 1.0f;     /* display.cl:209 */
			a.spectrum[di] = make_float2(xpos, y);
			a.spectrum[N + di] = make_float2(xpos, m);
		}
	} else if constexpr (LOAD == 1) {
		/* ================= producer (TMA) ================= */
		if (lane == 0) {
			const unsigned mask = (unsigned)a.wf_mask;
			const unsigned total = (unsigned)a.n_calls * (unsigned)(B / BOXR);
			const unsigned stage0 = cnt_smem_u32(stage);
			unsigned long long pol = 0ull;
			if (a.l2_hints)
				asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
			for (unsigned n = 0; n < total; n++) {
				const unsigned slot = n & (DEPTH - 1u);
				if (n >= DEPTH)
					mbar_wait_parity(empty0 + 8u * slot, ((n >> dlog) - 1u) & 1u);
				asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
				             ::"r"(full0 + 8u * slot), "r"((unsigned)C::BOX_BYTES) : "memory");
				const int row = (int)(((unsigned)a.wf_pos + n * BOXR) & mask);
				if (a.l2_hints)
					tma_load_2d_hint(stage0 + slot * (unsigned)C::BOX_BYTES, &tmap, col0, row, full0 + 8u * slot, pol);
				else
					tma_load_2d(stage0 + slot * (unsigned)C::BOX_BYTES, &tmap, col0, row, full0 + 8u * slot);
			}
		}
	}
}

/* first-use state, cl.c:406-465 */
__global__ void fill_kernel(float *p, size_t n, float v)
{
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
	     i += (size_t)gridDim.x * blockDim.x)
		p[i] = v;
}

/* The last `cnt` rows of the engine's log-power scratch ring -> the user-visible waterfall ring
 * (display.cl:141-146 writes them there directly; here only when somebody looks, engine.cu: publish). */
__global__ void publish_rows_kernel(const float4 *__restrict__ ring, float4 *__restrict__ wf, int n4, int cnt,
                                    int src0, int src_mask, int dst0, int dst_mask)
{
	const size_t total = (size_t)cnt * n4;
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
		const int r = (int)(i / n4), c = (int)(i - (size_t)r * n4);
		wf[(size_t)((dst0 + r) & dst_mask) * n4 + c] = __ldcs(&ring[(size_t)((src0 + r) & src_mask) * n4 + c]);
	}
}

/* max-hold trace (y values, display order) for the multi-GPU reduce */
__global__ void export_maxhold_kernel(const float2 *spectrum, int n, float *out)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n)
		out[i] = spectrum[n + i].y;
}

} /* namespace fosphor_b200 */
