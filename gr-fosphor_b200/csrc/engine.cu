/*
 * engine.cu - host side of libfosphor_b200.so: the parameterised fosphor_cu_*
 * engine (include/fosphor_b200.h).  It plays the role of the reference's
 * OpenCL host driver lib/fosphor/cl.c (state machine :95-99, buffer set-up
 * :648-731, per-call enqueue :870-968, finish/read-back :970-1061) for CUDA:
 * everything is enqueued on one stream, results stay in device arrays until
 * finish().  The kernels live in fft_power.cuh and accumulate.cuh.
 *
 * No CPU fallback: if CUDA is unusable create() fails (-ENODEV / -EIO).
 */
#include <cerrno>
#include <new>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/fosphor_b200.h"
#include "accumulate.cuh"
#include "fft_power.cuh"

using namespace fosphor_b200;

namespace {

enum { ST_BOOTING = 0, ST_PENDING, ST_READY };   /* cl.c:95-99 */

constexpr int N_TABLES = 8;      /* cached (weights, lut) sets, one per distinct batch size */
constexpr int MAX_SLICES = 128;  /* (call, row-split) slices folded by one count/update launch pair */
constexpr size_t CNT_BUDGET = (size_t)1 << 30;   /* bytes of u16 hit-count slices kept on the device (split kernels) */
constexpr int N_CHUNK_EV = 64;   /* chunks in flight tracked by the two-stream schedule */
constexpr size_t UPD_SMEM_MAX = 96 * 1024;   /* dynamic shared memory of update_kernel */
constexpr size_t UPD_LUT_SMEM_MAX = 64 * 1024;   /* ... of which the (d, e) table: batches up to 8191 rows */

struct BatchTables {
	int batch = -1;
	unsigned long long last_use = 0;
	float *d_weights = nullptr;   /* [batch_max]      */
	float2 *d_lut = nullptr;      /* [batch_max + 1]  */
	float *h_stage = nullptr;     /* pinned: weights then lut */
	cudaEvent_t uploaded = nullptr;
	float carry = 0.0f;
};

} /* namespace */

struct fosphor_cu {
	fosphor_cu_params p;
	int log2n = 0;
	int device = 0;
	int sm_count = 0;
	int await_slot = -1;                 /* staging slot whose H2D copy from caller memory the current call still has to wait for */
	size_t smem_optin = 0;               /* largest dynamic shared memory a CTA may ask for */

	cudaStream_t own_stream = nullptr;
	cudaStream_t stream = nullptr;       /* FFT kernel, copies to the host, everything the caller orders against */
	cudaStream_t acc_stream = nullptr;   /* count / update kernels: overlap the next chunk's FFT */
	cudaEvent_t fft_done[N_CHUNK_EV] = {};   /* per chunk, round robin */
	cudaEvent_t cnt_done[N_CHUNK_EV] = {};
	cudaEvent_t acc_done = nullptr;
	cudaEvent_t cols_fork = nullptr, cols_join = nullptr;   /* column update runs beside the cell update */
	int overlap = -1;                    /* env FOSPHOR_B200_OVERLAP.  Two-stream schedule: the accumulate kernel of chunk c
	                                      * runs on a second (higher priority) stream while the FFT of chunk c+1 runs.
	                                      * -1 (default) = automatic: on when the ring holds four chunks of >= 32 M samples
	                                      *    (N = 512 / 1024 streaming FFT + fused accumulate); chunk = ring / 4, full-size
	                                      *    accumulate CTAs, FFT at 3 CTAs/SM.  The accumulate CTAs take 128 SMs, the next
	                                      *    FFT starts on the 20 that are left and on every SM an accumulate CTA leaves:
	                                      *    cfg2, 256-call ring: 385 vs 345 Gsamples/s; DRAM traffic at 0.91 of the peak.
	                                      *    Smaller rings lose: chunks of 16 calls pay the launch ramp/tail 4x as often
	                                      *    (64-call ring: 327 vs 345).
	                                      *  0 = one stream.   1 = forced, with OVERLAP_CHUNK / ACC_SLIM / FFT_CTAS knobs. */
	bool two_streams_now = false;        /* set per process call */
	bool acc_pending = false;            /* accumulate work on acc_stream that `stream` has not been ordered after yet */
	long long chunk_seq = 0;             /* chunks issued by the two-stream schedule since the last join */
	unsigned long long two_stream_chunks = 0;   /* ... since create (diagnostics) */
	int seq_batch = 0, seq_chunk_calls = 0;   /* ... and their geometry (a change forces a join) */
	bool slim_now = false;               /* ... two-stream mode with the slim co-resident accumulate CTA */
	int overlap_chunk = 16;              /* env FOSPHOR_B200_OVERLAP_CHUNK: calls per chunk of the two-stream schedule */
	int acc_slim = 1;                    /* env FOSPHOR_B200_ACC_SLIM: two-stream mode uses the 14-warp fused kernel that
	                                      * is co-resident with two FFT CTAs per SM */
	int fft_r64 = 0;                     /* env FOSPHOR_B200_FFT_R64: two-pass radix-64 plans for N = 2048 / 4096 */
	int plan_key = 0;                    /* fft_len, +1 for the radix-64 plan */
	int fft_pf = -1;                     /* env FOSPHOR_B200_FFT_PF: L2 prefetch distance (spectra) of the plain FFT kernel;
	                                      * -1 = the number of resident CTAs, 0 = off */
	int fft_ctas_per_sm = 0;             /* env FOSPHOR_B200_FFT_CTAS: force the CTAs/SM of the persistent FFT
	                                      * kernel (0 = automatic: 3, or 2 when count runs beside it) */
	CUtensorMap wf_tmap;                 /* waterfall ring as a 2-D tensor, box = 16 rows x 32 columns */
	bool tmap_ok = false;
	int count_variant = 1;               /* 1: TMA-staged count kernel where applicable, 0: plain
	                                      * (env FOSPHOR_B200_COUNT_VARIANT) */
	int acc_mode = -1;                   /* env FOSPHOR_B200_ACC: 1 = fused accumulate kernel (count + update in one
	                                      * launch, state tile resident in shared memory), 0 = split count / update
	                                      * kernels, -1 = by shape: fused when the batch is at least as long as the
	                                      * bin count (measured: cfg2 +14 %, cfg4 +9 %, N=512 sweep +10 %), split
	                                      * when the per-call state update dominates (cfg3: B = K/2, even) */
	int chunk_calls = 0;                 /* env FOSPHOR_B200_CHUNK_CALLS: cap on the calls folded per launch (0 = ring) */
	int acc_roles = 0;                   /* env FOSPHOR_B200_ACC_ROLES: counter / updater warps 1 = 16/8, 2 = 8/16 (4-call groups only); 0 = by shape */
	int acc_stage_kb = 0;                /* env FOSPHOR_B200_ACC_STAGE_KB: cap of the fused kernel's stage ring (0 = what fits) */
	int acc_group = 0;                   /* env FOSPHOR_B200_ACC_GROUP: calls per counter -> updater hand-over (1 | 2 | 4); 0 = by batch size */
	int acc_cols = 8;                    /* columns per CTA of the fused kernel (env FOSPHOR_B200_ACC_COLS: 4 | 8) */
	int acc_box_max = 256;               /* largest TMA box in rows (env FOSPHOR_B200_ACC_BOX: 0 (plain loads) | 16 | 64 | 256) */
	int acc_sub_max = 64;                /* rows per unrolled body (env FOSPHOR_B200_ACC_SUB: 16 | 64) */
	CUtensorMap acc_tmap[3];             /* waterfall ring, box = 16 / 64 / 256 rows x acc_cols columns */
	bool acc_tmap_ok = false;

	float *d_win = nullptr;
	float2 *d_tw = nullptr;
	float *d_wf = nullptr;
	float *d_hist = nullptr;
	float2 *d_spec = nullptr;
	unsigned short *d_cnt = nullptr;     /* [max_slices][K][N] */
	float *d_part_live = nullptr, *d_part_max = nullptr;   /* [max_slices][N] */
	int max_slices = 0;

	/* host-sample staging (fosphor_cu_process_host*) */
	size_t stage_elems = 0;              /* complex samples per slot */
	float2 *h_in[2] = {nullptr, nullptr};
	float2 *d_in[2] = {nullptr, nullptr};
	cudaStream_t copy_stream = nullptr;  /* H2D of samples, overlaps the compute stream */
	cudaEvent_t copied[2] = {nullptr, nullptr};     /* H2D into slot done (copy stream)   */
	cudaEvent_t slot_free[2] = {nullptr, nullptr};  /* kernels that read the slot done    */
	int slot = 0;
	int last_slot = -1;
	float *h_win = nullptr;              /* pinned copy of the window */
	cudaEvent_t win_done = nullptr;

	float histo_scale = 0.0f, histo_ofs = 0.0f;   /* cl.c:811 memset, :1087-1088 */
	int wf_pos = 0;
	int state = ST_BOOTING;

	BatchTables tables[N_TABLES];
	unsigned long long use_clock = 0;
	unsigned long long launches = 0;
	int fft_variant = 2;                 /* env FOSPHOR_B200_FFT_VARIANT, N = 512/1024 with aligned input:
	                                      * 2: TMA-prefetching persistent kernel, twiddles in registers
	                                      * 1: same, twiddles fetched per spectrum   0: plain kernel
	                                      * 3: as 2, plus the CTA-level streaming kernel for N = 2048..8192 */

	/* optional per-kernel timing (bench.py roofline): event pairs around launches */
	bool profiling = false;
	std::vector<cudaEvent_t> prof_ev[3][2];   /* [kernel: 0 fft, 1 count, 2 update][begin/end] */
	size_t prof_used[3] = {0, 0, 0};

	char err[256] = {0};
};

namespace {

int fail(fosphor_cu *e, int rc, const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	if (e) {
		vsnprintf(e->err, sizeof(e->err), fmt, ap);
		fprintf(stderr, "[!] fosphor_b200: %s\n", e->err);   /* cl.c:109-114 style */
	} else {
		fprintf(stderr, "[!] fosphor_b200: ");
		vfprintf(stderr, fmt, ap);
		fprintf(stderr, "\n");
	}
	va_end(ap);
	return rc;
}

#define CU_CHECK(e, call)                                                        \
	do {                                                                     \
		cudaError_t err__ = (call);                                      \
		if (err__ != cudaSuccess)                                        \
			return fail((e), -EIO, "CUDA error %d (%s) at %s:%d: %s", (int)err__, \
			            cudaGetErrorString(err__), __FILE__, __LINE__, #call);    \
	} while (0)

/* ---- optional kernel timing ------------------------------------------------ */

void prof_mark(fosphor_cu *e, int kernel, int end, cudaStream_t st = nullptr)
{
	if (!e->profiling)
		return;
	auto &v = e->prof_ev[kernel][end];
	const size_t i = e->prof_used[kernel];
	if (i >= v.size()) {
		cudaEvent_t ev;
		if (cudaEventCreate(&ev) != cudaSuccess)
			return;
		v.push_back(ev);
	}
	cudaEventRecord(v[i], st ? st : e->stream);
	if (end)
		e->prof_used[kernel]++;
}

/* ---- FFT plan dispatch --------------------------------------------------- */

template <class P>
void build_twiddles(std::vector<float2> &tw)
{
	tw.resize(P::TW_ELEMS);
	/* pass 1: tw[t][k] = exp(-2 pi i t k / (R0*R1)), k < R0 */
	for (int t = 0; t < P::R1; t++)
		for (int k = 0; k < P::R0; k++) {
			const double a = -2.0 * M_PI * (double)t * (double)k / (double)(P::R0 * P::R1);
			tw[t * P::R0 + k] = make_float2((float)cos(a), (float)sin(a));
		}
	if (P::NPASS == 3) {
		const int p2 = P::R0 * P::R1;
		for (int t = 0; t < P::R1; t++)
			for (int k = 0; k < p2; k++) {
				const double a = -2.0 * M_PI * (double)t * (double)k / (double)P::N;
				tw[P::TW1 + t * p2 + k] = make_float2((float)cos(a), (float)sin(a));
			}
	}
}

template <class P>
cudaError_t plan_setup()
{
	cudaError_t err = cudaFuncSetAttribute(fft_power_kernel<P, false>,
		cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P::SMEM);
	if (err != cudaSuccess)
		return err;
	return cudaFuncSetAttribute(fft_power_kernel<P, true>,
		cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P::SMEM);
}

template <class P, bool CPLX>
cudaError_t plan_launch(fosphor_cu *e, const float2 *in, long long hop, int wf_pos,
                        float2 *cplx_out, int n_spectra)
{
	const int grid = (n_spectra + P::SPB - 1) / P::SPB;
	/* L2 prefetch distance of the one-spectrum-per-CTA plans: the CTAs resident at once */
	int pf = 0;
	const bool aligned = ((reinterpret_cast<unsigned long long>(in) & 15ull) == 0) && ((hop & 1) == 0);
	if (P::SPB == 1 && aligned && e->fft_pf != 0) {
		if (e->fft_pf > 0) {
			pf = e->fft_pf;
		} else {
			static int per_sm = 0;         /* per plan instantiation */
			if (per_sm == 0 &&
			    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fft_power_kernel<P, CPLX>,
			                                                  P::THREADS, P::SMEM) != cudaSuccess)
				per_sm = 1;
			pf = e->sm_count * (per_sm > 0 ? per_sm : 1);
		}
	}
	prof_mark(e, 0, 0);
	fft_power_kernel<P, CPLX><<<grid, P::THREADS, P::SMEM, e->stream>>>(
		in, hop, e->d_win, e->d_tw, e->d_wf, wf_pos, e->p.wf_rows - 1, cplx_out, n_spectra, pf);
	prof_mark(e, 0, 1);
	e->launches++;
	return cudaGetLastError();
}

#define PLAN_SWITCH(n, EXPR)                                   \
	switch (n) {                                           \
	case 512:   { using P = Plan512;   EXPR; } break;      \
	case 1024:  { using P = Plan1024;  EXPR; } break;      \
	case 2048:  { using P = Plan2048;  EXPR; } break;      \
	case 2049:  { using P = Plan2048R64; EXPR; } break;    \
	case 4097:  { using P = Plan4096R64; EXPR; } break;    \
	case 4096:  { using P = Plan4096;  EXPR; } break;      \
	case 8192:  { using P = Plan8192;  EXPR; } break;      \
	case 16384: { using P = Plan16384; EXPR; } break;      \
	default: break;                                        \
	}

bool plan_supported(int n)
{
	return n == 512 || n == 1024 || n == 2048 || n == 4096 || n == 8192 || n == 16384;
}

template <class P>
cudaError_t stream_setup()
{
	cudaError_t err = cudaFuncSetAttribute(fft_power_stream_kernel<P, false>,
		cudaFuncAttributeMaxDynamicSharedMemorySize, (int)StreamCfg<P>::SMEM);
	if (err != cudaSuccess)
		return err;
	return cudaFuncSetAttribute(fft_power_stream_kernel<P, true>,
		cudaFuncAttributeMaxDynamicSharedMemorySize, (int)StreamCfg<P>::SMEM_TWREG);
}

template <class P>
cudaError_t stream_launch(fosphor_cu *e, const float2 *in, long long hop, int wf_pos, int n_spectra)
{
	using C = StreamCfg<P>;
	const int units = (n_spectra + C::SPW - 1) / C::SPW;   /* a warp takes SPW spectra per iteration */
	int grid = (units + C::WARPS - 1) / C::WARPS;
	int per_sm = C::CTAS_PER_SM;
	if (e->fft_ctas_per_sm > 0 && e->fft_ctas_per_sm < per_sm)
		per_sm = e->fft_ctas_per_sm;
	else if (e->fft_ctas_per_sm == 0 && e->slim_now && per_sm > 2)
		per_sm = 2;                      /* leave room for the count kernel on the other stream */
	const int resident = e->sm_count * per_sm;
	if (grid > resident)
		grid = resident;                 /* persistent warps, grid-stride over spectra */
	prof_mark(e, 0, 0);
	if (e->fft_variant >= 2)
		fft_power_stream_kernel<P, true><<<grid, C::THREADS, C::SMEM_TWREG, e->stream>>>(
			in, hop, e->d_win, e->d_tw, e->d_wf, wf_pos, e->p.wf_rows - 1, n_spectra);
	else
		fft_power_stream_kernel<P, false><<<grid, C::THREADS, C::SMEM, e->stream>>>(
			in, hop, e->d_win, e->d_tw, e->d_wf, wf_pos, e->p.wf_rows - 1, n_spectra);
	prof_mark(e, 0, 1);
	e->launches++;
	return cudaGetLastError();
}

template <class P>
cudaError_t cta_stream_setup()
{
	return cudaFuncSetAttribute(fft_power_cta_stream_kernel<P>,
		cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CtaStreamCfg<P>::SMEM);
}

template <class P>
cudaError_t cta_stream_launch(fosphor_cu *e, const float2 *in, long long hop, int wf_pos, int n_spectra)
{
	using C = CtaStreamCfg<P>;
	int grid = n_spectra;
	const int resident = e->sm_count * C::CTAS_PER_SM;
	if (grid > resident)
		grid = resident;
	prof_mark(e, 0, 0);
	fft_power_cta_stream_kernel<P><<<grid, C::THREADS, C::SMEM, e->stream>>>(
		in, hop, e->d_win, e->d_tw, e->d_wf, wf_pos, e->p.wf_rows - 1, n_spectra);
	prof_mark(e, 0, 1);
	e->launches++;
	return cudaGetLastError();
}

template <class P>
cudaError_t half_stage_launch(fosphor_cu *e, const float2 *in, long long hop, int wf_pos, int n_spectra)
{
	using C = HalfStageCfg<P>;
	static bool configured = false;
	if (!configured) {
		cudaError_t err = cudaFuncSetAttribute(fft_power_half_stage_kernel<P>,
			cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
		if (err != cudaSuccess)
			return err;
		configured = true;
	}
	const int grid = n_spectra < e->sm_count ? n_spectra : e->sm_count;   /* persistent, one CTA per SM */
	prof_mark(e, 0, 0);
	fft_power_half_stage_kernel<P><<<grid, C::THREADS, C::SMEM, e->stream>>>(
		in, hop, e->d_win, e->d_tw, e->d_wf, wf_pos, e->p.wf_rows - 1, n_spectra);
	prof_mark(e, 0, 1);
	e->launches++;
	return cudaGetLastError();
}

cudaError_t launch_fft(fosphor_cu *e, const float2 *in, long long hop, int wf_pos, int n_spectra)
{
	/* TMA bulk copies need 16-byte aligned spectra */
	const bool aligned = ((reinterpret_cast<unsigned long long>(in) & 15ull) == 0) && ((hop & 1) == 0);
	if (aligned && e->fft_variant != 0) {
		if (e->p.fft_len == 1024)
			return stream_launch<Plan1024>(e, in, hop, wf_pos, n_spectra);
		if (e->p.fft_len == 512)
			return stream_launch<Plan512>(e, in, hop, wf_pos, n_spectra);
		if (e->p.fft_len == 16384 && e->fft_variant >= 2)
			return half_stage_launch<Plan16384>(e, in, hop, wf_pos, n_spectra);
		/* The CTA-level streaming kernel measured SLOWER than the plain one (N = 4096:
		 * 101 vs 86 us per 8192 spectra; fewer resident CTAs outweigh the prefetch), so
		 * it is only reachable as experiment variant 3. */
		if (e->fft_variant == 3 && e->plan_key == e->p.fft_len) {
			if (e->p.fft_len == 2048)
				return cta_stream_launch<Plan2048>(e, in, hop, wf_pos, n_spectra);
			if (e->p.fft_len == 4096)
				return cta_stream_launch<Plan4096>(e, in, hop, wf_pos, n_spectra);
			if (e->p.fft_len == 8192)
				return cta_stream_launch<Plan8192>(e, in, hop, wf_pos, n_spectra);
		}
	}
	cudaError_t err = cudaErrorInvalidValue;
	PLAN_SWITCH(e->plan_key, (err = plan_launch<P, false>(e, in, hop, wf_pos, nullptr, n_spectra)));
	return err;
}

/* ---- per-batch tables ---------------------------------------------------- */

/* display.cl:150,210,241-245 evaluated once per distinct batch size on the
 * host in f32 (same operation order as the reference / the oracle). */
int get_tables(fosphor_cu *e, int batch, BatchTables **out)
{
	BatchTables *t = nullptr, *lru = &e->tables[0];
	for (auto &c : e->tables) {
		if (c.batch == batch)
			t = &c;
		if (c.last_use < lru->last_use)
			lru = &c;
	}
	if (!t) {
		t = lru;
		const int bm = e->p.batch_max;
		CU_CHECK(e, cudaEventSynchronize(t->uploaded));   /* staging free again */
		float *w = t->h_stage;
		float2 *lut = reinterpret_cast<float2 *>(t->h_stage + bm);
		const float oma = 1.0f - e->p.live_alpha;         /* display.cl:99 */
		const float fb = (float)batch;
		const float rt0r = 1.0f / e->p.histo_t0r;
		const float rt0d = 1.0f / e->p.histo_t0d;
		for (int s = 0; s < batch; s++)
			w[s] = powf(oma, (float)(batch - s - 1));
		for (int hc = 0; hc <= batch; hc++) {
			const float a = (float)hc / fb;
			const float b = a * rt0r;
			const float c = b + rt0d;
			const float d = b * (1.0f / c);
			lut[hc] = make_float2(d, powf(1.0f - c, fb));
		}
		t->carry = powf(oma, fb);
		CU_CHECK(e, cudaMemcpyAsync(t->d_weights, w, sizeof(float) * batch,
		                            cudaMemcpyHostToDevice, e->stream));
		CU_CHECK(e, cudaMemcpyAsync(t->d_lut, lut, sizeof(float2) * (batch + 1),
		                            cudaMemcpyHostToDevice, e->stream));
		CU_CHECK(e, cudaEventRecord(t->uploaded, e->stream));
		t->batch = batch;
	}
	t->last_use = ++e->use_clock;
	*out = t;
	return 0;
}

/* How the calls of one launch are cut into slices (CTAs along the row axis).
 * Only hit counts (integers) and per-ROWBLOCK partial sums cross slice
 * boundaries, so the results are bit-identical for any slicing; the choice is
 * purely about keeping the SMs evenly busy: pick the split count whose CTA
 * grid quantises best against the number of resident CTAs (every slice also
 * pays a fixed cost for clearing and storing its hit tile). */
void choose_slicing(const fosphor_cu *e, int n_calls, int batch, int *splits, int *rows_per_split)
{
	const int tiles = e->p.fft_len / ACC_COLS;
	const int blocks = (batch + ROWBLOCK - 1) / ROWBLOCK;
	const size_t smem = sizeof(unsigned) * 32 * (size_t)e->p.n_bins + sizeof(CountStage) + 1024;
	int per_sm = (int)((size_t)227 * 1024 / smem);
	if (per_sm < 1) per_sm = 1;
	if (per_sm > 8) per_sm = 8;
	const long resident = (long)per_sm * e->sm_count;
	const int overhead_rows = 24 + e->p.n_bins / 16;       /* tile clear + store, in row-equivalents */
	int best_s = 1;
	double best_cost = 1e300;
	for (int s = 1; s <= blocks; s++) {
		if ((long)s * n_calls > e->max_slices)
			break;
		const int rows = (blocks + s - 1) / s * ROWBLOCK;
		const int real_s = (batch + rows - 1) / rows;
		const long ctas = (long)tiles * n_calls * real_s;
		const long waves = (ctas + resident - 1) / resident;
		/* a partially filled last wave still runs at per-CTA speed */
		const double cost = (double)waves * (rows + overhead_rows) *
		                    (ctas < resident ? (double)((ctas + e->sm_count - 1) / e->sm_count) / per_sm : 1.0);
		if (cost < best_cost - 1e-9) {
			best_cost = cost;
			best_s = real_s;
		}
	}
	const int rows = (blocks + best_s - 1) / best_s * ROWBLOCK;
	*rows_per_split = rows;
	*splits = (batch + rows - 1) / rows;
}

constexpr int ACC_UW = 8;         /* updater warps of the fused accumulate kernel */
constexpr int ACC_FW_SLIM = 8;    /* counter / updater warps of the slim variant that runs beside the FFT kernel: */
constexpr int ACC_UW_SLIM = 4;    /* 14 warps x 48 registers fit next to two FFT CTAs (2 x 4 warps x 168 registers)  */

template <int COLS, int FW, int UW, int BOXR, int SUBR, int LOAD, int GC>
cudaError_t fused_launch(fosphor_cu *e, AccumArgs a, cudaStream_t st)
{
	using C = FusedCfg<COLS, FW, UW, BOXR, GC>;
	/* stage ring: as many boxes as the SM has room for (FOSPHOR_B200_ACC_STAGE_KB caps it), never more than the launch has */
	size_t limit = e->smem_optin;
	if (e->acc_stage_kb > 0 && C::smem_fixed(a.n_bins, a.batch) + (size_t)e->acc_stage_kb * 1024 < limit)
		limit = C::smem_fixed(a.n_bins, a.batch) + (size_t)e->acc_stage_kb * 1024;
	int dlog = C::depth_log2(a.n_bins, a.batch, limit);
	const long long boxes = (long long)a.n_calls * (a.batch / BOXR);
	while (dlog > 0 && (1ll << (dlog - 1)) >= boxes)
		dlog--;
	a.depth_log2 = dlog;
	if constexpr (LOAD != 0) {
		/* parity waits on the stage ring are only sound for these shapes (accumulate.cuh: acc_ring_safe) */
		if (!acc_ring_safe(1ll << dlog, (long long)GC * (a.batch / BOXR), boxes))
			return fused_launch<COLS, FW, UW, 16, 16, 0, GC>(e, a, st);
	}
	const size_t smem = C::smem(a.n_bins, a.batch, LOAD != 0, dlog);
	static size_t configured = 0;          /* per kernel instantiation */
	if (smem > configured) {
		cudaError_t err = cudaFuncSetAttribute(accumulate_fused_kernel<COLS, FW, UW, BOXR, SUBR, LOAD, GC>,
			cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if (err != cudaSuccess)
			return err;
		configured = smem;
	}
	const CUtensorMap &tm = e->acc_tmap[BOXR == 256 ? 2 : (BOXR == 64 ? 1 : 0)];
	accumulate_fused_kernel<COLS, FW, UW, BOXR, SUBR, LOAD, GC><<<a.n / COLS, C::THREADS, smem, st>>>(a, tm);
	return cudaGetLastError();
}

template <int COLS, int FW, int UW, int GC = 1>
cudaError_t fused_dispatch(fosphor_cu *e, const AccumArgs &a, cudaStream_t st, int boxr, int subr)
{
	if (boxr == 256)
		return subr == 64 ? fused_launch<COLS, FW, UW, 256, 64, 1, GC>(e, a, st) : fused_launch<COLS, FW, UW, 256, 16, 1, GC>(e, a, st);
	if (boxr == 64)
		return subr == 64 ? fused_launch<COLS, FW, UW, 64, 64, 1, GC>(e, a, st) : fused_launch<COLS, FW, UW, 64, 16, 1, GC>(e, a, st);
	if (boxr == 16)
		return fused_launch<COLS, FW, UW, 16, 16, 1, GC>(e, a, st);
	return fused_launch<COLS, FW, UW, 16, 16, 0, GC>(e, a, st);
}

/* shared memory of the fused kernel for a shape with a minimal stage ring (two 256-row boxes) */
template <int COLS>
size_t fused_smem_need(int n_bins, int batch, int gc)
{
	if (gc == 4)
		return FusedCfg<COLS, 16, ACC_UW, 256, 4>::smem(n_bins, batch, true, 1);
	if (gc == 2)
		return FusedCfg<COLS, 16, ACC_UW, 256, 2>::smem(n_bins, batch, true, 1);
	return FusedCfg<COLS, 16, ACC_UW, 256, 1>::smem(n_bins, batch, true, 1);
}

/* one launch: count + rise/decay + live + max-hold of n_calls calls */
int launch_accumulate_fused(fosphor_cu *e, const BatchTables *t, cudaStream_t st, cudaEvent_t count_done,
                            int wf_pos, int n_calls, int batch)
{
	AccumArgs a;
	memset(&a, 0, sizeof(a));
	a.wf = e->d_wf;
	a.hist = e->d_hist;
	a.spectrum = e->d_spec;
	a.weights = t->d_weights;
	a.lut = t->d_lut;
	a.n = e->p.fft_len;
	a.n_bins = e->p.n_bins;
	a.wf_mask = e->p.wf_rows - 1;
	a.wf_pos = wf_pos;
	a.batch = batch;
	a.n_calls = n_calls;
	a.hscale = e->histo_scale;
	a.hofs = e->histo_ofs;
	a.alpha = e->p.live_alpha;
	a.live_carry = t->carry;
	a.mh_keep = e->p.maxhold_keep;
	a.mh_mix = e->p.maxhold_mix;
	a.rho_rg = powf(1.0f - e->p.live_alpha, (float)(32 / e->acc_cols));
	/* Largest TMA box (rows) that tiles the batch and the virtual warps' runs and never
	 * straddles the ring end; 0 = plain loads.  Transport only: the row -> lane
	 * assignment, hence every result bit, is the same for all of them. */
	int boxr = 0, subr = 16;
	if (e->acc_tmap_ok) {
		const int rv = acc_rows_per_vwarp(batch);
		for (int b = e->acc_box_max; b >= 16; b >>= 2)
			if (batch % b == 0 && wf_pos % b == 0 && e->p.wf_rows % b == 0 && (b % rv == 0 || rv % b == 0)) {
				boxr = b;
				const int ch = rv < b ? rv : b;
				subr = (ch % 64 == 0 && e->acc_sub_max >= 64) ? 64 : 16;
				break;
			}
	}
	prof_mark(e, 1, 0, st);
	cudaError_t err;
	const bool slim = e->slim_now;
	if (slim)
		subr = 16;      /* 46 registers (the 64-row body needs 56): 14 warps fit beside two FFT CTAs */
	/* calls per synchronisation group: with few rows per call the barrier round trips between
	 * counter and updater warps set the pace, so small batches hand over 2 or 4 calls at a time
	 * (transport only: same per-cell and per-column operation order) */
	int gc = e->acc_group;
	if (gc == 0)
		gc = batch <= 256 ? 4 : 1;        /* measured: cfg3 (B = 256) 115 -> 105 us per 32 calls; B = 1024 prefers the deeper stage ring */
	const size_t smem_max = e->smem_optin;
	while (gc > 1 && (e->acc_cols == 4 ? fused_smem_need<4>(a.n_bins, batch, gc)
	                                   : fused_smem_need<8>(a.n_bins, batch, gc)) > smem_max)
		gc >>= 1;
	/* warp roles (counters / cell updaters): 16 / 8, or 8 / 16 when a call has more cells to update
	 * than rows to count (ncu of cfg3, B = 256, K = 512: the counter warps spent half their time
	 * waiting for the updaters to hand the hit tiles back) */
	int roles = e->acc_roles;
	if (roles == 0)
		roles = (gc == 4 && a.n_bins >= batch) ? 2 : 1;
	if (e->acc_cols == 4)
		err = slim ? fused_dispatch<4, ACC_FW_SLIM, ACC_UW_SLIM>(e, a, st, boxr, subr)
		    : gc == 4 ? (roles == 2 ? fused_dispatch<4, 8, 16, 4>(e, a, st, boxr, subr)
		                            : fused_dispatch<4, 16, ACC_UW, 4>(e, a, st, boxr, subr))
		    : gc == 2 ? fused_dispatch<4, 16, ACC_UW, 2>(e, a, st, boxr, subr)
		    : roles == 2 ? fused_dispatch<4, 8, 16, 1>(e, a, st, boxr, subr)
		              : fused_dispatch<4, 16, ACC_UW, 1>(e, a, st, boxr, subr);
	else
		err = slim ? fused_dispatch<8, ACC_FW_SLIM, ACC_UW_SLIM>(e, a, st, boxr, subr)
		    : gc == 4 ? (roles == 2 ? fused_dispatch<8, 8, 16, 4>(e, a, st, boxr, subr)
		                            : fused_dispatch<8, 16, ACC_UW, 4>(e, a, st, boxr, subr))
		    : gc == 2 ? fused_dispatch<8, 16, ACC_UW, 2>(e, a, st, boxr, subr)
		    : roles == 2 ? fused_dispatch<8, 8, 16, 1>(e, a, st, boxr, subr)
		              : fused_dispatch<8, 16, ACC_UW, 1>(e, a, st, boxr, subr);
	prof_mark(e, 1, 1, st);
	e->launches++;
	CU_CHECK(e, err);
	if (count_done)
		CU_CHECK(e, cudaEventRecord(count_done, st));   /* the ring rows of this chunk are free again */
	return 0;
}

/* path choice depends on (B, K) only, never on the ring position or the launch
 * folding: the two paths add the live spectrum in different orders */
bool use_fused(const fosphor_cu *e, int batch)
{
	return e->acc_mode > 0 || (e->acc_mode < 0 && 2 * batch >= e->p.n_bins && (batch % 16) == 0 && e->acc_tmap_ok);
}

/* fold n_calls calls (rows wf_pos .. wf_pos + n_calls*batch of the ring) into the state */
int launch_accumulate(fosphor_cu *e, const BatchTables *t, cudaStream_t st, cudaEvent_t count_done,
                      int wf_pos, int n_calls, int batch)
{
	/* path choice depends on (B, K) only, never on the ring position or the launch
	 * folding: the two paths add the live spectrum in different orders */
	if (use_fused(e, batch))
		return launch_accumulate_fused(e, t, st, count_done, wf_pos, n_calls, batch);
	AccumArgs a;
	a.wf = e->d_wf;
	a.hist = e->d_hist;
	a.spectrum = e->d_spec;
	a.cnt = e->d_cnt;
	a.part_live = e->d_part_live;
	a.part_max = e->d_part_max;
	a.weights = t->d_weights;
	a.lut = t->d_lut;
	a.n = e->p.fft_len;
	a.n_bins = e->p.n_bins;
	a.wf_mask = e->p.wf_rows - 1;
	a.wf_pos = wf_pos;
	a.batch = batch;
	a.n_calls = n_calls;
	choose_slicing(e, n_calls, batch, &a.splits, &a.rows_per_split);
	a.hscale = e->histo_scale;
	a.hofs = e->histo_ofs;
	a.alpha = e->p.live_alpha;
	a.live_carry = t->carry;
	a.mh_keep = e->p.maxhold_keep;
	a.mh_mix = e->p.maxhold_mix;

	const dim3 grid(e->p.fft_len / ACC_COLS, n_calls * a.splits);
	const size_t smem = sizeof(unsigned) * 32 * (size_t)e->p.n_bins;
	const bool use_tma = e->tmap_ok && e->count_variant != 0 &&
	                     (batch % TMA_ROWS) == 0 && (wf_pos % TMA_ROWS) == 0;
	prof_mark(e, 1, 0, st);
	if (use_tma)
		count_tma_kernel<<<grid, ACC_THREADS, smem + sizeof(CountStage), st>>>(a, e->wf_tmap);
	else
		count_kernel<<<grid, ACC_THREADS, smem, st>>>(a);
	prof_mark(e, 1, 1, st);
	if (count_done)
		CU_CHECK(e, cudaEventRecord(count_done, st));   /* the ring rows of this chunk are free again */
	const size_t cells = (size_t)e->p.n_bins * e->p.fft_len;
	const size_t per_block = (size_t)UPD_THREADS * UPD_CELLS;   /* N is a multiple of 512: no straddling */
	const int cell_blocks = (int)((cells + per_block - 1) / per_block);
	const int col_blocks = (e->p.fft_len + UPD_COLS - 1) / UPD_COLS;
	a.lut_staged = sizeof(float2) * (size_t)(batch + 1) <= UPD_LUT_SMEM_MAX;
	const size_t lut_smem = a.lut_staged ? sizeof(float2) * (size_t)(batch + 1) : 0;
	/* partials staged per pass by the column blocks */
	const int blocks_per_call = (batch + ROWBLOCK - 1) / ROWBLOCK;
	const int cap = blocks_per_call > UPD_PARTS ? blocks_per_call : UPD_PARTS / blocks_per_call * blocks_per_call;
	const size_t part_smem = sizeof(float) * 2 * (size_t)cap * UPD_COLS;
	const size_t upd_smem = part_smem > lut_smem ? part_smem : lut_smem;
	prof_mark(e, 2, 0, st);
	update_kernel<<<cell_blocks + col_blocks, UPD_THREADS, upd_smem, st>>>(a, cell_blocks, cap);
	prof_mark(e, 2, 1, st);
	e->launches += 2;
	CU_CHECK(e, cudaGetLastError());
	return 0;
}

/* Two-stream schedule: order `stream` after everything issued on acc_stream.  Process calls do not
 * do this on return - the FFT of the next call may start while the last accumulate launch of this
 * one still runs - so every other consumer of the state on `stream` does it first. */
int join_accumulate(fosphor_cu *e)
{
	if (e->acc_pending) {
		CU_CHECK(e, cudaStreamWaitEvent(e->stream, e->acc_done, 0));
		e->acc_pending = false;
		e->chunk_seq = 0;
	}
	return 0;
}

int clear_buffers(fosphor_cu *e)
{
	/* cl.c:406-465: spectrum (all 4N floats) and waterfall = -power.offset
	 * (== -histo_ofs, fosphor.c:149-151), histogram = 0 */
	const size_t n = e->p.fft_len;
	const float nf = -e->histo_ofs;
	fill_kernel<<<64, 256, 0, e->stream>>>(reinterpret_cast<float *>(e->d_spec), 4 * n, nf);
	fill_kernel<<<4 * e->sm_count, 256, 0, e->stream>>>(e->d_wf, (size_t)e->p.wf_rows * n, nf);
	e->launches += 2;
	CU_CHECK(e, cudaGetLastError());
	CU_CHECK(e, cudaMemsetAsync(e->d_hist, 0, sizeof(float) * (size_t)e->p.n_bins * n, e->stream));
	return 0;
}

int validate_batch(const fosphor_cu *e, int n_spectra)
{
	/* cl.c:881-886 */
	if (n_spectra < 0 || (n_spectra % e->p.batch_mult) || n_spectra > e->p.batch_max)
		return -EINVAL;
	return 0;
}

/* n_calls calls of `batch` spectra, input already on the device */
int process_device_calls(fosphor_cu *e, const float2 *in, int n_calls, int batch, long long hop)
{
	if (validate_batch(e, batch) || n_calls < 0 || hop < 1)
		return -EINVAL;
	if (e->state == ST_BOOTING) {         /* cl.c:930-934 */
		int rc = clear_buffers(e);
		if (rc)
			return rc;
	}
	if (batch > 0 && n_calls > 0) {
		BatchTables *t;
		int rc = get_tables(e, batch, &t);       /* uploads (if new) are ordered before the first FFT */
		if (rc)
			return rc;
		/* A chunk = calls folded by one FFT + count + update launch triple.  Its rows
		 * must fit the ring and its slices the count buffer; with room for two chunks
		 * in the ring the count/update of chunk c (acc_stream) overlap the FFT of
		 * chunk c+1 (main stream). */
		const int ring_calls = e->p.wf_rows / batch;         /* >= 1: wf_rows >= batch_max */
		int calls_per_chunk = ring_calls;
		int ov_chunk = 1;
		bool two_streams = false;
		e->slim_now = false;
		if (e->overlap > 0) {                        /* forced, knobs from the environment */
			ov_chunk = e->overlap_chunk < ring_calls / 2 ? e->overlap_chunk : ring_calls / 2;
			if (ov_chunk < 1) ov_chunk = 1;
			two_streams = ring_calls >= 2 && n_calls >= 2 * ov_chunk;
			e->slim_now = two_streams && e->acc_slim;
		} else if (e->overlap < 0) {                 /* automatic, see the comment at `overlap` */
			const bool stream_fft = (e->p.fft_len == 512 || e->p.fft_len == 1024) && e->fft_variant >= 2;
			ov_chunk = ring_calls / 4;
			two_streams = stream_fft && use_fused(e, batch) && ov_chunk >= 1 && ov_chunk <= e->max_slices &&
			              (long long)ov_chunk * batch * e->p.fft_len >= (32ll << 20) && n_calls >= 2 * ov_chunk;
		}
		if (two_streams)
			calls_per_chunk = ov_chunk;
		e->two_streams_now = two_streams;
		if (calls_per_chunk > e->max_slices)
			calls_per_chunk = e->max_slices;
		if (e->chunk_calls > 0 && calls_per_chunk > e->chunk_calls)
			calls_per_chunk = e->chunk_calls;
		cudaStream_t acc = two_streams ? e->acc_stream : e->stream;
		/* a one-stream call, or another chunk geometry, first waits for what is still on acc_stream */
		if (e->acc_pending && (!two_streams || e->seq_batch != batch || e->seq_chunk_calls != calls_per_chunk)) {
			rc = join_accumulate(e);
			if (rc)
				return rc;
		}
		e->seq_batch = batch;
		e->seq_chunk_calls = calls_per_chunk;
		/* The FFT of chunk q overwrites ring rows last read by the accumulate launch of a
		 * chunk no younger than q - lag (every chunk has at most calls_per_chunk calls):
		 * waiting for that one - the acc stream is in order - frees them.  The chunk
		 * sequence runs on across process calls until something joins the streams. */
		int lag = ring_calls / calls_per_chunk;
		if (lag > N_CHUNK_EV) lag = N_CHUNK_EV;
		for (int c0 = 0; c0 < n_calls; c0 += calls_per_chunk) {
			const int nc = n_calls - c0 < calls_per_chunk ? n_calls - c0 : calls_per_chunk;
			const long long q = two_streams ? e->chunk_seq++ : 0;
			e->two_stream_chunks += two_streams ? 1 : 0;
			const int pp = (int)(q % N_CHUNK_EV);
			if (two_streams && q >= lag)
				CU_CHECK(e, cudaStreamWaitEvent(e->stream, e->cnt_done[(q - lag) % N_CHUNK_EV], 0));
			CU_CHECK(e, launch_fft(e, in + (long long)c0 * batch * hop, hop, e->wf_pos, nc * batch));
			if (two_streams) {
				CU_CHECK(e, cudaEventRecord(e->fft_done[pp], e->stream));
				CU_CHECK(e, cudaStreamWaitEvent(acc, e->fft_done[pp], 0));
			}
			rc = launch_accumulate(e, t, acc, two_streams ? e->cnt_done[pp] : nullptr, e->wf_pos, nc, batch);
			if (rc)
				return rc;
			e->wf_pos = (e->wf_pos + nc * batch) & (e->p.wf_rows - 1);   /* cl.c:954, nc times */
		}
		if (two_streams) {                       /* what a later join waits for */
			CU_CHECK(e, cudaEventRecord(e->acc_done, acc));
			e->acc_pending = true;
		}
	}
	e->state = ST_PENDING;                /* cl.c:957 */
	return 0;
}

bool is_pinned_host(const void *p)
{
	cudaPointerAttributes attr;
	if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
		cudaGetLastError();
		return false;
	}
	return attr.type == cudaMemoryTypeHost;
}

/* Host samples -> device slot.  The source buffer is free when this returns
 * (reference contract, base_sink_c_impl.cc:170-174): pageable sources are
 * copied into a pinned staging slot by the CPU; page-locked sources
 * (cudaHostAlloc / cudaHostRegister, e.g. a pinned FIFO) are DMA'd directly and
 * only that copy is waited for.  Either way the H2D runs on its own stream and
 * overlaps the kernels of the previous call. */
int upload_staged(fosphor_cu *e, const float2 *src, size_t n_samples, float2 **dev_out)
{
	const int s = e->slot;
	e->slot ^= 1;
	const size_t bytes = sizeof(float2) * n_samples;
	CU_CHECK(e, cudaEventSynchronize(e->slot_free[s]));     /* previous reader of d_in[s] done */
	if (is_pinned_host(src)) {
		/* DMA straight from the caller's page-locked memory.  The caller may recycle it as soon as the
		 * process call returns (base_sink_c_impl.cc:170-174), so the call waits for this copy - but
		 * only at its very end (await_uploads), after the kernels that consume it were enqueued behind
		 * the `copied` event: the launch overhead hides under the copy instead of following it. */
		CU_CHECK(e, cudaMemcpyAsync(e->d_in[s], src, bytes, cudaMemcpyHostToDevice, e->copy_stream));
		CU_CHECK(e, cudaEventRecord(e->copied[s], e->copy_stream));
		e->await_slot = s;
	} else {
		memcpy(e->h_in[s], src, bytes);
		CU_CHECK(e, cudaMemcpyAsync(e->d_in[s], e->h_in[s], bytes, cudaMemcpyHostToDevice, e->copy_stream));
		CU_CHECK(e, cudaEventRecord(e->copied[s], e->copy_stream));
	}
	CU_CHECK(e, cudaStreamWaitEvent(e->stream, e->copied[s], 0));
	e->last_slot = s;
	*dev_out = e->d_in[s];
	return 0;
}

/* the caller's page-locked samples have been read (the copy stream is in order: the last copy
 * issued is the last to complete) */
int await_uploads(fosphor_cu *e)
{
	if (e->await_slot >= 0) {
		const int s = e->await_slot;
		e->await_slot = -1;
		CU_CHECK(e, cudaEventSynchronize(e->copied[s]));
	}
	return 0;
}

int release_slot(fosphor_cu *e)
{
	if (e->last_slot >= 0) {
		CU_CHECK(e, cudaEventRecord(e->slot_free[e->last_slot], e->stream));
		e->last_slot = -1;
	}
	return 0;
}

} /* namespace */

/* ------------------------------------------------------------------------ */
/* C ABI                                                                     */
/* ------------------------------------------------------------------------ */

extern "C" {

void fosphor_cu_default_params(struct fosphor_cu_params *p)
{
	p->fft_len = 1024;          /* private.h:21-22 */
	p->n_bins = 128;            /* display.cl:96 */
	p->wf_rows = 1024;          /* cl.c:430-432 */
	p->batch_mult = 16;         /* private.h:24 */
	p->batch_max = 1024;        /* private.h:25 */
	p->histo_t0r = 16.0f;       /* cl.c:714 */
	p->histo_t0d = 1024.0f;     /* cl.c:715 */
	p->live_alpha = 0.002f;     /* cl.c:716 */
	p->maxhold_keep = 0.999f;   /* display.cl:303 */
	p->maxhold_mix = 0.001f;
	p->device = -1;
}

void fosphor_cu_destroy(struct fosphor_cu *e)
{
	if (!e)
		return;
	if (e->stream)
		cudaStreamSynchronize(e->stream);
	cudaFree(e->d_win); cudaFree(e->d_tw); cudaFree(e->d_wf); cudaFree(e->d_hist);
	cudaFree(e->d_spec); cudaFree(e->d_cnt); cudaFree(e->d_part_live);
	cudaFree(e->d_part_max);
	for (int i = 0; i < 2; i++) {
		cudaFreeHost(e->h_in[i]);
		cudaFree(e->d_in[i]);
		if (e->copied[i]) cudaEventDestroy(e->copied[i]);
		if (e->slot_free[i]) cudaEventDestroy(e->slot_free[i]);
	}
	if (e->copy_stream) { cudaStreamSynchronize(e->copy_stream); cudaStreamDestroy(e->copy_stream); }
	cudaFreeHost(e->h_win);
	if (e->win_done) cudaEventDestroy(e->win_done);
	for (auto &t : e->tables) {
		cudaFree(t.d_weights); cudaFree(t.d_lut); cudaFreeHost(t.h_stage);
		if (t.uploaded) cudaEventDestroy(t.uploaded);
	}
	for (int k = 0; k < 3; k++)
		for (int j = 0; j < 2; j++)
			for (cudaEvent_t ev : e->prof_ev[k][j])
				cudaEventDestroy(ev);
	if (e->acc_stream) { cudaStreamSynchronize(e->acc_stream); cudaStreamDestroy(e->acc_stream); }
	for (int i = 0; i < N_CHUNK_EV; i++) {
		if (e->fft_done[i]) cudaEventDestroy(e->fft_done[i]);
		if (e->cnt_done[i]) cudaEventDestroy(e->cnt_done[i]);
	}
	if (e->acc_done) cudaEventDestroy(e->acc_done);
	if (e->cols_fork) cudaEventDestroy(e->cols_fork);
	if (e->cols_join) cudaEventDestroy(e->cols_join);
	if (e->own_stream) cudaStreamDestroy(e->own_stream);
	delete e;
}

int fosphor_cu_create(struct fosphor_cu **out, const struct fosphor_cu_params *pp)
{
	if (!out || !pp)
		return -EINVAL;
	*out = nullptr;
	const fosphor_cu_params &p = *pp;
	if (!plan_supported(p.fft_len) || p.n_bins < 2 || p.n_bins > 4096 ||
	    p.wf_rows < 1 || (p.wf_rows & (p.wf_rows - 1)) ||
	    p.batch_mult < 1 || p.batch_max < p.batch_mult || p.batch_max % p.batch_mult || p.batch_max > 32768 ||
	    p.wf_rows < p.batch_max || !(p.histo_t0r > 0.0f) || !(p.histo_t0d > 0.0f))
		return fail(nullptr, -EINVAL, "unsupported engine parameters (N=%d K=%d W=%d batch %d/%d)",
		            p.fft_len, p.n_bins, p.wf_rows, p.batch_mult, p.batch_max);

	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
		return fail(nullptr, -ENODEV, "no CUDA device (this library has no CPU fallback)");

	fosphor_cu *e = new (std::nothrow) fosphor_cu;
	if (!e)
		return -ENOMEM;
	e->p = p;
	e->log2n = ilog2c(p.fft_len);

#define CREATE_CHECK(call)                                                         \
	do {                                                                       \
		cudaError_t err__ = (call);                                        \
		if (err__ != cudaSuccess) {                                        \
			int rc__ = fail(e, err__ == cudaErrorMemoryAllocation ? -ENOMEM : -EIO, \
			                "CUDA error %d (%s) at %s:%d: %s", (int)err__,  \
			                cudaGetErrorString(err__), __FILE__, __LINE__, #call); \
			fosphor_cu_destroy(e);                                     \
			return rc__;                                               \
		}                                                                  \
	} while (0)

	if (p.device >= 0)
		CREATE_CHECK(cudaSetDevice(p.device));
	CREATE_CHECK(cudaGetDevice(&e->device));
	cudaDeviceProp prop;
	CREATE_CHECK(cudaGetDeviceProperties(&prop, e->device));
	if (prop.major < 10) {
		fail(e, -ENODEV, "device %s is sm_%d%d; this library is built for sm_100a only",
		     prop.name, prop.major, prop.minor);
		fosphor_cu_destroy(e);
		return -ENODEV;
	}
	e->sm_count = prop.multiProcessorCount;
	e->smem_optin = prop.sharedMemPerBlockOptin;
	CREATE_CHECK(cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking));
	e->stream = e->own_stream;
	{
		/* the accumulate stream outranks the FFT stream: in the two-stream schedule the accumulate CTAs
		 * of chunk c (one per SM, most of its shared memory) must get their SMs before the FFT CTAs of
		 * chunk c+1 refill them */
		int prio_lo = 0, prio_hi = 0;
		CREATE_CHECK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
		CREATE_CHECK(cudaStreamCreateWithPriority(&e->acc_stream, cudaStreamNonBlocking, prio_hi));
	}
	for (int i = 0; i < N_CHUNK_EV; i++) {
		CREATE_CHECK(cudaEventCreateWithFlags(&e->fft_done[i], cudaEventDisableTiming));
		CREATE_CHECK(cudaEventCreateWithFlags(&e->cnt_done[i], cudaEventDisableTiming));
	}
	CREATE_CHECK(cudaEventCreateWithFlags(&e->acc_done, cudaEventDisableTiming));
	CREATE_CHECK(cudaEventCreateWithFlags(&e->cols_fork, cudaEventDisableTiming));
	CREATE_CHECK(cudaEventCreateWithFlags(&e->cols_join, cudaEventDisableTiming));
	if (const char *v = getenv("FOSPHOR_B200_OVERLAP"))
		e->overlap = atoi(v);
	if (const char *v = getenv("FOSPHOR_B200_OVERLAP_CHUNK"))
		e->overlap_chunk = atoi(v) > 0 ? atoi(v) : 1;
	if (const char *v = getenv("FOSPHOR_B200_ACC_SLIM"))
		e->acc_slim = atoi(v);
	if (const char *v = getenv("FOSPHOR_B200_ACC_ROLES"))
		e->acc_roles = atoi(v);
	if (const char *v = getenv("FOSPHOR_B200_ACC_STAGE_KB"))
		e->acc_stage_kb = atoi(v);
	if (const char *v = getenv("FOSPHOR_B200_ACC_GROUP"))
		e->acc_group = (atoi(v) == 1 || atoi(v) == 2 || atoi(v) == 4) ? atoi(v) : 0;

	const size_t n = p.fft_len, k = p.n_bins, w = p.wf_rows;
	CREATE_CHECK(cudaMalloc(&e->d_win, sizeof(float) * n));
	CREATE_CHECK(cudaMalloc(&e->d_wf, sizeof(float) * w * n));
	CREATE_CHECK(cudaMalloc(&e->d_hist, sizeof(float) * k * n));
	CREATE_CHECK(cudaMalloc(&e->d_spec, sizeof(float2) * 2 * n));
	{
		/* slices: at least what one call needs to fill the chip, at most MAX_SLICES,
		 * within CNT_BUDGET bytes of u16 counts */
		const int tiles = p.fft_len / ACC_COLS;
		int need = (2 * e->sm_count + tiles - 1) / tiles;
		if (need < 1) need = 1;
		size_t fit = CNT_BUDGET / (sizeof(unsigned short) * k * n);
		int ms = fit > (size_t)MAX_SLICES ? MAX_SLICES : (int)fit;
		if (ms < need) ms = need;
		e->max_slices = ms;
	}
	CREATE_CHECK(cudaMalloc(&e->d_cnt, sizeof(unsigned short) * k * n * e->max_slices));
	{
		const size_t blocks = (size_t)(p.batch_max + ROWBLOCK - 1) / ROWBLOCK * e->max_slices;
		CREATE_CHECK(cudaMalloc(&e->d_part_live, sizeof(float) * blocks * n));
		CREATE_CHECK(cudaMalloc(&e->d_part_max, sizeof(float) * blocks * n));
	}
	CREATE_CHECK(cudaMemset(e->d_hist, 0, sizeof(float) * k * n));
	CREATE_CHECK(cudaMemset(e->d_wf, 0, sizeof(float) * w * n));
	CREATE_CHECK(cudaMemset(e->d_spec, 0, sizeof(float2) * 2 * n));

	/* window defaults to all ones until one is loaded (cl.c leaves it undefined) */
	CREATE_CHECK(cudaMallocHost(&e->h_win, sizeof(float) * n));
	CREATE_CHECK(cudaEventCreateWithFlags(&e->win_done, cudaEventDisableTiming));
	for (size_t i = 0; i < n; i++)
		e->h_win[i] = 1.0f;
	CREATE_CHECK(cudaMemcpy(e->d_win, e->h_win, sizeof(float) * n, cudaMemcpyHostToDevice));

	{
		std::vector<float2> tw;
		if (const char *v = getenv("FOSPHOR_B200_FFT_R64"))
			e->fft_r64 = atoi(v);
		e->plan_key = p.fft_len + ((e->fft_r64 && (p.fft_len == 2048 || p.fft_len == 4096)) ? 1 : 0);
		PLAN_SWITCH(e->plan_key, build_twiddles<P>(tw));
		CREATE_CHECK(cudaMalloc(&e->d_tw, sizeof(float2) * tw.size()));
		CREATE_CHECK(cudaMemcpy(e->d_tw, tw.data(), sizeof(float2) * tw.size(), cudaMemcpyHostToDevice));
		cudaError_t perr = cudaErrorInvalidValue;
		PLAN_SWITCH(e->plan_key, (perr = plan_setup<P>()));
		CREATE_CHECK(perr);
		if (p.fft_len == 1024)
			CREATE_CHECK(stream_setup<Plan1024>());
		if (p.fft_len == 512)
			CREATE_CHECK(stream_setup<Plan512>());
		if (p.fft_len == 2048)
			CREATE_CHECK(cta_stream_setup<Plan2048>());
		if (p.fft_len == 4096)
			CREATE_CHECK(cta_stream_setup<Plan4096>());
		if (p.fft_len == 8192)
			CREATE_CHECK(cta_stream_setup<Plan8192>());
		if (const char *v = getenv("FOSPHOR_B200_FFT_VARIANT"))
			e->fft_variant = atoi(v);
		if (const char *v = getenv("FOSPHOR_B200_FFT_CTAS"))
			e->fft_ctas_per_sm = atoi(v);
		if (const char *v = getenv("FOSPHOR_B200_FFT_PF"))
			e->fft_pf = atoi(v);
	}
	{
		size_t upd = sizeof(float2) * (size_t)(p.batch_max + 1);
		if (upd > UPD_LUT_SMEM_MAX)
			upd = UPD_LUT_SMEM_MAX;       /* larger tables are read from global memory */
		const size_t parts = sizeof(float) * 2 * UPD_COLS * (size_t)((p.batch_max + ROWBLOCK - 1) / ROWBLOCK);
		if (upd < parts) upd = parts;
		if (upd < UPD_SMEM_MAX) upd = UPD_SMEM_MAX;
		CREATE_CHECK(cudaFuncSetAttribute(update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)upd));
	}
	CREATE_CHECK(cudaFuncSetAttribute(count_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
	                                  (int)(sizeof(unsigned) * 32 * k + sizeof(CountStage))));
	if (const char *v = getenv("FOSPHOR_B200_COUNT_VARIANT"))
		e->count_variant = atoi(v);
	{
		/* 2-D tensor map over the waterfall ring for the TMA-staged count kernel; the
		 * encoder lives in the driver and is fetched through the runtime */
		typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
		                              const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
		                              CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
		                              CUtensorMapFloatOOBfill);
		void *fn = nullptr;
		cudaDriverEntryPointQueryResult qres;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
		    qres == cudaDriverEntryPointSuccess && fn) {
			const cuuint64_t gdim[2] = {(cuuint64_t)n, (cuuint64_t)w};
			const cuuint64_t gstride[1] = {(cuuint64_t)n * sizeof(float)};
			const cuuint32_t box[2] = {ACC_COLS, TMA_ROWS};
			const cuuint32_t estr[2] = {1, 1};
			CUresult cr = reinterpret_cast<encode_fn>(fn)(&e->wf_tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, e->d_wf,
				gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
				CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
			e->tmap_ok = (cr == CUDA_SUCCESS) && (w % TMA_ROWS) == 0;
			if (const char *v = getenv("FOSPHOR_B200_ACC"))
				e->acc_mode = atoi(v);
			/* fused kernel: narrow tiles so that N / cols CTAs fill the chip */
			e->acc_cols = (p.fft_len <= 512 || p.n_bins > 2048) ? 4 : 8;
			if (const char *v = getenv("FOSPHOR_B200_ACC_COLS"))
				e->acc_cols = atoi(v) == 4 ? 4 : 8;
			if (const char *v = getenv("FOSPHOR_B200_CHUNK_CALLS"))
				e->chunk_calls = atoi(v);
			if (const char *v = getenv("FOSPHOR_B200_ACC_BOX")) {
				const int b = atoi(v);
				e->acc_box_max = b >= 256 ? 256 : (b >= 64 ? 64 : (b >= 16 ? 16 : 0));
			}
			if (const char *v = getenv("FOSPHOR_B200_ACC_SUB"))
				e->acc_sub_max = atoi(v) >= 64 ? 64 : 16;
			e->acc_tmap_ok = true;
			for (int i = 0; i < 3; i++) {
				const cuuint32_t abox[2] = {(cuuint32_t)e->acc_cols, (cuuint32_t)(16 << (2 * i))};
				cr = reinterpret_cast<encode_fn>(fn)(&e->acc_tmap[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, e->d_wf,
					gdim, gstride, abox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
					CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
				if (cr != CUDA_SUCCESS)
					e->acc_tmap_ok = false;
			}
		} else {
			cudaGetLastError();
		}
		if (!e->tmap_ok)
			fprintf(stderr, "[w] fosphor_b200: tensor-map encode unavailable, using the plain count kernel\n");
	}
	CREATE_CHECK(cudaFuncSetAttribute(count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
	                                  (int)(sizeof(unsigned) * 32 * k)));

	for (auto &t : e->tables) {
		CREATE_CHECK(cudaMalloc(&t.d_weights, sizeof(float) * p.batch_max));
		CREATE_CHECK(cudaMalloc(&t.d_lut, sizeof(float2) * (p.batch_max + 1)));
		CREATE_CHECK(cudaMallocHost(&t.h_stage, sizeof(float) * p.batch_max + sizeof(float2) * (p.batch_max + 1)));
		CREATE_CHECK(cudaEventCreateWithFlags(&t.uploaded, cudaEventDisableTiming));
	}

	e->stage_elems = (size_t)p.batch_max * n;
	for (int i = 0; i < 2; i++) {
		CREATE_CHECK(cudaMallocHost(&e->h_in[i], sizeof(float2) * e->stage_elems));
		CREATE_CHECK(cudaMalloc(&e->d_in[i], sizeof(float2) * e->stage_elems));
		CREATE_CHECK(cudaEventCreateWithFlags(&e->copied[i], cudaEventDisableTiming));
		CREATE_CHECK(cudaEventCreateWithFlags(&e->slot_free[i], cudaEventDisableTiming));
	}
	CREATE_CHECK(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
#undef CREATE_CHECK

	*out = e;
	return 0;
}

int fosphor_cu_set_stream(struct fosphor_cu *e, void *cuda_stream)
{
	if (!e)
		return -EINVAL;
	if (int rc = join_accumulate(e))
		return rc;
	CU_CHECK(e, cudaStreamSynchronize(e->stream));
	e->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : e->own_stream;
	return 0;
}

int fosphor_cu_load_fft_window(struct fosphor_cu *e, const float *win_host)
{
	if (!e || !win_host)
		return -EINVAL;
	CU_CHECK(e, cudaEventSynchronize(e->win_done));
	memcpy(e->h_win, win_host, sizeof(float) * e->p.fft_len);
	CU_CHECK(e, cudaMemcpyAsync(e->d_win, e->h_win, sizeof(float) * e->p.fft_len,
	                            cudaMemcpyHostToDevice, e->stream));
	CU_CHECK(e, cudaEventRecord(e->win_done, e->stream));
	return 0;
}

int fosphor_cu_set_histogram_range(struct fosphor_cu *e, float scale, float offset)
{
	if (!e)
		return -EINVAL;
	e->histo_scale = scale * (float)e->p.n_bins;    /* cl.c:1087 */
	e->histo_ofs = offset;
	return 0;
}

int fosphor_cu_process_device(struct fosphor_cu *e, const void *samples_dev,
                              int n_spectra, long long hop)
{
	if (!e || (!samples_dev && n_spectra > 0))
		return -EINVAL;
	return process_device_calls(e, static_cast<const float2 *>(samples_dev), 1, n_spectra, hop);
}

int fosphor_cu_process_device_multi(struct fosphor_cu *e, const void *samples_dev,
                                    int n_calls, int batch, long long hop)
{
	if (!e || (!samples_dev && n_calls > 0 && batch > 0))
		return -EINVAL;
	return process_device_calls(e, static_cast<const float2 *>(samples_dev), n_calls, batch, hop);
}

int fosphor_cu_process_host(struct fosphor_cu *e, const void *samples_host, int len)
{
	if (!e)
		return -EINVAL;
	const int n = e->p.fft_len;
	/* cl.c:881-886 */
	if (len < 0 || (len % (e->p.batch_mult * n)) || len > e->p.batch_max * n)
		return -EINVAL;
	if (len > 0 && !samples_host)
		return -EINVAL;
	float2 *dev = nullptr;
	if (len > 0) {
		int rc = upload_staged(e, static_cast<const float2 *>(samples_host), (size_t)len, &dev);
		if (rc)
			return rc;
	}
	int rc = process_device_calls(e, dev, 1, len / n, n);
	if (!rc)
		rc = release_slot(e);
	const int rc2 = await_uploads(e);     /* also on the error path: the source is the caller's again on return */
	return rc ? rc : rc2;
}

int fosphor_cu_process_host_raw(struct fosphor_cu *e, const void *raw_host,
                                int n_calls, int batch, long long hop)
{
	if (!e || validate_batch(e, batch) || n_calls < 0 || hop < 1 || hop > e->p.fft_len)
		return -EINVAL;
	if (n_calls == 0 || batch == 0)
		return process_device_calls(e, nullptr, n_calls, batch, hop);
	if (!raw_host)
		return -EINVAL;
	const float2 *raw = static_cast<const float2 *>(raw_host);
	const long long n = e->p.fft_len;
	/* calls per staged chunk: (c*batch - 1)*hop + N <= stage_elems */
	long long cpc = (((long long)e->stage_elems - n) / hop + 1) / batch;
	if (cpc < 1)
		cpc = 1;    /* batch*hop <= batch_max*N always fits one call */
	for (long long c0 = 0; c0 < n_calls; c0 += cpc) {
		const long long nc = n_calls - c0 < cpc ? n_calls - c0 : cpc;
		const size_t samples = (size_t)((nc * batch - 1) * hop + n);
		float2 *dev = nullptr;
		int rc = upload_staged(e, raw + c0 * batch * hop, samples, &dev);
		if (!rc)
			rc = process_device_calls(e, dev, (int)nc, batch, hop);
		if (!rc)
			rc = release_slot(e);
		if (rc) {
			await_uploads(e);
			return rc;
		}
	}
	return await_uploads(e);              /* the whole raw buffer stays the caller's until here: copies run back to back */
}

int fosphor_cu_finish(struct fosphor_cu *e, float *waterfall_host,
                      float *histogram_host, float *spectrum_host)
{
	if (!e)
		return -EINVAL;
	if (e->state == ST_READY)             /* cl.c:978-979 */
		return 0;
	if (e->state == ST_BOOTING) {         /* cl.c:982-994 */
		int rc = clear_buffers(e);
		if (rc)
			return rc;
	}
	if (int rc = join_accumulate(e))
		return rc;
	const size_t n = e->p.fft_len;
	/* cl.c:1012-1048 */
	if (waterfall_host)
		CU_CHECK(e, cudaMemcpyAsync(waterfall_host, e->d_wf, sizeof(float) * e->p.wf_rows * n,
		                            cudaMemcpyDeviceToHost, e->stream));
	if (histogram_host)
		CU_CHECK(e, cudaMemcpyAsync(histogram_host, e->d_hist, sizeof(float) * e->p.n_bins * n,
		                            cudaMemcpyDeviceToHost, e->stream));
	if (spectrum_host)
		CU_CHECK(e, cudaMemcpyAsync(spectrum_host, e->d_spec, sizeof(float2) * 2 * n,
		                            cudaMemcpyDeviceToHost, e->stream));
	CU_CHECK(e, cudaStreamSynchronize(e->stream));   /* cl.c:1052 */
	e->state = ST_READY;
	return 1;
}

int fosphor_cu_sync(struct fosphor_cu *e)
{
	if (!e)
		return -EINVAL;
	if (int rc = join_accumulate(e))
		return rc;
	CU_CHECK(e, cudaStreamSynchronize(e->stream));
	return 0;
}

unsigned long long fosphor_cu_two_stream_chunks(const struct fosphor_cu *e)
{
	return e ? e->two_stream_chunks : 0;
}

int fosphor_cu_flush(struct fosphor_cu *e)
{
	if (!e)
		return -EINVAL;
	return join_accumulate(e);
}

int fosphor_cu_get_waterfall_position(const struct fosphor_cu *e)
{
	return e ? e->wf_pos : -EINVAL;       /* cl.c:1073-1079 */
}

float *fosphor_cu_device_waterfall(struct fosphor_cu *e) { return e ? e->d_wf : nullptr; }
float *fosphor_cu_device_histogram(struct fosphor_cu *e) { return e ? e->d_hist : nullptr; }
float *fosphor_cu_device_spectrum(struct fosphor_cu *e)
{
	return e ? reinterpret_cast<float *>(e->d_spec) : nullptr;
}

int fosphor_cu_export_maxhold(struct fosphor_cu *e, float *out_dev)
{
	if (!e || !out_dev)
		return -EINVAL;
	const int n = e->p.fft_len;
	if (int rc = join_accumulate(e))
		return rc;
	export_maxhold_kernel<<<(n + 255) / 256, 256, 0, e->stream>>>(e->d_spec, n, out_dev);
	e->launches++;
	CU_CHECK(e, cudaGetLastError());
	return 0;
}

int fosphor_cu_debug_fft(struct fosphor_cu *e, const void *samples_dev,
                         int n_spectra, long long hop, void *out_dev)
{
	if (!e || !samples_dev || !out_dev || n_spectra < 0 || hop < 1)
		return -EINVAL;
	cudaError_t err = cudaErrorInvalidValue;
	PLAN_SWITCH(e->plan_key, (err = plan_launch<P, true>(e, static_cast<const float2 *>(samples_dev), hop, 0,
	                                                      static_cast<float2 *>(out_dev), n_spectra)));
	CU_CHECK(e, err);
	return 0;
}

int fosphor_cu_profile(struct fosphor_cu *e, int enable)
{
	if (!e)
		return -EINVAL;
	if (int rc = join_accumulate(e))
		return rc;
	CU_CHECK(e, cudaStreamSynchronize(e->stream));
	e->profiling = enable != 0;
	e->prof_used[0] = e->prof_used[1] = e->prof_used[2] = 0;
	return 0;
}

int fosphor_cu_profile_read(struct fosphor_cu *e, double *ms_out, unsigned long long *launches_out)
{
	if (!e)
		return -EINVAL;
	if (int rc = join_accumulate(e))
		return rc;
	CU_CHECK(e, cudaStreamSynchronize(e->stream));
	double ms[3] = {0.0, 0.0, 0.0};
	for (int k = 0; k < 3; k++)
		for (size_t i = 0; i < e->prof_used[k]; i++) {
			float t = 0.0f;
			CU_CHECK(e, cudaEventElapsedTime(&t, e->prof_ev[k][0][i], e->prof_ev[k][1][i]));
			ms[k] += t;
		}
	for (int k = 0; k < 3; k++) {
		if (ms_out) ms_out[k] = ms[k];
		if (launches_out) launches_out[k] = e->prof_used[k];
		e->prof_used[k] = 0;
	}
	return 0;
}

unsigned long long fosphor_cu_launch_count(const struct fosphor_cu *e)
{
	return e ? e->launches : 0;
}

const char *fosphor_cu_last_error(const struct fosphor_cu *e)
{
	return e ? e->err : "null engine";
}

} /* extern "C" */
