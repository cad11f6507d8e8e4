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
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <unordered_map>
#include <vector>

#include <fcntl.h>
#include <unistd.h>

#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/fosphor_b200.h"
#include "../host/copy_pool.h"
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
constexpr int MAX_STAGE_SLOTS = 8;               /* host-sample staging slots (page-locked + device) */
constexpr size_t STAGE_SLOT_SMALL = (size_t)32 << 20;   /* slots up to this size: 4 of them, larger: 2 */
constexpr size_t SCRATCH_BUDGET = (size_t)1 << 30;   /* automatic log-power scratch ring: at most this many bytes ... */
constexpr int OUT_CHUNKS = 16;                   /* result read-back: D2H / copy-out pipeline depth */
constexpr size_t HOSTREG_BUDGET = (size_t)256 << 20; /* caller memory page-locked on the fly (opt-in), at most */

/* Experiment knobs.  Every default is the measured best; the environment is read in
 * ONE place (tuning_from_env) at create so that tools/ab_accumulate.py and the bit-identity test can
 * reach every transport variant.  None of them changes a result bit except acc_mode = 0 (the split
 * kernels add the live spectrum in canonical 128-row blocks: within the parity tolerance). */
struct Tuning {
	int overlap = -1;          /* OVERLAP: two-stream schedule, -1 automatic / 0 off / 1 forced */
	int overlap_chunk = 16;    /* OVERLAP_CHUNK: calls per chunk when forced */
	int acc_slim = 1;          /* ACC_SLIM: forced mode uses the 14-warp co-resident accumulate CTA */
	int acc_roles = 0;         /* ACC_ROLES: counter / updater warps 1 = 16/8, 2 = 8/16; 0 = by shape */
	int acc_stage_kb = 0;      /* ACC_STAGE_KB: cap of the fused kernel's stage ring (0 = what fits) */
	int acc_group = 0;         /* ACC_GROUP: calls per counter -> updater hand-over (1 | 2 | 4); 0 = by batch */
	int acc_mode = -1;         /* ACC: 1 fused, 0 split kernels, -1 by shape */
	int acc_cols = 0;          /* ACC_COLS: columns per fused CTA (4 | 8); 0 = by shape */
	int acc_box_max = 256;     /* ACC_BOX: largest TMA box in rows (0 plain loads | 16 | 64 | 256) */
	int acc_sub_max = 64;      /* ACC_SUB: rows per unrolled body (16 | 64) */
	int chunk_calls = 0;       /* CHUNK_CALLS: cap on the calls folded per launch (0 = ring) */
	int count_variant = 1;     /* COUNT_VARIANT: 1 TMA-staged count kernel, 0 plain (split path) */
	int fft_variant = 2;       /* FFT_VARIANT: 0 plain kernels; 1 TMA stream (N <= 1024); 2 TMA stream with twiddles in
	                            * registers (N <= 1024), half-staged persistent kernel (16384); 3 as 2 + CTA-level streaming
	                            * for 2048..8192; 4 as 2 but the grouped kernel (warp-local late passes) for 8192 / 16384 */
	int fft_r64 = 0;           /* FFT_R64: two-pass radix-64 plans for N = 2048 / 4096 */
	int fft_pf = -1;           /* FFT_PF: L2 prefetch distance of the plain FFT kernel (-1 = resident CTAs) */
	int fft_ctas_per_sm = 0;   /* FFT_CTAS: CTAs/SM of the persistent FFT kernel (0 = automatic) */
	int hostreg = -1;          /* HOSTREG: page-lock pageable caller buffers where they lie (see hostreg_cover):
	                            * 1 on first sight, trusting the caller to keep them alive; 0 never (always stage);
	                            * -1 automatic: on SECOND sight, and only where the physical page numbers can be read
	                            * (/proc/self/pagemap, i.e. a privileged process) so that every use can be checked
	                            * against a buffer that was freed and reallocated meanwhile */
	int copy_threads = 0;      /* COPY_THREADS: staging-copy threads (0 = automatic) */
	int copy_nt = 0;           /* COPY_NT: non-temporal stores into the staging slots.  Off: with plain stores the
	                            * slots (4 x 8 MiB) stay in the host's last-level cache and the DMA engine reads them
	                            * there (calls only: 6.29 vs 5.70 Gsamples/s, tools/e2e_probe.py) */
	int stage_piece_kb = 4096; /* STAGE_PIECE_KB: pageable sources are staged and DMA'd in pieces of about this size
	                            * (1 / 2 / 4 / 8 MiB: 4.88 / 5.15 / 5.30 / 5.31 Gsamples/s with a finish per frame) */
	int stage_slots = 0;       /* STAGE_SLOTS: staging slots (0 = 4, or 2 when a slot is larger than 32 MiB) */
	int l2_hints = 0;          /* L2_HINTS: N <= 1024 stream kernel reads samples evict-first and writes log-power rows
	                            * evict-last; the fused accumulate kernel reads them evict-first */
};

int env_int(const char *name, int dflt)
{
	char key[64];
	snprintf(key, sizeof(key), "FOSPHOR_B200_%s", name);
	const char *v = getenv(key);
	return v ? atoi(v) : dflt;
}

Tuning tuning_from_env()
{
	Tuning t;
	t.overlap = env_int("OVERLAP", t.overlap);
	t.overlap_chunk = env_int("OVERLAP_CHUNK", t.overlap_chunk);
	if (t.overlap_chunk < 1) t.overlap_chunk = 1;
	t.acc_slim = env_int("ACC_SLIM", t.acc_slim);
	t.acc_roles = env_int("ACC_ROLES", t.acc_roles);
	t.acc_stage_kb = env_int("ACC_STAGE_KB", t.acc_stage_kb);
	t.acc_group = env_int("ACC_GROUP", 0);
	if (t.acc_group != 1 && t.acc_group != 2 && t.acc_group != 4) t.acc_group = 0;
	t.acc_mode = env_int("ACC", t.acc_mode);
	t.acc_cols = env_int("ACC_COLS", 0);
	if (t.acc_cols != 4 && t.acc_cols != 8 && t.acc_cols != 16) t.acc_cols = 0;
	{
		const int b = env_int("ACC_BOX", t.acc_box_max);
		t.acc_box_max = b >= 256 ? 256 : (b >= 64 ? 64 : (b >= 16 ? 16 : 0));
	}
	t.acc_sub_max = env_int("ACC_SUB", t.acc_sub_max) >= 64 ? 64 : 16;
	t.chunk_calls = env_int("CHUNK_CALLS", t.chunk_calls);
	t.count_variant = env_int("COUNT_VARIANT", t.count_variant);
	t.fft_variant = env_int("FFT_VARIANT", t.fft_variant);
	t.fft_r64 = env_int("FFT_R64", t.fft_r64);
	t.fft_pf = env_int("FFT_PF", t.fft_pf);
	t.fft_ctas_per_sm = env_int("FFT_CTAS", t.fft_ctas_per_sm);
	t.hostreg = env_int("HOSTREG", t.hostreg);
	t.copy_threads = env_int("COPY_THREADS", t.copy_threads);
	t.copy_nt = env_int("COPY_NT", t.copy_nt);
	t.stage_piece_kb = env_int("STAGE_PIECE_KB", t.stage_piece_kb);
	if (t.stage_piece_kb < 64) t.stage_piece_kb = 64;
	t.stage_slots = env_int("STAGE_SLOTS", t.stage_slots);
	t.l2_hints = env_int("L2_HINTS", t.l2_hints) != 0;
	if (t.stage_slots < 0 || t.stage_slots > MAX_STAGE_SLOTS) t.stage_slots = 0;
	return t;
}

struct HostRange {             /* caller memory page-locked by the engine */
	uintptr_t lo, hi;
	std::vector<uint64_t> pfn; /* physical page numbers at registration (automatic mode), one per 4 KiB page */
};

/* Every range any engine of this process has page-locked.  cudaPointerGetAttributes cannot tell
 * "page-locked by the caller" from "page-locked by another engine of this library" - and only the
 * engine that registered a range can vouch for it (it holds the page numbers).  An engine that
 * meets somebody else's registration stages instead of trusting it. */
struct GlobalReg {
	uintptr_t lo, hi;
	const void *owner;
};
std::mutex g_reg_mutex;
std::vector<GlobalReg> g_regs;

void global_reg_add(const void *owner, uintptr_t lo, uintptr_t hi)
{
	std::lock_guard<std::mutex> lk(g_reg_mutex);
	g_regs.push_back({lo, hi, owner});
}

void global_reg_remove(const void *owner, uintptr_t lo)
{
	std::lock_guard<std::mutex> lk(g_reg_mutex);
	for (size_t i = 0; i < g_regs.size(); i++)
		if (g_regs[i].owner == owner && g_regs[i].lo == lo) {
			g_regs.erase(g_regs.begin() + (long)i);
			return;
		}
}

bool global_reg_foreign(const void *self, uintptr_t a, size_t bytes)
{
	std::lock_guard<std::mutex> lk(g_reg_mutex);
	for (const GlobalReg &r : g_regs)
		if (r.owner != self && a < r.hi && a + bytes > r.lo)
			return true;
	return false;
}

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
	Tuning tn;
	int log2n = 0;
	int device = 0;
	int sm_count = 0;
	size_t smem_optin = 0;               /* largest dynamic shared memory a CTA may ask for */
	/* cudaFuncSetAttribute / occupancy results are per DEVICE: kept per engine, never in statics
	 * (several engines on several GPUs may live in one process) */
	std::unordered_map<const void *, size_t> func_smem;
	std::unordered_map<const void *, int> func_per_sm;

	cudaStream_t own_stream = nullptr;
	cudaStream_t stream = nullptr;       /* FFT kernel, copies to the host, everything the caller orders against */
	cudaStream_t acc_stream = nullptr;   /* accumulate kernels of the two-stream schedule: overlap the next chunk's FFT */
	cudaEvent_t fft_done[N_CHUNK_EV] = {};   /* per chunk, round robin */
	cudaEvent_t cnt_done[N_CHUNK_EV] = {};
	cudaEvent_t acc_done = nullptr;
	cudaEvent_t side_ev = nullptr;       /* hand-over to / from a caller's side stream (export_maxhold_on) */
	bool side_pending = false;           /* a side stream still reads the spectrum state: the next accumulate waits for side_ev */
	/* Two-stream schedule (tn.overlap): the accumulate kernel of chunk c runs on a second, higher
	 * priority stream while the FFT of chunk c+1 runs.  Automatic (-1): on when the log-power ring
	 * holds four chunks of >= 32 M samples (N = 512 / 1024 streaming FFT + fused accumulate); chunk =
	 * ring / 4, full-size accumulate CTAs, FFT at 3 CTAs/SM.  The accumulate CTAs take 128 SMs, the
	 * next FFT starts on the 20 that are left and on every SM an accumulate CTA leaves: cfg2,
	 * 256-call ring: 385 vs 345 Gsamples/s.  Smaller rings lose: chunks of 16 calls pay the launch
	 * ramp/tail 4x as often (64-call ring: 327 vs 345). */
	bool two_streams_now = false;        /* set per process call */
	bool acc_pending = false;            /* accumulate work on acc_stream that `stream` has not been ordered after yet */
	long long chunk_seq = 0;             /* chunks issued by the two-stream schedule since the last join */
	unsigned long long two_stream_chunks = 0;   /* ... since create (diagnostics) */
	int seq_batch = 0, seq_chunk_calls = 0;   /* ... and their geometry (a change forces a join) */
	bool slim_now = false;               /* ... two-stream mode with the slim co-resident accumulate CTA */
	int plan_key = 0;                    /* fft_len, +1 for the radix-64 plan */
	int acc_cols = 8;                    /* columns per CTA of the fused kernel (by shape, or tn.acc_cols) */
	CUtensorMap wf_tmap;                 /* log-power ring as a 2-D tensor, box = 16 rows x 32 columns (split path) */
	bool tmap_ok = false;
	CUtensorMap acc_tmap[3];             /* log-power ring, box = 16 / 64 / 256 rows x acc_cols columns */
	bool acc_tmap_ok = false;
	void *tmap_encode = nullptr;         /* cuTensorMapEncodeTiled, fetched through the runtime */

	float *d_win = nullptr;
	float2 *d_tw = nullptr;
	float2 *d_twg = nullptr;             /* grouped FFT kernel: pass-2 twiddles [group][t][lane] (N = 8192, 16384) */
	/* The user-visible waterfall (W rows, the reference's 1024, cl.c:430-432) and the ring the
	 * kernels work in are two things: the FFT kernel writes log-power rows into d_ring (ring_rows
	 * >= W, a power of two) and the accumulate kernel reads them there.  How many calls one launch
	 * pair can fold - what amortises ramp, tail and state-tile traffic - depends on ring_rows only;
	 * W stays whatever the display wants.  While ring_rows == W the ring IS the waterfall (no
	 * copy); a deeper scratch ring is allocated the first time a multi-call launch asks for it
	 * (p.scratch_rows), and publish() then copies the rows written since the last publish (at most W)
	 * into the waterfall when somebody looks (finish / flush). */
	float *d_wf = nullptr;               /* [W][N] */
	float *d_ring = nullptr;             /* [ring_rows][N]; == d_wf while no scratch ring exists */
	int ring_rows = 0;
	long long scratch_limit = 0;         /* rows a scratch ring may grow to (0 = not determined yet) */
	int ring_pos = 0;                    /* ring row of the next spectrum */
	long long unpublished = 0;           /* rows written into a scratch ring since the last publish */
	float *d_hist = nullptr;
	float2 *d_spec = nullptr;
	unsigned short *d_cnt = nullptr;     /* [max_slices][K][N]: split path only, allocated on first use */
	float *d_part_live = nullptr, *d_part_max = nullptr;   /* [max_slices][N] */
	int max_slices = 0;

	/* host-sample staging (fosphor_cu_process_host*), allocated on first use */
	size_t stage_elems = 0;              /* complex samples per slot */
	int n_slots = 0;
	float2 *h_in[MAX_STAGE_SLOTS] = {};
	float2 *d_in[MAX_STAGE_SLOTS] = {};
	cudaStream_t copy_stream = nullptr;  /* H2D of samples, overlaps the compute stream */
	cudaEvent_t copied[MAX_STAGE_SLOTS] = {};     /* H2D into slot done (copy stream)   */
	cudaEvent_t slot_free[MAX_STAGE_SLOTS] = {};  /* kernels that read the slot done    */
	int slot = 0;
	int last_slot = -1;
	int await_slot = -1;                 /* slot whose H2D copy from CALLER memory the current call still has to wait for */
	std::unique_ptr<copy_pool> pool;     /* staging-copy threads, started on the first pageable call */
	std::vector<HostRange> hostreg;      /* caller memory page-locked on the fly (tn.hostreg) */
	size_t hostreg_bytes = 0;
	std::vector<HostRange> seen;         /* automatic mode: pageable call ranges staged recently (second sight promotes) */
	int pagemap_fd = -2;                 /* /proc/self/pagemap: -2 not tried, -1 unusable (no PFNs for this process) */
	int hostreg_stale = 0;               /* registrations found stale: after a few the engine stops registering */
	unsigned long long staged_calls = 0, direct_calls = 0;   /* diagnostics: how the host samples travelled */
	/* results on their way to pageable caller memory (finish): D2H into page-locked memory, then the copy pool */
	float *h_out = nullptr;
	size_t h_out_bytes = 0;
	std::vector<cudaEvent_t> out_ev;
	long long rows_since_finish = -1;    /* waterfall rows written since the last finish; -1 = everything is new */
	float *h_win = nullptr;              /* pinned copy of the window */
	cudaEvent_t win_done = nullptr;

	float histo_scale = 0.0f, histo_ofs = 0.0f;   /* cl.c:811 memset, :1087-1088 */
	int wf_pos = 0;
	int state = ST_BOOTING;

	BatchTables tables[N_TABLES];
	unsigned long long use_clock = 0;
	unsigned long long launches = 0;

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

/* Every C-ABI entry runs with the engine's device current and restores the caller's on return:
 * several engines on several GPUs may be driven from one thread (the reference allows many sink
 * instances per process, lib/base_sink_c_impl.cc:46,97). */
struct DevGuard {
	int prev = -1;
	bool switched = false;
	explicit DevGuard(const fosphor_cu *e) : DevGuard(e ? e->device : -1) {}
	explicit DevGuard(int dev)
	{
		if (dev < 0 || cudaGetDevice(&prev) != cudaSuccess)
			return;
		if (prev != dev)
			switched = cudaSetDevice(dev) == cudaSuccess;
	}
	~DevGuard()
	{
		if (switched)
			cudaSetDevice(prev);
	}
	DevGuard(const DevGuard &) = delete;
	DevGuard &operator=(const DevGuard &) = delete;
};

/* opt a kernel in to `bytes` of dynamic shared memory on THIS engine's device (the attribute is per
 * device; remembered per engine) */
template <class K>
cudaError_t ensure_smem(fosphor_cu *e, K kernel, size_t bytes)
{
	const void *key = reinterpret_cast<const void *>(kernel);
	auto it = e->func_smem.find(key);
	if (it != e->func_smem.end() && it->second >= bytes)
		return cudaSuccess;
	cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
	if (err == cudaSuccess)
		e->func_smem[key] = bytes;
	return err;
}

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

/* pass-2 twiddles of the grouped kernel: the very values of the table above, tw[TW1 + t*P2 + k2]
 * with k2 = R0*lane + g, laid out [g][t][lane] so that a warp reads 256 consecutive bytes */
template <class P>
void build_grouped_twiddles(std::vector<float2> &twg)
{
	constexpr int R0 = P::R0;
	twg.resize((size_t)R0 * 32 * 32);
	for (int g = 0; g < R0; g++)
		for (int t = 0; t < 32; t++)
			for (int l = 0; l < 32; l++) {
				const int k = R0 * l + g;
				const double a = -2.0 * M_PI * (double)t * (double)k / (double)P::N;
				twg[((size_t)g * 32 + t) * 32 + l] = make_float2((float)cos(a), (float)sin(a));
			}
}

template <class P>
cudaError_t grouped_launch(fosphor_cu *e, const float2 *in, long long hop, int wf_pos, int n_spectra)
{
	using C = GroupedCfg<P>;
	if (cudaError_t err = ensure_smem(e, fft_power_grouped_kernel<P>, C::SMEM))
		return err;
	const int units = (n_spectra + C::SPC - 1) / C::SPC;
	const int grid = units < e->sm_count ? units : e->sm_count;      /* persistent, one CTA per SM */
	const bool aligned = ((reinterpret_cast<unsigned long long>(in) & 15ull) == 0) && ((hop & 1) == 0);
	prof_mark(e, 0, 0);
	fft_power_grouped_kernel<P><<<grid, C::THREADS, C::SMEM, e->stream>>>(
		in, hop, e->d_win, e->d_tw, e->d_twg, e->d_ring, wf_pos, e->ring_rows - 1, n_spectra, aligned ? 1 : 0);
	prof_mark(e, 0, 1);
	e->launches++;
	return cudaGetLastError();
}

template <class P>
cudaError_t plan_setup(fosphor_cu *e)
{
	cudaError_t err = ensure_smem(e, fft_power_kernel<P, false>, P::SMEM);
	if (err != cudaSuccess)
		return err;
	return ensure_smem(e, fft_power_kernel<P, true>, P::SMEM);
}

template <class P, bool CPLX>
cudaError_t plan_launch(fosphor_cu *e, const float2 *in, long long hop, int wf_pos,
                        float2 *cplx_out, int n_spectra)
{
	const int grid = (n_spectra + P::SPB - 1) / P::SPB;
	/* L2 prefetch distance of the one-spectrum-per-CTA plans: the CTAs resident at once */
	int pf = 0;
	const bool aligned = ((reinterpret_cast<unsigned long long>(in) & 15ull) == 0) && ((hop & 1) == 0);
	if (P::SPB == 1 && aligned && e->tn.fft_pf != 0) {
		if (e->tn.fft_pf > 0) {
			pf = e->tn.fft_pf;
		} else {
			int &per_sm = e->func_per_sm[reinterpret_cast<const void *>(fft_power_kernel<P, CPLX>)];
			if (per_sm == 0 &&
			    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fft_power_kernel<P, CPLX>,
			                                                  P::THREADS, P::SMEM) != cudaSuccess)
				per_sm = 1;
			pf = e->sm_count * (per_sm > 0 ? per_sm : 1);
		}
	}
	prof_mark(e, 0, 0);
	fft_power_kernel<P, CPLX><<<grid, P::THREADS, P::SMEM, e->stream>>>(
		in, hop, e->d_win, e->d_tw, e->d_ring, wf_pos, e->ring_rows - 1, cplx_out, n_spectra, pf);
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
cudaError_t stream_setup(fosphor_cu *e)
{
	cudaError_t err = ensure_smem(e, fft_power_stream_kernel<P, false>, StreamCfg<P>::SMEM);
	if (err != cudaSuccess)
		return err;
	return ensure_smem(e, fft_power_stream_kernel<P, true>, StreamCfg<P>::SMEM_TWREG);
}

template <class P>
cudaError_t stream_launch(fosphor_cu *e, const float2 *in, long long hop, int wf_pos, int n_spectra)
{
	using C = StreamCfg<P>;
	const int units = (n_spectra + C::SPW - 1) / C::SPW;   /* a warp takes SPW spectra per iteration */
	int grid = (units + C::WARPS - 1) / C::WARPS;
	int per_sm = C::CTAS_PER_SM;
	if (e->tn.fft_ctas_per_sm > 0 && e->tn.fft_ctas_per_sm < per_sm)
		per_sm = e->tn.fft_ctas_per_sm;
	else if (e->tn.fft_ctas_per_sm == 0 && e->slim_now && per_sm > 2)
		per_sm = 2;                      /* leave room for the count kernel on the other stream */
	const int resident = e->sm_count * per_sm;
	if (grid > resident)
		grid = resident;                 /* persistent warps, grid-stride over spectra */
	prof_mark(e, 0, 0);
	if (e->tn.fft_variant >= 2)
		fft_power_stream_kernel<P, true><<<grid, C::THREADS, C::SMEM_TWREG, e->stream>>>(
			in, hop, e->d_win, e->d_tw, e->d_ring, wf_pos, e->ring_rows - 1, n_spectra, e->tn.l2_hints);
	else
		fft_power_stream_kernel<P, false><<<grid, C::THREADS, C::SMEM, e->stream>>>(
			in, hop, e->d_win, e->d_tw, e->d_ring, wf_pos, e->ring_rows - 1, n_spectra, e->tn.l2_hints);
	prof_mark(e, 0, 1);
	e->launches++;
	return cudaGetLastError();
}

template <class P>
cudaError_t cta_stream_setup(fosphor_cu *e)
{
	return ensure_smem(e, fft_power_cta_stream_kernel<P>, CtaStreamCfg<P>::SMEM);
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
		in, hop, e->d_win, e->d_tw, e->d_ring, wf_pos, e->ring_rows - 1, n_spectra);
	prof_mark(e, 0, 1);
	e->launches++;
	return cudaGetLastError();
}

template <class P>
cudaError_t half_stage_launch(fosphor_cu *e, const float2 *in, long long hop, int wf_pos, int n_spectra)
{
	using C = HalfStageCfg<P>;
	if (cudaError_t err = ensure_smem(e, fft_power_half_stage_kernel<P>, C::SMEM))
		return err;
	const int grid = n_spectra < e->sm_count ? n_spectra : e->sm_count;   /* persistent, one CTA per SM */
	prof_mark(e, 0, 0);
	fft_power_half_stage_kernel<P><<<grid, C::THREADS, C::SMEM, e->stream>>>(
		in, hop, e->d_win, e->d_tw, e->d_ring, wf_pos, e->ring_rows - 1, n_spectra);
	prof_mark(e, 0, 1);
	e->launches++;
	return cudaGetLastError();
}

cudaError_t launch_fft(fosphor_cu *e, const float2 *in, long long hop, int wf_pos, int n_spectra)
{
	/* N = 8192 / 16384, experiment variant 4: warp-local late passes (plain loads: any alignment).
	 * Measured SLOWER than the kernels it was meant to replace (N = 16384, r = 1: 1126 us per 16384
	 * spectra vs 926 half-staged / 947 plain; N = 8192: 1093 vs 755 us per 32768): fewer barriers and no
	 * bank conflicts do not help a kernel whose bound is the l1tex pipe (64 B per sample through
	 * LDS/STS/LDG either way) - and one phase-locked CTA per SM overlaps less than three plain ones. */
	if (e->d_twg && e->tn.fft_variant == 4) {
		if (e->p.fft_len == 16384)
			return grouped_launch<Plan16384>(e, in, hop, wf_pos, n_spectra);
		if (e->p.fft_len == 8192)
			return grouped_launch<Plan8192>(e, in, hop, wf_pos, n_spectra);
	}
	/* TMA bulk copies need 16-byte aligned spectra */
	const bool aligned = ((reinterpret_cast<unsigned long long>(in) & 15ull) == 0) && ((hop & 1) == 0);
	if (aligned && e->tn.fft_variant != 0) {
		if (e->p.fft_len == 1024)
			return stream_launch<Plan1024>(e, in, hop, wf_pos, n_spectra);
		if (e->p.fft_len == 512)
			return stream_launch<Plan512>(e, in, hop, wf_pos, n_spectra);
		if (e->p.fft_len == 16384 && e->tn.fft_variant >= 2 && e->tn.fft_variant != 4)
			return half_stage_launch<Plan16384>(e, in, hop, wf_pos, n_spectra);
		/* The CTA-level streaming kernel measured SLOWER than the plain one (N = 4096:
		 * 101 vs 86 us per 8192 spectra; fewer resident CTAs outweigh the prefetch), so
		 * it is only reachable as experiment variant 3. */
		if (e->tn.fft_variant == 3 && e->plan_key == e->p.fft_len) {
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

/* ---- log-power ring, tensor maps, split-path buffers ------------------------ */

/* 2-D tensor maps over the log-power ring (re-encoded whenever the ring is replaced) */
void encode_tensor_maps(fosphor_cu *e)
{
	typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
	                              const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
	                              CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
	                              CUtensorMapFloatOOBfill);
	e->tmap_ok = e->acc_tmap_ok = false;
	if (!e->tmap_encode)
		return;
	encode_fn enc = reinterpret_cast<encode_fn>(e->tmap_encode);
	const cuuint64_t gdim[2] = {(cuuint64_t)e->p.fft_len, (cuuint64_t)e->ring_rows};
	const cuuint64_t gstride[1] = {(cuuint64_t)e->p.fft_len * sizeof(float)};
	const cuuint32_t box[2] = {ACC_COLS, TMA_ROWS};
	const cuuint32_t estr[2] = {1, 1};
	CUresult cr = enc(&e->wf_tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, e->d_ring, gdim, gstride, box, estr,
	                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
	                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	e->tmap_ok = (cr == CUDA_SUCCESS) && (e->ring_rows % TMA_ROWS) == 0;
	e->acc_tmap_ok = true;
	for (int i = 0; i < 3; i++) {
		const cuuint32_t abox[2] = {(cuuint32_t)e->acc_cols, (cuuint32_t)(16 << (2 * i))};
		cr = enc(&e->acc_tmap[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, e->d_ring, gdim, gstride, abox, estr,
		         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
		         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
		if (cr != CUDA_SUCCESS)
			e->acc_tmap_ok = false;
	}
}

int join_accumulate(fosphor_cu *e);
int publish(fosphor_cu *e);

/* rows a scratch ring may have: p.scratch_rows > 0 as given, 0 automatic (SCRATCH_BUDGET bytes, an
 * eighth of the free device memory at most), < 0 never */
long long scratch_rows_limit(const fosphor_cu *e)
{
	const long long w = e->p.wf_rows;
	if (e->p.scratch_rows < 0)
		return w;
	if (e->p.scratch_rows > 0) {
		long long r = w;
		while (r < e->p.scratch_rows)
			r <<= 1;
		return r;
	}
	size_t budget = SCRATCH_BUDGET, free_b = 0, total_b = 0;
	if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && free_b / 8 < budget)
		budget = free_b / 8;
	long long r = w;
	while ((size_t)(2 * r) * e->p.fft_len * sizeof(float) <= budget)
		r <<= 1;
	return r;
}

/* A launch pair folds as many calls as the ring holds.  When a process call brings more rows than
 * the ring has, replace the ring by a deeper scratch ring (once per size: the ring only grows). */
int ensure_ring(fosphor_cu *e, long long rows_wanted)
{
	if (rows_wanted <= e->ring_rows)
		return 0;
	if (e->scratch_limit == 0)
		e->scratch_limit = scratch_rows_limit(e);    /* once: cudaMemGetInfo is far too slow for the call path */
	const long long limit = e->scratch_limit;
	long long r = e->ring_rows;
	while (r < rows_wanted && r < limit)
		r <<= 1;
	if (r <= e->ring_rows)
		return 0;
	/* everything that reads or writes the old ring has to be done; rows not yet published move first */
	if (int rc = join_accumulate(e))
		return rc;
	if (int rc = publish(e))
		return rc;
	CU_CHECK(e, cudaStreamSynchronize(e->stream));
	float *nr = nullptr;
	if (cudaMalloc(&nr, sizeof(float) * (size_t)r * e->p.fft_len) != cudaSuccess) {
		cudaGetLastError();
		return 0;                        /* no memory for a deeper ring: keep folding by the old one */
	}
	if (e->d_ring != e->d_wf)
		cudaFree(e->d_ring);
	e->d_ring = nr;
	e->ring_rows = (int)r;
	e->ring_pos = 0;
	e->unpublished = 0;
	encode_tensor_maps(e);
	return 0;
}

/* Bring the user-visible waterfall up to date: the last min(unpublished, W) rows of the scratch ring
 * go to waterfall rows ending at wf_pos.  No-op while the ring is the waterfall. */
int publish(fosphor_cu *e)
{
	if (e->d_ring == e->d_wf || e->unpublished == 0)
		return 0;
	const long long w = e->p.wf_rows;
	const int cnt = (int)(e->unpublished < w ? e->unpublished : w);
	const size_t n4 = (size_t)e->p.fft_len / 4;
	const size_t total = (size_t)cnt * n4;
	int grid = (int)((total + 255) / 256);
	if (grid > 8 * e->sm_count)
		grid = 8 * e->sm_count;
	publish_rows_kernel<<<grid, 256, 0, e->stream>>>(
		reinterpret_cast<const float4 *>(e->d_ring), reinterpret_cast<float4 *>(e->d_wf), (int)n4, cnt,
		(e->ring_pos - cnt) & (e->ring_rows - 1), e->ring_rows - 1,
		(e->wf_pos - cnt) & (e->p.wf_rows - 1), e->p.wf_rows - 1);
	e->launches++;
	CU_CHECK(e, cudaGetLastError());
	e->unpublished = 0;
	return 0;
}

/* hit-count slices and live / max partials of the split kernels: most engines never leave the fused
 * path, so these (up to CNT_BUDGET bytes) are allocated when the split path first runs */
int ensure_split_buffers(fosphor_cu *e)
{
	if (e->d_cnt)
		return 0;
	const size_t n = e->p.fft_len, k = e->p.n_bins;
	if (cudaMalloc(&e->d_cnt, sizeof(unsigned short) * k * n * e->max_slices) != cudaSuccess) {
		cudaGetLastError();
		return fail(e, -ENOMEM, "no device memory for the split-path hit counts");
	}
	const size_t blocks = (size_t)(e->p.batch_max + ROWBLOCK - 1) / ROWBLOCK * e->max_slices;
	if (cudaMalloc(&e->d_part_live, sizeof(float) * blocks * n) != cudaSuccess ||
	    cudaMalloc(&e->d_part_max, sizeof(float) * blocks * n) != cudaSuccess) {
		cudaGetLastError();
		return fail(e, -ENOMEM, "no device memory for the split-path partials");
	}
	CU_CHECK(e, ensure_smem(e, count_tma_kernel, sizeof(unsigned) * 32 * k + sizeof(CountStage)));
	CU_CHECK(e, ensure_smem(e, count_kernel, sizeof(unsigned) * 32 * k));
	size_t upd = sizeof(float2) * (size_t)(e->p.batch_max + 1);
	if (upd > UPD_LUT_SMEM_MAX)
		upd = UPD_LUT_SMEM_MAX;       /* larger tables are read from global memory */
	const size_t parts = sizeof(float) * 2 * UPD_COLS * (size_t)((e->p.batch_max + ROWBLOCK - 1) / ROWBLOCK);
	if (upd < parts) upd = parts;
	if (upd < UPD_SMEM_MAX) upd = UPD_SMEM_MAX;
	CU_CHECK(e, ensure_smem(e, update_kernel, upd));
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
	if (e->tn.acc_stage_kb > 0 && C::smem_fixed(a.n_bins, a.batch) + (size_t)e->tn.acc_stage_kb * 1024 < limit)
		limit = C::smem_fixed(a.n_bins, a.batch) + (size_t)e->tn.acc_stage_kb * 1024;
	int dlog = C::depth_log2(a.n_bins, a.batch, limit);
	const long long boxes = (long long)a.n_calls * (a.batch / BOXR);
	while (dlog > 0 && (1ll << (dlog - 1)) >= boxes)
		dlog--;
	a.depth_log2 = dlog;
	a.l2_hints = e->tn.l2_hints;
	if constexpr (LOAD != 0) {
		/* parity waits on the stage ring are only sound for these shapes (accumulate.cuh: acc_ring_safe) */
		if (!acc_ring_safe(1ll << dlog, (long long)GC * (a.batch / BOXR), boxes))
			return fused_launch<COLS, FW, UW, 16, 16, 0, GC>(e, a, st);
	}
	const size_t smem = C::smem(a.n_bins, a.batch, LOAD != 0, dlog);
	if (cudaError_t err = ensure_smem(e, accumulate_fused_kernel<COLS, FW, UW, BOXR, SUBR, LOAD, GC>, smem))
		return err;
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

size_t fused_smem_need_cols(int cols, int n_bins, int batch, int gc)
{
	switch (cols) {
	case 4:  return fused_smem_need<4>(n_bins, batch, gc);
	case 16: return fused_smem_need<16>(n_bins, batch, gc);
	default: return fused_smem_need<8>(n_bins, batch, gc);
	}
}

/* one launch: count + rise/decay + live + max-hold of n_calls calls */
int launch_accumulate_fused(fosphor_cu *e, const BatchTables *t, cudaStream_t st, cudaEvent_t count_done,
                            int wf_pos, int n_calls, int batch)
{
	AccumArgs a;
	memset(&a, 0, sizeof(a));
	a.wf = e->d_ring;
	a.hist = e->d_hist;
	a.spectrum = e->d_spec;
	a.weights = t->d_weights;
	a.lut = t->d_lut;
	a.n = e->p.fft_len;
	a.n_bins = e->p.n_bins;
	a.wf_mask = e->ring_rows - 1;
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
		for (int b = e->tn.acc_box_max; b >= 16; b >>= 2)
			if (batch % b == 0 && wf_pos % b == 0 && e->ring_rows % b == 0 && (b % rv == 0 || rv % b == 0)) {
				boxr = b;
				const int ch = rv < b ? rv : b;
				subr = (ch % 64 == 0 && e->tn.acc_sub_max >= 64) ? 64 : 16;
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
	int gc = e->tn.acc_group;
	if (gc == 0)
		gc = batch <= 256 ? 4 : 1;        /* measured: cfg3 (B = 256) 115 -> 105 us per 32 calls; B = 1024 prefers the deeper stage ring */
	const size_t smem_max = e->smem_optin;
	while (gc > 1 && fused_smem_need_cols(e->acc_cols, a.n_bins, batch, gc) > smem_max)
		gc >>= 1;
	/* 16-column tiles (experiment knob ACC_COLS=16: N/16 CTAs leave more of the chip to the FFT kernel
	 * beside them; cfg2 +1.7 %, within the run-to-run spread): one variant.  32-column tiles were 9x
	 * SLOWER (32 CTAs: the per-SM L2 -> shared-memory path and a 2-box stage ring cannot feed them). */
	if (e->acc_cols > 8)
		gc = 1;
	/* warp roles (counters / cell updaters): 16 / 8, or 8 / 16 when a call has more cells to update
	 * than rows to count (ncu of cfg3, B = 256, K = 512: the counter warps spent half their time
	 * waiting for the updaters to hand the hit tiles back) */
	int roles = e->tn.acc_roles;
	if (roles == 0)
		roles = (gc == 4 && a.n_bins >= batch) ? 2 : 1;
	if (e->acc_cols == 16)
		err = fused_dispatch<16, 16, ACC_UW, 1>(e, a, st, boxr, subr);
	else if (e->acc_cols == 4)
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
	/* the fused CTA must fit: state tile + hit tiles + partials + table + a minimal stage ring */
	const size_t need = fused_smem_need_cols(e->acc_cols, e->p.n_bins, batch, 1);
	if (need > e->smem_optin)
		return false;
	return e->tn.acc_mode > 0 || (e->tn.acc_mode < 0 && 2 * batch >= e->p.n_bins && (batch % 16) == 0 && e->acc_tmap_ok);
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
	if (int rc = ensure_split_buffers(e))
		return rc;
	a.wf = e->d_ring;
	a.hist = e->d_hist;
	a.spectrum = e->d_spec;
	a.cnt = e->d_cnt;
	a.part_live = e->d_part_live;
	a.part_max = e->d_part_max;
	a.weights = t->d_weights;
	a.lut = t->d_lut;
	a.n = e->p.fft_len;
	a.n_bins = e->p.n_bins;
	a.wf_mask = e->ring_rows - 1;
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
	const bool use_tma = e->tmap_ok && e->tn.count_variant != 0 &&
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
	e->unpublished = 0;                   /* nothing of a scratch ring belongs to the cleared waterfall */
	e->rows_since_finish = -1;            /* every row of the waterfall is new to the host */
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
		/* more rows than the ring holds: fold deeper (scratch ring, grows once per size) */
		if (n_calls > 1 && (long long)n_calls * batch > e->ring_rows &&
		    (e->scratch_limit == 0 || e->ring_rows < e->scratch_limit)) {
			rc = ensure_ring(e, (long long)n_calls * batch);
			if (rc)
				return rc;
		}
		/* A chunk = calls folded by one FFT + accumulate launch pair.  Its rows must fit the
		 * ring (and its slices the count buffer of the split path); with room for several
		 * chunks in the ring the accumulate kernel of chunk c (acc_stream) overlaps the FFT
		 * of chunk c+1 (main stream). */
		const int ring_calls = e->ring_rows / batch;         /* >= 1: ring_rows >= wf_rows >= batch_max */
		int calls_per_chunk = ring_calls;
		int ov_chunk = 1;
		bool two_streams = false;
		e->slim_now = false;
		if (e->tn.overlap > 0) {                        /* forced, knobs from the environment */
			ov_chunk = e->tn.overlap_chunk < ring_calls / 2 ? e->tn.overlap_chunk : ring_calls / 2;
			if (ov_chunk < 1) ov_chunk = 1;
			two_streams = ring_calls >= 2 && n_calls >= 2 * ov_chunk;
			e->slim_now = two_streams && e->tn.acc_slim;
		} else if (e->tn.overlap < 0) {                 /* automatic, see the comment in struct fosphor_cu */
			const bool stream_fft = (e->p.fft_len == 512 || e->p.fft_len == 1024) && e->tn.fft_variant >= 2;
			ov_chunk = ring_calls / 4;
			two_streams = stream_fft && use_fused(e, batch) && ov_chunk >= 1 && ov_chunk <= e->max_slices &&
			              (long long)ov_chunk * batch * e->p.fft_len >= (32ll << 20) && n_calls >= 2 * ov_chunk;
		}
		if (two_streams)
			calls_per_chunk = ov_chunk;
		e->two_streams_now = two_streams;
		if (calls_per_chunk > e->max_slices)
			calls_per_chunk = e->max_slices;
		if (e->tn.chunk_calls > 0 && calls_per_chunk > e->tn.chunk_calls)
			calls_per_chunk = e->tn.chunk_calls;
		cudaStream_t acc = two_streams ? e->acc_stream : e->stream;
		/* a one-stream call, or another chunk geometry, first waits for what is still on acc_stream */
		if (e->acc_pending && (!two_streams || e->seq_batch != batch || e->seq_chunk_calls != calls_per_chunk)) {
			rc = join_accumulate(e);
			if (rc)
				return rc;
		}
		e->seq_batch = batch;
		e->seq_chunk_calls = calls_per_chunk;
		/* a caller's side stream still reads the spectrum state (export_maxhold_on): the accumulate
		 * launches below overwrite it */
		if (e->side_pending) {
			CU_CHECK(e, cudaStreamWaitEvent(acc, e->side_ev, 0));
			e->side_pending = false;
		}
		/* The FFT of chunk q overwrites ring rows last read by the accumulate launch of a
		 * chunk no younger than q - lag (every chunk has at most calls_per_chunk calls):
		 * waiting for that one - the acc stream is in order - frees them.  The chunk
		 * sequence runs on across process calls until something joins the streams. */
		int lag = ring_calls / calls_per_chunk;
		if (lag > N_CHUNK_EV) lag = N_CHUNK_EV;
		const bool scratch = e->d_ring != e->d_wf;
		for (int c0 = 0; c0 < n_calls; c0 += calls_per_chunk) {
			const int nc = n_calls - c0 < calls_per_chunk ? n_calls - c0 : calls_per_chunk;
			const long long q = two_streams ? e->chunk_seq++ : 0;
			e->two_stream_chunks += two_streams ? 1 : 0;
			const int pp = (int)(q % N_CHUNK_EV);
			if (two_streams && q >= lag)
				CU_CHECK(e, cudaStreamWaitEvent(e->stream, e->cnt_done[(q - lag) % N_CHUNK_EV], 0));
			CU_CHECK(e, launch_fft(e, in + (long long)c0 * batch * hop, hop, e->ring_pos, nc * batch));
			if (two_streams) {
				CU_CHECK(e, cudaEventRecord(e->fft_done[pp], e->stream));
				CU_CHECK(e, cudaStreamWaitEvent(acc, e->fft_done[pp], 0));
			}
			rc = launch_accumulate(e, t, acc, two_streams ? e->cnt_done[pp] : nullptr, e->ring_pos, nc, batch);
			if (rc)
				return rc;
			e->ring_pos = (e->ring_pos + nc * batch) & (e->ring_rows - 1);
			e->wf_pos = (e->wf_pos + nc * batch) & (e->p.wf_rows - 1);   /* cl.c:954, nc times */
			if (scratch)
				e->unpublished += (long long)nc * batch;
		}
		if (e->rows_since_finish >= 0)
			e->rows_since_finish += (long long)n_calls * batch;
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

/* first AND last byte page-locked: a buffer that merely starts inside somebody's registration is not */
bool is_pinned_range(const void *p, size_t bytes)
{
	return is_pinned_host(p) && (bytes < 2 || is_pinned_host(static_cast<const char *>(p) + bytes - 1));
}

/* staging slots (page-locked host + device), the copy stream and the copy threads: first host-fed call */
int ensure_staging(fosphor_cu *e)
{
	if (e->n_slots)
		return 0;
	e->stage_elems = (size_t)e->p.batch_max * e->p.fft_len;
	const size_t bytes = sizeof(float2) * e->stage_elems;
	const int want = e->tn.stage_slots ? e->tn.stage_slots : (bytes <= STAGE_SLOT_SMALL ? 4 : 2);
	CU_CHECK(e, cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
	for (int i = 0; i < want; i++) {
		CU_CHECK(e, cudaMallocHost(&e->h_in[i], bytes));
		CU_CHECK(e, cudaMalloc(&e->d_in[i], bytes));
		CU_CHECK(e, cudaEventCreateWithFlags(&e->copied[i], cudaEventDisableTiming));
		CU_CHECK(e, cudaEventCreateWithFlags(&e->slot_free[i], cudaEventDisableTiming));
		e->n_slots = i + 1;
	}
	return 0;
}

copy_pool *get_pool(fosphor_cu *e)
{
	if (!e->pool)
		e->pool.reset(new (std::nothrow) copy_pool(e->tn.copy_threads));
	return e->pool.get();
}

/* ---- page-locking the caller's buffer where it lies --------------------------------------------
 * cudaHostRegister once, DMA straight out of the caller's memory from then on: no CPU copy, one trip
 * through host DRAM instead of three (8 GPUs on one host: 21 vs 5.6 Gsamples/s, the staged path is
 * host-memory bound there).  Right for callers whose sample memory lives as long as the engine - the
 * reference sink's FIFO does (one 16 MiB ring allocated in the constructor, lib/base_sink_c_impl.cc:58,
 * freed after the worker thread that owns the engine has been joined) - and DANGEROUS for buffers
 * that are freed while registered: the driver keeps the old pages pinned, and a later allocation at
 * the same address would be read stale.  Therefore:
 *   HOSTREG=1   the caller vouches for its buffers: register on first sight;
 *   automatic   register a call range on SECOND sight, record the physical page numbers of the range
 *               (/proc/self/pagemap - readable with PFNs only for a privileged process; otherwise the
 *               automatic mode never registers) and compare a few of them before EVERY use: a freed and
 *               reallocated buffer has other pages, is unregistered and staged (three strikes and
 *               the engine stops registering);
 *   HOSTREG=0   never.
 * Ranges are registered page-wise, never overlapping, merged when they touch. */
uint64_t page_frame(fosphor_cu *e, uintptr_t addr)
{
	if (e->pagemap_fd == -2) {
		e->pagemap_fd = open("/proc/self/pagemap", O_RDONLY | O_CLOEXEC);
		if (e->pagemap_fd >= 0) {
			int probe = 1;                    /* a page of our own stack: present by construction */
			uint64_t ent = 0;
			const uintptr_t a = reinterpret_cast<uintptr_t>(&probe);
			if (pread(e->pagemap_fd, &ent, 8, (off_t)(a / 4096 * 8)) != 8 || !(ent >> 63) ||
			    (ent & ((1ull << 55) - 1)) == 0) {
				close(e->pagemap_fd);             /* PFNs are hidden from this process */
				e->pagemap_fd = -1;
			}
		} else {
			e->pagemap_fd = -1;
		}
		if (e->pagemap_fd == -1 && e->tn.hostreg < 0)
			fprintf(stderr, "[+] fosphor_b200: pageable sample buffers are staged through copy threads (this process "
			        "cannot read its page frame numbers, so page-locking them in place cannot be checked); set "
			        "FOSPHOR_B200_HOSTREG=1 if the buffers outlive the engine (zero-copy DMA)\n");
	}
	if (e->pagemap_fd < 0)
		return 0;
	uint64_t ent = 0;
	if (pread(e->pagemap_fd, &ent, 8, (off_t)(addr / 4096 * 8)) != 8 || !(ent >> 63))
		return 0;
	return ent & ((1ull << 55) - 1);
}

void hostreg_drop(fosphor_cu *e, size_t idx)
{
	HostRange &r = e->hostreg[idx];
	global_reg_remove(e, r.lo);
	if (cudaHostUnregister(reinterpret_cast<void *>(r.lo)) != cudaSuccess)
		cudaGetLastError();
	e->hostreg_bytes -= r.hi - r.lo;
	e->hostreg.erase(e->hostreg.begin() + (long)idx);
}

/* true iff [p, p + bytes) is inside ONE range registered by this engine - and, in automatic mode,
 * still backed by the pages that were registered */
bool hostreg_cover(fosphor_cu *e, const void *p, size_t bytes)
{
	const uintptr_t page = 4096;
	uintptr_t lo = reinterpret_cast<uintptr_t>(p) & ~(page - 1);
	uintptr_t hi = (reinterpret_cast<uintptr_t>(p) + bytes + page - 1) & ~(page - 1);
	const bool automatic = e->tn.hostreg < 0;
	if (automatic && (e->hostreg_stale >= 3 || (e->pagemap_fd == -1)))
		return false;
	for (size_t i = 0; i < e->hostreg.size(); i++) {
		HostRange &r = e->hostreg[i];
		if (lo < r.lo || hi > r.hi)
			continue;
		if (!r.pfn.empty()) {
			/* first, last and two inner pages of THIS call's range: a buffer that was freed and
			 * reallocated has none of its old pages (they are still pinned by the registration) */
			const uintptr_t n = (hi - lo) / page;
			const uintptr_t probe[4] = {lo, lo + (n / 3) * page, lo + (2 * n / 3) * page, hi - page};
			for (uintptr_t a : probe)
				if (page_frame(e, a) != r.pfn[(a - r.lo) / page]) {
					hostreg_drop(e, i);
					e->hostreg_stale++;
					return false;                 /* staged this time; it may be registered afresh later */
				}
		}
		return true;
	}
	if (automatic) {
		/* second sight promotes: one-shot buffers are never registered (1.3 ms per 8 MiB) */
		if (page_frame(e, lo) == 0)
			return false;
		bool known = false;
		for (const HostRange &r : e->seen)
			if (r.lo == lo && r.hi == hi)
				known = true;
		if (!known) {
			if (e->seen.size() >= 16)
				e->seen.erase(e->seen.begin());
			e->seen.push_back({lo, hi, {}});
			return false;
		}
	}
	/* grow to the hull of everything it touches: the sink's ring is one allocation, its chunks abut */
	for (size_t i = 0; i < e->hostreg.size();) {
		const HostRange &r = e->hostreg[i];
		if (r.hi < lo || r.lo > hi) {
			i++;
			continue;
		}
		lo = r.lo < lo ? r.lo : lo;
		hi = r.hi > hi ? r.hi : hi;
		hostreg_drop(e, i);
	}
	if (e->hostreg_bytes + (hi - lo) > HOSTREG_BUDGET)
		return false;
	if (cudaHostRegister(reinterpret_cast<void *>(lo), hi - lo, cudaHostRegisterDefault) != cudaSuccess) {
		cudaGetLastError();
		return false;
	}
	HostRange nr{lo, hi, {}};
	if (automatic) {
		nr.pfn.resize((hi - lo) / page);
		std::vector<uint64_t> ent(nr.pfn.size());
		const ssize_t want = (ssize_t)(8 * ent.size());
		if (pread(e->pagemap_fd, ent.data(), (size_t)want, (off_t)(lo / page * 8)) != want) {
			cudaHostUnregister(reinterpret_cast<void *>(lo));
			return false;
		}
		for (size_t i = 0; i < ent.size(); i++)
			nr.pfn[i] = (ent[i] >> 63) ? (ent[i] & ((1ull << 55) - 1)) : 0;
	}
	e->hostreg.push_back(std::move(nr));
	e->hostreg_bytes += hi - lo;
	global_reg_add(e, lo, hi);
	return true;
}

/* May the DMA engine work on [p, p + bytes) of the caller's memory in place?  Page-locked by the
 * caller: yes.  Inside a range this engine registered: after re-validation.  Otherwise, if the
 * policy allows and the range is worth it, register it now. */
bool host_range_direct(fosphor_cu *e, const void *p, size_t bytes, size_t min_bytes)
{
	bool own = false;
	const uintptr_t a = reinterpret_cast<uintptr_t>(p);
	for (const HostRange &r : e->hostreg)
		if (a < r.hi && a + bytes > r.lo)
			own = true;
	if (own)
		return hostreg_cover(e, p, bytes);   /* inside one of this engine's own ranges (and re-validated): page-locked */
	if (global_reg_foreign(e, a, bytes))
		return false;                        /* another engine's registration: only it can vouch for it */
	if (is_pinned_range(p, bytes))
		return true;                         /* page-locked by the caller (pinned FIFO, cudaHostRegister) */
	return e->tn.hostreg != 0 && bytes >= min_bytes && hostreg_cover(e, p, bytes) && is_pinned_range(p, bytes);
}

/* Host samples -> device slot.  The source buffer is free when the process call returns (reference
 * contract, base_sink_c_impl.cc:170-174).
 *   page-locked source (cudaHostAlloc / cudaHostRegister, e.g. the pinned FIFO, or a range the
 *     engine registered itself): DMA straight from the caller's memory; the call waits for that copy
 *     only at its very end (await_uploads), after the kernels that consume it were enqueued behind
 *     the `copied` event, so the launch overhead hides under the copy.
 *   pageable source: the copy threads move it into a page-locked slot piece by piece and the DMA of
 *     piece p starts while they are on piece p+1; the call returns when the last piece has been
 *     READ (its DMA, and the DMA of earlier calls, run on behind it).
 * Either way the H2D runs on its own stream and overlaps the kernels of earlier calls. */
int upload_staged(fosphor_cu *e, const float2 *src, size_t n_samples, float2 **dev_out)
{
	if (int rc = ensure_staging(e))
		return rc;
	const int s = e->slot;
	e->slot = (e->slot + 1) % e->n_slots;
	const size_t bytes = sizeof(float2) * n_samples;
	/* previous reader of d_in[s] done (which implies the H2D that filled it, hence h_in[s] too) */
	CU_CHECK(e, cudaEventSynchronize(e->slot_free[s]));
	bool direct = host_range_direct(e, src, bytes, (size_t)1 << 20);
	if (direct && cudaMemcpyAsync(e->d_in[s], src, bytes, cudaMemcpyHostToDevice, e->copy_stream) != cudaSuccess) {
		cudaGetLastError();               /* e.g. a range stitched from two registrations: stage it instead */
		direct = false;
	}
	if (direct) {
		CU_CHECK(e, cudaEventRecord(e->copied[s], e->copy_stream));
		e->await_slot = s;
		e->direct_calls++;
	} else {
		/* with nothing in flight on the copy stream (the first call after a finish) the DMA engine waits
		 * for the first piece: make that one small; behind a running DMA larger pieces cost fewer requests */
		size_t piece = (size_t)e->tn.stage_piece_kb << 10;
		if (piece > ((size_t)1 << 20)) {
			if (cudaStreamQuery(e->copy_stream) == cudaSuccess)
				piece = (size_t)1 << 20;
			else
				cudaGetLastError();       /* cudaErrorNotReady is an answer, not an error to find later */
		}
		copy_pool *pool = bytes >= ((size_t)1 << 20) ? get_pool(e) : nullptr;
		if (pool) {
			int pieces = (int)((bytes + piece - 1) / piece);
			if (pieces > 16) pieces = 16;     /* large-N calls: 16 DMA requests are plenty */
			pool->start(e->h_in[s], src, bytes, pieces, e->tn.copy_nt != 0);
			char *hp = reinterpret_cast<char *>(e->h_in[s]), *dp = reinterpret_cast<char *>(e->d_in[s]);
			cudaError_t err = cudaSuccess;
			for (int p = 0; p < pieces; p++) {
				pool->wait_piece(p);
				if (err == cudaSuccess && pool->piece_offset(p) < bytes)
					err = cudaMemcpyAsync(dp + pool->piece_offset(p), hp + pool->piece_offset(p),
					                      pool->piece_size(p), cudaMemcpyHostToDevice, e->copy_stream);
			}
			CU_CHECK(e, err);
		} else {
			memcpy(e->h_in[s], src, bytes);
			CU_CHECK(e, cudaMemcpyAsync(e->d_in[s], e->h_in[s], bytes, cudaMemcpyHostToDevice, e->copy_stream));
		}
		CU_CHECK(e, cudaEventRecord(e->copied[s], e->copy_stream));
		e->staged_calls++;
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

/* Results for pageable caller memory: D2H into a page-locked bounce buffer (13 -> 55 GB/s on the
 * box), then the copy threads.  Segments are (device source, host destination, bytes). */
struct OutSeg {
	const void *dev;
	void *host;
	size_t bytes;
};

int download(fosphor_cu *e, const OutSeg *segs, int n)
{
	/* the caller's result images are long-lived (self->img_*, fosphor.c:52-54): page-locked where they
	 * lie under the same policy as the sample buffers, the D2H then lands in them directly; what is
	 * left (small or unregistered segments) goes through the bounce buffer */
	bool direct[8] = {};
	size_t bounce = 0;
	for (int i = 0; i < n && i < 8; i++) {
		direct[i] = segs[i].bytes == 0 || host_range_direct(e, segs[i].host, segs[i].bytes, (size_t)8 << 10);
		if (!direct[i])
			bounce += segs[i].bytes;
	}
	for (int i = 0; i < n; i++)
		if (direct[i] && segs[i].bytes)
			CU_CHECK(e, cudaMemcpyAsync(segs[i].host, segs[i].dev, segs[i].bytes, cudaMemcpyDeviceToHost, e->stream));
	if (bounce == 0) {
		CU_CHECK(e, cudaStreamSynchronize(e->stream));   /* cl.c:1052 */
		return 0;
	}
	copy_pool *pool = bounce < ((size_t)256 << 10) ? nullptr : get_pool(e);
	if (e->h_out_bytes < bounce) {
		if (e->h_out)
			cudaFreeHost(e->h_out);
		e->h_out = nullptr;
		e->h_out_bytes = 0;
		CU_CHECK(e, cudaMallocHost(&e->h_out, bounce));
		e->h_out_bytes = bounce;
	}
	/* cut into at most OUT_CHUNKS chunks of >= 1 MiB; the D2H of chunk c+1 runs while the copy
	 * threads move chunk c out of the bounce buffer */
	size_t chunk = (bounce + OUT_CHUNKS - 1) / OUT_CHUNKS;
	if (chunk < ((size_t)1 << 20))
		chunk = (size_t)1 << 20;
	struct Piece { void *host; size_t off, bytes; };
	Piece pieces[OUT_CHUNKS + 8];
	int np = 0;
	char *hp = reinterpret_cast<char *>(e->h_out);
	size_t off = 0;
	for (int i = 0; i < n; i++) {
		if (direct[i])
			continue;
		for (size_t o = 0; o < segs[i].bytes; o += chunk) {
			const size_t nb = segs[i].bytes - o < chunk ? segs[i].bytes - o : chunk;
			while ((int)e->out_ev.size() <= np) {
				cudaEvent_t ev;
				CU_CHECK(e, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
				e->out_ev.push_back(ev);
			}
			CU_CHECK(e, cudaMemcpyAsync(hp + off, static_cast<const char *>(segs[i].dev) + o, nb,
			                            cudaMemcpyDeviceToHost, e->stream));
			CU_CHECK(e, cudaEventRecord(e->out_ev[np], e->stream));
			pieces[np++] = {static_cast<char *>(segs[i].host) + o, off, nb};
			off += nb;
		}
	}
	for (int c = 0; c < np; c++) {
		CU_CHECK(e, cudaEventSynchronize(e->out_ev[c]));
		if (pool)
			pool->copy(pieces[c].host, hp + pieces[c].off, pieces[c].bytes);
		else
			memcpy(pieces[c].host, hp + pieces[c].off, pieces[c].bytes);
	}
	CU_CHECK(e, cudaStreamSynchronize(e->stream));       /* cl.c:1052 */
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
	p->scratch_rows = 0;        /* automatic */
}

void fosphor_cu_destroy(struct fosphor_cu *e)
{
	if (!e)
		return;
	DevGuard guard(e);
	if (e->acc_stream)
		cudaStreamSynchronize(e->acc_stream);
	if (e->copy_stream)
		cudaStreamSynchronize(e->copy_stream);
	if (e->stream)
		cudaStreamSynchronize(e->stream);
	e->pool.reset();
	for (const HostRange &r : e->hostreg) {
		global_reg_remove(e, r.lo);
		if (cudaHostUnregister(reinterpret_cast<void *>(r.lo)) != cudaSuccess)
			cudaGetLastError();
	}
	if (e->pagemap_fd >= 0)
		close(e->pagemap_fd);
	if (e->d_ring != e->d_wf)
		cudaFree(e->d_ring);
	cudaFree(e->d_twg);
	cudaFree(e->d_win); cudaFree(e->d_tw); cudaFree(e->d_wf); cudaFree(e->d_hist);
	cudaFree(e->d_spec); cudaFree(e->d_cnt); cudaFree(e->d_part_live);
	cudaFree(e->d_part_max);
	for (int i = 0; i < MAX_STAGE_SLOTS; i++) {
		cudaFreeHost(e->h_in[i]);
		cudaFree(e->d_in[i]);
		if (e->copied[i]) cudaEventDestroy(e->copied[i]);
		if (e->slot_free[i]) cudaEventDestroy(e->slot_free[i]);
	}
	if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
	cudaFreeHost(e->h_out);
	for (cudaEvent_t ev : e->out_ev)
		cudaEventDestroy(ev);
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
	if (e->acc_stream) cudaStreamDestroy(e->acc_stream);
	for (int i = 0; i < N_CHUNK_EV; i++) {
		if (e->fft_done[i]) cudaEventDestroy(e->fft_done[i]);
		if (e->cnt_done[i]) cudaEventDestroy(e->cnt_done[i]);
	}
	if (e->acc_done) cudaEventDestroy(e->acc_done);
	if (e->side_ev) cudaEventDestroy(e->side_ev);
	if (e->own_stream) cudaStreamDestroy(e->own_stream);
	delete e;
}

int fosphor_cu_create(struct fosphor_cu **out, const struct fosphor_cu_params *pp)
{
	if (!out || !pp)
		return -EINVAL;
	*out = nullptr;
	const fosphor_cu_params &p = *pp;
	/* n_bins: both accumulate paths keep a [K][cols] hit tile in shared memory; 1024 bins (BASELINE
	 * configs[3]) is what fits every path */
	if (!plan_supported(p.fft_len) || p.n_bins < 2 || p.n_bins > FOSPHOR_CU_MAX_BINS ||
	    p.wf_rows < 1 || (p.wf_rows & (p.wf_rows - 1)) ||
	    p.batch_mult < 1 || p.batch_max < p.batch_mult || p.batch_max % p.batch_mult || p.batch_max > 32768 ||
	    p.wf_rows < p.batch_max || !(p.histo_t0r > 0.0f) || !(p.histo_t0d > 0.0f))
		return fail(nullptr, -EINVAL, "unsupported engine parameters (N=%d K=%d W=%d batch %d/%d)",
		            p.fft_len, p.n_bins, p.wf_rows, p.batch_mult, p.batch_max);

	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
		return fail(nullptr, -ENODEV, "no CUDA device (this library has no CPU fallback)");
	if (p.device >= ndev)
		return fail(nullptr, -ENODEV, "CUDA device %d does not exist (%d visible)", p.device, ndev);

	fosphor_cu *e = new (std::nothrow) fosphor_cu;
	if (!e)
		return -ENOMEM;
	e->p = p;
	e->tn = tuning_from_env();
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

	if (p.device >= 0) {
		e->device = p.device;
	} else if (cudaGetDevice(&e->device) != cudaSuccess) {
		delete e;
		return fail(nullptr, -ENODEV, "no current CUDA device");
	}
	DevGuard guard(e->device);           /* the caller's current device is restored on return */
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
	CREATE_CHECK(cudaEventCreateWithFlags(&e->side_ev, cudaEventDisableTiming));

	const size_t n = p.fft_len, k = p.n_bins, w = p.wf_rows;
	CREATE_CHECK(cudaMalloc(&e->d_win, sizeof(float) * n));
	CREATE_CHECK(cudaMalloc(&e->d_wf, sizeof(float) * w * n));
	e->d_ring = e->d_wf;                 /* until a multi-call launch wants a deeper ring (ensure_ring) */
	e->ring_rows = p.wf_rows;
	CREATE_CHECK(cudaMalloc(&e->d_hist, sizeof(float) * k * n));
	CREATE_CHECK(cudaMalloc(&e->d_spec, sizeof(float2) * 2 * n));
	{
		/* split path: slices per launch - at least what one call needs to fill the chip, at most
		 * MAX_SLICES, within CNT_BUDGET bytes of u16 counts (allocated on first use) */
		const int tiles = p.fft_len / ACC_COLS;
		int need = (2 * e->sm_count + tiles - 1) / tiles;
		if (need < 1) need = 1;
		size_t fit = CNT_BUDGET / (sizeof(unsigned short) * k * n);
		int ms = fit > (size_t)MAX_SLICES ? MAX_SLICES : (int)fit;
		if (ms < need) ms = need;
		e->max_slices = ms;
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
		e->plan_key = p.fft_len + ((e->tn.fft_r64 && (p.fft_len == 2048 || p.fft_len == 4096)) ? 1 : 0);
		PLAN_SWITCH(e->plan_key, build_twiddles<P>(tw));
		CREATE_CHECK(cudaMalloc(&e->d_tw, sizeof(float2) * tw.size()));
		CREATE_CHECK(cudaMemcpy(e->d_tw, tw.data(), sizeof(float2) * tw.size(), cudaMemcpyHostToDevice));
		cudaError_t perr = cudaErrorInvalidValue;
		PLAN_SWITCH(e->plan_key, (perr = plan_setup<P>(e)));
		CREATE_CHECK(perr);
		if (p.fft_len == 1024)
			CREATE_CHECK(stream_setup<Plan1024>(e));
		if (p.fft_len == 512)
			CREATE_CHECK(stream_setup<Plan512>(e));
		if (p.fft_len == 2048)
			CREATE_CHECK(cta_stream_setup<Plan2048>(e));
		if (p.fft_len == 4096)
			CREATE_CHECK(cta_stream_setup<Plan4096>(e));
		if (p.fft_len == 8192)
			CREATE_CHECK(cta_stream_setup<Plan8192>(e));
		if ((p.fft_len == 8192 || p.fft_len == 16384) && e->tn.fft_variant == 4) {
			std::vector<float2> twg;
			if (p.fft_len == 8192)
				build_grouped_twiddles<Plan8192>(twg);
			else
				build_grouped_twiddles<Plan16384>(twg);
			CREATE_CHECK(cudaMalloc(&e->d_twg, sizeof(float2) * twg.size()));
			CREATE_CHECK(cudaMemcpy(e->d_twg, twg.data(), sizeof(float2) * twg.size(), cudaMemcpyHostToDevice));
		}
	}
	{
		/* fused kernel: narrow tiles for short spectra so that N / cols CTAs still fill the chip */
		e->acc_cols = p.fft_len <= 512 ? 4 : 8;
		if (e->tn.acc_cols)
			e->acc_cols = e->tn.acc_cols;
		/* the tensor-map encoder lives in the driver and is fetched through the runtime */
		void *fn = nullptr;
		cudaDriverEntryPointQueryResult qres;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
		    qres == cudaDriverEntryPointSuccess && fn)
			e->tmap_encode = fn;
		else
			cudaGetLastError();
		encode_tensor_maps(e);
		if (!e->tmap_ok)
			fprintf(stderr, "[w] fosphor_b200: tensor-map encode unavailable, using the plain-load kernels\n");
	}

	for (auto &t : e->tables) {
		CREATE_CHECK(cudaMalloc(&t.d_weights, sizeof(float) * p.batch_max));
		CREATE_CHECK(cudaMalloc(&t.d_lut, sizeof(float2) * (p.batch_max + 1)));
		CREATE_CHECK(cudaMallocHost(&t.h_stage, sizeof(float) * p.batch_max + sizeof(float2) * (p.batch_max + 1)));
		CREATE_CHECK(cudaEventCreateWithFlags(&t.uploaded, cudaEventDisableTiming));
	}
#undef CREATE_CHECK

	*out = e;
	return 0;
}

int fosphor_cu_set_stream(struct fosphor_cu *e, void *cuda_stream)
{
	if (!e)
		return -EINVAL;
	DevGuard guard(e);
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
	DevGuard guard(e);
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

void fosphor_cu_default_window(int fft_len, float *win)
{
	/* fosphor.c:113-118 with FOSPHOR_FFT_LEN -> fft_len: periodic Hamming x 1.855, pi truncated to
	 * 3.141592f, everything in f32 */
	for (int i = 0; i < fft_len; i++) {
		const float ft = (float)fft_len;
		const float fp = (float)i;
		win[i] = (0.54f - 0.46f * cosf((2.0f * 3.141592f * fp) / ft)) * 1.855f;
	}
}

void fosphor_cu_power_range(int fft_len, int db_ref, int db_per_div, float *scale, float *offset)
{
	/* fosphor.c:131-152; its constant k = log10f(FOSPHOR_FFT_LEN) becomes log10f(fft_len) */
	const int db0 = db_ref - 10 * db_per_div;
	const int db1 = db_ref;
	const float k = log10f((float)fft_len);
	if (offset) *offset = -(k + ((float)db0 / 20.0f));
	if (scale) *scale = 20.0f / (float)(db1 - db0);
}

int fosphor_cu_process_device(struct fosphor_cu *e, const void *samples_dev,
                              int n_spectra, long long hop)
{
	if (!e || (!samples_dev && n_spectra > 0))
		return -EINVAL;
	DevGuard guard(e);
	return process_device_calls(e, static_cast<const float2 *>(samples_dev), 1, n_spectra, hop);
}

int fosphor_cu_process_device_multi(struct fosphor_cu *e, const void *samples_dev,
                                    int n_calls, int batch, long long hop)
{
	if (!e || (!samples_dev && n_calls > 0 && batch > 0))
		return -EINVAL;
	DevGuard guard(e);
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
	DevGuard guard(e);
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
	DevGuard guard(e);
	if (n_calls == 0 || batch == 0)
		return process_device_calls(e, nullptr, n_calls, batch, hop);
	if (!raw_host)
		return -EINVAL;
	if (int rc = ensure_staging(e))
		return rc;
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

/* cl.c:970-1061.  rows_only_new: copy only the waterfall rows written since the previous finish
 * (the host image is the caller's persistent copy of the ring, like self->img_waterfall). */
static int finish_common(struct fosphor_cu *e, float *waterfall_host, float *histogram_host,
                         float *spectrum_host, bool rows_only_new, int *first_row, int *n_rows)
{
	if (first_row) *first_row = 0;
	if (n_rows) *n_rows = 0;
	if (!e)
		return -EINVAL;
	if (e->state == ST_READY)             /* cl.c:978-979 */
		return 0;
	DevGuard guard(e);
	if (e->state == ST_BOOTING) {         /* cl.c:982-994 */
		int rc = clear_buffers(e);
		if (rc)
			return rc;
	}
	if (int rc = join_accumulate(e))
		return rc;
	if (int rc = publish(e))
		return rc;
	const size_t n = e->p.fft_len;
	const int w = e->p.wf_rows;
	/* cl.c:1012-1048 */
	OutSeg segs[4];
	int ns = 0;
	int row0 = 0, cnt = w;
	if (rows_only_new && e->rows_since_finish >= 0 && e->rows_since_finish < w) {
		cnt = (int)e->rows_since_finish;
		row0 = (e->wf_pos - cnt) & (w - 1);
	}
	if (waterfall_host && cnt > 0) {
		const int c1 = row0 + cnt <= w ? cnt : w - row0;          /* up to the ring end, then the wrap */
		segs[ns++] = {e->d_wf + (size_t)row0 * n, waterfall_host + (size_t)row0 * n, sizeof(float) * (size_t)c1 * n};
		if (c1 < cnt)
			segs[ns++] = {e->d_wf, waterfall_host, sizeof(float) * (size_t)(cnt - c1) * n};
	}
	if (histogram_host)
		segs[ns++] = {e->d_hist, histogram_host, sizeof(float) * e->p.n_bins * n};
	if (spectrum_host)
		segs[ns++] = {e->d_spec, spectrum_host, sizeof(float2) * 2 * n};
	if (int rc = download(e, segs, ns))   /* synchronises the stream, cl.c:1052 */
		return rc;
	if (first_row) *first_row = row0;
	if (n_rows) *n_rows = cnt;
	if (waterfall_host)
		e->rows_since_finish = 0;
	e->state = ST_READY;
	return 1;
}

int fosphor_cu_finish(struct fosphor_cu *e, float *waterfall_host,
                      float *histogram_host, float *spectrum_host)
{
	return finish_common(e, waterfall_host, histogram_host, spectrum_host, false, nullptr, nullptr);
}

int fosphor_cu_finish_new_rows(struct fosphor_cu *e, float *waterfall_host,
                               float *histogram_host, float *spectrum_host,
                               int *first_row, int *n_rows)
{
	return finish_common(e, waterfall_host, histogram_host, spectrum_host, true, first_row, n_rows);
}

int fosphor_cu_sync(struct fosphor_cu *e)
{
	if (!e)
		return -EINVAL;
	DevGuard guard(e);
	if (int rc = join_accumulate(e))
		return rc;
	if (int rc = publish(e))
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
	DevGuard guard(e);
	if (int rc = join_accumulate(e))
		return rc;
	return publish(e);
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
	DevGuard guard(e);
	const int n = e->p.fft_len;
	if (int rc = join_accumulate(e))
		return rc;
	export_maxhold_kernel<<<(n + 255) / 256, 256, 0, e->stream>>>(e->d_spec, n, out_dev);
	e->launches++;
	CU_CHECK(e, cudaGetLastError());
	return 0;
}

int fosphor_cu_export_maxhold_on(struct fosphor_cu *e, float *out_dev, void *side_stream)
{
	if (!e || !out_dev || !side_stream)
		return -EINVAL;
	DevGuard guard(e);
	const int n = e->p.fft_len;
	cudaStream_t side = static_cast<cudaStream_t>(side_stream);
	/* order the side stream after the last accumulate launch, wherever it runs - WITHOUT ordering the
	 * engine's own stream after it: the next call's FFT keeps overlapping that accumulate launch */
	if (e->acc_pending) {
		CU_CHECK(e, cudaStreamWaitEvent(side, e->acc_done, 0));
	} else {
		CU_CHECK(e, cudaEventRecord(e->side_ev, e->stream));
		CU_CHECK(e, cudaStreamWaitEvent(side, e->side_ev, 0));
	}
	export_maxhold_kernel<<<(n + 255) / 256, 256, 0, side>>>(e->d_spec, n, out_dev);
	e->launches++;
	CU_CHECK(e, cudaGetLastError());
	/* the next accumulate launch overwrites the max-hold trace: it waits for this read */
	CU_CHECK(e, cudaEventRecord(e->side_ev, side));
	e->side_pending = true;
	return 0;
}

int fosphor_cu_debug_fft(struct fosphor_cu *e, const void *samples_dev,
                         int n_spectra, long long hop, void *out_dev)
{
	if (!e || !samples_dev || !out_dev || n_spectra < 0 || hop < 1)
		return -EINVAL;
	DevGuard guard(e);
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
	DevGuard guard(e);
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
	DevGuard guard(e);
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

int fosphor_cu_host_feed_stats(const struct fosphor_cu *e, unsigned long long *staged_calls,
                               unsigned long long *direct_calls, int *copy_threads, int *ring_rows)
{
	if (!e)
		return -EINVAL;
	if (staged_calls) *staged_calls = e->staged_calls;
	if (direct_calls) *direct_calls = e->direct_calls;
	if (copy_threads) *copy_threads = e->pool ? e->pool->threads() : 0;
	if (ring_rows) *ring_rows = e->ring_rows;
	return 0;
}

const char *fosphor_cu_last_error(const struct fosphor_cu *e)
{
	return e ? e->err : "null engine";
}

} /* extern "C" */
