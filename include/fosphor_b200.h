/*
 * fosphor_b200.h - C ABI of libfosphor_b200.so, the B200-native (sm_100a CUDA)
 * replacement for gr-fosphor's OpenCL compute path.
 *
 * Two layers, both plain C (pointers and sizes only, no CUDA/torch types):
 *
 *  1. The DROP-IN boundary: the seven fosphor_cl_* entry points of the
 *     reference's lib/fosphor/cl.h:22-32, same signatures, return codes and
 *     state machine as lib/fosphor/cl.c:795-1089.  Linking libfosphor against
 *     this library instead of cl.c + cl_compat.c (+ fft.cl / display.cl) leaves
 *     fosphor.c, the GL renderer and the GNU Radio sink untouched
 *     (INTEGRATION.md).  Fixed problem size N=1024 / 128 bins / 1024 rows
 *     (lib/fosphor/private.h:21-25).
 *
 *  2. The parameterised engine fosphor_cu_*: any power-of-two N in
 *     512..16384, any bin count, device-resident input, in-engine overlap
 *     (hop addressing), many calls per launch, results as plain device arrays.
 *
 * All functions return 0 on success or a negative errno like the reference
 * (-EINVAL bad size cl.c:881-886, -EIO runtime failure cl.c:842,967,1060,
 * -ENOMEM cl.c:809, -ENODEV no usable device cl.c:321) unless stated.
 * There is NO CPU fallback: without a CUDA device init fails with -ENODEV/-EIO.
 */
#ifndef FOSPHOR_B200_H
#define FOSPHOR_B200_H

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------ */
/* 1. Drop-in boundary (reference: lib/fosphor/cl.h)                         */
/* ------------------------------------------------------------------------ */

struct fosphor; /* lib/fosphor/private.h:30-55, restated in fosphor_private_abi.h */

/* cl.h:22 / cl.c:799-843.  Allocates the engine into self->cl, leaves
 * FLG_FOSPHOR_USE_CLGL_SHARING clear so fosphor_init() allocates the host
 * result images (fosphor.c:50-62).  0, -ENOMEM, -EIO. */
int fosphor_cl_init(struct fosphor *self);

/* cl.h:23 / cl.c:845-868.  NULL-state safe, sets self->cl = NULL. */
void fosphor_cl_release(struct fosphor *self);

/* cl.h:25-26 / cl.c:870-968.  samples: host cf32, len complex samples,
 * len % (16*1024) == 0 and len <= 1024*1024 else -EINVAL.  Asynchronous; the
 * sample buffer has been fully consumed (staged) when the call returns, so the
 * caller may recycle it at once (base_sink_c_impl.cc:170-174). */
int fosphor_cl_process(struct fosphor *self, void *samples, int len);

/* cl.h:27 / cl.c:970-1061.  1: results copied to self->img_waterfall /
 * img_histogram / buf_spectrum; 0: nothing new; <0 error. */
int fosphor_cl_finish(struct fosphor *self);

/* cl.h:29 / cl.c:1064-1071.  Pointer retained, read at next process. */
void fosphor_cl_load_fft_window(struct fosphor *self, float *win);

/* cl.h:30 / cl.c:1073-1079 */
int fosphor_cl_get_waterfall_position(struct fosphor *self);

/* cl.h:31-32 / cl.c:1081-1089.  Stores scale * 128 and offset. */
void fosphor_cl_set_histogram_range(struct fosphor *self, float scale, float offset);

/* ------------------------------------------------------------------------ */
/* 2. Parameterised engine                                                    */
/* ------------------------------------------------------------------------ */

#define FOSPHOR_CU_MAX_BINS 1024

struct fosphor_cu_params {
	int fft_len;        /* N: 512, 1024, 2048, 4096, 8192 or 16384 (ref: private.h:21-22) */
	int n_bins;         /* K power bins, 2..FOSPHOR_CU_MAX_BINS (ref: 128, display.cl:96)  */
	int wf_rows;        /* W waterfall ring rows, power of two >= batch_max (ref: 1024)   */
	int batch_mult;     /* spectra per call must be a multiple of this (ref: 16)          */
	int batch_max;      /* and at most this                              (ref: 1024)      */
	float histo_t0r;    /* rise time constant   (ref 16.0,   cl.c:714) */
	float histo_t0d;    /* decay time constant  (ref 1024.0, cl.c:715) */
	float live_alpha;   /* live IIR alpha       (ref 0.002,  cl.c:716) */
	float maxhold_keep; /* ref 0.999, display.cl:303 */
	float maxhold_mix;  /* ref 0.001, display.cl:303 */
	int device;         /* CUDA device ordinal, -1 = current device */
	int scratch_rows;   /* The engine's internal log-power ring, which decides how many calls
	                     * fosphor_cu_process_device_multi / _host_raw fold into one launch pair
	                     * (independent of wf_rows, which is only what the display shows):
	                     * 0 = automatic (grown on demand up to 1 GiB, an eighth of the free
	                     * device memory at most), > 0 = at most this many rows (rounded up to
	                     * a power of two), < 0 = never (the waterfall itself is the ring: one
	                     * launch pair per wf_rows rows, as in round 1). */
};

struct fosphor_cu; /* opaque */

void fosphor_cu_default_params(struct fosphor_cu_params *p);
int  fosphor_cu_create(struct fosphor_cu **out, const struct fosphor_cu_params *p);
void fosphor_cu_destroy(struct fosphor_cu *e);

/* Enqueue all subsequent work on this CUDA stream (a cudaStream_t passed as
 * void*); NULL restores the engine's own non-blocking stream. */
int fosphor_cu_set_stream(struct fosphor_cu *e, void *cuda_stream);

/* Copies N floats now (cl.c:889-900 uploads lazily; here the copy is taken
 * at call time, the upload is stream ordered). */
int fosphor_cu_load_fft_window(struct fosphor_cu *e, const float *win_host);
/* Stores scale * n_bins and offset (cl.c:1081-1089). */
int fosphor_cu_set_histogram_range(struct fosphor_cu *e, float scale, float offset);

/* What the caller of the boundary computes for the reference's fixed N = 1024, for any N
 * (SURVEY.md 8a T2): the default window of fosphor.c:108-121 (periodic Hamming x 1.855, pi
 * truncated to 3.141592f) and the power range -> (scale, offset) math of fosphor.c:131-152
 * (offset = -(log10f(N) + db0/20), scale = 20/(db1 - db0), db0 = db_ref - 10 db_per_div). */
void fosphor_cu_default_window(int fft_len, float *win_out);
void fosphor_cu_power_range(int fft_len, int db_ref, int db_per_div, float *scale, float *offset);

/* One reference-style call on HOST samples: len complex samples = n_spectra
 * pre-overlapped windows (cl.c:870-968).  The source buffer is free when the call
 * returns (base_sink_c_impl.cc:170-174).  How the samples travel:
 *   - page-locked source (cudaHostAlloc / cudaHostRegister / fosphor_fifo_*): DMA in place;
 *   - pageable source (the unmodified sink's FIFO, lib/fifo.cc:17-21): staged by a few copy
 *     threads, piece by piece, the DMA of a piece overlapping the copy of the next
 *     (FOSPHOR_B200_COPY_THREADS, default min(4, cores/2)) - or, better, page-locked by the
 *     engine where it lies and DMA'd in place from then on.  FOSPHOR_B200_HOSTREG selects:
 *       unset  automatic: a call range is registered the second time it is seen, and only if the
 *              process can read its physical page numbers (/proc/self/pagemap: privileged
 *              processes); every later use compares a few of them, so a buffer that was freed and
 *              reallocated at the same address is detected, unregistered and staged instead;
 *       1      register on first sight without that check - for callers whose sample memory
 *              outlives the engine, as the sink's FIFO does (the setting for an unprivileged sink);
 *       0      always stage. */
int fosphor_cu_process_host(struct fosphor_cu *e, const void *samples_host, int len);

/* One call on DEVICE-resident samples.  Spectrum s is read at
 * samples_dev + s*hop complex samples; hop == fft_len is the pre-overlapped
 * layout, hop = fft_len/overlap does the overlap block's job
 * (lib/overlap_cc_impl.cc:64-79) without materialising it.  Asynchronous. */
int fosphor_cu_process_device(struct fosphor_cu *e, const void *samples_dev,
                              int n_spectra, long long hop);

/* n_calls consecutive calls of `batch` spectra each, results identical to
 * n_calls fosphor_cu_process_device() calls (the per-call semantics of
 * display.cl:241-245,303 are kept) but the FFT pass is launched once per
 * wf_rows spectra.  Asynchronous. */
int fosphor_cu_process_device_multi(struct fosphor_cu *e, const void *samples_dev,
                                    int n_calls, int batch, long long hop);

/* HOST raw (not pre-overlapped) stream: n_spectra windows hopping `hop`,
 * (n_spectra-1)*hop + fft_len complex samples are read.  n_spectra is split
 * into calls of `batch`.  The copy moves each raw sample once. */
int fosphor_cu_process_host_raw(struct fosphor_cu *e, const void *raw_host,
                                int n_calls, int batch, long long hop);

/* cl.c:970-1061 semantics: 1 = results copied into the given host arrays
 * (any may be NULL): waterfall [wf_rows][N], histogram [n_bins][N],
 * spectrum float2 live[N] + float2 max[N]; 0 = nothing new. Synchronises. */
int fosphor_cu_finish(struct fosphor_cu *e, float *waterfall_host,
                      float *histogram_host, float *spectrum_host);
/* Same, but only the waterfall rows written since the previous finish are copied (all of them the
 * first time and after >= wf_rows new rows): right when waterfall_host is the caller's persistent
 * image of the ring, as self->img_waterfall is (fosphor.c:52, cl.c:1012-1021 re-reads all 4 MiB
 * per frame).  first_row / n_rows (may be NULL) report the ring rows refreshed (n_rows may wrap). */
int fosphor_cu_finish_new_rows(struct fosphor_cu *e, float *waterfall_host,
                               float *histogram_host, float *spectrum_host,
                               int *first_row, int *n_rows);
/* Wait for all enqueued work (no copies). */
int fosphor_cu_sync(struct fosphor_cu *e);
/* Order the engine's stream after all process work enqueued so far, without waiting on the host.
 * Large device-resident batches run their accumulate kernels on a second internal stream and a
 * process call does NOT join it on return (the next call's FFT may start while the last accumulate
 * launch still runs).  finish / sync / export_maxhold / set_stream join by themselves; callers
 * that read the device arrays below from their own work on the engine's stream call this first. */
int fosphor_cu_flush(struct fosphor_cu *e);
/* Diagnostics: chunks that went through the two-stream schedule since create. */
unsigned long long fosphor_cu_two_stream_chunks(const struct fosphor_cu *e);

int fosphor_cu_get_waterfall_position(const struct fosphor_cu *e);

/* "Plain device arrays" (BASELINE.json north_star): valid until destroy; contents are
 * stream ordered after fosphor_cu_flush(). */
float *fosphor_cu_device_waterfall(struct fosphor_cu *e);  /* [wf_rows][N]      */
float *fosphor_cu_device_histogram(struct fosphor_cu *e);  /* [n_bins][N]       */
float *fosphor_cu_device_spectrum(struct fosphor_cu *e);   /* [2][N][2]         */

/* Write the max-hold trace (N floats, display order) to a device buffer;
 * input to the multi-GPU ncclMax reduce.  Asynchronous on the engine stream. */
int fosphor_cu_export_maxhold(struct fosphor_cu *e, float *out_dev);

/* Same on a stream of the caller's (cudaStream_t as void*), e.g. the one its NCCL all-reduce runs
 * on: that stream is ordered after the engine's last accumulate launch, the engine's own stream is
 * NOT - the next process call's FFT still overlaps that launch (export_maxhold would join the two
 * streams and drain the pipeline once per reduction).  The next accumulate launch waits for this
 * read.  Work the caller enqueues on side_stream afterwards sees the exported trace. */
int fosphor_cu_export_maxhold_on(struct fosphor_cu *e, float *out_dev, void *side_stream);

/* Test hook: windowed forward FFT only, cf32 [n_spectra][N] out (device). */
int fosphor_cu_debug_fft(struct fosphor_cu *e, const void *samples_dev,
                         int n_spectra, long long hop, void *out_dev);

/* Per-kernel device timing for the roofline report: when enabled, every launch
 * of the three hot kernels is bracketed by CUDA events on the engine stream.
 * profile_read() synchronises, fills ms[3] / launches[3] (0 = fft_power,
 * 1 = count, 2 = update) with the summed durations and launch counts since
 * the last read, and resets the counters. */
int fosphor_cu_profile(struct fosphor_cu *e, int enable);
int fosphor_cu_profile_read(struct fosphor_cu *e, double *ms, unsigned long long *launches);

/* Diagnostics of the host-fed path: process calls whose samples were staged by the copy threads /
 * DMA'd straight from page-locked caller memory, copy threads running, rows of the log-power ring. */
int fosphor_cu_host_feed_stats(const struct fosphor_cu *e, unsigned long long *staged_calls,
                               unsigned long long *direct_calls, int *copy_threads, int *ring_rows);
/* Number of kernel launches issued by this engine so far. */
unsigned long long fosphor_cu_launch_count(const struct fosphor_cu *e);
const char *fosphor_cu_last_error(const struct fosphor_cu *e);

/* ------------------------------------------------------------------------ */
/* 3. Page-locked sample FIFO (sink side)                    */
/* ------------------------------------------------------------------------ */
/* C handle of fosphor_b200::pinned_fifo (gr-fosphor_b200/host/pinned_fifo.h):
 * the reference's gr::fosphor::fifo (lib/fifo.h:20-46) with page-locked
 * storage, so process() DMAs straight out of the ring.  length: power of two,
 * in complex samples.  write_prepare / read_peek return NULL when wait == 0
 * and the request cannot be met. */
void *fosphor_fifo_create(int length);
void  fosphor_fifo_destroy(void *fifo);
int   fosphor_fifo_is_pinned(void *fifo);
int   fosphor_fifo_free(void *fifo);
int   fosphor_fifo_used(void *fifo);
int   fosphor_fifo_write_max_size(void *fifo);
int   fosphor_fifo_read_max_size(void *fifo);
void *fosphor_fifo_write_prepare(void *fifo, int size, int wait);
void  fosphor_fifo_write_commit(void *fifo, int size);
void *fosphor_fifo_read_peek(void *fifo, int size, int wait);
void  fosphor_fifo_read_discard(void *fifo, int size);

/* Host-only self test of the staging copy pool (gr-fosphor_b200/host/copy_pool.h): jobs of `bytes`
 * bytes in `pieces` in-order pieces on `threads` workers; 0 = every copy exact.  Needs no GPU. */
int fosphor_host_copy_selftest(int threads, unsigned long long bytes, int pieces, int rounds);

/* ------------------------------------------------------------------------ */
/* 4. FFT window generator                                  */
/* ------------------------------------------------------------------------ */
/* What the sink obtains from gr::fft::window::build(type, 1024, 6.76)
 * (lib/base_sink_c_impl.cc:251-255); type uses gr::fft::window::win_type
 * numbering: 0 Hamming, 1 Hann, 2 Blackman, 3 rectangular, 4 Kaiser(beta),
 * 5 Blackman-harris, 6 Bartlett, 7 flat-top.  0 / -1. */
int fosphor_window_build(int type, int n, double beta, float *out);

#ifdef __cplusplus
}
#endif
#endif
