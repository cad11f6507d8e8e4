/*
 * fosphor_oracle.h - CPU restatement of fosphor's spectral hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (gr-fosphor_b200/,
 * include/) may include, link or call this.  Only tests/, bench.py's
 * cpu_baseline / --impl reference leg and __graft_entry__.smoke() use it,
 * and only as the checker.
 *
 * PARITY STATUS: "parity unpinned".  The reference (osmocom/gr-fosphor
 * @74d54fc) ships no tests, golden vectors or fixtures for this path and its
 * device code is OpenCL C that cannot be compiled in the build container
 * (no OpenCL compiler / ICD).  The pins of this oracle are the
 * analytic known-answer tests in tests/test_oracle_kat.py, derived from the
 * reference sources cited below.
 *
 * What is restated (reference file:line, all under lib/fosphor/):
 *   fft.cl:397-466      window multiply in f32, forward unnormalised DFT,
 *                       natural bin order
 *   display.cl:130-178  log10(hypot) power, waterfall rows, live weights,
 *                       bin mapping + clamp, hit counting
 *   display.cl:186-214  live spectrum IIR merge
 *   display.cl:217-254  histogram rise / decay
 *   display.cl:257-310  max-hold with decay (MAX_HOLD_DECAY variant)
 *   cl.c:406-465        first-use clears
 *   cl.c:870-968        process(): validation, ring advance, state machine
 *   cl.c:970-1061       finish(): state machine / return codes
 *   cl.c:1064-1089      window / range setters (scale * n_bins)
 *   overlap_cc_impl.cc:64-79 (lib/)  overlap addressing (hop = N / overlap)
 */
#ifndef FOSPHOR_ORACLE_H
#define FOSPHOR_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

struct fosphor_oracle_params {
	int fft_len;     /* N, power of two (reference: 1024, private.h:21-22)   */
	int n_bins;      /* K power bins    (reference: 128, display.cl:96)      */
	int wf_rows;     /* W waterfall rows, power of two (reference: 1024)     */
	int batch_mult;  /* spectra per call multiple (reference: 16)            */
	int batch_max;   /* max spectra per call      (reference: 1024)          */
	float histo_t0r; /* rise  time constant (reference 16.0,   cl.c:714)     */
	float histo_t0d; /* decay time constant (reference 1024.0, cl.c:715)     */
	float live_alpha;/* live IIR alpha      (reference 0.002,  cl.c:716)     */
	float maxhold_keep; /* 0.999 (display.cl:303) */
	float maxhold_mix;  /* 0.001 (display.cl:303) */
	int fft_f32;     /* 0: double-precision FFT rounded to f32 (parity)      *
	                  * 1: f32 radix FFT (CPU-baseline timing variant)       */
};

struct fosphor_oracle;

void fosphor_oracle_default_params(struct fosphor_oracle_params *p);

struct fosphor_oracle *fosphor_oracle_create(const struct fosphor_oracle_params *p);
void fosphor_oracle_destroy(struct fosphor_oracle *o);

/* cl.c:1064-1071 (copies N floats instead of keeping the pointer) */
void fosphor_oracle_load_fft_window(struct fosphor_oracle *o, const float *win);
/* fosphor.c:108-121 default window generalised to N */
void fosphor_oracle_default_window(int fft_len, float *win);
/* fosphor.c:131-152 */
void fosphor_oracle_power_range(int fft_len, int db_ref, int db_per_div,
                                float *scale, float *offset);
/* cl.c:1081-1089: stores scale * n_bins, offset */
void fosphor_oracle_set_histogram_range(struct fosphor_oracle *o,
                                        float scale, float offset);

/* cl.c:870-968.  samples: interleaved cf32, len complex samples, windows
 * already laid out back to back (pre-overlapped).  0 / -EINVAL. */
int fosphor_oracle_process(struct fosphor_oracle *o, const float *samples, int len);
/* Same but windows taken from a raw stream with hop complex samples between
 * the starts of consecutive spectra (overlap_cc_impl.cc:64-79). */
int fosphor_oracle_process_hop(struct fosphor_oracle *o, const float *raw,
                               int n_spectra, int hop);
/* cl.c:970-1061: 1 = new results, 0 = nothing new */
int fosphor_oracle_finish(struct fosphor_oracle *o);
int fosphor_oracle_get_waterfall_position(const struct fosphor_oracle *o);

/* Result arrays (valid after finish or process; host memory, row major):
 *   waterfall [wf_rows][N], histogram [n_bins][N],
 *   spectrum  float2 live[N] then float2 max[N] in display order */
const float *fosphor_oracle_waterfall(const struct fosphor_oracle *o);
const float *fosphor_oracle_histogram(const struct fosphor_oracle *o);
const float *fosphor_oracle_spectrum(const struct fosphor_oracle *o);

/* Intermediates of the most recent process call, for stage-wise parity:
 *   fft_out  cf32 [B][N]  (what fft.cl leaves in mem_fft_out)
 *   hits     u32  [n_bins][N] hit counts (display.cl:170-177) */
const float *fosphor_oracle_last_fft(const struct fosphor_oracle *o);
const unsigned *fosphor_oracle_last_hits(const struct fosphor_oracle *o);
int fosphor_oracle_last_batch(const struct fosphor_oracle *o);

int fosphor_oracle_threads(void);

#ifdef __cplusplus
}
#endif
#endif
