/*
 * fosphor_oracle.c - CPU restatement of fosphor's spectral hot path.
 *
 * TEST INFRASTRUCTURE ONLY (pinned by golden vectors from the reference itself,
 * see fosphor_oracle.h).
 * Written from the behaviour of the reference sources cited inline
 * (paths relative to the reference's lib/fosphor/ unless noted); it is not
 * a translation of the OpenCL kernels: the device code there is organised
 * around 16x16 work-groups and local memory, here every frequency column is
 * simply walked in the same arithmetic order.
 */
#include <errno.h>
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "fosphor_oracle.h"

/* The reference's display kernel interleaves the spectra of a call over 16
 * work-item rows (display.cl:67 reqd_work_group_size(16,16,1), :130); the
 * f32 summation order of the live spectrum and the "sum / 16" fallback
 * (display.cl:206-207) depend on it. */
#define ROWS 16

enum { ST_BOOTING = 0, ST_PENDING, ST_READY }; /* cl.c:95-99 */

struct fosphor_oracle {
	struct fosphor_oracle_params p;
	int log2n;
	float *win;        /* [N] */
	float histo_scale; /* scale * K (cl.c:1087) */
	float histo_ofs;
	int wf_pos;
	int state;

	float *waterfall;  /* [W][N] */
	float *histogram;  /* [K][N] */
	float *spectrum;   /* [2][N][2] */

	float *fft_out;    /* [batch_max][N][2] */
	unsigned *hits;    /* [K][N] */
	int last_batch;

	double *tw_d;      /* [N/2][2] forward twiddles, double */
	float *tw_f;       /* [N/2][2] forward twiddles, f32 */
	unsigned *brev;    /* [N] bit reversal */
};

void
fosphor_oracle_default_params(struct fosphor_oracle_params *p)
{
	p->fft_len = 1024;      /* private.h:21-22 */
	p->n_bins = 128;        /* display.cl:96 */
	p->wf_rows = 1024;      /* cl.c:430-432 */
	p->batch_mult = 16;     /* private.h:24 */
	p->batch_max = 1024;    /* private.h:25 */
	p->histo_t0r = 16.0f;   /* cl.c:714 */
	p->histo_t0d = 1024.0f; /* cl.c:715 */
	p->live_alpha = 0.002f; /* cl.c:716 */
	p->maxhold_keep = 0.999f; /* display.cl:303 */
	p->maxhold_mix = 0.001f;
	p->fft_f32 = 0;
}

static int
ilog2(int v)
{
	int l = 0;
	while ((1 << l) < v)
		l++;
	return l;
}

struct fosphor_oracle *
fosphor_oracle_create(const struct fosphor_oracle_params *p)
{
	struct fosphor_oracle *o;
	int n = p->fft_len, i, b;

	if (n < 16 || (n & (n - 1)) || p->n_bins < 2 ||
	    p->wf_rows < 1 || (p->wf_rows & (p->wf_rows - 1)) ||
	    p->batch_mult < 1 || p->batch_max < p->batch_mult)
		return NULL;

	o = calloc(1, sizeof(*o));
	if (!o)
		return NULL;
	o->p = *p;
	o->log2n = ilog2(n);
	o->state = ST_BOOTING;

	o->win = malloc(sizeof(float) * n);
	o->waterfall = malloc(sizeof(float) * (size_t)p->wf_rows * n);
	o->histogram = malloc(sizeof(float) * (size_t)p->n_bins * n);
	o->spectrum = malloc(sizeof(float) * 4 * n);
	o->fft_out = malloc(sizeof(float) * 2 * (size_t)p->batch_max * n);
	o->hits = malloc(sizeof(unsigned) * (size_t)p->n_bins * n);
	o->tw_d = malloc(sizeof(double) * n);
	o->tw_f = malloc(sizeof(float) * n);
	o->brev = malloc(sizeof(unsigned) * n);
	if (!o->win || !o->waterfall || !o->histogram || !o->spectrum ||
	    !o->fft_out || !o->hits || !o->tw_d || !o->tw_f || !o->brev) {
		fosphor_oracle_destroy(o);
		return NULL;
	}

	for (i = 0; i < n; i++)
		o->win[i] = 1.0f;
	for (i = 0; i < n / 2; i++) {
		double a = -2.0 * M_PI * (double)i / (double)n;
		o->tw_d[2 * i] = cos(a);
		o->tw_d[2 * i + 1] = sin(a);
		o->tw_f[2 * i] = (float)o->tw_d[2 * i];
		o->tw_f[2 * i + 1] = (float)o->tw_d[2 * i + 1];
	}
	for (i = 0; i < n; i++) {
		unsigned r = 0;
		for (b = 0; b < o->log2n; b++)
			if (i & (1 << b))
				r |= 1u << (o->log2n - 1 - b);
		o->brev[i] = r;
	}

	/* cl.c zero-initialises histo_scale / histo_offset (memset, :811);
	 * fosphor_init then calls fosphor_set_power_range(0, 10). */
	o->histo_scale = 0.0f;
	o->histo_ofs = 0.0f;
	memset(o->hits, 0, sizeof(unsigned) * (size_t)p->n_bins * n);
	return o;
}

void
fosphor_oracle_destroy(struct fosphor_oracle *o)
{
	if (!o)
		return;
	free(o->win);
	free(o->waterfall);
	free(o->histogram);
	free(o->spectrum);
	free(o->fft_out);
	free(o->hits);
	free(o->tw_d);
	free(o->tw_f);
	free(o->brev);
	free(o);
}

void
fosphor_oracle_load_fft_window(struct fosphor_oracle *o, const float *win)
{
	memcpy(o->win, win, sizeof(float) * o->p.fft_len);
}

void
fosphor_oracle_default_window(int fft_len, float *win)
{
	/* fosphor.c:113-118: periodic Hamming x 1.855, pi truncated to 3.141592f */
	int i;
	for (i = 0; i < fft_len; i++) {
		float ft = (float)fft_len;
		float fp = (float)i;
		win[i] = (0.54f - 0.46f * cosf((2.0f * 3.141592f * fp) / ft)) * 1.855f;
	}
}

void
fosphor_oracle_power_range(int fft_len, int db_ref, int db_per_div,
                           float *scale, float *offset)
{
	/* fosphor.c:131-152 */
	int db0 = db_ref - 10 * db_per_div;
	int db1 = db_ref;
	float k = log10f((float)fft_len);
	*offset = -(k + ((float)db0 / 20.0f));
	*scale = 20.0f / (float)(db1 - db0);
}

void
fosphor_oracle_set_histogram_range(struct fosphor_oracle *o, float scale, float offset)
{
	o->histo_scale = scale * (float)o->p.n_bins; /* cl.c:1087 */
	o->histo_ofs = offset;
}

static void
clear_buffers(struct fosphor_oracle *o)
{
	/* cl.c:406-465: spectrum (all 4N floats) and waterfall = -power.offset,
	 * histogram = 0 */
	float nf = -o->histo_ofs;
	size_t i, n = o->p.fft_len;
	for (i = 0; i < 4 * n; i++)
		o->spectrum[i] = nf;
	for (i = 0; i < (size_t)o->p.wf_rows * n; i++)
		o->waterfall[i] = nf;
	memset(o->histogram, 0, sizeof(float) * (size_t)o->p.n_bins * n);
}

/* ---- FFT ---------------------------------------------------------------- */

/* fft.cl:416-417 then :419-462: x*win in f32, forward unnormalised DFT,
 * natural order.  Parity variant: transform in double, round once. */
static void
fft_double(const struct fosphor_oracle *o, const float *in, float *out, double *buf)
{
	int n = o->p.fft_len, i, len, j, k;

	for (i = 0; i < n; i++) {
		float re = in[2 * i] * o->win[i];
		float im = in[2 * i + 1] * o->win[i];
		buf[2 * o->brev[i]] = re;
		buf[2 * o->brev[i] + 1] = im;
	}
	for (len = 2; len <= n; len <<= 1) {
		int half = len >> 1, step = n / len;
		for (j = 0; j < n; j += len) {
			for (k = 0; k < half; k++) {
				double wr = o->tw_d[2 * k * step], wi = o->tw_d[2 * k * step + 1];
				double *a = &buf[2 * (j + k)], *b = &buf[2 * (j + k + half)];
				double tr = b[0] * wr - b[1] * wi;
				double ti = b[0] * wi + b[1] * wr;
				b[0] = a[0] - tr;
				b[1] = a[1] - ti;
				a[0] += tr;
				a[1] += ti;
			}
		}
	}
	for (i = 0; i < 2 * n; i++)
		out[i] = (float)buf[i];
}

/* Timing variant: same transform in f32 (reference-class arithmetic). */
static void
fft_float(const struct fosphor_oracle *o, const float *in, float *out)
{
	int n = o->p.fft_len, i, len, j, k;

	for (i = 0; i < n; i++) {
		float w = o->win[i];
		out[2 * o->brev[i]] = in[2 * i] * w;
		out[2 * o->brev[i] + 1] = in[2 * i + 1] * w;
	}
	for (len = 2; len <= n; len <<= 1) {
		int half = len >> 1, step = n / len;
		for (j = 0; j < n; j += len) {
			float *a = &out[2 * j], *b = &out[2 * (j + half)];
			for (k = 0; k < half; k++) {
				float wr = o->tw_f[2 * k * step], wi = o->tw_f[2 * k * step + 1];
				float br = b[2 * k], bi = b[2 * k + 1];
				float tr = br * wr - bi * wi;
				float ti = br * wi + bi * wr;
				b[2 * k] = a[2 * k] - tr;
				b[2 * k + 1] = a[2 * k + 1] - ti;
				a[2 * k] += tr;
				a[2 * k + 1] += ti;
			}
		}
	}
}

/* ---- display ------------------------------------------------------------ */

static inline float
log_power(const struct fosphor_oracle *o, float re, float im)
{
	/* display.cl:136  pwr = log10(hypot(x, y)); magnitude, not power */
	if (o->p.fft_f32)
		return log10f(hypotf(re, im));
	return (float)log10(hypot((double)re, (double)im));
}

static inline int
map_bin(float x, int k)
{
	/* display.cl:161-165: (int)round(), half away from zero, clamp to
	 * [0, K-1].  Non-finite: NaN / -inf -> 0, +inf -> K-1 (the OpenCL
	 * conversion is implementation-defined there; DESIGN.md fixes this). */
	float r = roundf(x);
	if (!(r > 0.0f))
		return 0;
	if (r > (float)(k - 1))
		return k - 1;
	return (int)r;
}

/* pwr_in == NULL: log-power from the complex spectra of this call (the path);
 * pwr_in != NULL: [batch][N] log-power rows given by the caller - the same
 * display arithmetic on somebody else's waterfall rows (stage-wise parity:
 * fosphor_oracle_process_pwr) */
static void
display(struct fosphor_oracle *o, int batch, const float *pwr_in)
{
	const int n = o->p.fft_len, kb = o->p.n_bins, wmask = o->p.wf_rows - 1;
	const float alpha = o->p.live_alpha;
	const float oma = 1.0f - alpha;         /* display.cl:99 */
	const float fb = (float)batch;
	const float rt0r = 1.0f / o->p.histo_t0r; /* native_recip */
	const float rt0d = 1.0f / o->p.histo_t0d;
	const float live_carry = powf(oma, fb);  /* display.cl:210 */
	float *wts = malloc(sizeof(float) * batch);
	int s;

	for (s = 0; s < batch; s++) /* display.cl:150 */
		wts[s] = powf(oma, (float)(batch - s - 1));

	memset(o->hits, 0, sizeof(unsigned) * (size_t)kb * n);

#pragma omp parallel for schedule(static)
	for (int f = 0; f < n; f++) {
		float live_buf[ROWS], max_buf[ROWS];
		float sum, y, m;
		int r, i, b;

		for (r = 0; r < ROWS; r++) {
			live_buf[r] = 0.0f;    /* display.cl:113 */
			max_buf[r] = -1000.0f; /* display.cl:91 */
		}

		for (int sp = 0; sp < batch; sp++) {
			const float *x = &o->fft_out[2 * ((size_t)sp * n + f)];
			float pwr = pwr_in ? pwr_in[(size_t)sp * n + f] : log_power(o, x[0], x[1]);
			r = sp % ROWS;
			max_buf[r] = fmaxf(max_buf[r], pwr);                 /* :139 */
			o->waterfall[(size_t)((o->wf_pos + sp) & wmask) * n + f] = pwr; /* :141-146 */
			live_buf[r] += pwr * wts[sp];                        /* :149-150 */
			b = map_bin(o->histo_scale * (pwr + o->histo_ofs), kb); /* :161 */
			o->hits[(size_t)b * n + f]++;                        /* :170-177 */
		}

		/* Live spectrum, display.cl:186-214 */
		sum = 0.0f;
		for (r = 0; r < ROWS; r++)
			sum += live_buf[r];
		i = f ^ (n >> 1);
		y = o->spectrum[2 * i + 1];
		if (!isfinite(y))
			y = sum / (float)ROWS;
		o->spectrum[2 * i] = ((float)i / (float)(n >> 1)) - 1.0f;
		o->spectrum[2 * i + 1] = y * live_carry + sum * alpha;

		/* Histogram rise / decay, display.cl:217-254 */
		for (b = 0; b < kb; b++) {
			float hv = o->histogram[(size_t)b * n + f];
			unsigned hc = o->hits[(size_t)b * n + f];
			float a, bb, c, d, e;
			if (hv <= 0.01f && hc == 0)
				continue;
			a = (float)hc / fb;
			bb = a * rt0r;
			c = bb + rt0d;
			d = bb * (1.0f / c);
			e = powf(1.0f - c, fb);
			hv = (hv - d) * e + d;
			hv = fminf(fmaxf(hv, 0.0f), 1.0f);
			o->histogram[(size_t)b * n + f] = hv;
		}

		/* Max hold with decay, display.cl:257-310 */
		m = o->spectrum[2 * (n + i) + 1];
		if (!isfinite(m))
			m = -FLT_MAX;
		m = m * o->p.maxhold_keep + o->p.maxhold_mix * o->spectrum[2 * i + 1];
		for (r = 0; r < ROWS; r++)
			m = fmaxf(m, max_buf[r]);
		o->spectrum[2 * (n + i)] = ((float)i / (float)(n >> 1)) - 1.0f;
		o->spectrum[2 * (n + i) + 1] = m;
	}

	free(wts);
}

static int
process_common(struct fosphor_oracle *o, const float *src, int n_spectra, int hop)
{
	const int n = o->p.fft_len;

	/* cl.c:881-886 */
	if (n_spectra % o->p.batch_mult)
		return -EINVAL;
	if (n_spectra > o->p.batch_max)
		return -EINVAL;

#pragma omp parallel
	{
		double *buf = o->p.fft_f32 ? NULL : malloc(sizeof(double) * 2 * n);
#pragma omp for schedule(static)
		for (int s = 0; s < n_spectra; s++) {
			const float *in = &src[2 * (size_t)s * hop];
			float *out = &o->fft_out[2 * (size_t)s * n];
			if (o->p.fft_f32)
				fft_float(o, in, out);
			else
				fft_double(o, in, out, buf);
		}
		free(buf);
	}

	if (o->state == ST_BOOTING) /* cl.c:930-934 */
		clear_buffers(o);

	if (n_spectra > 0)
		display(o, n_spectra, NULL);
	o->last_batch = n_spectra;

	o->wf_pos = (o->wf_pos + n_spectra) & (o->p.wf_rows - 1); /* cl.c:954 */
	o->state = ST_PENDING;
	return 0;
}

int
fosphor_oracle_process_pwr(struct fosphor_oracle *o, const float *pwr_rows, int n_spectra)
{
	/* cl.c:881-886, then display.cl:141-310 on the given rows */
	if (n_spectra < 0 || (n_spectra % o->p.batch_mult) || n_spectra > o->p.batch_max)
		return -EINVAL;
	if (o->state == ST_BOOTING) /* cl.c:930-934 */
		clear_buffers(o);
	if (n_spectra > 0)
		display(o, n_spectra, pwr_rows);
	o->last_batch = n_spectra;
	o->wf_pos = (o->wf_pos + n_spectra) & (o->p.wf_rows - 1); /* cl.c:954 */
	o->state = ST_PENDING;
	return 0;
}

int
fosphor_oracle_process(struct fosphor_oracle *o, const float *samples, int len)
{
	/* cl.c:881: len must be a multiple of batch_mult * N */
	if (len < 0 || (len % (o->p.batch_mult * o->p.fft_len)))
		return -EINVAL;
	return process_common(o, samples, len / o->p.fft_len, o->p.fft_len);
}

int
fosphor_oracle_process_hop(struct fosphor_oracle *o, const float *raw,
                           int n_spectra, int hop)
{
	if (n_spectra < 0 || hop < 1)
		return -EINVAL;
	return process_common(o, raw, n_spectra, hop);
}

int
fosphor_oracle_finish(struct fosphor_oracle *o)
{
	if (o->state == ST_READY) /* cl.c:978-979 */
		return 0;
	if (o->state == ST_BOOTING) /* cl.c:982-994 */
		clear_buffers(o);
	o->state = ST_READY;
	return 1;
}

int
fosphor_oracle_get_waterfall_position(const struct fosphor_oracle *o)
{
	return o->wf_pos;
}

const float *fosphor_oracle_waterfall(const struct fosphor_oracle *o) { return o->waterfall; }
const float *fosphor_oracle_histogram(const struct fosphor_oracle *o) { return o->histogram; }
const float *fosphor_oracle_spectrum(const struct fosphor_oracle *o) { return o->spectrum; }
const float *fosphor_oracle_last_fft(const struct fosphor_oracle *o) { return o->fft_out; }
const unsigned *fosphor_oracle_last_hits(const struct fosphor_oracle *o) { return o->hits; }
int fosphor_oracle_last_batch(const struct fosphor_oracle *o) { return o->last_batch; }

int
fosphor_oracle_threads(void)
{
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}
