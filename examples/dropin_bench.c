/*
 * The host-fed drop-in figure of bench.py (`e2e`) from plain C, no Python in the loop: sink frames as
 * lib/base_sink_c_impl.cc:130-175 produces them - 8 x fosphor_cl_process() of 1024 spectra out of
 * an ordinary malloc()ed ring, then fosphor_cl_finish() into the malloc()ed images fosphor_init()
 * owns (fosphor.c:52-54).
 *
 *   gcc -std=c99 -O2 -Wall -pedantic -D_POSIX_C_SOURCE=199309L -Iinclude examples/dropin_bench.c \
 *       -o /tmp/dropin_bench -Lgr-fosphor_b200 -lfosphor_b200 -Wl,-rpath,$PWD/gr-fosphor_b200 -lm
 *   /tmp/dropin_bench [frames]          (FOSPHOR_B200_HOSTREG=0 / 1 to force staging / registration)
 */
#include <errno.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "fosphor_b200.h"
#include "fosphor_private_abi.h"

static double now(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int main(int argc, char **argv)
{
	struct fosphor self;
	const int n = FOSPHOR_FFT_LEN, b = FOSPHOR_FFT_MAX_BATCH, calls = 8;
	const int frames = argc > 1 ? atoi(argv[1]) : 400;
	const size_t call_len = (size_t)n * b;
	float *ring;
	double t0, el;
	int rc, i, f, c;

	memset(&self, 0, sizeof(self));
	self.img_waterfall = malloc(sizeof(float) * 1024 * n);
	self.img_histogram = malloc(sizeof(float) * 128 * n);
	self.buf_spectrum = malloc(sizeof(float) * 2 * 2 * n);
	for (i = 0; i < n; i++)                      /* fosphor.c:108-121 */
		self.fft_win[i] = (0.54f - 0.46f * cosf(2.0f * 3.141592f * i / n)) * 1.855f;
	rc = fosphor_cl_init(&self);
	if (rc) {
		fprintf(stderr, "fosphor_cl_init: %d (%s)\n", rc, rc == -EIO ? "-EIO: no usable CUDA device" : "error");
		return rc == -EIO ? 2 : 1;
	}
	fosphor_cl_load_fft_window(&self, self.fft_win);
	fosphor_cl_set_histogram_range(&self, 0.2f, 1.9896998f);

	ring = malloc(sizeof(float) * 2 * call_len * calls);   /* 64 MiB of pageable samples */
	for (i = 0; i < (int)(call_len * calls); i++) {
		ring[2 * i] = 0.3f * cosf(0.37f * (float)(i & 0xffff)) + 0.01f * (float)((i * 2654435761u >> 20) & 255) / 255.0f;
		ring[2 * i + 1] = 0.3f * sinf(0.37f * (float)(i & 0xffff));
	}
	for (f = -5, t0 = 0.0; f < frames; f++) {
		if (f == 0)
			t0 = now();
		for (c = 0; c < calls; c++)
			if ((rc = fosphor_cl_process(&self, ring + 2 * call_len * c, (int)call_len)) != 0)
				return 1;
		if (fosphor_cl_finish(&self) != 1)
			return 1;
	}
	el = now() - t0;
	printf("{\"frames\": %d, \"ms_per_frame\": %.4f, \"Msamples_per_s\": %.1f, \"PCIe_GBps\": %.1f}\n", frames,
	       el / frames * 1e3, (double)frames * calls * (double)call_len / el / 1e6,
	       (double)frames * calls * (double)call_len * 8.0 / el / 1e9);
	fosphor_cl_release(&self);
	free(ring);
	free(self.img_waterfall);
	free(self.img_histogram);
	free(self.buf_spectrum);
	return 0;
}
