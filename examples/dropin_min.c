/*
 * Minimal C caller of the drop-in boundary (include/fosphor_b200.h): what lib/fosphor/fosphor.c does
 * with the seven fosphor_cl_* entry points (fosphor.c:47,86,95,101,103,120,127,151), without GL.
 *
 *   gcc -std=c99 -Wall -pedantic -Iinclude examples/dropin_min.c -o /tmp/dropin_min \
 *       -Lgr-fosphor_b200 -lfosphor_b200 -Wl,-rpath,$PWD/gr-fosphor_b200 -lm
 *
 * Exit code 0: one 64-spectrum burst processed and read back.  Exit code 2: no usable CUDA device -
 * fosphor_cl_init reported -EIO (there is no CPU fallback).  tests/test_abi.py builds and runs this.
 */
#include <errno.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fosphor_b200.h"
#include "fosphor_private_abi.h"

int main(void)
{
	struct fosphor self;
	const int n = FOSPHOR_FFT_LEN, b = 64;
	float *x;
	int rc, i;

	memset(&self, 0, sizeof(self));
	/* what fosphor_init() provides when CL/GL sharing is off (fosphor.c:52-62) */
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

	x = malloc(sizeof(float) * 2 * n * b);
	for (i = 0; i < n * b; i++) {                /* one tone */
		x[2 * i] = 0.5f * cosf(2.0f * 3.14159265f * 200.25f * i / n);
		x[2 * i + 1] = 0.5f * sinf(2.0f * 3.14159265f * 200.25f * i / n);
	}
	rc = fosphor_cl_process(&self, x, n * b);
	if (rc == 0)
		rc = fosphor_cl_finish(&self) == 1 ? 0 : 1;
	printf("waterfall position %d, live[200] = %f\n", fosphor_cl_get_waterfall_position(&self),
	       self.buf_spectrum[2 * ((200 ^ (n / 2))) + 1]);
	fosphor_cl_release(&self);
	free(x);
	free(self.img_waterfall);
	free(self.img_histogram);
	free(self.buf_spectrum);
	return rc;
}
