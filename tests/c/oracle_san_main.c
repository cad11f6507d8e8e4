/* Driver for the sanitizer build of the CPU oracle (tests/test_host_sanitizers.py): a burst, a few
 * calls of different sizes with a ring wrap, the display stage on given rows, the error paths. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "fosphor_oracle.h"

int main(void)
{
	struct fosphor_oracle_params p;
	struct fosphor_oracle *o;
	const int n = 1024, sizes[] = {64, 16, 1024, 48, 256};
	float *x, *win, scale, offset;
	double sum = 0.0;
	int i, c, rc;

	fosphor_oracle_default_params(&p);
	p.wf_rows = 1024;
	o = fosphor_oracle_create(&p);
	if (!o)
		return 1;
	win = malloc(sizeof(float) * n);
	fosphor_oracle_default_window(n, win);
	fosphor_oracle_load_fft_window(o, win);
	fosphor_oracle_power_range(n, 0, 10, &scale, &offset);
	fosphor_oracle_set_histogram_range(o, scale, offset);
	x = malloc(sizeof(float) * 2 * n * 1024);
	for (i = 0; i < n * 1024; i++) {
		x[2 * i] = 0.4f * cosf(0.7f * (float)(i % 4099)) + 0.01f * (float)((i * 2654435761u >> 16) & 255) / 255.0f;
		x[2 * i + 1] = 0.4f * sinf(0.7f * (float)(i % 4099));
	}
	if (fosphor_oracle_finish(o) != 1)             /* BOOTING -> clears */
		return 2;
	for (c = 0; c < 5; c++) {
		rc = fosphor_oracle_process(o, x, sizes[c] * n);
		if (rc)
			return 3;
		if (fosphor_oracle_finish(o) != 1)
			return 4;
	}
	if (fosphor_oracle_process(o, x, 15 * n) == 0)  /* cl.c:881-886 */
		return 5;
	/* display stage on the rows just produced (ring rows 0..63 hold the last call's tail) */
	if (fosphor_oracle_process_pwr(o, fosphor_oracle_waterfall(o), 64) != 0)
		return 6;
	if (fosphor_oracle_process_hop(o, x, 32, 256) != 0)
		return 7;
	fosphor_oracle_finish(o);
	for (i = 0; i < 128 * n; i++)
		sum += fosphor_oracle_histogram(o)[i];
	printf("histogram mass %.6f, wf_pos %d\n", sum, fosphor_oracle_get_waterfall_position(o));
	fosphor_oracle_destroy(o);
	free(x);
	free(win);
	return isfinite(sum) && sum > 0.0 ? 0 : 8;
}
