/*
 * fosphor_private_abi.h - the part of the reference's private state that the
 * fosphor_cl_* boundary shares with its caller, restated so the drop-in can be
 * built without the reference tree.  Must stay layout-identical to
 * lib/fosphor/private.h:21-55 (osmocom/gr-fosphor @74d54fc); when building
 * inside the reference tree, include its own private.h instead.
 */
#ifndef FOSPHOR_PRIVATE_ABI_H
#define FOSPHOR_PRIVATE_ABI_H

#define FOSPHOR_FFT_LEN_LOG     10                         /* private.h:21 */
#define FOSPHOR_FFT_LEN         (1 << FOSPHOR_FFT_LEN_LOG) /* private.h:22 */
#define FOSPHOR_FFT_MULT_BATCH  16                         /* private.h:24 */
#define FOSPHOR_FFT_MAX_BATCH   1024                       /* private.h:25 */

/* not named in private.h but fixed by its users */
#define FOSPHOR_B200_N_BINS     128   /* display.cl:96, fosphor.c:53 */
#define FOSPHOR_B200_WF_ROWS    1024  /* cl.c:430-432, fosphor.c:52 */

struct fosphor_cl_state;
struct fosphor_gl_state;

struct fosphor {                      /* private.h:30-55 */
	struct fosphor_cl_state *cl;  /* the drop-in keeps its engine here */
	struct fosphor_gl_state *gl;

#define FLG_FOSPHOR_USE_CLGL_SHARING (1 << 0)
	int flags;

	float fft_win[FOSPHOR_FFT_LEN];

	float *img_waterfall;         /* [1024][1024] f32, fosphor.c:52 */
	float *img_histogram;         /* [128][1024]  f32, fosphor.c:53 */
	float *buf_spectrum;          /* 2 x 1024 float2,  fosphor.c:54 */

	struct {
		int db_ref;
		int db_per_div;
		float scale;
		float offset;         /* first-use clears use -offset, cl.c:415 */
	} power;

	struct {
		double center;
		double span;
	} frequency;
};

#endif
