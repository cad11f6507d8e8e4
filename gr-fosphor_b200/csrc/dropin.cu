/*
 * dropin.cu - the reference's compute boundary on top of the CUDA engine.
 *
 * Exports exactly the seven symbols of lib/fosphor/cl.h:22-32 with the
 * behaviour of lib/fosphor/cl.c:795-1089, so libfosphor can link this library
 * in place of cl.c + cl_compat.c + the two OpenCL programs.
 * The engine handle is stored in self->cl (cl.c:810); CL/GL sharing is never
 * advertised, so fosphor_init() allocates the host result images and
 * fosphor_gl_refresh() uploads them as in the reference's non-shared mode
 * (fosphor.c:50-62, gl.c:342-349).
 */
#include <cerrno>
#include <new>
#include <cstdio>
#include <cstdlib>

#include "../../include/fosphor_b200.h"
#include "../../include/fosphor_private_abi.h"

namespace {

struct DropinState {
	fosphor_cu *eng;
	float *fft_win;          /* cl.c:84-85: pointer retained until next process */
	int fft_win_updated;
};

inline DropinState *st(struct fosphor *self)
{
	return reinterpret_cast<DropinState *>(self->cl);
}

} /* namespace */

extern "C" {

int fosphor_cl_init(struct fosphor *self)
{
	DropinState *s = new (std::nothrow) DropinState();
	if (!s)
		return -ENOMEM;                    /* cl.c:808-809 */
	self->cl = reinterpret_cast<struct fosphor_cl_state *>(s);

	fosphor_cu_params p;
	fosphor_cu_default_params(&p);             /* N=1024, 128 bins, 1024 rows, 16/1024 */
	p.scratch_rows = -1;                       /* one call per launch pair: the waterfall is the ring */
	/* cl.c:279 lets FOSPHOR_CL_DEV=<platform>:<device> pick the OpenCL device; here
	 * FOSPHOR_CUDA_DEV=<ordinal> picks the CUDA device (default: the current one) */
	if (const char *dev = getenv("FOSPHOR_CUDA_DEV"))
		p.device = atoi(dev);
	int rc = fosphor_cu_create(&s->eng, &p);
	if (rc) {
		fprintf(stderr, "[!] No suitable CUDA device / engine init failed (%d)\n", rc);
		fosphor_cl_release(self);          /* cl.c:839-842 */
		return -EIO;
	}
	fprintf(stderr, "[+] fosphor_b200: CUDA engine ready (sm_100a)\n");   /* cl.c:824 */
	self->flags &= ~FLG_FOSPHOR_USE_CLGL_SHARING;
	return 0;
}

void fosphor_cl_release(struct fosphor *self)
{
	DropinState *s = st(self);
	if (!s)                                    /* cl.c:850-852 */
		return;
	fosphor_cu_destroy(s->eng);
	delete s;
	self->cl = nullptr;                        /* cl.c:867 */
}

int fosphor_cl_process(struct fosphor *self, void *samples, int len)
{
	DropinState *s = st(self);

	/* cl.c:881-886 (before any side effect, like the reference) */
	if (len & ((FOSPHOR_FFT_MULT_BATCH * FOSPHOR_FFT_LEN) - 1))
		return -EINVAL;
	if (len > (FOSPHOR_FFT_LEN * FOSPHOR_FFT_MAX_BATCH))
		return -EINVAL;

	if (s->fft_win_updated) {                  /* cl.c:889-900 */
		if (fosphor_cu_load_fft_window(s->eng, s->fft_win))
			return -EIO;
		s->fft_win_updated = 0;
	}
	int rc = fosphor_cu_process_host(s->eng, samples, len);
	return rc == 0 ? 0 : (rc == -EINVAL ? -EINVAL : -EIO);
}

int fosphor_cl_finish(struct fosphor *self)
{
	DropinState *s = st(self);
	/* self->img_* are the caller's persistent images (fosphor.c:52-54): only the waterfall rows
	 * written since the last finish travel (cl.c:1012-1021 re-reads the whole ring every frame) */
	int rc = fosphor_cu_finish_new_rows(s->eng, self->img_waterfall, self->img_histogram, self->buf_spectrum,
	                                    nullptr, nullptr);
	return rc < 0 ? -EIO : rc;                 /* cl.c:1057,1060 */
}

void fosphor_cl_load_fft_window(struct fosphor *self, float *win)
{
	DropinState *s = st(self);
	s->fft_win = win;                          /* cl.c:1069-1070 */
	s->fft_win_updated = 1;
}

int fosphor_cl_get_waterfall_position(struct fosphor *self)
{
	return fosphor_cu_get_waterfall_position(st(self)->eng);
}

void fosphor_cl_set_histogram_range(struct fosphor *self, float scale, float offset)
{
	fosphor_cu_set_histogram_range(st(self)->eng, scale, offset);   /* scale * 128, cl.c:1087 */
}

} /* extern "C" */
