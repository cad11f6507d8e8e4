/*
 * gl_stub.c - headless stand-in for the reference's OpenGL renderer
 * (lib/fosphor/gl.c, interface lib/fosphor/gl.h:25-39), so that the reference's
 * UNMODIFIED lib/fosphor/fosphor.c can be linked against libfosphor_b200.so and
 * driven without a GL context.  TEST INFRASTRUCTURE ONLY (oracle/ref_link/):
 * link-level proof that the CUDA library is a drop-in for cl.c behind
 * fosphor.c.  What the renderer would do with the results is
 * out of scope; the stub only records that it was asked to.
 */
#include <stdlib.h>

#include "gl.h"
#include "private.h"

struct fosphor_gl_state {
	int refreshes; /* fosphor_gl_refresh(): "upload img_* to the textures" (gl.c:337-350) */
	int draws;
};

int
fosphor_gl_init(struct fosphor *self)
{
	self->gl = calloc(1, sizeof(struct fosphor_gl_state));
	return self->gl ? 0 : -1;
}

void
fosphor_gl_release(struct fosphor *self)
{
	free(self->gl);
	self->gl = NULL;
}

GLuint
fosphor_gl_get_shared_id(struct fosphor *self, enum fosphor_gl_id id)
{
	(void)self; (void)id;
	return 0; /* never asked for: the drop-in leaves FLG_FOSPHOR_USE_CLGL_SHARING clear */
}

void
fosphor_gl_refresh(struct fosphor *self)
{
	self->gl->refreshes++;
}

void
fosphor_gl_draw(struct fosphor *self, struct fosphor_render *render)
{
	(void)render;
	self->gl->draws++;
}

/* test accessors */
int fosphor_stub_refreshes(struct fosphor *self) { return self->gl ? self->gl->refreshes : -1; }
int fosphor_stub_draws(struct fosphor *self) { return self->gl ? self->gl->draws : -1; }
