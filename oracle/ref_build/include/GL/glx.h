/* stub GLX: no GL context exists in the headless harness */
#ifndef FOSPHOR_STUB_GLX_H
#define FOSPHOR_STUB_GLX_H
#include <GL/gl.h>
void *glXGetCurrentContext(void);
void *glXGetCurrentDisplay(void);
#endif
