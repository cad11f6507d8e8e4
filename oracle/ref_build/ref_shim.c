/*
 * ref_shim.c - glue that lets the UNMODIFIED reference host driver
 * (lib/fosphor/cl.c + cl_compat.c, compiled from /root/reference where they
 * lie) run headless on a box that has an OpenCL vendor library but no ICD
 * registration, no GL and no headers.  TEST INFRASTRUCTURE ONLY: the result
 * (oracle/_ref/libfosphor_ref.so) is the "reference itself" arm used to pin
 * the CPU oracle and as the bench's reference baseline.
 *
 * Provides:
 *   - cl* entry points as dlopen() trampolines into the vendor library
 *     (the Khronos loader libOpenCL.so.1 pointed at libnvidia-opencl.so.1 through
 *     OCL_ICD_FILENAMES on the GPU box; override with FOSPHOR_REF_OPENCL_LIB)
 *   - resource_get()/resource_put() (reference API lib/fosphor/resource.h:20-21)
 *     serving the kernel text packed at build time by pack_kernels.py
 *   - fosphor_gl_get_shared_id() / glX stubs (never used: gl sharing is hidden
 *     from the device extension string so cl.c takes its non-shared path,
 *     cl.c:626-642 / cl_init_buffers_nogl :507-568)
 */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <CL/cl.h>
#include <CL/cl_gl.h>
#include <GL/glx.h>

/* ---- vendor library ------------------------------------------------------ */

static void *g_lib;

static void *
ocl_sym(const char *name)
{
	void *p;

	if (!g_lib) {
		/* The GPU box has the vendor library (libnvidia-opencl.so.1, an ICD
		 * that only exports the clIcd* entry points) and the Khronos loader
		 * from the CUDA toolkit (libOpenCL.so.1), but no /etc/OpenCL/vendors
		 * registration: point the loader at the vendor library directly. */
		const char *cands[] = {
			getenv("FOSPHOR_REF_OPENCL_LIB"),
			"libOpenCL.so.1",
			"/usr/local/cuda/lib64/libOpenCL.so.1",
			"libOpenCL.so",
		};
		setenv("OCL_ICD_FILENAMES", "libnvidia-opencl.so.1", 0);
		for (unsigned i = 0; i < sizeof(cands) / sizeof(cands[0]) && !g_lib; i++)
			if (cands[i] && cands[i][0])
				g_lib = dlopen(cands[i], RTLD_NOW | RTLD_LOCAL);
		if (!g_lib) {
			fprintf(stderr, "[ref_shim] no OpenCL library could be opened: %s\n", dlerror());
			return NULL;
		}
	}
	p = dlsym(g_lib, name);
	if (!p)
		fprintf(stderr, "[ref_shim] OpenCL symbol %s not found\n", name);
	return p;
}

int
fosphor_ref_opencl_available(void)
{
	return ocl_sym("clGetPlatformIDs") != NULL;
}

#define CL_ERR_NOLIB (-1001) /* CL_PLATFORM_NOT_FOUND_KHR */

#define TRAMP_INT(name, params, args)                                  \
	cl_int name params                                             \
	{                                                              \
		static cl_int (*fn) params;                            \
		if (!fn) *(void **)&fn = ocl_sym(#name);               \
		if (!fn) return CL_ERR_NOLIB;                          \
		return fn args;                                        \
	}

#define TRAMP_OBJ(rtype, name, params, args, errp)                     \
	rtype name params                                              \
	{                                                              \
		static rtype (*fn) params;                             \
		if (!fn) *(void **)&fn = ocl_sym(#name);               \
		if (!fn) { if (errp) *(errp) = CL_ERR_NOLIB; return NULL; } \
		return fn args;                                        \
	}

TRAMP_INT(clGetPlatformIDs, (cl_uint a, cl_platform_id *b, cl_uint *c), (a, b, c))
TRAMP_INT(clGetPlatformInfo, (cl_platform_id a, cl_platform_info b, size_t c, void *d, size_t *e), (a, b, c, d, e))
TRAMP_INT(clGetDeviceIDs, (cl_platform_id a, cl_device_type b, cl_uint c, cl_device_id *d, cl_uint *e), (a, b, c, d, e))
TRAMP_OBJ(cl_context, clCreateContext,
	(const cl_context_properties *a, cl_uint b, const cl_device_id *c,
	 void (*d)(const char *, const void *, size_t, void *), void *e, cl_int *f),
	(a, b, c, d, e, f), f)
TRAMP_INT(clReleaseContext, (cl_context a), (a))
TRAMP_OBJ(cl_command_queue, clCreateCommandQueue,
	(cl_context a, cl_device_id b, cl_command_queue_properties c, cl_int *d), (a, b, c, d), d)
TRAMP_INT(clReleaseCommandQueue, (cl_command_queue a), (a))
TRAMP_OBJ(cl_mem, clCreateBuffer, (cl_context a, cl_mem_flags b, size_t c, void *d, cl_int *e), (a, b, c, d, e), e)
TRAMP_OBJ(cl_mem, clCreateImage2D,
	(cl_context a, cl_mem_flags b, const cl_image_format *c, size_t d, size_t e, size_t f, void *g, cl_int *h),
	(a, b, c, d, e, f, g, h), h)
TRAMP_INT(clReleaseMemObject, (cl_mem a), (a))
TRAMP_INT(clGetImageInfo, (cl_mem a, cl_image_info b, size_t c, void *d, size_t *e), (a, b, c, d, e))
TRAMP_OBJ(cl_program, clCreateProgramWithSource,
	(cl_context a, cl_uint b, const char **c, const size_t *d, cl_int *e), (a, b, c, d, e), e)
TRAMP_INT(clBuildProgram,
	(cl_program a, cl_uint b, const cl_device_id *c, const char *d, void (*e)(cl_program, void *), void *f),
	(a, b, c, d, e, f))
TRAMP_INT(clGetProgramBuildInfo,
	(cl_program a, cl_device_id b, cl_program_build_info c, size_t d, void *e, size_t *f), (a, b, c, d, e, f))
TRAMP_INT(clGetProgramInfo, (cl_program a, cl_program_info b, size_t c, void *d, size_t *e), (a, b, c, d, e))
TRAMP_INT(clReleaseProgram, (cl_program a), (a))
TRAMP_OBJ(cl_kernel, clCreateKernel, (cl_program a, const char *b, cl_int *c), (a, b, c), c)
TRAMP_INT(clReleaseKernel, (cl_kernel a), (a))
TRAMP_INT(clSetKernelArg, (cl_kernel a, cl_uint b, size_t c, const void *d), (a, b, c, d))
TRAMP_INT(clEnqueueNDRangeKernel,
	(cl_command_queue a, cl_kernel b, cl_uint c, const size_t *d, const size_t *e, const size_t *f,
	 cl_uint g, const cl_event *h, cl_event *i), (a, b, c, d, e, f, g, h, i))
TRAMP_INT(clEnqueueWriteBuffer,
	(cl_command_queue a, cl_mem b, cl_bool c, size_t d, size_t e, const void *f,
	 cl_uint g, const cl_event *h, cl_event *i), (a, b, c, d, e, f, g, h, i))
TRAMP_INT(clEnqueueReadBuffer,
	(cl_command_queue a, cl_mem b, cl_bool c, size_t d, size_t e, void *f,
	 cl_uint g, const cl_event *h, cl_event *i), (a, b, c, d, e, f, g, h, i))
TRAMP_INT(clEnqueueReadImage,
	(cl_command_queue a, cl_mem b, cl_bool c, const size_t *d, const size_t *e, size_t f, size_t g,
	 void *h, cl_uint i, const cl_event *j, cl_event *k), (a, b, c, d, e, f, g, h, i, j, k))
TRAMP_INT(clEnqueueWriteImage,
	(cl_command_queue a, cl_mem b, cl_bool c, const size_t *d, const size_t *e, size_t f, size_t g,
	 const void *h, cl_uint i, const cl_event *j, cl_event *k), (a, b, c, d, e, f, g, h, i, j, k))
TRAMP_INT(clFinish, (cl_command_queue a), (a))
TRAMP_OBJ(cl_mem, clCreateFromGLBuffer, (cl_context a, cl_mem_flags b, GLuint c, cl_int *d), (a, b, c, d), d)
TRAMP_OBJ(cl_mem, clCreateFromGLTexture2D,
	(cl_context a, cl_mem_flags b, GLenum c, GLint d, GLuint e, cl_int *f), (a, b, c, d, e, f), f)
TRAMP_INT(clEnqueueAcquireGLObjects,
	(cl_command_queue a, cl_uint b, const cl_mem *c, cl_uint d, const cl_event *e, cl_event *f), (a, b, c, d, e, f))
TRAMP_INT(clEnqueueReleaseGLObjects,
	(cl_command_queue a, cl_uint b, const cl_mem *c, cl_uint d, const cl_event *e, cl_event *f), (a, b, c, d, e, f))

/* clGetDeviceInfo: forwarded, but cl_khr_gl_sharing is blanked out of the
 * extension list so the reference picks its own non-shared (host read-back)
 * path -- there is no GL context in this harness. */
cl_int
clGetDeviceInfo(cl_device_id dev, cl_device_info param, size_t sz, void *val, size_t *ret)
{
	static cl_int (*fn)(cl_device_id, cl_device_info, size_t, void *, size_t *);
	cl_int err;

	if (!fn) *(void **)&fn = ocl_sym("clGetDeviceInfo");
	if (!fn) return CL_ERR_NOLIB;
	err = fn(dev, param, sz, val, ret);
	if (err == CL_SUCCESS && param == CL_DEVICE_EXTENSIONS && val && sz) {
		char *s = val, *p;
		s[sz - 1] = 0;
		while ((p = strstr(s, "_gl_sharing")) != NULL)
			memcpy(p, "_XX_XXXXXXX", 11);
	}
	return err;
}

/* ---- packed kernel text (generated by pack_kernels.py) ------------------- */

struct fosphor_ref_blob {
	const char *name;
	const unsigned char *data;
	int len;
};
extern const struct fosphor_ref_blob fosphor_ref_blobs[];

const void *
resource_get(const char *name, int *len)
{
	for (const struct fosphor_ref_blob *b = fosphor_ref_blobs; b->name; b++) {
		if (!strcmp(b->name, name)) {
			if (len)
				*len = b->len;
			return b->data; /* NUL terminated by the packer */
		}
	}
	return NULL;
}

void
resource_put(const void *r)
{
	(void)r;
}

/* ---- GL stubs ------------------------------------------------------------ */

struct fosphor;

GLuint
fosphor_gl_get_shared_id(struct fosphor *self, int id)
{
	(void)self; (void)id;
	return 0;
}

void *glXGetCurrentContext(void) { return NULL; }
void *glXGetCurrentDisplay(void) { return NULL; }
