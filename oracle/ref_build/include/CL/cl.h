/*
 * Minimal OpenCL declarations (Khronos OpenCL 1.2 API, public standard) --
 * only what the reference's lib/fosphor/cl.c and cl_compat.c use.  The build
 * container has no OpenCL headers; this lets the UNMODIFIED reference host
 * driver be compiled (see ../Makefile).  Test infrastructure only.
 */
#ifndef FOSPHOR_STUB_CL_H
#define FOSPHOR_STUB_CL_H

#include <stddef.h>
#include <stdint.h>

#define CL_API_ENTRY
#define CL_API_CALL
#define CL_VERSION_1_0 1
#define CL_VERSION_1_1 1
#define CL_VERSION_1_2 1

typedef int32_t  cl_int;
typedef uint32_t cl_uint;
typedef uint64_t cl_ulong;
typedef float    cl_float;
typedef cl_uint  cl_bool;
typedef cl_ulong cl_bitfield;
typedef cl_bitfield cl_device_type;
typedef cl_bitfield cl_mem_flags;
typedef cl_bitfield cl_command_queue_properties;
typedef cl_uint  cl_platform_info;
typedef cl_uint  cl_device_info;
typedef cl_uint  cl_image_info;
typedef cl_uint  cl_program_info;
typedef cl_uint  cl_program_build_info;
typedef cl_uint  cl_channel_order;
typedef cl_uint  cl_channel_type;
typedef cl_uint  cl_mem_object_type;
typedef intptr_t cl_context_properties;

typedef struct _cl_platform_id   *cl_platform_id;
typedef struct _cl_device_id     *cl_device_id;
typedef struct _cl_context       *cl_context;
typedef struct _cl_command_queue *cl_command_queue;
typedef struct _cl_mem           *cl_mem;
typedef struct _cl_program       *cl_program;
typedef struct _cl_kernel        *cl_kernel;
typedef struct _cl_event         *cl_event;

typedef struct _cl_image_format {
	cl_channel_order image_channel_order;
	cl_channel_type  image_channel_data_type;
} cl_image_format;

typedef struct _cl_image_desc {
	cl_mem_object_type image_type;
	size_t image_width;
	size_t image_height;
	size_t image_depth;
	size_t image_array_size;
	size_t image_row_pitch;
	size_t image_slice_pitch;
	cl_uint num_mip_levels;
	cl_uint num_samples;
	cl_mem buffer;
} cl_image_desc;

#define CL_SUCCESS                      0
#define CL_OUT_OF_RESOURCES             -5
#define CL_IMAGE_FORMAT_NOT_SUPPORTED   -10
#define CL_INVALID_VALUE                -30

#define CL_FALSE 0
#define CL_TRUE  1

#define CL_PLATFORM_VERSION             0x0901

#define CL_DEVICE_TYPE_GPU              (1 << 2)
#define CL_DEVICE_TYPE_ALL              0xFFFFFFFF

#define CL_DEVICE_TYPE                  0x1000
#define CL_DEVICE_IMAGE_SUPPORT         0x1016
#define CL_DEVICE_LOCAL_MEM_SIZE        0x1023
#define CL_DEVICE_NAME                  0x102B
#define CL_DEVICE_VENDOR                0x102C
#define CL_DEVICE_VERSION               0x102F
#define CL_DEVICE_EXTENSIONS            0x1030

#define CL_CONTEXT_PLATFORM             0x1084

#define CL_MEM_READ_WRITE               (1 << 0)
#define CL_MEM_WRITE_ONLY               (1 << 1)
#define CL_MEM_READ_ONLY                (1 << 2)

#define CL_R                            0x10B0
#define CL_FLOAT                        0x10DE
#define CL_MEM_OBJECT_IMAGE2D           0x10F1
#define CL_IMAGE_FORMAT                 0x1110

#define CL_PROGRAM_BINARY_SIZES         0x1165
#define CL_PROGRAM_BINARIES             0x1166
#define CL_PROGRAM_BUILD_LOG            0x1183

cl_int clGetPlatformIDs(cl_uint, cl_platform_id *, cl_uint *);
cl_int clGetPlatformInfo(cl_platform_id, cl_platform_info, size_t, void *, size_t *);
cl_int clGetDeviceIDs(cl_platform_id, cl_device_type, cl_uint, cl_device_id *, cl_uint *);
cl_int clGetDeviceInfo(cl_device_id, cl_device_info, size_t, void *, size_t *);
cl_context clCreateContext(const cl_context_properties *, cl_uint, const cl_device_id *,
	void (*)(const char *, const void *, size_t, void *), void *, cl_int *);
cl_int clReleaseContext(cl_context);
cl_command_queue clCreateCommandQueue(cl_context, cl_device_id, cl_command_queue_properties, cl_int *);
cl_int clReleaseCommandQueue(cl_command_queue);
cl_mem clCreateBuffer(cl_context, cl_mem_flags, size_t, void *, cl_int *);
cl_mem clCreateImage2D(cl_context, cl_mem_flags, const cl_image_format *, size_t, size_t, size_t, void *, cl_int *);
cl_mem clCreateImage(cl_context, cl_mem_flags, const cl_image_format *, const cl_image_desc *, void *, cl_int *);
cl_int clReleaseMemObject(cl_mem);
cl_int clGetImageInfo(cl_mem, cl_image_info, size_t, void *, size_t *);
cl_program clCreateProgramWithSource(cl_context, cl_uint, const char **, const size_t *, cl_int *);
cl_int clBuildProgram(cl_program, cl_uint, const cl_device_id *, const char *,
	void (*)(cl_program, void *), void *);
cl_int clGetProgramBuildInfo(cl_program, cl_device_id, cl_program_build_info, size_t, void *, size_t *);
cl_int clGetProgramInfo(cl_program, cl_program_info, size_t, void *, size_t *);
cl_int clReleaseProgram(cl_program);
cl_kernel clCreateKernel(cl_program, const char *, cl_int *);
cl_int clReleaseKernel(cl_kernel);
cl_int clSetKernelArg(cl_kernel, cl_uint, size_t, const void *);
cl_int clEnqueueNDRangeKernel(cl_command_queue, cl_kernel, cl_uint, const size_t *, const size_t *,
	const size_t *, cl_uint, const cl_event *, cl_event *);
cl_int clEnqueueWriteBuffer(cl_command_queue, cl_mem, cl_bool, size_t, size_t, const void *,
	cl_uint, const cl_event *, cl_event *);
cl_int clEnqueueReadBuffer(cl_command_queue, cl_mem, cl_bool, size_t, size_t, void *,
	cl_uint, const cl_event *, cl_event *);
cl_int clEnqueueFillBuffer(cl_command_queue, cl_mem, const void *, size_t, size_t, size_t,
	cl_uint, const cl_event *, cl_event *);
cl_int clEnqueueReadImage(cl_command_queue, cl_mem, cl_bool, const size_t *, const size_t *,
	size_t, size_t, void *, cl_uint, const cl_event *, cl_event *);
cl_int clEnqueueWriteImage(cl_command_queue, cl_mem, cl_bool, const size_t *, const size_t *,
	size_t, size_t, const void *, cl_uint, const cl_event *, cl_event *);
cl_int clEnqueueFillImage(cl_command_queue, cl_mem, const void *, const size_t *, const size_t *,
	cl_uint, const cl_event *, cl_event *);
cl_int clFinish(cl_command_queue);

#endif
