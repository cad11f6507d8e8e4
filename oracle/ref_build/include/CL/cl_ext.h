/* stub: nothing from cl_ext.h is needed beyond what cl_compat.h defines itself */
#ifndef FOSPHOR_STUB_CL_EXT_H
#define FOSPHOR_STUB_CL_EXT_H
#endif
