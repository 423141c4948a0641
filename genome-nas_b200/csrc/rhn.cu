// placeholder until the RHN kernels land
#include "common.cuh"
#include "../../include/gnas.h"
using namespace gnas;
extern "C" int gnas_rhn_workspace(const GnasRhnDesc*, int64_t*, int64_t*) { set_error("rhn: not built"); return GNAS_ERR_UNSUPPORTED; }
extern "C" int gnas_rhn_forward(const GnasRhnDesc*, const float*, const float*, const float*, const float*, const float*, const float*, const float*, float*, float*, float*, void*) { set_error("rhn: not built"); return GNAS_ERR_UNSUPPORTED; }
extern "C" int gnas_rhn_backward(const GnasRhnDesc*, const float*, const float*, const float*, const float*, const float*, const float*, const float*, const float*, float*, float*, float*, float*, float*, float*, float*, int, void*) { set_error("rhn: not built"); return GNAS_ERR_UNSUPPORTED; }
