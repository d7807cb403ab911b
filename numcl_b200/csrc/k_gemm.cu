// k_gemm.cu -- K4 dispatch (contractions).  Kernel families land in k_gemm_*.cu.
#include "ncl_internal.h"
int try_gemm(const Nest&) { return NCL_ERR_UNSUPPORTED; }
