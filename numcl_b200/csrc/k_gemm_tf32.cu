// k_gemm_tf32.cu -- placeholder until the tcgen05 3xTF32 kernel lands (next commit)
#include "k_gemm.cuh"
int gemm_f32_tf32x3(const GemmParams&) { return NCL_ERR_UNSUPPORTED; }
