#include "k_elementwise.cuh"
EwKernel ew_pick_i32(int shape, int E) { return E == 16 ? pick_int32<16>(shape) : E == 8 ? pick_int32<8>(shape) : E == 4 ? pick_int32<4>(shape) : nullptr; }
