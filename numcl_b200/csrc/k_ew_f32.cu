#include "k_elementwise.cuh"
EwKernel ew_pick_f32(int shape, int E) { return E == 4 ? pick_float<float, 4>(shape) : (E == 1 ? pick_float<float, 1>(shape) : nullptr); }
