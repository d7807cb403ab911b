#include "k_elementwise.cuh"
EwKernel ew_pick_f64(int shape, int E) { return E == 2 ? pick_float<double, 2>(shape) : (E == 1 ? pick_float<double, 1>(shape) : nullptr); }
