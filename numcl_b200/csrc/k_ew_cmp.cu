#include "k_elementwise.cuh"
EwKernel ew_pick_cmp(int shape, int cls) {
  return cls == CLS_F ? pick_cmp<float, 32>(shape) : cls == CLS_D ? pick_cmp<double, 32>(shape) : pick_cmp<long long, 32>(shape);
}
