#include "k_elementwise.cuh"
EwKernel ew_pick_cmp(int shape, int cls) {
  // cls 4: integer operands that all fit 32-bit lanes
  return cls == CLS_F ? pick_cmp<float, 32>(shape) : cls == CLS_D ? pick_cmp<double, 32>(shape) : cls == 4 ? pick_cmp<int, 32>(shape) : pick_cmp<long long, 32>(shape);
}
