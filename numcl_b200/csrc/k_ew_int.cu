#include "k_elementwise.cuh"
EwKernel ew_pick_int(int shape, int E) {
  return E == 16 ? pick_int<16>(shape) : E == 8 ? pick_int<8>(shape) : E == 4 ? pick_int<4>(shape) : E == 2 ? pick_int<2>(shape) : E == 1 ? pick_int<1>(shape) : nullptr;
}
EwKernel ew_pick_idiv(int E) { return E == 4 ? kern<long long, FnIDiv, 2, 4>() : (E == 1 ? kern<long long, FnIDiv, 2, 1>() : nullptr); }
