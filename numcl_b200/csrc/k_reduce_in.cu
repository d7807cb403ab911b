// k_reduce_in.cu -- integer instantiations over 8/16/32-bit slots (E = 16, 8, 4) (k_reduce.cuh)
#include "k_reduce.cuh"
RedKernel red_pick_int_narrow(int E, int op, int which) {
  return E == 16 ? kget_op<long long, 16>(op, which) : E == 8 ? kget_op<long long, 8>(op, which) : kget_op<long long, 4>(op, which);
}
RedKernel red_pick_int_narrow_dot(int E, int which) {
  return E == 16 ? kget_dot<long long, 16>(which) : E == 8 ? kget_dot<long long, 8>(which) : kget_dot<long long, 4>(which);
}
