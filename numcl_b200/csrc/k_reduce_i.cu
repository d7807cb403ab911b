// k_reduce_i.cu -- integer instantiations over 64-bit slots (E = 2, 1) and the register row kernels (k_reduce.cuh)
#include "k_reduce.cuh"
RedKernel red_pick_int_wide(int E, int op, int which) { return E == 2 ? kget_op<long long, 2>(op, which) : kget_op<long long, 1>(op, which); }
RedKernel red_pick_int_wide_dot(int E, int which) { return E == 2 ? kget_dot<long long, 2>(which) : kget_dot<long long, 1>(which); }
RedKernel red_rows_reg_pick(int cls, int slot_bits, int op, int c) {
  return cls == CLS_F ? rows_reg_pick<float, 4>(op, c) : cls == CLS_D ? rows_reg_pick<double, 2>(op, c)
       : slot_bits == 32 ? rows_reg_pick<long long, 4>(op, c) : rows_reg_pick<long long, 2>(op, c);
}
