// k_reduce_f.cu -- single- and double-float instantiations of the reduction kernels (k_reduce.cuh)
#include "k_reduce.cuh"
RedKernel red_pick_fd(int cls, int E, int op, int which) {
  if (cls == CLS_F) return E == 4 ? kget_op<float, 4>(op, which) : kget_op<float, 1>(op, which);
  return E == 2 ? kget_op<double, 2>(op, which) : kget_op<double, 1>(op, which);
}
RedKernel red_pick_fd_dot(int cls, int E, int which) {
  if (cls == CLS_F) return E == 4 ? kget_dot<float, 4>(which) : kget_dot<float, 1>(which);
  return E == 2 ? kget_dot<double, 2>(which) : kget_dot<double, 1>(which);
}
RedKernel red_staged_pick(int cls, int op) { return cls == CLS_F ? staged_pick<float>(op) : cls == CLS_D ? staged_pick<double>(op) : staged_pick<long long>(op); }
void red_dot_matvec_launch(int cls, unsigned blocks, cudaStream_t stream, const RedParams& p) {
  if (cls == CLS_F) dot_matvec_kernel<float, 4><<<blocks, 256, 0, stream>>>(p); else dot_matvec_kernel<double, 2><<<blocks, 256, 0, stream>>>(p);
}
