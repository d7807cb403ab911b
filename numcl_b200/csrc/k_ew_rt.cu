// k_ew_rt.cu -- instantiations of the elementwise engine for the runtime-selected function families (k_elementwise.cuh: FnRt*).
#include "k_elementwise.cuh"

template <class TI, class F, int NIN> static EwKernel pick_e(int E) {
  constexpr bool wide = sizeof(TI) == 8 && !std::is_integral<TI>::value;      // double: chunks of 2
  switch (E) {
    case 1: return kern<TI, F, NIN, 1>();
    case 2: if constexpr (std::is_integral<TI>::value || wide) return kern<TI, F, NIN, 2>(); else return nullptr;
    case 4: if constexpr (!wide) return kern<TI, F, NIN, 4>(); else return nullptr;
    case 8: if constexpr (std::is_integral<TI>::value) return kern<TI, F, NIN, 8>(); else return nullptr;
    case 16: if constexpr (std::is_integral<TI>::value) return kern<TI, F, NIN, 16>(); else return nullptr;
    default: return nullptr;
  }
}
EwKernel ew_pick_rt(int family, int E) {
  switch (family) {
    case RT_BIN_I: return pick_e<long long, FnRtBinI, 2>(E);
    case RT_UN_I: return pick_e<long long, FnRtUnI, 1>(E);
    case RT_BIN_F: return pick_e<float, FnRtBinF<float>, 2>(E);
    case RT_UN_F: return pick_e<float, FnRtUnF<float>, 1>(E);
    case RT_BIN_D: return pick_e<double, FnRtBinF<double>, 2>(E);
    case RT_UN_D: return pick_e<double, FnRtUnF<double>, 1>(E);
    default: return nullptr;
  }
}
