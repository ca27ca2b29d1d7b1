// linear_bias variant 0: MMA tile / cluster / 2-SM = (256, 128, 2, 1, true), scheduler = void
// (one translation unit per instantiation so they compile in parallel)
#include "tc_gemm.h"

namespace dear_tc {

at::Tensor linear_bias_v0(const at::Tensor& x, const at::Tensor& w, const at::Tensor& bias) {
  using G = TcGemm<ColMajor, FusionBias, 256, 128, 2, 1, true, void>;
  return linear_bias_impl<G>(x, w, bias);
}

}  // namespace dear_tc
